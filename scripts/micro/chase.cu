// Dependent-load latency vs footprint (TLB reach): one thread chases hashed indices.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void chase(const uint32_t* buf, size_t n_words, int iters, uint32_t* out, long long* cycles) {
  uint64_t idx = 12345;
  long long t0 = clock64();
  uint32_t acc = 0;
  for (int i = 0; i < iters; i++) {
    uint32_t v = __ldcg(buf + idx);
    acc += v;
    uint64_t h = (idx + v + 0x9E3779B97F4A7C15ull) * 0xBF58476D1CE4E5B9ull;
    h ^= h >> 29;
    idx = h % n_words;
  }
  long long t1 = clock64();
  out[0] = acc;
  cycles[0] = t1 - t0;
}
int main() {
  size_t sizes_mb[] = {32, 128, 512, 2048, 16384, 65536, 120000};
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 4); cudaMalloc(&cyc, 8);
  for (size_t mb : sizes_mb) {
    uint32_t* buf;
    size_t bytes = mb << 20;
    if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("%zu MB: alloc failed\n", mb); continue; }
    cudaMemset(buf, 0, bytes);
    int iters = 20000;
    chase<<<1, 1>>>(buf, bytes / 4, iters, out, cyc);
    chase<<<1, 1>>>(buf, bytes / 4, iters, out, cyc);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%8zu MB: %.0f cycles per dependent load (incl ~40 cycles of hash)\n", mb, (double)c / iters);
    cudaFree(buf);
  }
  return 0;
}
