// landmass_b200 — pieces shared by the A* kernels (k_astar.cu, k_astar_quad.cu): the exact
// emulation of Rust's BinaryHeap<Reverse<NodeRef>> (astar.rs:56-83),
// edge-record loads and the per-query phase codes.
#pragma once
#include "lmb_device.cuh"
#include "lmb_kernels.h"

namespace lmb {
namespace astar_detail {

constexpr float F_INF = __builtin_huge_valf();
constexpr uint32_t FULL = 0xffffffffu;

struct HeapEntry {
  float cost, estimate;
  uint32_t index;
  uint32_t state;
  uint32_t depth;
};

// NodeRef::partial_cmp (astar.rs:67-77) as `a <= b`
__device__ __forceinline__ bool noderef_le(const HeapEntry& a, const HeapEntry& b) {
  // Less, or Equal-estimates with Reverse(a.cost) <= Reverse(b.estimate); unordered (NaN) -> false
  return (a.estimate < b.estimate) | ((a.estimate == b.estimate) & (b.estimate <= a.cost));
}
// Reverse<T>::le(x, y) = y.0 <= x.0
__device__ __forceinline__ bool rev_le(const HeapEntry& x, const HeapEntry& y) { return noderef_le(y, x); }

// Open list.  The SP = false members touch shared memory only (valid while every index they
// can reach is < AT_HEAP_SM); SP = true adds the in-place spill to the slot's arena.  The
// caller picks per operation from the current length, so the common case carries no
// per-access branch.
template <uint32_t STRIDE>
struct Heap {
  uint4* sm;      // this query's column: entry i at sm[i * STRIDE]
  uint32_t* smd;  // depth column
  uint4* gm;      // spill: entry i (>= AT_HEAP_SM) at gm[i - AT_HEAP_SM]
  uint32_t* gmd;
  uint32_t len;
  template <bool SP>
  __device__ __forceinline__ HeapEntry get(uint32_t i) const {
    uint4 v;
    uint32_t d;
    if (!SP || i < AT_HEAP_SM) {
      v = sm[i * STRIDE];
      d = smd[i * STRIDE];
    } else {
      v = gm[i - AT_HEAP_SM];
      d = gmd[i - AT_HEAP_SM];
    }
    return {__uint_as_float(v.x), __uint_as_float(v.y), v.z, v.w, d};
  }
  template <bool SP>
  __device__ __forceinline__ void set(uint32_t i, const HeapEntry& e) {
    uint4 v = make_uint4(__float_as_uint(e.cost), __float_as_uint(e.estimate), e.index, e.state);
    if (!SP || i < AT_HEAP_SM) {
      sm[i * STRIDE] = v;
      smd[i * STRIDE] = e.depth;
    } else {
      gm[i - AT_HEAP_SM] = v;
      gmd[i - AT_HEAP_SM] = e.depth;
    }
  }
  template <bool SP>
  __device__ __forceinline__ void sift_up(uint32_t pos, const HeapEntry& hole) {
    while (pos > 0) {
      uint32_t parent = (pos - 1) >> 1;
      HeapEntry p = get<SP>(parent);
      if (rev_le(hole, p)) break;
      set<SP>(pos, p);
      pos = parent;
    }
    set<SP>(pos, hole);
  }
  template <bool SP>
  __device__ __forceinline__ void push_t(const HeapEntry& e) {
    uint32_t old_len = len++;
    sift_up<SP>(old_len, e);
  }
  __device__ __forceinline__ void push(const HeapEntry& e) {
    if (len < AT_HEAP_SM)
      push_t<false>(e);
    else
      push_t<true>(e);
  }
  template <bool SP>
  __device__ __forceinline__ HeapEntry pop_t() {
    HeapEntry item = get<SP>(--len);
    if (len > 0) {
      HeapEntry top = get<SP>(0);
      HeapEntry hole = item;  // swap(item, data[0]); the hole element is the old last
      item = top;
      uint32_t end = len, pos = 0, child = 1;
      while (child + 1 < end) {  // child <= end - 2
        HeapEntry a = get<SP>(child), b = get<SP>(child + 1);
        bool right = rev_le(a, b);
        set<SP>(pos, right ? b : a);
        pos = child + (right ? 1u : 0u);
        child = 2 * pos + 1;
      }
      if (child + 1 == end) {
        set<SP>(pos, get<SP>(child));
        pos = child;
      }
      sift_up<SP>(pos, hole);
    }
    return item;
  }
  __device__ __forceinline__ HeapEntry pop() {
    if (len <= AT_HEAP_SM) return pop_t<false>();
    return pop_t<true>();
  }
};

template <class T>
__device__ __forceinline__ T sel4(uint32_t j, T a0, T a1, T a2, T a3) {
  return j == 0 ? a0 : (j == 1 ? a1 : (j == 2 ? a2 : a3));
}

__device__ __forceinline__ EdgeRec ld_rec(const EdgeRec* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  EdgeRec r;
  r.mx = __uint_as_float(a.x);
  r.my = __uint_as_float(a.y);
  r.mz = __uint_as_float(a.z);
  r.twin = a.w;
  r.poly = b.x;
  r.first_edge = b.y;
  r.info = b.z;
  r.aux = b.w;
  return r;
}

enum : uint32_t { PH_IDLE = 0, PH_SEARCH = 1, PH_WALK = 2, PH_FINISH = 3, PH_EXIT = 4 };


}  // namespace astar_detail
}  // namespace lmb
