#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native landmass agent step.

Metric (BASELINE.json): agent-updates/s of one `Archipelago::update`, repaths included.
Workload at N=1 (configs[1]): 256x256-room grid maze (131 071 polygons), 100 000 agents with
uniform random positions/targets; every timed step is a full update in which EVERY agent
re-plans (the device-side path is dropped before each step), i.e. 100k point samples x2,
100k A* queries, funnel + reached test, and ORCA avoidance.  N>1: the same per GPU (weak
scaling), agents sharded by contiguous ranges, navmesh replicated, one NCCL all-gather of
the 32-byte avoidance records per step.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

`value`  : device-timed throughput with inputs resident in HBM (CUDA events on the step stream).
`e2e`    : the same step through lmb_step with HOST buffers (H2D + D2H inside the timed region).
`roofline`: the dominant kernel (k_astar) against the measured HBM copy bandwidth.
`cpu_baseline`: the CPU oracle (single-threaded C++ port of the reference; the Rust crate
            cannot be built here) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROOMS = 256
AGENTS_PER_GPU = 100_000
DT = 1.0 / 60.0
B_STEP = 240.0  # algorithmic bytes per steady agent-update
# DRAM bytes per explored node of k_astar, from the one `ncu --set full` capture of a fully loaded
# launch (profiles/r01_ncu_astar_loaded.md: dram read + write / explored nodes).  The bench-size
# launch itself cannot be captured: kernel replay would have to save and restore its ~100 GB arena.
NCU_DRAM_BYTES_PER_EXPLORED = 114.4


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.lines = []
        self.proc = None
        self.device_index = device_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device_index}", f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_workload(rooms: int, n_agents: int, rank: int):
    from landmass_b200 import synth

    mesh = synth.maze_mesh(rooms)
    flat = synth.single_island_flat(mesh)
    ag = synth.make_agents(mesh, n_agents, seed=synth.SEED + 7919 * rank)
    return mesh, flat, ag


def run_reference(args):
    """`--impl reference`: the reference's own CPU path.  The Rust crate cannot be compiled in
    this image (no cargo/rustc), so this times the oracle port (single-threaded, like the
    reference) on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from landmass_b200 import synth
    from oracle.oracle_py import OracleBackend

    sample = args.ref_sample
    mesh, flat, ag = build_workload(args.rooms, sample, 0)
    cpu = OracleBackend()
    cpu.navdata_upload(flat)
    opts = synth.default_options()
    times = []
    for step in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        cpu.step(ag, None, opts, DT)  # RESET flag set: every agent re-plans
        t1 = time.perf_counter()
        if step >= args.warmup:
            times.append(t1 - t0)
    ms = 1e3 * float(np.mean(times))
    value = sample / (ms * 1e-3)
    all_cores = reference_all_cores(args, sample)
    line = {
        "impl": "reference", "metric": "agent-updates/s", "value": value, "unit": "agent-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus, sample_note=f"{sample} agents of the workload per step"),
        "cpu_baseline": {"value": value, "unit": "agent-updates/s", "cores": 1, "kind": "port",
                         "sample": f"{sample} agents (every agent re-plans) per step on the full "
                                   f"{args.rooms}x{args.rooms} maze; single thread like the reference"},
        "e2e": {"value": value, "unit": "agent-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        # not reference behaviour (`Archipelago::update` is serial): what the host could do by running
        # one independent archipelago per core, for readers who want the harsher ratio
        "all_cores": all_cores,
    }
    print(json.dumps(line))


def _reference_worker(rank, rooms, sample, steps, barrier, out):
    from landmass_b200 import synth
    from oracle.oracle_py import OracleBackend

    mesh, flat, ag = build_workload(rooms, sample, rank)
    cpu = OracleBackend()
    cpu.navdata_upload(flat)
    opts = synth.default_options()
    barrier.wait()
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu.step(ag, None, opts, DT)
    out.put((rank, time.perf_counter() - t0))


def reference_all_cores(args, sample):
    """One oracle process per host core, each with its own `sample` agents on its own copy of the maze,
    started together; aggregate agent-updates/s = processes x sample x steps / slowest process."""
    import multiprocessing as mp

    procs = max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    procs = min(procs, 64)  # bounded: memory and start-up time of the forked interpreters
    steps = min(args.steps, 2)
    ctx = mp.get_context("fork")
    barrier, out = ctx.Barrier(procs), ctx.Queue()
    workers = [ctx.Process(target=_reference_worker, args=(r, args.rooms, sample, steps, barrier, out))
               for r in range(procs)]
    for w in workers:
        w.start()
    times = [out.get()[1] for _ in workers]
    for w in workers:
        w.join()
    return {"value": procs * sample * steps / max(times), "unit": "agent-updates/s", "processes": procs,
            "sample": f"{sample} agents per process, every agent re-plans, {steps} step(s)"}


def workload_config(args, n_gpus, sample_note=None):
    cfg = {
        "workload": f"configs[1]: {args.rooms}x{args.rooms} grid-maze navmesh, {args.agents} agents per GPU, "
                    "full Archipelago::update with every agent re-planning each step (sample x2, A*, "
                    "funnel, ORCA avoidance)",
        "agents_per_gpu": args.agents, "polygons": 2 * args.rooms * args.rooms - 1,
        "parallelism": f"agents sharded x{n_gpus}, navmesh replicated, NCCL all-gather of 32 B/agent",
        "cache": "inputs and A* working set (one arena slot per query in flight, ~100 GB) far exceed the "
                 "126 MB L2; no flush needed",
    }
    if sample_note:
        cfg["sample"] = sample_note
    return cfg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rooms", type=int, default=ROOMS)
    ap.add_argument("--agents", type=int, default=AGENTS_PER_GPU)
    ap.add_argument("--cpu-sample", type=int, default=1000,
                    help="agents in the CPU-baseline sample (one step; ~15-25 s)")
    ap.add_argument("--ref-sample", type=int, default=150,
                    help="agents per step of the --impl reference arm")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")

    # keep stdout to the ONE JSON line: everything else any library writes to fd 1 (NCCL prints its
    # version banner there) goes to stderr; the JSON line is written to the saved descriptor
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch  # plumbing only: device selection, NCCL all-gather, barriers

    from landmass_b200 import _abi, synth
    from landmass_b200._native import NativeBackend

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA GPU: landmass_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n = args.agents
    mesh, flat, ag = build_workload(args.rooms, n, rank)
    opts = synth.default_options()
    gpu = NativeBackend(max_agents=n, device=local_rank)
    gpu.navdata_upload(flat)

    n_total, first = n * world, n * rank

    halo = None

    def one_step_resident():
        nonlocal halo
        if world == 1:
            gpu.step_resident(opts, DT)
            return
        gpu.step_begin(opts, DT, n_total, first)
        ptr, rec_bytes, n_rec = gpu.halo_buffer()
        if halo is None or halo[0] != ptr:
            # wrap the library's device buffer (no copy); torch is only the NCCL plumbing here
            class _Buf:
                __cuda_array_interface__ = {"shape": (int(n_rec * rec_bytes),), "typestr": "|u1",
                                            "data": (int(ptr), False), "version": 2}
            halo = (ptr, torch.as_tensor(_Buf(), device=torch.device("cuda", local_rank)))
        buf = halo[1]
        mine = buf[first * rec_bytes:(first + n) * rec_bytes]
        dist.all_gather_into_tensor(buf, mine)
        torch.cuda.synchronize()
        gpu.step_end(opts, DT)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (`value`) ----
    gpu.agents_upload(ag, None)
    for _ in range(args.warmup):
        one_step_resident()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    dev_ms, astar_ms, phases = [], [], []
    explored = path_nodes = n_repath = launches = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_step_resident()
        T = gpu.step_timings()
        dev_ms.append(T.total_ms)
        astar_ms.append(T.astar_ms)
        phases.append((T.sample_ms, T.repath_decide_ms, T.astar_ms, T.follow_ms, T.avoidance_ms))
        explored += T.explored_nodes
        path_nodes += T.path_nodes
        n_repath += T.n_repath
        launches += T.kernel_launches
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0) / args.steps
    clocks = sampler.stop()
    print(f"[bench rank {rank}] per-step device ms: {[round(x, 1) for x in dev_ms]}", file=sys.stderr)
    out = gpu.agents_download()
    step_ms = float(np.mean(dev_ms)) if world == 1 else wall_ms  # multi-GPU: wall between barriers incl. all-gather
    red = torch.tensor([step_ms, float(np.mean(astar_ms))], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
    step_ms_max, astar_ms_max = float(red[0]), float(red[1])

    # ---- context figure: steady frames of the same agents (no re-plan forced) ----
    # Not the headline: the agents keep the corridors of the last timed step and only sample,
    # check their corridor, follow it (funnel) and avoid each other.
    steady = None
    if world == 1:
        ag_steady = dict(ag)
        ag_steady["flags"] = ag["flags"] & ~np.uint32(_abi.AGENT_RESET)
        gpu.agents_upload(ag_steady, None)
        gpu.step_resident(opts, DT)
        s_ms, s_rep = [], 0
        for _ in range(5):
            gpu.step_resident(opts, DT)
            T = gpu.step_timings()
            s_ms.append(T.total_ms)
            s_rep += T.n_repath
        steady = {"value": n / (float(np.mean(s_ms)) * 1e-3), "unit": "agent-updates/s",
                  "ms_per_step": float(np.mean(s_ms)), "repaths_per_step": s_rep / 5.0,
                  "note": "same agents keeping their corridors: sample x2, corridor check, funnel, ORCA"}
        gpu.agents_upload(ag, None)

    # ---- end-to-end timing through lmb_step with host buffers (`e2e`) ----
    e2e = None
    if not args.skip_e2e and world == 1:
        gpu.step(ag, None, opts, DT)
        barrier()
        t0 = time.perf_counter()
        for _ in range(max(2, args.steps // 2)):
            out = gpu.step(ag, None, opts, DT)
            _ = float(out["desired_velocity"][0, 0])
        barrier()
        e2e_ms = 1e3 * (time.perf_counter() - t0) / max(2, args.steps // 2)
        h2d = sum(int(np.asarray(v).nbytes) for k, v in ag.items() if k != "n")
        d2h = sum(int(v.nbytes) for v in out.values())
        e2e = {"value": n * world / (e2e_ms * 1e-3), "unit": "agent-updates/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}
    elif world > 1:
        # multi-GPU e2e: upload + sharded step + download per step, wall clock, max over ranks
        barrier()
        t0 = time.perf_counter()
        reps = max(2, args.steps // 2)
        for _ in range(reps):
            gpu.agents_upload(ag, None)
            one_step_resident()
            out = gpu.agents_download()
            _ = float(out["desired_velocity"][0, 0])
        barrier()
        e2e_ms = 1e3 * (time.perf_counter() - t0) / reps
        r2 = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(r2, op=dist.ReduceOp.MAX)
        h2d = sum(int(np.asarray(v).nbytes) for k, v in ag.items() if k != "n")
        d2h = sum(int(v.nbytes) for v in out.values())
        e2e = {"value": n * world / (float(r2[0]) * 1e-3), "unit": "agent-updates/s",
               "ms_per_step": float(r2[0]), "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        # algorithmic bytes of the A* launch: 32 per query + 8 per corridor node + 96 per explored node
        astar_bytes = (32.0 * n_repath + 8.0 * path_nodes + 96.0 * explored) / args.steps
        achieved = astar_bytes / (astar_ms_max * 1e-3) / 1e9 if astar_ms_max > 0 else 0.0
        mean_phase = np.mean(np.array(phases), axis=0)
        line = {
            "metric": "agent-updates/s", "value": n * world / (step_ms_max * 1e-3), "unit": "agent-updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms_max,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, world),
            "clocks": clocks, "gpu_launches": int(launches),
            "astar_queries_per_s": n_repath / args.steps * world / (astar_ms_max * 1e-3),
            "explored_nodes_per_s": explored / args.steps * world / (astar_ms_max * 1e-3),
            "phase_ms": {"sample": float(mean_phase[0]), "repath_decide": float(mean_phase[1]),
                         "astar": float(mean_phase[2]), "follow": float(mean_phase[3]),
                         "avoidance": float(mean_phase[4]), "wall_per_step": wall_ms},
            "roofline": {"bound": "hbm", "kernel": "k_astar", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_DRAM_BYTES_PER_EXPLORED * explored / args.steps,
                         "traffic_source": "ncu dram bytes per explored node of a fully loaded smaller "
                                           "launch (profiles/r01_ncu_astar_loaded.md) x explored nodes of this "
                                           "launch; the bench-size launch is not capturable (~100 GB arena)",
                         "algorithmic_bytes": astar_bytes, "peak_source": peak_src,
                         "note": "latency-bound graph search: one dependent chain per query (pop -> "
                                 "records -> best[] probes -> push), random 32-byte sectors of a per-query "
                                 "arena; algorithmic bytes = 32/query + 8/corridor node + 96/explored node"},
        }
        if e2e:
            line["e2e"] = e2e
        if steady:
            line["steady_state"] = steady
        if not args.skip_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args)
        states = np.bincount(out["state"], minlength=9).tolist()
        line["state_histogram"] = states
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    gpu.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(args) -> dict:
    from landmass_b200 import synth
    from oracle.oracle_py import OracleBackend

    sample = args.cpu_sample
    _, flat, ag = build_workload(args.rooms, sample, 0)
    cpu = OracleBackend()
    cpu.navdata_upload(flat)
    opts = synth.default_options()
    t0 = time.perf_counter()
    cpu.step(ag, None, opts, DT)
    dt = time.perf_counter() - t0
    cpu.close()
    return {"value": sample / dt, "unit": "agent-updates/s", "cores": 1, "kind": "port",
            "seconds": dt,
            "sample": f"one full update of {sample} agents (every agent re-plans) on the same "
                      f"{args.rooms}x{args.rooms} maze; oracle = single-threaded C++ port of the reference "
                      "(the Rust crate cannot be built here)"}


if __name__ == "__main__":
    main()
