"""N>1 host logic on CPU: world_size 2 over gloo (127.0.0.1).  Each rank owns a contiguous
range of agents, runs phases A-C locally, all-gathers the avoidance records, runs phase D for
its range.  The concatenated result must equal the single-process step exactly (same
arithmetic, neighbour order is by global agent index).  The backend here is the CPU oracle —
this tests the sharding / halo-exchange logic that bench.py uses with NCCL on GPUs."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, frames, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from landmass_b200 import _abi, sharded, synth
    from oracle.oracle_py import OracleBackend

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mesh = synth.maze_mesh(8)
    flat = synth.single_island_flat(mesh)
    ag = synth.make_agents(mesh, n)
    opts = synth.default_options()
    lo, hi = sharded.shard_range(n, world, rank)
    be = OracleBackend()
    be.navdata_upload(flat)
    outs, results = [], []
    for frame in range(frames):
        mine = sharded.slice_agents(ag, lo, hi)
        be.step_begin_host(mine, None, opts, 1.0 / 30.0, n, lo)
        halo = torch.from_numpy(be.halo_numpy())
        parts = [halo[slice(*sharded.shard_range(n, world, r))] for r in range(world)]
        dist.all_gather(parts, halo[lo:hi].clone())
        out = be.step_end_host(opts, 1.0 / 30.0)
        results.append(sharded.gather_pathing_results(be.pathing_results(), lo))
        # every rank needs everyone's velocities to integrate the global state
        vel = torch.zeros((n, 3), dtype=torch.float32)
        vparts = [vel[slice(*sharded.shard_range(n, world, r))] for r in range(world)]
        dist.all_gather(vparts, torch.from_numpy(out["desired_velocity"]).clone())
        st = torch.zeros(n, dtype=torch.int64)
        sparts = [st[slice(*sharded.shard_range(n, world, r))] for r in range(world)]
        dist.all_gather(sparts, torch.from_numpy(out["state"].astype(np.int64)).clone())
        outs.append((vel.numpy().copy(), st.numpy().copy()))
        ag["flags"] = ag["flags"] & ~np.uint32(_abi.AGENT_RESET)
        ag["velocity"] = vel.numpy().copy()
        ag["position"] = ag["position"] + vel.numpy() * np.float32(1.0 / 30.0)
    all_results = [None] * world
    dist.all_gather_object(all_results, results)
    if rank == 0:
        merged = [sum((all_results[r][f] for r in range(world)), []) for f in range(frames)]
        q.put((outs, merged))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_step_equals_single_process():
    import torch.multiprocessing as mp

    from landmass_b200 import _abi, synth
    from oracle.oracle_py import OracleBackend

    n, frames, world = 240, 3, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs, merged = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    mesh = synth.maze_mesh(8)
    flat = synth.single_island_flat(mesh)
    ag = synth.make_agents(mesh, n)
    opts = synth.default_options()
    be = OracleBackend()
    be.navdata_upload(flat)
    for frame in range(frames):
        out = be.step(ag, None, opts, 1.0 / 30.0)
        assert np.array_equal(out["state"].astype(np.int64), outs[frame][1]), f"frame {frame} states"
        assert np.array_equal(out["desired_velocity"].view(np.uint32), outs[frame][0].view(np.uint32)), \
            f"frame {frame} velocities"
        assert be.pathing_results() == merged[frame]
        ag["flags"] = ag["flags"] & ~np.uint32(_abi.AGENT_RESET)
        ag["velocity"] = out["desired_velocity"].copy()
        ag["position"] = ag["position"] + out["desired_velocity"] * np.float32(1.0 / 30.0)


def test_shard_range_covers_everything():
    from landmass_b200.sharded import shard_range

    for n in (0, 1, 7, 100, 1001):
        for world in (1, 2, 3, 8):
            edges = [shard_range(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
