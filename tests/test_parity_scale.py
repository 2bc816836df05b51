"""Parity at the sizes and on the code paths the benchmark runs ("parity gaps"):

* the benchmarked 256x256-room maze itself: 300 find_paths, the large-batch launch shape forced on the
  same mesh, and one 300-agent full step, CUDA vs oracle, bit-exact;
* a 3x3 tiled, typed, multi-island world (boundary links, three type costs) with the hashed
  best_estimates layout, CUDA vs oracle;
* the SHARDED step on the GPU: two contexts on one device act as ranks 0 and 1 (phases A-C each,
  records exchanged device-to-device, phase D each) — the concatenated result must be bit-equal to one
  context stepping all agents, and match the oracle (avoidance.rs:84-130: neighbours come from ALL
  agents; lib.rs:406-410: PathingResult order = agent order);
* the library-owned exchange (lmb_step_sharded, NCCL opened by the library) on a one-rank communicator.
"""
import numpy as np
import pytest

from landmass_b200 import _abi, sharded, synth

pytestmark = pytest.mark.gpu

NAMES = ["success", "explored", "begin", "nodes", "portals"]


def _backends(flat, max_agents=4096, flags=0, **kw):
    from landmass_b200._native import NativeBackend
    from oracle.oracle_py import OracleBackend

    gpu = NativeBackend(max_agents=max_agents, flags=flags, **kw)
    cpu = OracleBackend(flags=flags & _abi.CONFIG_NO_AVOIDANCE)
    gpu.navdata_upload(flat)
    cpu.navdata_upload(flat)
    return gpu, cpu


@pytest.fixture(scope="module")
def maze256():
    mesh = synth.maze_mesh(256)
    return mesh, synth.single_island_flat(mesh)


@pytest.mark.parametrize("shape", ["pair", "warp", "thread", "large"])
def test_find_paths_on_the_256_room_maze(maze256, shape):
    """300 queries on the bench's own mesh (131 071 polygons, searches of ~65 000 explored nodes) through
    each kernel family: warp pair per query, warp-per-query, thread-per-query small-batch shapes, and the large-batch shape the
    bench launches (<16, 40, 5, quad, forest, SIMPLE>) on 192 slots."""
    mesh, flat = maze256
    flags = {"pair": 0, "warp": _abi.CONFIG_ASTAR_NO_PAIR, "thread": _abi.CONFIG_ASTAR_THREAD_ONLY,
             "large": _abi.CONFIG_ASTAR_LARGE_SHAPES}[shape]
    gpu, cpu = _backends(flat, flags=flags, astar_slots=192 if shape == "large" else 0)
    rng = np.random.Generator(np.random.PCG64(256))
    nq = 300
    sp, sn = synth.random_points_on_mesh(mesh, nq, rng)
    ep, en = synth.random_points_on_mesh(mesh, nq, rng)
    g = gpu.find_paths(sn, sp, en, ep, path_capacity=1 << 23)
    c = cpu.find_paths(sn, sp, en, ep, path_capacity=1 << 23)
    for a, b, name in zip(g, c, NAMES):
        assert np.array_equal(a, b), name
    assert g[0].all() and g[1].mean() > 30000
    gpu.close()
    cpu.close()


def test_full_step_on_the_256_room_maze(maze256):
    """One full update (sample x2, A*, funnel, avoidance) of 300 agents on the bench's mesh, then two
    steady frames: states, PathingResults, corridors exact, velocities 1e-4."""
    mesh, flat = maze256
    n = 300
    gpu, cpu = _backends(flat, max_agents=n)
    ag = synth.make_agents(mesh, n, seed=synth.SEED + 1)
    opts = synth.default_options()
    dt = 1.0 / 60.0
    for frame in range(3):
        og = gpu.step(ag, None, opts, dt)
        oc = cpu.step(ag, None, opts, dt)
        assert np.array_equal(og["state"], oc["state"]), f"frame {frame}"
        assert gpu.pathing_results() == cpu.pathing_results(), f"frame {frame}"
        np.testing.assert_allclose(og["desired_velocity"], oc["desired_velocity"], rtol=1e-4, atol=1e-6)
        for slot in range(0, n, 37):
            gp, cp = gpu.agent_path(slot), cpu.agent_path(slot)
            assert np.array_equal(gp[0], cp[0]) and np.array_equal(gp[1], cp[1])
        ag["flags"] = ag["flags"] & ~np.uint32(_abi.AGENT_RESET)
        ag["velocity"] = oc["desired_velocity"].copy()
        ag["position"] = ag["position"] + oc["desired_velocity"] * np.float32(dt)
    gpu.close()
    cpu.close()


@pytest.mark.parametrize("kernel", [0, _abi.CONFIG_ASTAR_NO_PAIR, _abi.CONFIG_ASTAR_THREAD_ONLY,
                                    _abi.CONFIG_ASTAR_LARGE_SHAPES])
@pytest.mark.parametrize("layout", ["hashed", "dense"])
def test_tiled_typed_multi_island_world(kernel, layout):
    """BASELINE config 4 in small: 3 x 3 islands of 32 x 32 quads, type costs {1, 2, 5}, boundary links.
    The full-size world (1 M polygons) picks the hashed layout by itself; here it is forced."""
    g, k = 32, 3
    flat = synth.tiled_archipelago_flat(k, g, type_block=8)
    lay = _abi.CONFIG_ASTAR_HASHED if layout == "hashed" else 0
    gpu, cpu = _backends(flat, flags=kernel | lay, astar_slots=192 if kernel == _abi.CONFIG_ASTAR_LARGE_SHAPES else 0)
    n = 600
    ag = synth.uniform_agents(n, k * g, k * g, seed=4, target_reach=1.5 * g)
    opts = synth.default_options()
    # batch queries between sampled points
    gp, gn = gpu.sample_points(ag["position"], opts)
    tp, tn = gpu.sample_points(ag["target"], opts)
    cp_, cn = cpu.sample_points(ag["position"], opts)
    assert np.array_equal(gn, cn) and np.array_equal(gp.view(np.uint32), cp_.view(np.uint32))
    ok = (gn != _abi.NONE) & (tn != _abi.NONE)
    gq = gpu.find_paths(gn[ok], gp[ok], tn[ok], tp[ok], path_capacity=1 << 21)
    cq = cpu.find_paths(gn[ok], gp[ok], tn[ok], tp[ok], path_capacity=1 << 21)
    for a, b, name in zip(gq, cq, NAMES):
        assert np.array_equal(a, b), name
    assert gq[0].all()
    crossings = int(((gq[4] & _abi.PORTAL_LINK) != 0).sum() - (gq[4] == _abi.NONE).sum())
    assert crossings > 100, "corridors should cross island borders through boundary links"
    # and three full frames
    dt = 1.0 / 60.0
    for frame in range(3):
        og = gpu.step(ag, None, opts, dt)
        oc = cpu.step(ag, None, opts, dt)
        assert np.array_equal(og["state"], oc["state"]), f"frame {frame}"
        assert gpu.pathing_results() == cpu.pathing_results(), f"frame {frame}"
        np.testing.assert_allclose(og["desired_velocity"], oc["desired_velocity"], rtol=1e-4, atol=1e-6)
        ag["flags"] = ag["flags"] & ~np.uint32(_abi.AGENT_RESET)
        ag["velocity"] = oc["desired_velocity"].copy()
        ag["position"] = ag["position"] + oc["desired_velocity"] * np.float32(dt)
    gpu.close()
    cpu.close()


def _wrap(ptr, nbytes):
    import torch

    class _Buf:
        __cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}

    return torch.as_tensor(_Buf(), device=torch.device("cuda", 0))


@pytest.mark.parametrize("order", [pytest.param(0, id="index-order"), pytest.param(_abi.CONFIG_ORDER_ALWAYS, id="work-order")])
@pytest.mark.parametrize("world_kind", ["maze", "tiled"])
def test_sharded_step_on_the_gpu_equals_the_single_context_step(world_kind, order):
    """`order`: from 8192 agents on k_follow takes a rank's agents longest corridor first and k_avoidance in the
    grid's cell order — with sharded agents through a flag / scan / scatter over the gathered records;
    LMB_CONFIG_ORDER_ALWAYS runs those paths at this size (results do not depend on the order)."""
    import torch

    from landmass_b200._native import NativeBackend
    from oracle.oracle_py import OracleBackend

    if world_kind == "maze":
        mesh = synth.maze_mesh(10)
        flat = synth.single_island_flat(mesh)
        n = 481  # uneven split: 241 + 240; crowded enough that ranks see each other's agents
        ag = synth.make_agents(mesh, n)
    else:
        flat = synth.tiled_archipelago_flat(2, 16, type_block=4)
        n = 600
        ag = synth.uniform_agents(n, 32, 32, seed=9)
    opts = synth.default_options()
    world = 2
    ranges = [sharded.shard_range(n, world, r) for r in range(world)]
    ranks = [NativeBackend(max_agents=n, device=0, flags=order) for _ in range(world)]
    single = NativeBackend(max_agents=n, device=0, flags=order)
    cpu = OracleBackend()
    for b in ranks + [single, cpu]:
        b.navdata_upload(flat)
    dt = 1.0 / 30.0
    cross_rank_neighbours = 0
    for frame in range(4):
        # --- two ranks on one GPU ---
        for r, be in enumerate(ranks):
            lo, hi = ranges[r]
            be.agents_upload(sharded.slice_agents(ag, lo, hi), None)
            be.step_begin(opts, dt, n, lo)
        bufs = []
        for be in ranks:
            ptr, rec_bytes, n_rec = be.halo_buffer()
            assert n_rec == n and rec_bytes == 32
            bufs.append(_wrap(ptr, n_rec * rec_bytes))
        # the exchange: every rank receives the other's record range (what ncclAllGather does in place)
        for r in range(world):
            for o in range(world):
                if o != r:
                    lo, hi = ranges[o]
                    bufs[r][lo * 32:hi * 32].copy_(bufs[o][lo * 32:hi * 32])
        torch.cuda.synchronize()
        outs, results = [], []
        for r, be in enumerate(ranks):
            be.step_end(opts, dt)
            outs.append(be.agents_download())
            results += sharded.gather_pathing_results(be.pathing_results(), ranges[r][0])
        vel = np.concatenate([o["desired_velocity"] for o in outs])
        state = np.concatenate([o["state"] for o in outs])
        # --- one context, all agents ---
        single.agents_upload(ag, None)
        single.step_resident(opts, dt)
        so = single.agents_download()
        assert np.array_equal(state, so["state"]), f"frame {frame}: states"
        assert np.array_equal(vel.view(np.uint32), so["desired_velocity"].view(np.uint32)), f"frame {frame}: velocities"
        assert results == single.pathing_results(), f"frame {frame}: PathingResult order"
        # --- oracle ---
        oc = cpu.step(ag, None, opts, dt)
        assert np.array_equal(state, oc["state"])
        assert results == cpu.pathing_results()
        np.testing.assert_allclose(vel, oc["desired_velocity"], rtol=1e-4, atol=1e-6)
        # the test is only meaningful if avoidance couples the ranks: agents near the split see each other
        p = ag["position"]
        lo1 = ranges[1][0]
        d = np.linalg.norm(p[:lo1, None, :2] - p[None, lo1:, :2], axis=2)
        cross_rank_neighbours += int((d < 5.0).sum())
        ag["flags"] = ag["flags"] & ~np.uint32(_abi.AGENT_RESET)
        ag["velocity"] = oc["desired_velocity"].copy()
        ag["position"] = ag["position"] + oc["desired_velocity"] * np.float32(dt)
    assert cross_rank_neighbours > 100
    for b in ranks + [single, cpu]:
        b.close()


def test_library_owned_exchange_on_a_one_rank_communicator():
    """lmb_step_sharded = phases A-C, ncclAllGather on the library stream, phase D.  With one rank the
    gather is the identity, so the result must be bit-equal to lmb_step_resident — with the rank's share
    padded (agents_per_rank > n: the tail records are invalid)."""
    from landmass_b200._native import NativeBackend

    mesh = synth.maze_mesh(10)
    flat = synth.single_island_flat(mesh)
    n = 300
    ag = synth.make_agents(mesh, n)
    opts = synth.default_options()
    a, b = NativeBackend(max_agents=n), NativeBackend(max_agents=n)
    for be in (a, b):
        be.navdata_upload(flat)
    a.comm_init(a.comm_unique_id(), 1, 0)
    dt = 1.0 / 60.0
    for frame in range(3):
        a.agents_upload(ag, None)
        b.agents_upload(ag, None)
        a.step_sharded(opts, dt, n + 20)
        b.step_resident(opts, dt)
        oa, ob = a.agents_download(), b.agents_download()
        assert np.array_equal(oa["state"], ob["state"])
        assert np.array_equal(oa["desired_velocity"].view(np.uint32), ob["desired_velocity"].view(np.uint32))
        assert a.pathing_results() == b.pathing_results()
        assert a.step_timings().allgather_ms >= 0.0
        ag["flags"] = ag["flags"] & ~np.uint32(_abi.AGENT_RESET)
        ag["velocity"] = ob["desired_velocity"].copy()
        ag["position"] = ag["position"] + ob["desired_velocity"] * np.float32(dt)
    a.close()
    b.close()


def test_queries_between_resident_steps_leave_the_step_state_alone():
    """lmb_sample_points used the agents' device buffers and lmb_find_paths reset the step's
    timings — a query between two lmb_step_resident calls must change neither the next step nor
    lmb_pathing_results of the last one."""
    from landmass_b200._native import NativeBackend

    mesh = synth.maze_mesh(10)
    flat = synth.single_island_flat(mesh)
    n = 200
    ag = synth.make_agents(mesh, n)
    opts = synth.default_options()
    a, b = NativeBackend(max_agents=n), NativeBackend(max_agents=n)
    for be in (a, b):
        be.navdata_upload(flat)
        be.agents_upload(ag, None)
        be.step_resident(opts, 1.0 / 60.0)
    want = b.pathing_results()
    rng = np.random.Generator(np.random.PCG64(5))
    pts, nodes = synth.random_points_on_mesh(mesh, 64, rng)
    a.sample_points(pts, opts)
    a.find_paths(nodes[:32], pts[:32], nodes[32:], pts[32:])
    assert a.pathing_results() == want and len(want) == n
    assert a.query_timings().n_repath == 32 and a.query_timings().explored_nodes > 0
    steady = dict(ag)
    steady["flags"] = ag["flags"] & ~np.uint32(_abi.AGENT_RESET)
    for be in (a, b):
        be.step_resident(opts, 1.0 / 60.0)  # still the agents uploaded before the queries
    oa, ob = a.agents_download(), b.agents_download()
    assert np.array_equal(oa["state"], ob["state"])
    assert np.array_equal(oa["desired_velocity"].view(np.uint32), ob["desired_velocity"].view(np.uint32))
    a.close()
    b.close()
