"""CUDA-vs-oracle parity of the A* kernel's data-layout variants (k_astar.cu, k_astar_common.cuh):
CSR vs padded state ids, dense vs compact best_estimates (incl. the overflow table and the
fallback to dense), slot reuse (both restore paths), arena-overflow retries and the
longest-first query order.  Everything here is bit-exact: corridors, portal edge indices,
explored_nodes."""
import numpy as np
import pytest

from landmass_b200 import _abi, synth

pytestmark = pytest.mark.gpu

NAMES = ["success", "explored", "begin", "nodes", "portals"]


KERNELS = [pytest.param(0, id="warp-pair-per-query"),
           pytest.param(_abi.CONFIG_ASTAR_NO_PAIR, id="warp-per-query"),
           pytest.param(_abi.CONFIG_ASTAR_THREAD_ONLY, id="thread-per-query")]


@pytest.fixture(params=KERNELS)
def kernel(request):
    """Small batches go to the warp kernel (k_astar_warp.cu) by default — a search warp plus a heap warp per query
    up to four queries per SM, one warp per query (NO_PAIR forces it) beyond; THREAD_ONLY keeps them on the
    thread-per-query kernel (k_astar.cu).  Every layout test runs on all three."""
    return request.param


def backends(flat, max_agents=4096, flags=0, **kw):
    from landmass_b200._native import NativeBackend
    from oracle.oracle_py import OracleBackend

    gpu = NativeBackend(max_agents=max_agents, flags=flags, **kw)
    cpu = OracleBackend(flags=flags & _abi.CONFIG_NO_AVOIDANCE)
    gpu.navdata_upload(flat)
    cpu.navdata_upload(flat)
    return gpu, cpu


def assert_same_paths(gpu, cpu, sn, sp, en, ep, cap=None):
    g = gpu.find_paths(sn, sp, en, ep, path_capacity=cap)
    c = cpu.find_paths(sn, sp, en, ep, path_capacity=cap)
    for a, b, name in zip(g, c, NAMES):
        assert np.array_equal(a, b), name
    return g


def centre_queries(mesh, pairs):
    sn = np.array([a for a, _ in pairs], dtype=np.uint32)
    en = np.array([b for _, b in pairs], dtype=np.uint32)
    return sn, mesh.poly_center[sn].astype(np.float32), en, mesh.poly_center[en].astype(np.float32)


def test_csr_state_ids_polygon_with_12_edges(kernel):
    """A 12-gon makes max edges > 8: the kernel falls back to CSR edge ids (the state's own record
    is read before the polygon's records can be addressed)."""
    mesh = synth.rosette_mesh(12, 3)
    flat = synth.single_island_flat(mesh)
    gpu, cpu = backends(flat, flags=kernel)
    pairs = [(a, b) for a in range(mesh.n_polys) for b in range(mesh.n_polys)]
    g = assert_same_paths(gpu, cpu, *centre_queries(mesh, pairs))
    assert g[0].all()


def test_padded_ids_with_8_record_polygons(kernel):
    """Hexagon + ring: max edges in (4, 8] -> 8 records per polygon, two chunks per expansion."""
    mesh = synth.rosette_mesh(7, 4)
    flat = synth.single_island_flat(mesh)
    gpu, cpu = backends(flat, flags=kernel)
    pairs = [(a, b) for a in range(mesh.n_polys) for b in range(mesh.n_polys)]
    g = assert_same_paths(gpu, cpu, *centre_queries(mesh, pairs))
    assert g[0].all()


@pytest.mark.parametrize("extra", [0, 6, 40])
def test_compact_best_with_overflow_and_slot_reuse(extra, capfd, kernel):
    """Perfect mazes (forests) use the compact best_estimates layout.  With extra openings the
    layout is only used when forced; short searches then complete through the overflow table while
    long ones fill it and are re-run dense.  128 slots for 1500 queries re-use every slot ~12
    times (scatter and dense restore)."""
    mesh = synth.loopy_maze_mesh(28, extra)
    flat = synth.single_island_flat(mesh)
    gpu, cpu = backends(flat, astar_slots=128, flags=kernel | _abi.CONFIG_DEBUG_LOG | _abi.CONFIG_ASTAR_COMPACT)
    rng = np.random.Generator(np.random.PCG64(17 + extra))
    nq = 1500
    sp, sn = synth.random_points_on_mesh(mesh, nq, rng)
    ep, en = synth.random_points_on_mesh(mesh, nq, rng)
    # short queries (scatter restore) next to long ones (dense restore)
    en[:300] = sn[:300]
    ep[:300] = sp[:300] + np.float32(0.05)
    g = assert_same_paths(gpu, cpu, sn, sp, en, ep, cap=1 << 21)
    assert g[0].all()
    err = capfd.readouterr().err
    assert "compact best" in err and (extra > 0 or "dense best" not in err)


def test_forest_detection_picks_layout(capfd, kernel):
    """A forest needs no best_estimates at all ("no best"); LMB_CONFIG_ASTAR_NO_FOREST falls back to the
    per-polygon layout; one extra opening (a cycle) or an open grid is dense."""
    for mesh, want, fl in [(synth.maze_mesh(12), "no best", 0),
                           (synth.maze_mesh(12), "compact best", _abi.CONFIG_ASTAR_NO_FOREST),
                           (synth.loopy_maze_mesh(12, 1), "dense best", 0),
                           (synth.grid_mesh(8, 8), "dense best", 0)]:
        gpu, cpu = backends(synth.single_island_flat(mesh), flags=kernel | fl | _abi.CONFIG_DEBUG_LOG)
        assert_same_paths(gpu, cpu, *centre_queries(mesh, [(0, mesh.n_polys - 1), (3, 5)]))
        assert want in capfd.readouterr().err
        gpu.close()


def test_compact_best_falls_back_to_dense(capfd, kernel):
    """Compact layout forced on a maze with cycles: long searches enter more than 512 polygons a
    second time, the overflow table fills, the queries are re-run and the context switches to the
    dense layout — results unchanged."""
    rooms = 64
    base = synth.maze_mesh(rooms)
    mesh = synth.loopy_maze_mesh(rooms, base.n_polys // 17)
    flat = synth.single_island_flat(mesh)
    gpu, cpu = backends(flat, flags=kernel | _abi.CONFIG_DEBUG_LOG | _abi.CONFIG_ASTAR_COMPACT)
    rng = np.random.Generator(np.random.PCG64(23))
    nq = 200
    sp, sn = synth.random_points_on_mesh(mesh, nq, rng)
    ep, en = synth.random_points_on_mesh(mesh, nq, rng)
    g = assert_same_paths(gpu, cpu, sn, sp, en, ep, cap=1 << 21)
    assert g[0].all()
    err = capfd.readouterr().err
    assert "compact best" in err and "dense best" in err
    # and again on the now-dense context
    assert_same_paths(gpu, cpu, sn[::-1].copy(), sp[::-1].copy(), en[::-1].copy(), ep[::-1].copy(), cap=1 << 21)


@pytest.mark.parametrize("mesh_kind", ["grid", "maze", "rosette"])
def test_hashed_best_with_table_growth(mesh_kind, capfd, kernel):
    """Hashed best_estimates (chosen for meshes above 2M states; forced here, with a 2048-node
    arena): probe chains, the half-full limit -> arena retry with a doubled table, slot re-use
    (whole-table clear), CSR ids (rosette)."""
    if mesh_kind == "rosette":
        mesh = synth.rosette_mesh(12, 6)
        pairs = [(a, b) for a in range(mesh.n_polys) for b in range(0, mesh.n_polys, 3)]
        sn, sp, en, ep = centre_queries(mesh, pairs)
    else:
        mesh = synth.grid_mesh(90, 90) if mesh_kind == "grid" else synth.loopy_maze_mesh(40, 30)
        rng = np.random.Generator(np.random.PCG64(41))
        nq = 700
        sp, sn = synth.random_points_on_mesh(mesh, nq, rng)
        ep, en = synth.random_points_on_mesh(mesh, nq, rng)
        en[:100] = sn[:100]
        ep[:100] = sp[:100] + np.float32(0.05)
    flat = synth.single_island_flat(mesh)
    gpu, cpu = backends(flat, astar_slots=96, flags=kernel | _abi.CONFIG_DEBUG_LOG | _abi.CONFIG_ASTAR_HASHED)
    g = assert_same_paths(gpu, cpu, sn, sp, en, ep, cap=1 << 21)
    assert g[0].all()
    err = capfd.readouterr().err
    assert "hashed best" in err
    if mesh_kind != "rosette":
        assert err.count("arena:") >= 2  # the 2048-node table overflowed at least once


def test_dense_best_slot_reuse_and_arena_retry(capfd, kernel):
    """Open grid (dense layout): corner-to-corner searches outgrow the initial node list
    (n_states / 4) -> retry with a doubled arena; 64 slots re-used by 400 queries."""
    mesh = synth.grid_mesh(120, 120)
    flat = synth.single_island_flat(mesh)
    gpu, cpu = backends(flat, astar_slots=64, flags=kernel | _abi.CONFIG_DEBUG_LOG)
    rng = np.random.Generator(np.random.PCG64(29))
    nq = 400
    sp, sn = synth.random_points_on_mesh(mesh, nq, rng)
    ep, en = synth.random_points_on_mesh(mesh, nq, rng)
    sp[:2] = [[0.5, 0.5, 0], [119.5, 0.5, 0]]
    sn[:2] = [0, 119]
    ep[:2] = [[119.5, 119.5, 0], [0.5, 119.5, 0]]
    en[:2] = [120 * 120 - 1, 119 * 120]
    en[100:200] = sn[100:200]
    ep[100:200] = sp[100:200] + np.float32(0.05)
    assert_same_paths(gpu, cpu, sn, sp, en, ep, cap=1 << 21)
    err = capfd.readouterr().err
    assert err.count("arena:") >= 2 and "compact best" not in err


@pytest.mark.parametrize("shape", [0, _abi.CONFIG_ASTAR_NO_PAIR, _abi.CONFIG_ASTAR_THREAD_ONLY,
                                   _abi.CONFIG_ASTAR_LARGE_SHAPES])
@pytest.mark.parametrize("mesh_kind", ["grid", "maze", "forest"])
def test_open_list_spill_overflow_and_retry(mesh_kind, shape, capfd):
    """the spill region of a slot must hold heap_cap entries for the SMALLEST number of shared-memory
    entries any launch shape keeps (the <16,40,5> shape wrote past a region sized for 48).  LMB_CONFIG_ASTAR_SMALL_HEAP
    starts every query with a 64-entry open list, so that at test sizes the open lists spill in place, overflow
    (status 7) and are re-run with doubled capacities — on an open grid several times over — through the warp
    kernel, the small thread-per-query shapes and the large-batch shapes; results stay bit-exact."""
    mesh = {"grid": lambda: synth.grid_mesh(64, 64), "maze": lambda: synth.loopy_maze_mesh(24, 40),
            "forest": lambda: synth.maze_mesh(32)}[mesh_kind]()
    flat = synth.single_island_flat(mesh)
    gpu, cpu = backends(flat, astar_slots=96, flags=shape | _abi.CONFIG_DEBUG_LOG | _abi.CONFIG_ASTAR_SMALL_HEAP)
    rng = np.random.Generator(np.random.PCG64(77))
    nq = 600
    sp, sn = synth.random_points_on_mesh(mesh, nq, rng)
    ep, en = synth.random_points_on_mesh(mesh, nq, rng)
    g = assert_same_paths(gpu, cpu, sn, sp, en, ep, cap=1 << 21)
    assert g[0].all()
    err = capfd.readouterr().err
    if mesh_kind == "grid":
        assert err.count("arena:") >= 3, err  # 64 -> 80 -> 112 -> ... entries
    # and again on the grown arena (slot re-use after a spilled search)
    assert_same_paths(gpu, cpu, sn[::-1].copy(), sp[::-1].copy(), en[::-1].copy(), ep[::-1].copy(), cap=1 << 21)


def test_step_more_agents_than_slots_longest_first_order(kernel):
    """The step orders a batch longest-first from the previous frame's explored_nodes when it has
    more queries than slots.  The order must not change any result: every agent re-plans on
    three consecutive frames (hints absent, then present) and matches the oracle each time."""
    mesh = synth.maze_mesh(24)
    flat = synth.single_island_flat(mesh)
    n = 1200
    gpu, cpu = backends(flat, flags=_abi.CONFIG_NO_AVOIDANCE | kernel, astar_slots=128)
    ag = synth.make_agents(mesh, n)
    opts = synth.default_options()
    for frame in range(3):
        og = gpu.step(ag, None, opts, 1.0 / 60.0)
        oc = cpu.step(ag, None, opts, 1.0 / 60.0)
        assert np.array_equal(og["state"], oc["state"])
        assert gpu.pathing_results() == cpu.pathing_results()
        assert len(gpu.pathing_results()) == n  # AGENT_RESET stays set: everyone re-plans
        np.testing.assert_allclose(og["desired_velocity"], oc["desired_velocity"], rtol=1e-4, atol=1e-6)
        for slot in range(0, n, 97):
            gp, cp = gpu.agent_path(slot), cpu.agent_path(slot)
            assert np.array_equal(gp[0], cp[0]) and np.array_equal(gp[1], cp[1])


@pytest.mark.parametrize("case", ["maze_simple", "maze_compact", "maze_overrides", "grid_simple", "grid_types",
                                  "loopy_maze_types"])
def test_large_batch_shapes_against_the_oracle(case):
    """The launch shapes of large batches — `<16, 40, 5, SIMPLE>` (the benchmarked kernel: one polygon
    type, no overrides), `<16, 48, 4>` (cost tables / per-query overrides in play) and `<8, 96, 4>` (large
    open lists) — only run above ~19 000 queries, which the oracle cannot check in seconds.
    LMB_CONFIG_ASTAR_LARGE_SHAPES forces them here: 2 500 queries on 192 slots, bit-exact.  `maze_simple` runs
    the best_estimates-free forest layout (the benchmarked one), `maze_compact` the per-polygon layout."""
    flags = _abi.CONFIG_ASTAR_LARGE_SHAPES | (_abi.CONFIG_ASTAR_NO_FOREST if case == "maze_compact" else 0)
    rng = np.random.Generator(np.random.PCG64(4242))
    nq = 2500
    overrides = None
    if case.startswith("maze"):
        types = None
        mesh = synth.maze_mesh(20)
        if case == "loopy_maze_types":
            mesh = synth.loopy_maze_mesh(20, 60)   # cycles: dense layout, larger open lists
    else:  # (not a forest: the first batch takes the large-open-list shape)
        types = (np.add.outer(np.arange(24) // 4, np.arange(24) // 4) % 3).astype(np.uint32)
        mesh = synth.grid_mesh(24, 24, types=types if case == "grid_types" else None)
    flat = synth.single_island_flat(mesh)
    if case in ("grid_types", "loopy_maze_types"):
        flat = dict(flat)
        if case == "loopy_maze_types":
            flat["poly_type"] = (np.arange(mesh.n_polys) % 3).astype(np.uint32)
        flat["n_type_costs"] = 3
        flat["type_cost_index"] = np.array([0, 1, 2], dtype=np.uint32)
        flat["type_cost_value"] = np.array([1.0, 2.5, 0.5], dtype=np.float32)
    gpu, cpu = backends(flat, astar_slots=192, flags=flags)
    sp, sn = synth.random_points_on_mesh(mesh, nq, rng)
    ep, en = synth.random_points_on_mesh(mesh, nq, rng)
    if case == "maze_overrides":
        overrides = [({0: float(1.0 + (i % 5))} if i % 2 else None) for i in range(nq)]
    g = gpu.find_paths(sn, sp, en, ep, overrides=overrides, path_capacity=1 << 21)
    c = cpu.find_paths(sn, sp, en, ep, overrides=overrides, path_capacity=1 << 21)
    for a, b, name in zip(g, c, NAMES):
        assert np.array_equal(a, b), name
    assert g[0].all()
    gpu.close()
    cpu.close()


@pytest.mark.parametrize("pair", [pytest.param(0, id="warp-pair"), pytest.param(_abi.CONFIG_ASTAR_NO_PAIR, id="one-warp")])
@pytest.mark.parametrize("case", ["maze", "grid_types", "loopy_maze_hashed"])
def test_every_warp_kernel_launch_shape_against_the_oracle(case, pair):
    """k_astar_warp.cu picks its launch shape by queries per SM — 1, up to 4, up to 8, up to 16 units per SM, each
    unit one warp or a search warp + heap warp pair, with less shared memory per open list every step.
    LMB_CONFIG_ASTAR_WARP_ONLY keeps batches of every size on it: 100 / 500 / 1 100 / 2 300 queries against the
    oracle, on a forest (no best_estimates), a typed grid (dense) and a maze with cycles under the hashed layout."""
    rng = np.random.Generator(np.random.PCG64(99))
    flags = _abi.CONFIG_ASTAR_WARP_ONLY | pair
    if case == "maze":
        mesh = synth.maze_mesh(20)
        flat = synth.single_island_flat(mesh)
    elif case == "grid_types":
        types = (np.add.outer(np.arange(24) // 4, np.arange(24) // 4) % 3).astype(np.uint32)
        mesh = synth.grid_mesh(24, 24, types=types)
        flat = dict(synth.single_island_flat(mesh))
        flat["n_type_costs"] = 3
        flat["type_cost_index"] = np.array([0, 1, 2], dtype=np.uint32)
        flat["type_cost_value"] = np.array([1.0, 2.5, 0.5], dtype=np.float32)
    else:
        mesh = synth.loopy_maze_mesh(20, 60)
        flat = synth.single_island_flat(mesh)
        flags |= _abi.CONFIG_ASTAR_HASHED
    gpu, cpu = backends(flat, flags=flags)
    for nq in (100, 500, 1100, 2300):
        sp, sn = synth.random_points_on_mesh(mesh, nq, rng)
        ep, en = synth.random_points_on_mesh(mesh, nq, rng)
        g = gpu.find_paths(sn, sp, en, ep, path_capacity=1 << 21)
        c = cpu.find_paths(sn, sp, en, ep, path_capacity=1 << 21)
        for a, b, name in zip(g, c, NAMES):
            assert np.array_equal(a, b), (nq, name)
        assert gpu.query_timings().astar_launch_kind == (2 if pair else 4)
    gpu.close()
    cpu.close()


BIG_TILED_POINTS = np.array([[100.5, 100.5, 0], [100.5, 500.5, 0], [300.5, 200.5, 0], [650.5, 300.5, 0],
                             [20.5, 400.5, 0], [22.5, 410.5, 0], [10.5, 10.5, 0], [300.5, 300.5, 0]], dtype=np.float32)


def test_hashed_layout_turns_dense_when_searches_are_not_local(capfd):
    """A mesh of more than 2^21 states starts with the hashed best_estimates (a search is expected to touch a small
    part of it).  Far-reaching searches make the node list — and with it the table — grow; once the table is no
    smaller than the dense array the context goes dense for good.  6 x 6 islands of 128 x 128 quads, 2.4 M states;
    LARGE_SHAPES keeps the four queries on the main arena (small batches would take the dense side arena)."""
    flat = synth.tiled_archipelago_flat(6, 128)
    gpu, cpu = backends(flat, flags=_abi.CONFIG_DEBUG_LOG | _abi.CONFIG_ASTAR_LARGE_SHAPES, astar_slots=32)
    opts = synth.default_options()
    p, nodes = gpu.sample_points(BIG_TILED_POINTS, opts)
    sn, sp, en, ep = nodes[0::2], p[0::2], nodes[1::2], p[1::2]
    g = assert_same_paths(gpu, cpu, sn, sp, en, ep, cap=1 << 22)
    assert g[0].all() and g[1].max() > 300_000
    err = capfd.readouterr().err
    assert "hashed best" in err and "dense best" in err and "side arena" not in err, err
    assert err.rindex("dense best") > err.rindex("hashed best")
    # and it stays dense: the same queries again, no reallocation
    assert_same_paths(gpu, cpu, sn, sp, en, ep, cap=1 << 22)
    assert "best" not in capfd.readouterr().err
    gpu.close()
    cpu.close()


def test_small_batches_on_a_hashed_mesh_take_the_dense_side_arena(capfd):
    """The hashed layout is for thousands of searches in flight; up to four queries per SM run on a side arena of
    dense slots instead (a plain load per probe), larger batches on the hashed main arena.  Same world as above."""
    flat = synth.tiled_archipelago_flat(6, 128)
    gpu, cpu = backends(flat, flags=_abi.CONFIG_DEBUG_LOG)
    opts = synth.default_options()
    p, nodes = gpu.sample_points(BIG_TILED_POINTS, opts)
    g = assert_same_paths(gpu, cpu, nodes[0::2], p[0::2], nodes[1::2], p[1::2], cap=1 << 22)
    assert g[0].all() and g[1].max() > 300_000
    err = capfd.readouterr().err
    assert "side arena" in err and "dense best" in err and "hashed best" not in err, err
    # 700 short queries: the main arena, hashed
    rng = np.random.Generator(np.random.PCG64(5))
    a = np.zeros((700, 3), dtype=np.float32)
    a[:, :2] = rng.uniform(40.0, 720.0, size=(700, 2))
    b = a.copy()
    b[:, :2] += rng.uniform(-25.0, 25.0, size=(700, 2))
    pa, na = gpu.sample_points(a, opts)
    pb, nb = gpu.sample_points(b, opts)
    g = assert_same_paths(gpu, cpu, na, pa, nb, pb, cap=1 << 22)
    assert g[0].all()
    err = capfd.readouterr().err
    assert "hashed best" in err and "side arena" not in err, err
    # ... and small batches keep using the side arena without reallocating it
    assert_same_paths(gpu, cpu, nodes[0::2], p[0::2], nodes[1::2], p[1::2], cap=1 << 22)
    assert "arena" not in capfd.readouterr().err
    gpu.close()
    cpu.close()
