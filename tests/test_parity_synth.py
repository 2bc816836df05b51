"""CUDA-vs-oracle parity on seeded synthetic navmeshes (SURVEY.md §8d configs at sizes the
oracle finishes in seconds), through the C ABI.  Bit-exact for node ids, corridors, portal
edge indices, explored_nodes and states; sampled points bit-exact; velocities 1e-4 relative."""
import numpy as np
import pytest

from landmass_b200 import _abi, synth

pytestmark = pytest.mark.gpu


def backends(flat, max_agents=4096, flags=0):
    from landmass_b200._native import NativeBackend
    from oracle.oracle_py import OracleBackend

    gpu = NativeBackend(max_agents=max_agents, flags=flags)
    cpu = OracleBackend(flags=flags)
    gpu.navdata_upload(flat)
    cpu.navdata_upload(flat)
    return gpu, cpu


@pytest.mark.parametrize("mesh_kind", ["grid", "maze"])
def test_sampling_bit_exact(mesh_kind):
    mesh = synth.grid_mesh(64, 64) if mesh_kind == "grid" else synth.maze_mesh(24)
    flat = synth.single_island_flat(mesh)
    gpu, cpu = backends(flat)
    rng = np.random.Generator(np.random.PCG64(7))
    n = 20000
    pts, _ = synth.random_points_on_mesh(mesh, n, rng, lo=-0.2, hi=1.2)  # some fall outside
    pts[:, 2] = rng.uniform(-0.8, 0.6, size=n).astype(np.float32)
    # lattice points exercise exact ties between neighbouring polygons
    lattice = np.stack([rng.integers(0, 50, n // 10), rng.integers(0, 50, n // 10), np.zeros(n // 10)], axis=1)
    pts = np.concatenate([pts, lattice.astype(np.float32)])
    opts = synth.default_options()
    gp, gn = gpu.sample_points(pts, opts)
    cp, cn = cpu.sample_points(pts, opts)
    assert np.array_equal(gn, cn)
    assert np.array_equal(gp.view(np.uint32), cp.view(np.uint32))
    assert (gn != _abi.NONE).sum() > n // 2 and (gn == _abi.NONE).sum() > 0


@pytest.mark.parametrize("mesh_kind,nq", [("grid", 300), ("maze", 300)])
def test_find_paths_bit_exact(mesh_kind, nq):
    mesh = synth.grid_mesh(48, 48) if mesh_kind == "grid" else synth.maze_mesh(24)
    flat = synth.single_island_flat(mesh)
    gpu, cpu = backends(flat)
    rng = np.random.Generator(np.random.PCG64(11))
    sp, sn = synth.random_points_on_mesh(mesh, nq, rng)
    ep, en = synth.random_points_on_mesh(mesh, nq, rng)
    g = gpu.find_paths(sn, sp, en, ep)
    c = cpu.find_paths(sn, sp, en, ep)
    for a, b, name in zip(g, c, ["success", "explored", "begin", "nodes", "portals"]):
        assert np.array_equal(a, b), name
    assert g[0].all()


def run_frames(gpu, cpu, ag, opts, frames, dt=1.0 / 60.0, retarget=None):
    n = ag["n"]
    for frame in range(frames):
        og = gpu.step(ag, None, opts, dt)
        oc = cpu.step(ag, None, opts, dt)
        assert np.array_equal(og["state"], oc["state"]), f"frame {frame}: states"
        assert gpu.pathing_results() == cpu.pathing_results(), f"frame {frame}: pathing results"
        np.testing.assert_allclose(og["desired_velocity"], oc["desired_velocity"], rtol=1e-4, atol=1e-6,
                                   err_msg=f"frame {frame}")
        for slot in range(0, n, max(1, n // 16)):
            gp, cp = gpu.agent_path(slot), cpu.agent_path(slot)
            assert np.array_equal(gp[0], cp[0]) and np.array_equal(gp[1], cp[1]), f"path of slot {slot}"
        ag["flags"] = ag["flags"] & ~np.uint32(_abi.AGENT_RESET)
        ag["velocity"] = oc["desired_velocity"].copy()
        ag["position"] = ag["position"] + oc["desired_velocity"] * np.float32(dt)
        if retarget is not None:
            retarget(frame, ag)
    return og


@pytest.mark.parametrize("mesh_kind", ["grid", "maze"])
def test_step_without_avoidance(mesh_kind):
    mesh = synth.grid_mesh(64, 64) if mesh_kind == "grid" else synth.maze_mesh(20)
    flat = synth.single_island_flat(mesh)
    gpu, cpu = backends(flat, flags=_abi.CONFIG_NO_AVOIDANCE)
    ag = synth.make_agents(mesh, 600)
    # a mix of agent situations: no target, paused, off-mesh agent, off-mesh target
    ag["flags"][0:20] &= ~np.uint32(_abi.AGENT_HAS_TARGET)
    ag["flags"][20:40] |= _abi.AGENT_PAUSED
    ag["position"][40:60, 2] = 50.0
    ag["target"][60:80, 0] = -100.0
    ag["reach_condition"] = np.zeros(600, dtype=np.uint32)
    ag["reach_condition"][100:200] = _abi.REACH_VISIBLE_AT_DISTANCE
    ag["reach_condition"][200:300] = _abi.REACH_STRAIGHT_PATH_DISTANCE
    ag["reach_distance"] = np.full(600, -1.0, dtype=np.float32)
    ag["reach_distance"][250:300] = 3.0
    opts = synth.default_options()
    rng = np.random.Generator(np.random.PCG64(5))

    def retarget(frame, a):
        if frame == 2:  # some agents get new targets -> repath or path reuse
            t, _ = synth.random_points_on_mesh(mesh, 100, rng)
            a["target"][300:400] = t
        if frame == 3:
            a["flags"][20:30] &= ~np.uint32(_abi.AGENT_PAUSED)

    og = run_frames(gpu, cpu, ag, opts, 6, dt=0.25, retarget=retarget)
    assert (og["state"] == _abi_state("Moving")).sum() > 100


def _abi_state(name):
    from landmass_b200 import AgentState

    return int(AgentState[name])


@pytest.mark.parametrize("mesh_kind,n", [("grid", 1500), ("maze", 800), ("grid64", 9000)])
def test_full_step_with_avoidance(mesh_kind, n):
    """All four phases.  Velocities within 1e-4 relative (north_star tolerance); states,
    pathing results and paths exact.  ORCA parity is CUDA-vs-oracle (dodgy_2d source absent).
    The 9 000-agent case is above the sizes from which k_follow takes the agents longest corridor first and
    k_avoidance in the grid's cell order."""
    mesh = {"grid": lambda: synth.grid_mesh(24, 24), "maze": lambda: synth.maze_mesh(10),
            "grid64": lambda: synth.grid_mesh(64, 64)}[mesh_kind]()
    flat = synth.single_island_flat(mesh)
    gpu, cpu = backends(flat, max_agents=max(4096, n))
    ag = synth.make_agents(mesh, n)   # dense: many neighbours per agent, borders nearby
    opts = synth.default_options()
    ch = {"n": 3,
          "position": np.array([[5.5, 5.5, 0.0], [9.5, 3.5, 0.0], [100.0, 0.0, 0.0]], dtype=np.float32),
          "velocity": np.array([[0.5, 0.0, 0.0], [0.0, -0.5, 0.0], [0.0, 0.0, 0.0]], dtype=np.float32),
          "radius": np.array([0.6, 0.4, 1.0], dtype=np.float32)}
    dt = 1.0 / 30.0
    for frame in range(2 if n > 5000 else 4):
        og = gpu.step(ag, ch, opts, dt)
        oc = cpu.step(ag, ch, opts, dt)
        assert np.array_equal(og["state"], oc["state"]), f"frame {frame}: states"
        assert gpu.pathing_results() == cpu.pathing_results()
        err = np.abs(og["desired_velocity"] - oc["desired_velocity"]).max(axis=1)
        scale = np.maximum(np.abs(oc["desired_velocity"]).max(axis=1), 1e-2)
        assert (err <= 1e-4 * scale + 1e-6).all(), f"frame {frame}: max rel err {(err / scale).max()}"
        ag["flags"] = ag["flags"] & ~np.uint32(_abi.AGENT_RESET)
        ag["velocity"] = oc["desired_velocity"].copy()
        ag["position"] = ag["position"] + oc["desired_velocity"] * np.float32(dt)
    moved = np.abs(oc["desired_velocity"]).max(axis=1) > 0
    assert moved.sum() > n // 2


def test_find_paths_open_list_spills_shared_memory():
    """Long queries on an open grid grow the open list past the shared-memory heap, so the
    first pass reports them for the spilling kernel variant; results must still be exact."""
    mesh = synth.grid_mesh(200, 200)
    flat = synth.single_island_flat(mesh)
    gpu, cpu = backends(flat)
    rng = np.random.Generator(np.random.PCG64(3))
    nq = 48
    sp, sn = synth.random_points_on_mesh(mesh, nq, rng)
    ep, en = synth.random_points_on_mesh(mesh, nq, rng)
    # force a few corner-to-corner queries
    sp[:4] = [[0.5, 0.5, 0], [199.5, 0.5, 0], [0.5, 199.5, 0], [199.5, 199.5, 0]]
    sn[:4] = [0, 199, 199 * 200, 200 * 200 - 1]
    ep[:4] = sp[:4][::-1]
    en[:4] = sn[:4][::-1]
    g = gpu.find_paths(sn, sp, en, ep, path_capacity=1 << 20)
    c = cpu.find_paths(sn, sp, en, ep, path_capacity=1 << 20)
    for a, b, name in zip(g, c, ["success", "explored", "begin", "nodes", "portals"]):
        assert np.array_equal(a, b), name


@pytest.mark.parametrize("mesh_kind", ["grid", "maze"])
def test_straight_paths_bit_exact(mesh_kind):
    """Batched query-API find_path (A* + repeated funnel): every waypoint bitwise equal."""
    mesh = synth.grid_mesh(40, 40) if mesh_kind == "grid" else synth.maze_mesh(16)
    flat = synth.single_island_flat(mesh)
    gpu, cpu = backends(flat)
    rng = np.random.Generator(np.random.PCG64(21))
    nq = 200
    sp, sn = synth.random_points_on_mesh(mesh, nq, rng)
    ep, en = synth.random_points_on_mesh(mesh, nq, rng)
    g = gpu.find_straight_paths(sn, sp, en, ep)
    c = cpu.find_straight_paths(sn, sp, en, ep)
    assert np.array_equal(g[0], c[0]) and np.array_equal(g[1], c[1]) and np.array_equal(g[2], c[2])
    assert np.array_equal(g[3].view(np.uint32), c[3].view(np.uint32))
    assert np.array_equal(g[4], c[4])
    assert (np.diff(g[1].astype(np.int64)) >= 2).all()


@pytest.mark.gpu
def test_queries_between_resident_steps_leave_the_agents_alone():
    """lmb_find_paths / lmb_find_straight_paths carry their own override and permitted-kind arrays;
    the copies lmb_agents_upload left on the device (used by every lmb_step_resident until the next
    upload) must not change when a query with different filters runs in between."""
    from landmass_b200._native import NativeBackend

    types = (np.add.outer(np.arange(24) // 4, np.arange(24) // 4) % 3).astype(np.uint32)
    mesh = synth.grid_mesh(24, 24, types=types)
    flat = synth.single_island_flat(mesh)
    flat["n_type_costs"] = 3
    flat["type_cost_index"] = np.array([0, 1, 2], dtype=np.uint32)
    flat["type_cost_value"] = np.array([1.0, 2.0, 5.0], dtype=np.float32)
    n = 200
    ag = synth.make_agents(mesh, n)
    ag["override_begin"] = np.arange(n + 1, dtype=np.uint32)
    ag["override_type"] = np.full(n, 2, dtype=np.uint32)
    ag["override_cost"] = np.full(n, 0.25, dtype=np.float32)   # type 2 becomes the cheapest
    opts = synth.default_options()

    def run(with_query):
        gpu = NativeBackend(max_agents=n, flags=_abi.CONFIG_NO_AVOIDANCE)
        gpu.navdata_upload(flat)
        gpu.agents_upload(ag, None)
        if with_query:
            sn = np.arange(50, dtype=np.uint32)
            en = np.arange(500, 550, dtype=np.uint32)
            gpu.find_paths(sn, mesh.poly_center[sn], en, mesh.poly_center[en], overrides=[{1: 9.0, 0: 3.0}] * 50,
                           permitted=[[7]] * 50)
            gpu.find_straight_paths(sn, mesh.poly_center[sn], en, mesh.poly_center[en], overrides=[{2: 40.0}] * 50)
        gpu.step_resident(opts, 1.0 / 60.0)
        out = gpu.agents_download()
        res = gpu.pathing_results_array().copy()
        paths = [gpu.agent_path(s) for s in range(0, n, 17)]
        gpu.close()
        return out, res, paths

    (o1, r1, p1), (o2, r2, p2) = run(False), run(True)
    assert np.array_equal(r1, r2)
    assert np.array_equal(o1["desired_velocity"], o2["desired_velocity"])
    for a, b in zip(p1, p2):
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_avoidance_constraints_export():
    """`lmb_agent_avoidance_constraints` (debug.rs:415-503): constraint lines of a dense crowd, CUDA vs
    oracle — counts and Satisfied/Fallback exact, lines within 1e-4 — for two agents in three (the others
    did not ask and report nothing)."""
    mesh = synth.grid_mesh(12, 12)
    flat = synth.single_island_flat(mesh)
    gpu, cpu = backends(flat)
    n = 90  # dense enough that the 2-D program fails for about half of the agents (Fallback)
    ag = synth.make_agents(mesh, n)
    asked = np.arange(n) % 3 != 1
    ag["flags"] = ag["flags"] | np.where(asked, np.uint32(_abi.AGENT_KEEP_AVOIDANCE_DATA), np.uint32(0))
    opts = synth.default_options()
    dt = 1.0 / 30.0
    n_fallback = n_lines = 0
    for frame in range(2):
        og = gpu.step(ag, None, opts, dt)
        oc = cpu.step(ag, None, opts, dt)
        assert np.array_equal(og["state"], oc["state"])
        for i in range(n):
            ran_g, orig_g, fb_g, fell_g = gpu.agent_avoidance_constraints(i)
            ran_c, orig_c, fb_c, fell_c = cpu.agent_avoidance_constraints(i)
            if not asked[i]:
                assert not ran_g and len(orig_g) == 0 and len(fb_g) == 0
                continue
            assert (ran_g, fell_g) == (ran_c, fell_c), f"agent {i}"
            assert orig_g.shape == orig_c.shape and fb_g.shape == fb_c.shape, f"agent {i}"
            for g, c in ((orig_g, orig_c), (fb_g, fb_c)):
                if len(c):
                    scale = np.maximum(np.abs(c).max(axis=1, keepdims=True), 1e-2)
                    assert (np.abs(g - c) <= 1e-4 * scale + 1e-6).all(), f"agent {i}: {np.abs(g - c).max()}"
            n_fallback += int(fell_c)
            n_lines += len(orig_c)
        ag["flags"] = ag["flags"] & ~np.uint32(_abi.AGENT_RESET)
        ag["velocity"] = oc["desired_velocity"].copy()
        ag["position"] = ag["position"] + oc["desired_velocity"] * np.float32(dt)
    assert 0 < n_fallback < 2 * asked.sum() and n_lines > 10 * asked.sum()
    gpu.close()
    cpu.close()
