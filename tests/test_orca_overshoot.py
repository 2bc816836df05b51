"""ORCA leaves a few agents of a dense crowd above `max_speed`.

What is pinned here (avoidance.rs:141-177, dodgy_2d `compute_avoiding_velocity`):
* the reference call site applies NO clamp to what dodgy_2d returns, so neither do the oracle and the kernel;
* the overshoot is a property of the published RVO2 algorithm, not of this port: it only ever happens to an
  agent that OVERLAPS a neighbour (centre distance < r_i + r_j).  The collision branch of the ORCA line then
  pushes with 1/delta_time, the 2-D program is infeasible, and the 3-D fallback (`linear_program_3`) projects
  nearly antiparallel constraints: the intersection point of two such lines lies far from the origin and
  `point + t * direction` loses the bound of the speed disc to cancellation;
* agents that overlap nobody never exceed max_speed (to 1e-5 relative);
* CUDA and the oracle agree on every one of those agents to 1e-4 like on all others.
Whether the real dodgy_2d 0.5.4 crate overshoots in the same way cannot be checked here (its source is absent:
a14 stays "parity unpinned"); the decision is to keep the published behaviour rather than add a clamp the
reference does not have."""
import numpy as np
import pytest

from landmass_b200 import _abi, synth


def dense_crowd(n=3000, side=14):
    """Open field of side x side quads of 8 m at BASELINE config 3's density (0.24 agents / m^2, uniform random
    positions: about half of the agents start overlapping somebody), antipodal targets."""
    mesh = synth.grid_mesh(side, side, cell=8.0)
    flat = synth.single_island_flat(mesh)
    rng = np.random.Generator(np.random.PCG64(synth.SEED))
    pos, _ = synth.random_points_on_mesh(mesh, n, rng)
    ag = synth.make_agents(mesh, n)
    length = side * 8.0
    ag["position"] = pos
    ag["target"] = np.stack([length - pos[:, 0], length - pos[:, 1], pos[:, 2]], axis=1).astype(np.float32)
    return flat, ag


def nearest_neighbour_distance(p):
    from scipy.spatial import cKDTree

    d, _ = cKDTree(p[:, :2]).query(p[:, :2], k=2)
    return d[:, 1]


def test_overshoot_only_with_overlapping_neighbours(make_backend, backend_kind):
    from oracle.oracle_py import OracleBackend

    flat, ag = dense_crowd()
    be = make_backend(max_agents=ag["n"])
    be.navdata_upload(flat)
    ref = None
    if backend_kind != "oracle":
        ref = OracleBackend()
        ref.navdata_upload(flat)
    opts = synth.default_options()
    dt = 1.0 / 60.0
    n_over_total = 0
    for frame in range(4):
        nn = nearest_neighbour_distance(ag["position"])
        out = be.step(ag, None, opts, dt)
        v = out["desired_velocity"]
        speed = np.linalg.norm(v[:, :2], axis=1)
        assert np.isfinite(v).all()
        over = speed > ag["max_speed"] * (1 + 1e-4)
        free = nn >= ag["radius"] * 2  # overlaps nobody
        assert not (over & free).any(), "an agent that overlaps nobody exceeded max_speed"
        assert (speed[free] <= ag["max_speed"][free] * (1 + 1e-5)).all()
        assert over.mean() < 0.01, f"frame {frame}: {int(over.sum())} agents above max_speed"
        n_over_total += int(over.sum())
        if ref is not None:
            oc = ref.step(ag, None, opts, dt)
            assert np.array_equal(out["state"], oc["state"])
            np.testing.assert_allclose(v, oc["desired_velocity"], rtol=1e-4, atol=1e-5)
            v = oc["desired_velocity"]
        ag["flags"] = ag["flags"] & ~np.uint32(_abi.AGENT_RESET)
        ag["velocity"] = v.copy()
        ag["position"] = ag["position"] + v * np.float32(dt)
    # the effect exists (if this ever reads 0 the premise of the note above has changed: update DESIGN.md §4)
    assert n_over_total > 0
    if ref is not None:
        ref.close()
