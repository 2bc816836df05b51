"""CUDA-vs-oracle parity on geometry where no f32 operation is exact: jittered, non-planar meshes of
triangles and quads with holes (synth.irregular_mesh) on rotated, translated islands, three node
types with non-trivial costs, per-agent cost overrides, boundary links between rotated islands and
animation links.  The axis-aligned unit grids of test_parity_synth.py cannot tell a fused
multiply-add or a re-associated sum from the reference's arithmetic; these can.  Bit-exact:
sampled points and nodes, corridors, portal codes, explored_nodes, states, pathing results;
desired velocities within 1e-4 relative."""
import math

import numpy as np
import pytest

from landmass_b200 import AnimationLink, Island, Transform, _abi, synth
from landmass_b200.nav_data import NavigationData, transform_apply

pytestmark = pytest.mark.gpu

T1 = (np.array([13.3, -7.1, 2.5], dtype=np.float32), np.float32(0.7))
COSTS = {0: 1.0, 1: 2.3, 2: 0.7}


def world(points, transform):
    t, r = transform
    return np.stack([transform_apply(t, r, p) for p in points]).astype(np.float32)


def one_island(seed=3, links=True):
    mesh = synth.irregular_mesh(22, 18, seed=seed)
    nd = NavigationData()
    nd.add_island(Island(Transform(tuple(T1[0]), float(T1[1])), mesh))
    for t, c in COSTS.items():
        nd.set_type_index_cost(t, c)
    if links:
        for k, (a, b) in enumerate([((2.2, 2.5), (15.5, 12.5)), ((18.5, 3.5), (4.5, 14.5))]):
            s = world([(a[0], a[1], 0.0), (a[0] + 1.5, a[1] + 0.3, 0.0)], T1)
            e = world([(b[0], b[1], 0.0), (b[0] + 1.5, b[1] - 0.2, 0.0)], T1)
            nd.add_animation_link(AnimationLink(start_edge=(tuple(s[0]), tuple(s[1])), end_edge=(tuple(e[0]), tuple(e[1])),
                                                cost=1.5 + k, kind=k, bidirectional=bool(k)))
    nd.update(0.01, 1.5)
    return mesh, nd.flatten()


def backends(flat, max_agents=2048, flags=0):
    from landmass_b200._native import NativeBackend
    from oracle.oracle_py import OracleBackend

    gpu = NativeBackend(max_agents=max_agents, flags=flags)
    cpu = OracleBackend(flags=flags)
    gpu.navdata_upload(flat)
    cpu.navdata_upload(flat)
    return gpu, cpu


def test_sampling_bit_exact_on_irregular_geometry():
    mesh, flat = one_island()
    assert flat["n_links"] >= 2
    gpu, cpu = backends(flat)
    rng = np.random.Generator(np.random.PCG64(11))
    on, _ = synth.random_points_in_polygons(mesh, 6000, rng)
    on[:, 2] += rng.uniform(-0.6, 0.4, size=len(on)).astype(np.float32)      # above / below / too far below
    on[:2000, :2] += rng.normal(0.0, 0.15, size=(2000, 2)).astype(np.float32)  # off the polygon, into holes
    pts = world(on, T1)
    opts = synth.default_options()
    gp, gn = gpu.sample_points(pts, opts)
    cp, cn = cpu.sample_points(pts, opts)
    assert np.array_equal(gn, cn)
    assert np.array_equal(gp.view(np.uint32), cp.view(np.uint32))  # bitwise
    hit = gn != _abi.NONE
    assert 0.5 < hit.mean() < 0.98


def queries(mesh, n, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    sp, sn = synth.random_points_in_polygons(mesh, n, rng)
    ep, en = synth.random_points_in_polygons(mesh, n, rng)
    overrides = [None if i % 3 else {int(rng.integers(3)): float(rng.uniform(0.3, 6.0))} for i in range(n)]
    permitted = [None if i % 4 else sorted(set(rng.integers(0, 3, size=int(rng.integers(0, 3))).tolist()))
                 for i in range(n)]
    return sn.astype(np.uint32), world(sp, T1), en.astype(np.uint32), world(ep, T1), overrides, permitted


@pytest.mark.parametrize("seed", [3, 5, 8, 13])
def test_find_paths_bit_exact_on_irregular_geometry(seed):
    mesh, flat = one_island(seed=seed)
    gpu, cpu = backends(flat)
    sn, sp, en, ep, overrides, permitted = queries(mesh, 500, 21 + seed)
    g = gpu.find_paths(sn, sp, en, ep, overrides=overrides, permitted=permitted)
    c = cpu.find_paths(sn, sp, en, ep, overrides=overrides, permitted=permitted)
    for a, b, name in zip(g, c, ["success", "explored", "begin", "nodes", "portals"]):
        assert np.array_equal(a, b), name
    assert g[0].sum() > 350
    if seed == 3:
        assert (g[4][g[4] != _abi.NONE] & _abi.PORTAL_LINK).any()  # some corridors use the animation links


def test_straight_paths_bit_exact_on_irregular_geometry():
    mesh, flat = one_island()
    gpu, cpu = backends(flat)
    sn, sp, en, ep, overrides, permitted = queries(mesh, 300, 22)
    g = gpu.find_straight_paths(sn, sp, en, ep, overrides=overrides, permitted=permitted)
    c = cpu.find_straight_paths(sn, sp, en, ep, overrides=overrides, permitted=permitted)
    for a, b, name in zip(g, c, ["success", "begin", "kind", "points", "link"]):
        if name == "points":
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), name
        else:
            assert np.array_equal(a, b), name
    assert (g[2] == 1).any()


def make_agents(mesh, transform, n, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    pos, _ = synth.random_points_in_polygons(mesh, n, rng, lift=0.05)
    tgt, _ = synth.random_points_in_polygons(mesh, n, rng)
    ag = {
        "n": n,
        "slot": np.arange(n, dtype=np.uint32),
        "flags": np.full(n, _abi.AGENT_HAS_TARGET | _abi.AGENT_PERMIT_ALL_LINKS | _abi.AGENT_RESET, dtype=np.uint32),
        "position": world(pos, transform),
        "velocity": np.zeros((n, 3), dtype=np.float32),
        "target": world(tgt, transform),
        "radius": rng.uniform(0.2, 0.45, size=n).astype(np.float32),
        "desired_speed": rng.uniform(0.8, 1.6, size=n).astype(np.float32),
        "max_speed": np.full(n, 2.0, dtype=np.float32),
        "reach_condition": (np.arange(n) % 3).astype(np.uint32),
        "reach_distance": np.where(np.arange(n) % 2, -1.0, 1.2).astype(np.float32),
    }
    ob = np.zeros(n + 1, dtype=np.uint32)
    ot, oc = [], []
    for i in range(n):
        if i % 5 == 0:
            ot.append(int(rng.integers(3)))
            oc.append(float(rng.uniform(0.4, 4.0)))
        ob[i + 1] = len(ot)
    ag["override_begin"], ag["override_type"], ag["override_cost"] = ob, np.array(ot, dtype=np.uint32), np.array(oc, dtype=np.float32)
    pb = np.zeros(n + 1, dtype=np.uint32)
    pk = []
    for i in range(n):
        if i % 7 == 0:  # only kind 1 permitted
            ag["flags"][i] &= ~np.uint32(_abi.AGENT_PERMIT_ALL_LINKS)
            pk.append(1)
        pb[i + 1] = len(pk)
    ag["permitted_begin"], ag["permitted_kind"] = pb, np.array(pk + [0], dtype=np.uint32)
    return ag


def run_frames(gpu, cpu, ag, opts, frames, dt):
    n = ag["n"]
    moving = 0
    for frame in range(frames):
        og = gpu.step(ag, None, opts, dt)
        oc = cpu.step(ag, None, opts, dt)
        assert np.array_equal(og["state"], oc["state"]), f"frame {frame}: states"
        assert gpu.pathing_results() == cpu.pathing_results(), f"frame {frame}: pathing results"
        np.testing.assert_allclose(og["desired_velocity"], oc["desired_velocity"], rtol=1e-4, atol=1e-6,
                                   err_msg=f"frame {frame}")
        assert np.array_equal(og["reached_link_id"], oc["reached_link_id"])
        for slot in range(0, n, max(1, n // 40)):
            gp, cp = gpu.agent_path(slot), cpu.agent_path(slot)
            assert np.array_equal(gp[0], cp[0]) and np.array_equal(gp[1], cp[1]), f"path of slot {slot}"
        moving = max(moving, int((og["state"] == 4).sum()))
        ag["flags"] = ag["flags"] & ~np.uint32(_abi.AGENT_RESET)
        ag["velocity"] = oc["desired_velocity"].copy()
        ag["position"] = ag["position"] + oc["desired_velocity"] * np.float32(dt)
    return moving


@pytest.mark.parametrize("seed", [4, 6])
@pytest.mark.parametrize("avoidance", [False, True])
def test_step_on_irregular_geometry(avoidance, seed):
    mesh, flat = one_island(seed=seed)
    gpu, cpu = backends(flat, flags=0 if avoidance else _abi.CONFIG_NO_AVOIDANCE)
    ag = make_agents(mesh, T1, 700, 31 + seed)
    moving = run_frames(gpu, cpu, ag, synth.default_options(), 8, 0.2)
    assert moving > 400


def test_step_across_rotated_islands_with_boundary_links():
    """Four copies of a straight-bordered irregular mesh tiled 2 x 2 under one rotation; the island
    translations are the rotated offsets rounded to f32, so the shared borders coincide only to
    ~1e-6 and the boundary links come from edge_intersection with edge_link_distance 0.01."""
    mesh = synth.irregular_mesh(10, 10, seed=9, straight_border=True, hole_fraction=0.05)
    rot = np.float32(-1.1)
    base = np.array([-4.0, 9.0, 1.0], dtype=np.float32)
    nd = NavigationData()
    transforms = []
    for iy in range(2):
        for ix in range(2):
            t = transform_apply(base, rot, np.array([10.0 * ix, 10.0 * iy, 0.0], dtype=np.float32)).astype(np.float32)
            transforms.append((t, rot))
            nd.add_island(Island(Transform(tuple(t), float(rot)), mesh))
    for t, c in COSTS.items():
        nd.set_type_index_cost(t, c)
    nd.update(0.01, 0.25)
    flat = nd.flatten()
    assert flat["n_links"] >= 4 * 2 * 8 and flat["n_modified"] > 0
    gpu, cpu = backends(flat)
    parts = [make_agents(mesh, tr, 150, 40 + k) for k, tr in enumerate(transforms)]
    rng = np.random.Generator(np.random.PCG64(77))
    ag = {"n": 600}
    for key in parts[0]:
        if key in ("n", "override_begin", "override_type", "override_cost", "permitted_begin", "permitted_kind"):
            continue
        ag[key] = np.concatenate([p[key] for p in parts])
    ag["slot"] = np.arange(600, dtype=np.uint32)
    ag["flags"][:] = _abi.AGENT_HAS_TARGET | _abi.AGENT_PERMIT_ALL_LINKS | _abi.AGENT_RESET
    ag["target"] = ag["target"][rng.permutation(600)]  # most targets lie on another island
    moving = run_frames(gpu, cpu, ag, synth.default_options(), 8, 0.2)
    assert moving > 400
    crossing = sum(len(set(np.searchsorted(flat["island_poly_begin"], gpu.agent_path(s)[0], side="right"))) > 1
                   for s in range(0, 600, 15))
    assert crossing > 10
