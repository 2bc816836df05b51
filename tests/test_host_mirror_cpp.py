"""include/landmass_b200.hpp — the C++ host mirror of the per-frame API (Archipelago / Agent / Character /
ArchipelagoOptions / AgentState / DenseSlotMap over the C ABI) — against the Python mirror: the same scripted
game loop (agents added and removed, a re-used slot, a paused agent, a cost override, a character), frame by
frame, on the same kind of backend.  The harness (tests/cpp/host_mirror_harness.cpp) is compiled here with g++
and gets the backend's entry points as function pointers: the CPU oracle on CPU, the CUDA library on a GPU box.
Everything must be identical: both mirrors pack the same arrays for the same step function."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from landmass_b200 import (Agent, Archipelago, ArchipelagoOptions, Character, Island, TargetReachedCondition,
                           Transform, _abi)
from tests.helpers import valid_mesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FRAMES, DT, MAX_AGENTS = 40, 0.1, 16


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hm") / "libhost_mirror_harness.so")
    subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-Wall", "-Wextra", "-Werror", "-shared", "-fPIC",
                    "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "host_mirror_harness.cpp"),
                    "-o", out], check=True)
    lib = C.CDLL(out)
    lib.hm_run_script.restype = C.c_int32
    lib.hm_queries.restype = C.c_int32
    lib.hm_validate_flatten.restype = C.c_int32
    lib.hm_validate_flatten_height.restype = C.c_int32
    lib.hm_link_flatten.restype = C.c_int32
    lib.hm_run_world_script.restype = C.c_int32
    lib.hm_sample_edge.restype = C.c_int32
    return lib


def python_trace(backend, debug=False):
    """The script of host_mirror_harness.cpp on the Python mirror.  debug: agent C keeps its avoidance data and the
    debug read-backs of the last frame are returned as a fourth element."""
    arch = Archipelago(ArchipelagoOptions.from_agent_radius(0.5), backend=backend)
    arch.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), valid_mesh(dict(
        vertices=[(0.0, 0.0, 0.0), (15.0, 0.0, 0.0), (15.0, 15.0, 0.0), (0.0, 15.0, 0.0)],
        polygons=[[0, 1, 2, 3]], polygon_type_indices=[0]))))

    def make(pos, target, radius):
        a = Agent.create(np.array(pos, dtype=np.float32), np.zeros(3, dtype=np.float32), radius, 1.0, 2.0)
        a.current_target = target
        a.target_reached_condition = TargetReachedCondition.Distance(0.01)
        return a

    ids = [arch.add_agent(make((1.0, 1.0, 0.0), (11.0, 1.1, 0.0), 1.0)),
           arch.add_agent(make((11.0, 1.1, 0.0), (1.0, 1.0, 0.0), 1.0)), None, None]
    c = make((6.0, 8.0, 0.0), (6.0, 2.0, 0.0), 0.5)
    c.target_reached_condition = TargetReachedCondition.StraightPathDistance(0.25)
    assert c.override_type_index_cost(0, 3.0)
    c.keep_avoidance_data = debug
    ids[2] = arch.add_agent(c)
    alive = [True, True, True, False]
    ch = arch.add_character(Character(np.array((6.0, 4.0, 0.0), dtype=np.float32),
                                      np.array((0.2, 0.0, 0.0), dtype=np.float32), 0.5))
    dt = np.float32(DT)
    trace_f = np.zeros((FRAMES, 4, 6), dtype=np.float32)
    trace_u = np.zeros((FRAMES, 14), dtype=np.uint32)
    for f in range(FRAMES):
        if f == 10:
            arch.remove_agent(ids[0])
            alive[0] = False
        if f == 12:
            ids[3] = arch.add_agent(make((2.0, 12.0, 0.0), (12.0, 12.0, 0.0), 0.5))
            alive[3] = True
        if f == 20:
            arch.get_agent_mut(ids[1]).paused = True
        if f == 25:
            arch.get_agent_mut(ids[1]).paused = False
        arch.update(float(dt))
        for k in range(4):
            if not alive[k]:
                trace_f[f, k] = -1000.0
                trace_u[f, k] = 99
                continue
            a = arch.get_agent_mut(ids[k])
            a.velocity = np.asarray(a.get_desired_velocity(), dtype=np.float32).copy()
            a.position = (np.asarray(a.position, dtype=np.float32) + a.velocity * dt).astype(np.float32)
            trace_f[f, k, :3] = a.position
            trace_f[f, k, 3:] = a.velocity
            trace_u[f, k] = int(a.state())
        chr_ = arch.get_character_mut(ch) if hasattr(arch, "get_character_mut") else arch.characters.get(ch)
        chr_.position = (np.asarray(chr_.position, dtype=np.float32)
                         + np.asarray(chr_.velocity, dtype=np.float32) * dt).astype(np.float32)
        pr = arch.get_pathing_results()
        trace_u[f, 4] = len(pr)
        trace_u[f, 5] = sum(r.explored_nodes for r in pr)
        for k, r in enumerate(pr[:4]):
            trace_u[f, 6 + 2 * k] = r.agent.idx
            trace_u[f, 7 + 2 * k] = r.agent.version
    arch._sync_navdata()
    if not debug:
        return trace_f, trace_u, arch._flat
    path = arch.agent_path(ids[1])
    data = arch.get_agent(ids[2])._avoidance_data
    extras_u = [0 if path is None else len(path[0]), _abi.NONE if path is None else int(path[0][0]),
                len(data[1]), len(data[2]) if data[0] == "Fallback" else 0, int(data[0] == "Fallback")]
    lines = np.concatenate([np.asarray(data[1], dtype=np.float32).reshape(-1, 4)]
                           + ([np.asarray(data[2], dtype=np.float32).reshape(-1, 4)] if data[0] == "Fallback" else []))
    return trace_f, trace_u, arch._flat, (extras_u, lines)


def cpp_trace(harness, ctx, fns, desc, debug_fns=None):
    trace_f = np.zeros((FRAMES, 4, 6), dtype=np.float32)
    trace_u = np.zeros((FRAMES, 14), dtype=np.uint32)
    extras_u, extras_f = np.zeros(8, dtype=np.uint32), np.zeros(4 * 256, dtype=np.float32)
    err = C.create_string_buffer(512)
    st = harness.hm_run_script(ctx, *[C.cast(f, C.c_void_p) for f in fns],
                               C.byref(desc) if desc is not None else None, C.c_uint32(MAX_AGENTS),
                               C.c_uint32(FRAMES), C.c_float(DT), trace_f.ctypes.data_as(C.c_void_p),
                               trace_u.ctypes.data_as(C.c_void_p), err, C.c_uint32(512),
                               *([C.cast(f, C.c_void_p) for f in debug_fns] if debug_fns else [None, None]),
                               extras_u.ctypes.data_as(C.c_void_p), extras_f.ctypes.data_as(C.c_void_p))
    assert st == 0, err.value.decode()
    if debug_fns:
        return trace_f, trace_u, None, (extras_u[:5].tolist(), extras_f[:4 * int(extras_u[2] + extras_u[3])].reshape(-1, 4))
    return trace_f, trace_u


def check(py, cpp):
    assert np.array_equal(py[1], cpp[1]), "states / pathing results differ"
    assert np.array_equal(py[0].view(np.uint32), cpp[0].view(np.uint32)), "positions / velocities differ"
    states = py[1][:, :4]
    assert (states[:10, 0] != 99).all() and (states[10:, 0] == 99).all()      # A removed at frame 10
    assert (states[:12, 3] == 99).all() and (states[12:, 3] != 99).all()      # D added at frame 12 ...
    assert py[1][12, 4] >= 1 and (py[1][12, 6:8] == py[1][0, 6:8] + [0, 2]).all()  # ... in A's slot, next version
    assert (states[20:25, 1] == 8).all() and (states[25:, 1] != 8).all()      # B paused for five frames


KINDS = ["TypeIndicesHaveWrongLength", "NotEnoughVerticesInPolygon", "InvalidVertexIndexInPolygon",
         "DegenerateEdgeInPolygon", "DoublyConnectedEdge", "ConcavePolygon", "IncorrectNumberOfPolygons",
         "InvalidPolygonIndices", "InvalidIndexInTriangle", "ClockwiseTriangle"]


def cpp_validate_flatten(harness, vertices, polygons, types, translation=(0.0, 0.0, 0.0), rotation=0.0, height=None):
    v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
    begin = np.zeros(len(polygons) + 1, dtype=np.uint32)
    begin[1:] = np.cumsum([len(p) for p in polygons])
    pv = np.array([x for p in polygons for x in p], dtype=np.uint32)
    t = np.asarray(types, dtype=np.uint32)
    P, E, V = len(polygons), len(pv), len(v)
    out_u = np.zeros(7 * P + 3 * E + 8, dtype=np.uint32)
    out_f = np.zeros(3 * V + 9 * P + 1, dtype=np.float32)
    out_i = np.zeros(10, dtype=np.float32)
    err = np.zeros(2, dtype=np.uint64)
    tr = np.asarray(translation, dtype=np.float32)
    p_ = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    if height is None:
        st = harness.hm_validate_flatten(C.c_uint32(V), p_(v), C.c_uint32(P), p_(begin), p_(pv), C.c_uint32(len(t)), p_(t),
                                         p_(tr), C.c_float(rotation), p_(err), p_(out_u), p_(out_f), p_(out_i))
    else:
        hp = np.array([[q.base_vertex_index, q.vertex_count, q.base_triangle_index, q.triangle_count]
                       for q in height.polygons], dtype=np.uint32).reshape(-1, 4)
        hv = np.ascontiguousarray(height.vertices, dtype=np.float32).reshape(-1, 3)
        ht = np.ascontiguousarray(height.triangles, dtype=np.uint32).reshape(-1, 3)
        out_hu = np.zeros(2 + 4 * P + 3 * len(ht) + 4, dtype=np.uint32)
        out_hf = np.zeros(3 * len(hv) + 3, dtype=np.float32)
        st = harness.hm_validate_flatten_height(
            C.c_uint32(V), p_(v), C.c_uint32(P), p_(begin), p_(pv), C.c_uint32(len(t)), p_(t), p_(tr), C.c_float(rotation),
            C.c_uint32(len(hp)), p_(hp), C.c_uint32(len(hv)), p_(hv), C.c_uint32(len(ht)), p_(ht), p_(err), p_(out_u),
            p_(out_f), p_(out_i), p_(out_hu), p_(out_hf))
    assert st >= 0
    if st != 0:
        return KINDS[st - 1], int(err[0]), int(err[1])
    u = iter(np.split(out_u[:5 * P + 3 * E + 6], np.cumsum([P + 1, E, E, E, P, P, P + 1, P, 2, 2])[:-1]))
    f = iter(np.split(out_f[:3 * V + 9 * P], np.cumsum([3 * V, 6 * P])))
    return dict(poly_edge_begin=next(u), edge_vertex=next(u), edge_neighbor=next(u), edge_reverse=next(u),
                poly_type=next(u), poly_region=next(u), node_link_begin=next(u), node_region_number=next(u),
                island_poly_begin=next(u), island_vert_begin=next(u), vertices=next(f).reshape(-1, 3),
                poly_bounds=next(f).reshape(-1, 6), poly_center=next(f).reshape(-1, 3), island=out_i,
                n_edges=int(err[0]), n_other=int(err[1]),
                **({} if height is None else dict(
                    n_hverts=int(out_hu[0]), n_htris=int(out_hu[1]), hpoly=out_hu[2:2 + 4 * P].reshape(-1, 4),
                    htris=out_hu[2 + 4 * P:2 + 4 * P + 3 * int(out_hu[1])].reshape(-1, 3),
                    hverts=out_hf[:3 * int(out_hu[0])].reshape(-1, 3))))


@pytest.mark.parametrize("mesh_kind", ["grid", "maze", "irregular", "rosette12", "two_regions"])
def test_cpp_validate_and_flatten_match_the_python_mirror(harness, mesh_kind):
    """`NavigationMesh::validate` + the single-island flatten of the C++ mirror against landmass_b200.nav_mesh /
    nav_data.flatten: every array of the descriptor, bit for bit."""
    from landmass_b200 import synth

    if mesh_kind == "two_regions":
        quad = lambda x: [(x, 0.0, 0.0), (x + 1.0, 0.0, 0.5), (x + 1.0, 1.0, 0.0), (x, 1.0, 0.25)]  # noqa: E731
        mesh = valid_mesh(dict(vertices=quad(0.0) + quad(1.0)[1:3] + quad(5.0),
                               polygons=[[0, 1, 2, 3], [1, 4, 5, 2], [6, 7, 8, 9]], polygon_type_indices=[0, 2, 1]))
    else:
        mesh = {"grid": lambda: synth.grid_mesh(9, 7, types=(np.arange(63).reshape(7, 9) % 3).astype(np.uint32)),
                "maze": lambda: synth.maze_mesh(9), "irregular": lambda: synth.irregular_mesh(10, 8),
                "rosette12": lambda: synth.rosette_mesh(12, 3)}[mesh_kind]()
    polygons = [mesh.polygon_vertices(p).tolist() for p in range(mesh.n_polys)]
    translation, rotation = (3.0, -2.0, 0.5), 0.7
    from landmass_b200.nav_data import NavigationData

    nd = NavigationData()
    nd.add_island(Island(Transform(translation, rotation), mesh))
    nd.update(0.01, 0.25)
    flat = nd.flatten()
    got = cpp_validate_flatten(harness, mesh.vertices, polygons, mesh.poly_type, translation, rotation)
    assert isinstance(got, dict), got
    for name in ["poly_edge_begin", "edge_vertex", "edge_neighbor", "edge_reverse", "poly_type", "poly_region",
                 "node_link_begin", "node_region_number", "island_poly_begin", "island_vert_begin"]:
        assert np.array_equal(got[name], np.asarray(flat[name], dtype=np.uint32)), name
    for name in ["vertices", "poly_bounds", "poly_center"]:
        assert np.array_equal(got[name].view(np.uint32), np.asarray(flat[name], dtype=np.float32).view(np.uint32)), name
    want_island = np.concatenate([flat["island_translation"][0], flat["island_rotation"][:1], flat["island_bounds"][0]])
    assert np.array_equal(got["island"], want_island.astype(np.float32))
    assert got["n_edges"] == flat["n_edges"] and got["n_other"] == 0
    assert flat["n_links"] == 0 and flat["n_modified"] == 0 and flat["n_region_numbers"] == 0


def test_cpp_validation_errors_match_the_python_mirror(harness):
    """Every ValidationError of nav_mesh.rs:131-162 that does not involve a height mesh, same variant and payload."""
    from landmass_b200 import NavigationMesh, ValidationError

    sq = [(0.0, 0.0, 0.0), (1.0, 0.0, 0.0), (1.0, 1.0, 0.0), (0.0, 1.0, 0.0), (2.0, 0.0, 0.0), (2.0, 1.0, 0.0),
          (0.5, 0.5, 0.0)]
    cases = [
        (sq, [[0, 1, 2, 3]], [0, 1]),                            # TypeIndicesHaveWrongLength
        (sq, [[0, 1, 2, 3], [1, 4]], [0, 0]),                    # NotEnoughVerticesInPolygon
        (sq, [[0, 1, 2, 3], [1, 4, 9]], [0, 0]),                 # InvalidVertexIndexInPolygon
        (sq, [[0, 1, 2, 3], [1, 4, 4, 5, 2]], [0, 0]),           # DegenerateEdgeInPolygon
        (sq, [[0, 1, 2, 3], [1, 4, 5, 2], [2, 1, 6]], [0, 0, 0]),  # DoublyConnectedEdge
        (sq, [[0, 1, 6, 2, 3]], [0]),                            # ConcavePolygon
        (sq, [[0, 3, 2, 1]], [0]),                               # clockwise: ConcavePolygon
    ]
    seen = set()
    for vertices, polygons, types in cases:
        with pytest.raises(ValidationError) as e:
            NavigationMesh(vertices=vertices, polygons=polygons, polygon_type_indices=types).validate()
        got = cpp_validate_flatten(harness, vertices, polygons, types)
        assert isinstance(got, tuple), (polygons, got)
        want_args = tuple(int(a) for a in e.value.args_) + (0,) * (2 - len(e.value.args_))
        assert got == (e.value.kind,) + want_args, (polygons, got, e.value)
        seen.add(got[0])
    assert seen == set(KINDS[:6])


def test_cpp_height_mesh_validation_and_flatten_match_the_python_mirror(harness):
    """HeightNavigationMesh::validate (nav_mesh.rs:410-479): the flattened height arrays and the bounds taken from the
    height vertices, and the four height-mesh errors, against the Python mirror."""
    from landmass_b200 import HeightNavigationMesh, HeightPolygon, NavigationMesh, ValidationError
    from landmass_b200.nav_data import NavigationData
    from tests.helpers import create_height_mesh

    verts = [(0.0, 0.0, 0.0), (2.0, 0.0, 0.0), (2.0, 2.0, 0.0), (0.0, 2.0, 0.0), (4.0, 0.0, 0.0), (4.0, 2.0, 0.0)]
    polys = [[0, 1, 2, 3], [1, 4, 5, 2]]
    hsrc = [(0.0, 0.0, 0.3), (2.0, 0.0, 0.1), (2.0, 2.0, 0.4), (0.0, 2.0, 0.2), (1.0, 1.0, 1.5), (4.0, 0.0, -0.2),
            (4.0, 2.0, 0.6)]
    good = create_height_mesh(hsrc, [[[0, 1, 4], [1, 2, 4], [2, 3, 4], [3, 0, 4]], [[1, 5, 6, 2]]])
    mesh = NavigationMesh(vertices=verts, polygons=polys, polygon_type_indices=[0, 1], height_mesh=good).validate()
    nd = NavigationData()
    nd.add_island(Island(Transform((1.0, 2.0, 3.0), 0.0), mesh))
    nd.update(0.01, 0.25)
    flat = nd.flatten()
    got = cpp_validate_flatten(harness, verts, polys, [0, 1], (1.0, 2.0, 3.0), 0.0, height=good)
    assert isinstance(got, dict), got
    assert got["n_hverts"] == flat["n_hverts"] and got["n_htris"] == flat["n_htris"]
    assert np.array_equal(got["hpoly"], flat["hpoly"]) and np.array_equal(got["htris"], flat["htris"].astype(np.uint32))
    for name in ["hverts", "poly_bounds", "poly_center", "vertices"]:
        assert np.array_equal(got[name].view(np.uint32), np.asarray(flat[name], dtype=np.float32).view(np.uint32)), name
    assert np.array_equal(got["island"][4:], flat["island_bounds"][0]) and got["island"][9] == np.float32(1.5)

    one = HeightPolygon(0, 3, 0, 1)
    tri_v = [(0.0, 0.0, 0.0), (1.0, 0.0, 0.0), (0.0, 1.0, 0.0)]
    bad = [
        HeightNavigationMesh(polygons=[one], vertices=tri_v, triangles=[[0, 1, 2]]),                     # IncorrectNumberOfPolygons
        HeightNavigationMesh(polygons=[one, HeightPolygon(0, 3, 1, 1)], vertices=tri_v, triangles=[[0, 1, 2]]),  # InvalidPolygonIndices
        HeightNavigationMesh(polygons=[one, HeightPolygon(0, 2, 0, 1)], vertices=tri_v, triangles=[[0, 1, 2]]),  # InvalidIndexInTriangle
        HeightNavigationMesh(polygons=[one, HeightPolygon(0, 3, 1, 1)], vertices=tri_v,
                             triangles=[[0, 1, 2], [0, 2, 1]]),                                            # ClockwiseTriangle
    ]
    seen = set()
    for h in bad:
        with pytest.raises(ValidationError) as e:
            NavigationMesh(vertices=verts, polygons=polys, polygon_type_indices=[0, 1], height_mesh=h).validate()
        got = cpp_validate_flatten(harness, verts, polys, [0, 1], height=h)
        want_args = tuple(int(a) for a in e.value.args_) + (0,) * (2 - len(e.value.args_))
        assert got == (e.value.kind,) + want_args, (got, e.value)
        seen.add(got[0])
    assert seen == set(KINDS[6:])


def test_cpp_host_mirror_builds_its_own_navigation_data(harness):
    """The scripted game loop again, but with the room validated and flattened by the C++ mirror."""
    from oracle import oracle_py
    from oracle.oracle_py import OracleBackend

    py = python_trace(OracleBackend())
    desc, keep = _abi.make_navdata_desc(py[2])
    lib = oracle_py.load()
    lib.orc_create.restype = C.c_void_p
    ctx = C.c_void_p(lib.orc_create(C.byref(desc), C.c_uint32(0)))
    cpp = cpp_trace(harness, ctx, (lib.orc_navdata_upload, lib.orc_step, lib.orc_pathing_results), None)
    check(py, cpp)
    del keep


def _cpp_link_flatten(harness, islands):
    """islands: [(translation, rotation, ValidNavigationMesh)] -> dict of the descriptor arrays the C++ mirror
    (include/landmass_b200_linking.hpp) builds."""
    vb, pb, peb, verts, pv, types = [0], [0], [0], [], [], []
    for _, _, m in islands:
        verts.append(m.vertices)
        vb.append(vb[-1] + len(m.vertices))
        pb.append(pb[-1] + m.n_polys)
        for p in range(m.n_polys):
            v = m.polygon_vertices(p)
            pv.extend(v.tolist())
            peb.append(peb[-1] + len(v))
        types.extend(m.poly_type.tolist())
    arr = lambda x, dt: np.ascontiguousarray(x, dtype=dt)  # noqa: E731
    vb, pb, peb, pv, types = (arr(x, np.uint32) for x in (vb, pb, peb, pv, types))
    verts = arr(np.concatenate(verts), np.float32)
    tr = arr([i[0] for i in islands], np.float32)
    rot = arr([i[1] for i in islands], np.float32)
    ou, of = np.zeros(1 << 20, dtype=np.uint32), np.zeros(1 << 20, dtype=np.float32)
    nu, nf, err = C.c_uint32(), C.c_uint32(), C.create_string_buffer(256)
    p_ = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    st = harness.hm_link_flatten(C.c_uint32(len(islands)), p_(vb), p_(verts), p_(pb), p_(peb), p_(pv), p_(types), p_(tr),
                                 p_(rot), p_(ou), C.c_uint32(len(ou)), p_(of), C.c_uint32(len(of)), C.byref(nu),
                                 C.byref(nf), err, C.c_uint32(256))
    assert st == 0, err.value.decode()
    ou, of = ou[:nu.value], of[:nf.value]
    P, E, L, M, R, V = (int(x) for x in ou[:6])
    out, pos = dict(n_polys=P, n_edges=E, n_links=L, n_modified=M, n_region_numbers=R, n_vertices=V), 6
    n_isl = len(islands)
    for name, n in [("island_poly_begin", n_isl + 1), ("island_vert_begin", n_isl + 1), ("poly_edge_begin", P + 1),
                    ("edge_vertex", E), ("edge_neighbor", E), ("edge_reverse", E), ("poly_type", P), ("poly_region", P),
                    ("link_dst_node", L), ("link_dst_type", L), ("link_kind", L), ("link_reverse", L),
                    ("node_link_begin", P + 1), ("node_links", L), ("modified_node", M), ("modified_edge_begin", M + 1)]:
        out[name] = ou[pos:pos + n]
        pos += n
    ne = int(out["modified_edge_begin"][-1])
    out["modified_edges"] = ou[pos:pos + 2 * ne]
    pos += 2 * ne
    out["modified_vert_begin"] = ou[pos:pos + M + 1]
    pos += M + 1
    out["node_region_number"] = ou[pos:pos + P]
    pos += P
    out["region_root"] = ou[pos:pos + R]
    assert pos + R == len(ou)
    fp = 0
    for name, n in [("vertices", 3 * V), ("poly_bounds", 6 * P), ("poly_center", 3 * P), ("island_bounds", 6 * n_isl),
                    ("link_portal", 6 * L), ("modified_verts", 2 * int(out["modified_vert_begin"][-1]))]:
        out[name] = of[fp:fp + n]
        fp += n
    assert fp == len(of)
    return out


@pytest.mark.parametrize("world", ["misaligned_pair", "tiling_3x3", "apart", "rotated_pi", "rotated_half_pi"])
def test_cpp_island_linking_matches_the_python_mirror(harness, world):
    """include/landmass_b200_linking.hpp — boundary links (`edge_intersection`, link order and reverse indices),
    modified nodes (clipped boundaries, new vertices), region numbers / roots and the multi-island flatten — against
    landmass_b200.linking + nav_data.flatten: every array of the descriptor, floats bit for bit."""
    from landmass_b200 import synth
    from landmass_b200.nav_data import NavigationData

    types = (np.add.outer(np.arange(8) // 2, np.arange(8) // 2) % 3).astype(np.uint32)
    grid = synth.grid_mesh(8, 8, types=types)
    if world == "misaligned_pair":   # avoidance_test.rs:855-925: redundant vertices, clipped borders
        red = valid_mesh(dict(vertices=[(0.0, 0.0, 0.0), (1.0, 0.0, 0.0), (1.0, 2.0, 0.0), (0.0, 2.0, 0.0),
                                        (0.0, 0.6, 0.0), (0.0, 0.2, 0.0)],
                              polygons=[[0, 1, 2, 3, 4, 5]], polygon_type_indices=[0]))
        islands = [((0.0, 0.0, 0.0), 0.0, red), ((1.0, 0.25, 0.0), 0.0, red)]
    elif world == "tiling_3x3":      # BASELINE config 4 in small: typed grids tiled by pure translations
        islands = [((8.0 * x, 8.0 * y, 0.0), 0.0, grid) for y in range(3) for x in range(3)]
    elif world == "apart":
        islands = [((0.0, 0.0, 0.0), 0.0, synth.maze_mesh(5)), ((40.0, 0.0, 0.0), 0.0, grid)]
    elif world == "rotated_pi":
        islands = [((0.0, 0.0, 0.0), 0.0, grid), ((16.0, 8.0, 0.0), float(np.pi), grid)]
    else:
        islands = [((0.0, 0.0, 0.0), 0.0, grid), ((16.0, 0.0, 0.0), float(np.pi / 2), grid)]
    nd = NavigationData()
    for t, r, m in islands:
        nd.add_island(Island(Transform(t, r), m))
    nd.update(0.01, 0.25)
    flat = nd.flatten()
    got = _cpp_link_flatten(harness, islands)
    for name in ["n_polys", "n_edges", "n_links", "n_modified", "n_region_numbers", "n_vertices"]:
        assert got[name] == flat[name], name
    for name, value in got.items():
        if isinstance(value, int):
            continue
        want = np.asarray(flat[name]).ravel()
        assert len(value) == len(want), name
        if value.dtype == np.float32:
            assert np.array_equal(value.view(np.uint32), want.astype(np.float32).view(np.uint32)), name
        else:
            assert np.array_equal(value, want.astype(np.uint32)), name
    if world == "tiling_3x3":
        assert got["n_links"] == 192 and got["n_modified"] == 176 and got["n_region_numbers"] == 9
        assert len(set(got["region_root"].tolist())) == 1          # one connected archipelago
    if world == "apart":
        assert got["n_links"] == 0 and (got["node_region_number"] == _abi.NONE).all()
    if world == "misaligned_pair":
        assert got["n_links"] == 6 and len(got["modified_verts"]) > 0


WORLD_FRAMES, WORLD_DT = 30, 0.25


def _python_world_trace(backend):
    """The script of hm_run_world_script on the Python mirror."""
    from landmass_b200 import synth

    mesh = synth.grid_mesh(6, 6)
    arch = Archipelago(ArchipelagoOptions.from_agent_radius(0.5), backend=backend)
    island_at = lambda x, y: Island(Transform((x, y, 0.0), 0.0), mesh)  # noqa: E731
    arch.add_island(island_at(0.0, 0.0))
    b = arch.add_island(island_at(6.0, 0.0))
    c = arch.add_island(island_at(12.0, 0.0))

    def make(pos, target):
        a = Agent.create(np.array(pos, dtype=np.float32), np.zeros(3, dtype=np.float32), 0.4, 2.0, 3.0)
        a.current_target = target
        return a

    ids = [arch.add_agent(make((0.7, 0.8, 0.0), (17.2, 5.1, 0.0))), arch.add_agent(make((1.5, 3.1, 0.0), (16.5, 2.2, 0.0))),
           arch.add_agent(make((0.9, 5.2, 0.0), (13.4, 0.8, 0.0))), arch.add_agent(make((16.8, 3.3, 0.0), (0.6, 2.9, 0.0)))]
    dt = np.float32(WORLD_DT)
    trace_f = np.zeros((WORLD_FRAMES, 4, 6), dtype=np.float32)
    trace_u = np.zeros((WORLD_FRAMES, 7), dtype=np.uint32)
    for f in range(WORLD_FRAMES):
        if f == 8:
            arch.add_island(island_at(6.0, 6.0))
        if f == 14:
            arch.remove_island(c)
        if f == 18:
            arch.get_island_mut(b).set_nav_mesh(mesh)
        if f == 22:
            c = arch.add_island(island_at(12.0, 0.0))
        if f == 25:
            arch.set_type_index_cost(0, 2.0)
        arch.update(float(dt))
        for k in range(4):
            a = arch.get_agent_mut(ids[k])
            a.velocity = np.asarray(a.get_desired_velocity(), dtype=np.float32).copy()
            a.position = (np.asarray(a.position, dtype=np.float32) + a.velocity * dt).astype(np.float32)
            trace_f[f, k, :3] = a.position
            trace_f[f, k, 3:] = a.velocity
            trace_u[f, k] = int(a.state())
        pr = arch.get_pathing_results()
        trace_u[f, 4] = len(pr)
        trace_u[f, 5] = sum(r.explored_nodes for r in pr)
        trace_u[f, 6] = arch._flat["n_links"]
    return trace_f, trace_u


def _cpp_world_trace(harness, ctx, fns):
    trace_f = np.zeros((WORLD_FRAMES, 4, 6), dtype=np.float32)
    trace_u = np.zeros((WORLD_FRAMES, 7), dtype=np.uint32)
    err = C.create_string_buffer(512)
    st = harness.hm_run_world_script(ctx, *[C.cast(f, C.c_void_p) for f in fns], C.c_uint32(WORLD_FRAMES),
                                     C.c_float(WORLD_DT), trace_f.ctypes.data_as(C.c_void_p),
                                     trace_u.ctypes.data_as(C.c_void_p), err, C.c_uint32(512))
    assert st == 0, err.value.decode()
    return trace_f, trace_u


def _check_world(py, cpp):
    assert np.array_equal(py[1], cpp[1]), "states / pathing results / link counts differ"
    assert np.array_equal(py[0].view(np.uint32), cpp[0].view(np.uint32)), "positions / velocities differ"
    links = py[1][:, 6]
    assert links[0] == 24 and links[8] == 36 and links[14] == 24 and links[22] == 36   # 6 edges x 2 per border
    n_results = py[1][:, 4]
    assert n_results[0] == 4                       # everybody plans in the first frame
    assert n_results[8] == 0                       # a new island invalidates nothing
    assert n_results[14] >= 1 and n_results[18] >= 1 and n_results[25] == 0   # removal / touch re-plan, costs do not


def test_cpp_island_archipelago_matches_the_python_mirror_on_the_oracle(harness):
    """`BasicIslandArchipelago` (islands in a DenseSlotMap, re-link + lmb_navdata_update with the island / link remap
    on change) against the Python `Archipelago` through a changing world: islands added, removed, touched, type costs
    changed.  Identical traces mean identical descriptors, identical remaps (the same paths survive each change)
    and identical packing."""
    from oracle import oracle_py
    from oracle.oracle_py import OracleBackend

    py = _python_world_trace(OracleBackend())
    lib = oracle_py.load()
    lib.orc_create.restype = C.c_void_p
    # the oracle wants navigation data at creation: any descriptor will do, the script uploads its own first
    from landmass_b200 import synth

    desc, keep = _abi.make_navdata_desc(synth.single_island_flat(synth.grid_mesh(2, 2)))
    ctx = C.c_void_p(lib.orc_create(C.byref(desc), C.c_uint32(0)))
    cpp = _cpp_world_trace(harness, ctx, (lib.orc_navdata_upload, lib.orc_navdata_update, lib.orc_step,
                                         lib.orc_pathing_results))
    _check_world(py, cpp)
    del keep


def _cpp_clip(harness, tri, edge, mv):
    tri = np.ascontiguousarray(tri, dtype=np.float32).reshape(-1, 9)
    edge = np.ascontiguousarray(edge, dtype=np.float32).reshape(-1, 6)
    out, ok = np.zeros((len(tri), 2), dtype=np.float32), np.zeros(len(tri), dtype=np.uint32)
    p_ = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    harness.hm_clip_edges(C.c_uint32(len(tri)), p_(tri), p_(edge), C.c_float(mv), p_(out), p_(ok))
    return out, ok


def test_cpp_clip_edge_to_triangle_matches_goldens_and_python_mirror(harness):
    """include/landmass_b200_anim_links.hpp `clip_edge_to_triangle` (geometry.rs:63-170): the reference's known
    answers (geometry_test.rs:79-121, 369-412, 503-526) and 3 000 random cases — degenerate triangles and
    axis-parallel edges among them — against landmass_b200.anim_links, bit for bit (None included)."""
    import warnings

    from landmass_b200.anim_links import clip_edge_to_triangle

    flat = lambda z: [(1.0, 1.0, z), (11.0, 1.0, z), (6.0, 11.0, z)]  # noqa: E731
    out, ok = _cpp_clip(harness, [flat(7.0), flat(5.0), flat(5.0), [(1, 1, 0), (11, 1, 10), (6, 11, 5)]],
                        [[(0, 6, 7), (12, 6, 7)], [(1, 6, 1), (11, 6, 1)], [(1, 6, 1), (11, 6, 1)], [(1, 6, 5), (11, 6, 5)]], 1.0)
    assert ok[0] == 1 and tuple(out[0]) == (np.float32(0.29166666), np.float32(0.70833333))
    assert ok[1] == 0                                       # 4 below the triangle with max distance 1
    out5, ok5 = _cpp_clip(harness, [flat(5.0)], [[(1, 6, 1), (11, 6, 1)]], 5.0)
    assert ok5[0] == 1 and tuple(out5[0]) == (0.25, 0.75)
    out2, ok2 = _cpp_clip(harness, [[(1, 1, 0), (11, 1, 10), (6, 11, 5)]], [[(1, 6, 5), (11, 6, 5)]], 2.0)
    assert ok2[0] == 1 and tuple(out2[0]) == (np.float32(0.3), np.float32(0.7))

    rng = np.random.Generator(np.random.PCG64(3))
    n = 3000
    tri = rng.uniform(-3, 12, (n, 3, 3)).astype(np.float32)
    tri[:, :, 2] = rng.uniform(0, 4, (n, 3))
    edge = rng.uniform(-3, 12, (n, 2, 3)).astype(np.float32)
    edge[:, :, 2] = rng.uniform(0, 4, (n, 2))
    edge[:300, 1, 0] = edge[:300, 0, 0]
    edge[300:600, 1, 1] = edge[300:600, 0, 1]
    tri[600:700, 2] = tri[600:700, 1]
    edge[700:800, 1, 2] = edge[700:800, 0, 2]
    tri[700:800, :, 2] = 1.0
    out, ok = _cpp_clip(harness, tri, edge, 1.0)
    hits = 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(n):
            want = clip_edge_to_triangle([tri[i, 0], tri[i, 1], tri[i, 2]], [edge[i, 0], edge[i, 1]], 1.0)
            assert (want is not None) == bool(ok[i]), i
            if want is not None:
                hits += 1
                assert np.array_equal(np.array(want, dtype=np.float32).view(np.uint32), out[i].view(np.uint32)), i
    assert hits > 300


def _cpp_sample_edge(harness, vertices, polygons, height, edge, mv):
    v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
    begin = np.zeros(len(polygons) + 1, dtype=np.uint32)
    begin[1:] = np.cumsum([len(p) for p in polygons])
    pv = np.array([x for p in polygons for x in p], dtype=np.uint32)
    if height is not None:
        hp = np.array([[q.base_vertex_index, q.vertex_count, q.base_triangle_index, q.triangle_count]
                       for q in height.polygons], dtype=np.uint32).reshape(-1, 4)
        hv = np.ascontiguousarray(height.vertices, dtype=np.float32).reshape(-1, 3)
        ht = np.ascontiguousarray(height.triangles, dtype=np.uint32).reshape(-1, 3)
    else:
        hp, hv, ht = np.zeros((0, 4), np.uint32), np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32)
    e = np.ascontiguousarray(edge, dtype=np.float32).reshape(6)
    on, ot = np.zeros(4096, dtype=np.uint32), np.zeros((4096, 2), dtype=np.float32)
    p_ = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    k = harness.hm_sample_edge(C.c_uint32(len(v)), p_(v), C.c_uint32(len(polygons)), p_(begin), p_(pv), C.c_uint32(len(hp)),
                               p_(hp), C.c_uint32(len(hv)), p_(hv), C.c_uint32(len(ht)), p_(ht), p_(e), C.c_float(mv),
                               C.c_uint32(4096), p_(on), p_(ot))
    assert k >= 0
    return [(int(on[i]), ot[i].copy()) for i in range(k)]


def test_cpp_sample_edge_matches_the_python_mirror(harness):
    """`ValidNavigationMesh::sample_edge` (nav_mesh.rs:819-923) of the C++ mirror: random edges over grid, irregular,
    maze and heptagon meshes and over a height mesh — the same nodes and merged intervals, bit for bit."""
    import warnings

    from landmass_b200 import NavigationMesh, synth
    from landmass_b200.anim_links import sample_edge
    from tests.helpers import create_height_mesh

    rng = np.random.Generator(np.random.PCG64(11))
    portions = 0

    def compare(mesh, vertices, polygons, height, e, mv):
        nonlocal portions
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = sample_edge(mesh, (e[0], e[1]), mv)
        got = _cpp_sample_edge(harness, vertices, polygons, height, e, mv)
        assert [w[0] for w in want] == [g[0] for g in got]
        for w, g in zip(want, got):
            assert np.array_equal(np.array(w[1], dtype=np.float32).view(np.uint32), g[1].view(np.uint32))
        portions += len(want)

    for m in (synth.grid_mesh(7, 5), synth.irregular_mesh(9, 7), synth.maze_mesh(5), synth.rosette_mesh(7, 3)):
        polygons = [m.polygon_vertices(p).tolist() for p in range(m.n_polys)]
        lo, hi = m.vertices.min(axis=0), m.vertices.max(axis=0)
        for q in range(40):
            e = rng.uniform(lo - 1, hi + 1, (2, 3)).astype(np.float32)
            e[:, 2] = rng.uniform(-0.3, 0.3, 2)
            if q % 5 == 0:
                e[1, 0] = e[0, 0]
            if q % 7 == 0:
                e[1, 1] = e[0, 1]
            compare(m, m.vertices, polygons, None, e, 0.25)
    verts = [(0.0, 0.0, 0.0), (2.0, 0.0, 0.0), (2.0, 2.0, 0.0), (0.0, 2.0, 0.0), (4.0, 0.0, 0.0), (4.0, 2.0, 0.0)]
    polys = [[0, 1, 2, 3], [1, 4, 5, 2]]
    hsrc = [(0.0, 0.0, 0.3), (2.0, 0.0, 0.1), (2.0, 2.0, 0.4), (0.0, 2.0, 0.2), (1.0, 1.0, 1.5), (4.0, 0.0, -0.2),
            (4.0, 2.0, 0.6)]
    height = create_height_mesh(hsrc, [[[0, 1, 4], [1, 2, 4], [2, 3, 4], [3, 0, 4]], [[1, 5, 6, 2]]])
    hm = NavigationMesh(vertices=verts, polygons=polys, polygon_type_indices=[0, 0], height_mesh=height).validate()
    for q in range(40):
        e = np.stack([rng.uniform(-0.5, 4.5, 2), rng.uniform(-0.5, 2.5, 2), rng.uniform(-0.2, 1.6, 2)], axis=1).astype(np.float32)
        compare(hm, verts, polys, height, e, 0.3)
    assert portions > 400


def _query_case(backend):
    """A maze island and 40 (start, end) pairs: most on the mesh, some off it, some with a cost override."""
    from landmass_b200 import FindPathError, SamplePointError, synth

    mesh = synth.maze_mesh(8)
    arch = Archipelago(ArchipelagoOptions.from_agent_radius(0.5), backend=backend)
    arch.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), mesh))
    arch.update(0.1)
    rng = np.random.Generator(np.random.PCG64(5))
    n = 40
    starts, _ = synth.random_points_on_mesh(mesh, n, rng)
    ends, _ = synth.random_points_on_mesh(mesh, n, rng)
    starts[3] = (-5.0, -5.0, 0.0)   # off the mesh
    ends[7] = (100.0, 3.0, 0.0)
    cost = np.where(np.arange(n) % 4 == 1, 2.5, 0.0).astype(np.float32)
    want_n, want_steps = [], []
    for q in range(n):
        try:
            a = arch.sample_point(tuple(starts[q]))
        except SamplePointError:
            want_n.append(-1)
            continue
        try:
            b = arch.sample_point(tuple(ends[q]))
        except SamplePointError:
            want_n.append(-2)
            continue
        try:
            steps = arch.find_path(a, b, {0: float(cost[q])} if cost[q] > 0 else None)
        except FindPathError:
            want_n.append(-3)
            continue
        want_n.append(len(steps))
        for st in steps:
            assert st.kind == "Waypoint"
            want_steps.append([0.0] + [float(x) for x in np.asarray(st.point, dtype=np.float32)] + [0.0, 0.0, 0.0])
    arch._sync_navdata()
    return arch._flat, starts, ends, cost, np.array(want_n, dtype=np.int32), np.array(want_steps, dtype=np.float32)


def _cpp_queries(harness, ctx, fns, desc, starts, ends, cost):
    n = len(starts)
    out_n = np.zeros(n, dtype=np.int32)
    out_steps = np.zeros((4096, 7), dtype=np.float32)
    err = C.create_string_buffer(512)
    p_ = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    st = harness.hm_queries(ctx, *[C.cast(f, C.c_void_p) for f in fns], C.byref(desc), C.c_uint32(n),
                            p_(np.ascontiguousarray(starts, dtype=np.float32)),
                            p_(np.ascontiguousarray(ends, dtype=np.float32)), p_(cost), p_(out_n), p_(out_steps),
                            C.c_uint32(4096), err, C.c_uint32(512))
    assert st == 0, err.value.decode()
    return out_n, out_steps[:int(out_n[out_n > 0].sum())]


def test_cpp_queries_match_the_python_mirror_on_the_oracle(harness):
    """`Archipelago::sample_point` / `find_path` of the C++ mirror (query.rs:70-94, 185-273) against the Python
    mirror: same OutOfRange cases, same straight paths, bit for bit."""
    from oracle import oracle_py
    from oracle.oracle_py import OracleBackend

    flat, starts, ends, cost, want_n, want_steps = _query_case(OracleBackend())
    desc, keep = _abi.make_navdata_desc(flat)
    lib = oracle_py.load()
    lib.orc_create.restype = C.c_void_p
    ctx = C.c_void_p(lib.orc_create(C.byref(desc), C.c_uint32(0)))
    got_n, got_steps = _cpp_queries(harness, ctx, (lib.orc_navdata_upload, lib.orc_sample_points,
                                                   lib.orc_find_straight_paths), desc, starts, ends, cost)
    assert np.array_equal(got_n, want_n) and (want_n == -1).sum() == 1 and (want_n == -2).sum() == 1
    assert (want_n > 2).sum() > 10
    assert np.array_equal(got_steps.view(np.uint32), want_steps.view(np.uint32))
    del keep


@pytest.mark.gpu
def test_cpp_queries_match_the_python_mirror_on_cuda(harness):
    from landmass_b200._native import NativeBackend

    flat, starts, ends, cost, want_n, want_steps = _query_case(NativeBackend(max_agents=MAX_AGENTS))
    desc, keep = _abi.make_navdata_desc(flat)
    gpu = NativeBackend(max_agents=MAX_AGENTS)
    lib = gpu.lib
    got_n, got_steps = _cpp_queries(harness, gpu.ctx, (lib.lmb_navdata_upload, lib.lmb_sample_points,
                                                       lib.lmb_find_straight_paths), desc, starts, ends, cost)
    assert np.array_equal(got_n, want_n)
    assert np.array_equal(got_steps.view(np.uint32), want_steps.view(np.uint32))
    gpu.close()
    del keep


def test_cpp_debug_read_backs_match_the_python_mirror_on_the_oracle(harness):
    """`agent_path` (lmb_agent_path) and `Agent::avoidance_data` (feature debug-avoidance: lmb_agent_avoidance_constraints,
    kept per agent like agent.avoidance_data) of the C++ mirror after the scripted game loop."""
    from oracle import oracle_py
    from oracle.oracle_py import OracleBackend

    py = python_trace(OracleBackend(), debug=True)
    desc, keep = _abi.make_navdata_desc(py[2])
    lib = oracle_py.load()
    lib.orc_create.restype = C.c_void_p
    ctx = C.c_void_p(lib.orc_create(C.byref(desc), C.c_uint32(0)))
    cpp = cpp_trace(harness, ctx, (lib.orc_navdata_upload, lib.orc_step, lib.orc_pathing_results), desc,
                    debug_fns=(lib.orc_agent_path, lib.orc_agent_avoidance_constraints))
    check(py, cpp)
    assert cpp[3][0] == py[3][0] and py[3][0][0] >= 1 and py[3][0][2] >= 4   # a corridor; >= the room's four walls
    assert np.array_equal(cpp[3][1].view(np.uint32), py[3][1].view(np.uint32))
    del keep


def test_cpp_host_mirror_matches_python_mirror_on_the_oracle(harness):
    from oracle import oracle_py
    from oracle.oracle_py import OracleBackend

    py = python_trace(OracleBackend())
    desc, keep = _abi.make_navdata_desc(py[2])
    lib = oracle_py.load()
    lib.orc_create.restype = C.c_void_p
    ctx = C.c_void_p(lib.orc_create(C.byref(desc), C.c_uint32(0)))
    assert ctx
    cpp = cpp_trace(harness, ctx, (lib.orc_navdata_upload, lib.orc_step, lib.orc_pathing_results), desc)
    check(py, cpp)
    del keep


@pytest.mark.gpu
def test_cpp_host_mirror_matches_python_mirror_on_cuda(harness):
    from landmass_b200._native import NativeBackend

    py = python_trace(NativeBackend(max_agents=MAX_AGENTS))
    desc, keep = _abi.make_navdata_desc(py[2])
    gpu = NativeBackend(max_agents=MAX_AGENTS)   # a second context, driven by the C++ mirror
    lib = gpu.lib
    cpp = cpp_trace(harness, gpu.ctx, (lib.lmb_navdata_upload, lib.lmb_step, lib.lmb_pathing_results), desc)
    check(py, cpp)
    gpu.close()
    del keep


def test_cpp_mirror_links_against_the_cuda_library(tmp_path):
    """`landmass_b200::Archipelago` (= BasicArchipelago<CudaBackend>) compiles and every lmb_* entry point it
    calls resolves against liblandmass_b200.so.  Without a GPU the program can only report the ABI version, and
    `--run` fails at lmb_create: the library has no CPU fallback."""
    import torch

    csrc = os.path.join(ROOT, "landmass_b200", "csrc")
    exe = str(tmp_path / "link_check")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "link_check.cpp"), "-o", exe, "-L" + csrc, "-llandmass_b200",
                    "-Wl,-rpath," + csrc], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert out.strip() == f"abi {_abi.ABI_VERSION}"
    run = subprocess.run([exe, "--run"], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert run.returncode == 0 and "pathing result(s)" in run.stdout, run.stdout + run.stderr
    else:
        assert run.returncode == 1
