"""Funnel / straight-path known answers — transcribed from crates/landmass/src/path_test.rs:59-243
(explicit corridors on rotated islands; oracle only, the C ABI takes no hand-made corridors),
path_test.rs:919-1089 (`find_index_of_node[_rev]` on a corridor across three islands), geometry_test.rs:547-589
(`project_point_to_line_segment`, the funnel's projection onto an animation link's start edge) and
query_test.rs:273-327 (`Archipelago::find_path` across three islands; oracle and CUDA)."""
import math

import numpy as np
import pytest

from landmass_b200 import XY, XYZ, Island, Transform, _abi
from landmass_b200.nav_data import transform_apply
from tests.golden import meshes
from tests.helpers import new_archipelago, valid_mesh

ZIGZAG = dict(
    vertices=[(0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0), (1.0, 2.0), (0.0, 2.0), (1.0, 3.0), (0.0, 3.0),
              (1.0, 4.0), (0.0, 4.0), (1.0, 5.0), (2.0, 4.0), (2.0, 5.0), (3.0, 4.0), (3.0, 5.0), (4.0, 4.0),
              (4.0, 5.0), (5.0, 5.0), (5.0, 6.0), (4.0, 6.0), (5.0, 7.0), (4.0, 7.0), (4.0, 8.0), (-3.0, 8.0),
              (-3.0, 7.0), (-4.0, 8.0), (-3.0, 15.0), (-4.0, 15.0)],
    polygons=[[0, 1, 2, 3], [3, 2, 4, 5], [5, 4, 6, 7], [7, 6, 8, 9], [9, 8, 10], [10, 8, 11, 12],
              [12, 11, 13, 14], [14, 13, 15, 16], [16, 15, 17], [16, 17, 18, 19], [19, 18, 20, 21],
              [21, 20, 22], [21, 22, 23, 24], [24, 23, 25], [25, 23, 26, 27]],
    polygon_type_indices=[0] * 15)


def collect_straight_path(backend, nodes, portals, start, end, limit):
    """path_test.rs:19-57"""
    out = []
    cur_index, cur_point = start
    it = 0
    while cur_index != end[0] and it < limit:
        it += 1
        idx, kind, pt = backend.next_straight_point(nodes, portals, cur_index, cur_point, end[0], end[1])
        assert kind == 0
        out.append((idx, pt[:3].copy()))
        cur_index, cur_point = idx, pt[:3].copy()
    return out


@pytest.fixture
def oracle_backend():
    from oracle.oracle_py import OracleBackend

    b = OracleBackend()
    yield b
    b.close()


def test_finds_next_point_for_organic_map(oracle_backend):  # path_test.rs:59-134
    t = np.array([5.0, 9.0, 7.0], dtype=np.float32)
    rot = np.float32(math.pi) * np.float32(-0.35)
    arch = new_archipelago(oracle_backend)
    i = arch.add_island(Island(Transform(tuple(t), float(rot)), valid_mesh(meshes.ORGANIC)))
    arch._sync_navdata()
    ap = lambda p: transform_apply(t, rot, np.array(p, dtype=np.float32))
    nodes = [arch.node_id(i, 0), arch.node_id(i, 1), arch.node_id(i, 2)]
    portals = [4, 2, _abi.NONE]
    got = collect_straight_path(oracle_backend, nodes, portals, (0, ap((3.0, 1.5, 0.0))),
                                (2, ap((2.5, 4.5, 0.5))), 3)
    exp = [(0, ap((2.0, 3.0, 0.0))), (1, ap((2.0, 4.0, 0.0))), (2, ap((2.5, 4.5, 0.5)))]
    assert [g[0] for g in got] == [e[0] for e in exp]
    for g, e in zip(got, exp):
        np.testing.assert_allclose(g[1], e[1], rtol=0, atol=2e-6)


def test_finds_next_point_in_zig_zag(oracle_backend):  # path_test.rs:136-243
    t = np.array([-1.0, -3.0, 0.0], dtype=np.float32)
    rot = np.float32(math.pi) * np.float32(-1.8)
    arch = new_archipelago(oracle_backend, cs=XY)
    i = arch.add_island(Island(Transform((-1.0, -3.0), float(rot)), valid_mesh(ZIGZAG, cs=XY)))
    arch._sync_navdata()
    ap = lambda p: transform_apply(t, rot, np.array(p, dtype=np.float32))
    nodes = [arch.node_id(i, k) for k in range(15)]
    portals = [2, 2, 2, 2, 1, 2, 2, 2, 2, 2, 2, 2, 2, 1, _abi.NONE]
    got = collect_straight_path(oracle_backend, nodes, portals, (0, ap((0.5, 0.5, 0.0))),
                                (14, ap((-3.5, 14.0, 0.0))), 5)
    exp = [(4, ap((1.0, 4.0, 0.0))), (8, ap((4.0, 5.0, 0.0))), (11, ap((4.0, 7.0, 0.0))),
           (13, ap((-3.0, 8.0, 0.0))), (14, ap((-3.5, 14.0, 0.0)))]
    assert [g[0] for g in got] == [e[0] for e in exp]
    for g, e in zip(got, exp):
        np.testing.assert_allclose(g[1], e[1], rtol=0, atol=4e-6)


def test_starts_at_end_index_goes_to_end_point(oracle_backend):  # path_test.rs:261-310
    arch = new_archipelago(oracle_backend)
    i = arch.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), valid_mesh(dict(
        vertices=[(0.0, 0.0, 0.0), (1.0, 0.0, 0.0), (1.0, 1.0, 0.0), (0.0, 1.0, 0.0), (1.0, 2.0, 0.0), (0.0, 2.0, 0.0)],
        polygons=[[0, 1, 2, 3], [3, 2, 4, 5]], polygon_type_indices=[0, 0]))))
    arch._sync_navdata()
    nodes, portals = [arch.node_id(i, 0), arch.node_id(i, 1)], [2, _abi.NONE]
    idx, kind, pt = oracle_backend.next_straight_point(nodes, portals, 1, (0.25, 1.1, 0.0), 1, (0.75, 1.9, 0.0))
    assert (idx, kind) == (1, 0) and tuple(pt[:3]) == tuple(np.float32([0.75, 1.9, 0.0]))


def test_straight_path_includes_animation_link(oracle_backend):  # path_test.rs:312-427
    """Two mesh chunks joined by an animation link: the funnel reports the link as its own step
    (start point on the start edge straight ahead of the agent, end point on the end edge) and
    resumes behind it."""
    from landmass_b200 import AnimationLink

    arch = new_archipelago(oracle_backend, cs=XY)
    i = arch.add_island(Island(Transform((0.0, 0.0), 0.0), valid_mesh(dict(
        vertices=[(0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0), (1.0, 2.0), (0.0, 2.0), (1.0, 3.0), (0.0, 3.0),
                  (0.0, 5.0), (1.0, 5.0), (1.0, 6.0), (0.0, 6.0), (1.0, 7.0), (0.0, 7.0)],
        polygons=[[0, 1, 2, 3], [3, 2, 4, 5], [5, 4, 6, 7], [8, 9, 10, 11], [11, 10, 12, 13]],
        polygon_type_indices=[0] * 5), cs=XY)))
    link_key = arch.add_animation_link(AnimationLink(start_edge=((0.0, 3.0), (1.0, 3.0)),
                                                     end_edge=((0.0, 5.0), (1.0, 5.0)), cost=1.0, kind=0,
                                                     bidirectional=False))
    arch._sync_navdata()
    f = arch._flat
    links = [l for l in range(f["n_links"]) if f["link_kind"][l] == _abi.LINK_ANIMATION
             and f["link_anim_id"][l] == link_key.idx]
    assert len(links) == 1
    nodes = [arch.node_id(i, k) for k in (0, 1, 2, 3, 4)]
    portals = [2, 2, _abi.PORTAL_LINK | links[0], 2, _abi.NONE]
    # PathIndex {segment 1, corridor 0} = flat index 3, {segment 1, corridor 1} = flat index 4
    idx, kind, pt = oracle_backend.next_straight_point(nodes, portals, 0, (0.1, 0.1, 0.0), 4, (0.9, 6.9, 0.0))
    assert (idx, kind) == (3, 1)
    assert tuple(pt[:3]) == tuple(np.float32([0.1, 3.0, 0.0])) and tuple(pt[3:6]) == tuple(np.float32([0.1, 5.0, 0.0]))
    idx, kind, pt = oracle_backend.next_straight_point(nodes, portals, 3, pt[3:6].copy(), 4, (0.9, 6.9, 0.0))
    assert (idx, kind) == (4, 0) and tuple(pt[:3]) == tuple(np.float32([0.9, 6.9, 0.0]))


def collect_steps(backend, nodes, portals, start, end, limit):
    """path_test.rs:19-57 incl. animation-link steps: [(flat index, kind, start/waypoint, end)]."""
    out = []
    cur_index, cur_point = start
    it, force = 0, False
    while (force or cur_index != end[0]) and it < limit:
        force = False
        it += 1
        idx, kind, pt = backend.next_straight_point(nodes, portals, cur_index, cur_point, end[0], end[1])
        if kind == 1:
            out.append((idx, 1, tuple(float(x) for x in pt[:3]), tuple(float(x) for x in pt[3:6])))
            cur_index, cur_point = idx, pt[3:6].copy()  # after the link we stand at its end point
            force = True
        else:
            out.append((idx, 0, tuple(float(x) for x in pt[:3]), None))
            cur_index, cur_point = idx, pt[:3].copy()
    return out


def link_arch(oracle_backend, vertices, polygons, links):
    """One XY island + animation links [(start edge, end edge)]; returns (arch, node ids, off-mesh link ids)."""
    from landmass_b200 import AnimationLink

    arch = new_archipelago(oracle_backend, cs=XY)
    i = arch.add_island(Island(Transform((0.0, 0.0), 0.0), valid_mesh(dict(
        vertices=vertices, polygons=polygons, polygon_type_indices=[0] * len(polygons)), cs=XY)))
    keys = [arch.add_animation_link(AnimationLink(start_edge=a, end_edge=b, cost=1.0, kind=0, bidirectional=False))
            for a, b in links]
    arch._sync_navdata()
    f = arch._flat
    ids = []
    for k in keys:
        hit = [l for l in range(f["n_links"]) if f["link_kind"][l] == _abi.LINK_ANIMATION
               and f["link_anim_id"][l] == k.idx]
        assert len(hit) == 1
        ids.append(hit[0])
    return arch, [arch.node_id(i, k) for k in range(len(polygons))], ids


def test_multiple_animation_links_in_a_row(oracle_backend):  # path_test.rs:428-601
    arch, n, l = link_arch(
        oracle_backend,
        [(0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0), (0.0, 2.0), (1.0, 2.0), (1.0, 3.0), (0.0, 3.0),
         (0.0, 4.0), (1.0, 4.0), (1.0, 5.0), (0.0, 5.0), (0.0, 6.0), (1.0, 6.0), (1.0, 7.0), (0.0, 7.0),
         (1.0, 8.0), (0.0, 8.0)],
        [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 10, 11], [12, 13, 14, 15], [15, 14, 16, 17]],
        [(((0.0, 1.0), (1.0, 1.0)), ((0.0, 2.0), (1.0, 2.0))), (((0.0, 3.0), (1.0, 3.0)), ((0.0, 4.0), (1.0, 4.0))),
         (((0.0, 5.0), (1.0, 5.0)), ((0.0, 6.0), (1.0, 6.0)))])
    portals = [_abi.PORTAL_LINK | l[0], _abi.PORTAL_LINK | l[1], _abi.PORTAL_LINK | l[2], 2, _abi.NONE]
    got = collect_steps(oracle_backend, n, portals, (0, (0.5, 0.5, 0.0)), (4, (0.5, 7.5, 0.0)), 10)
    assert got == [(1, 1, (0.5, 1.0, 0.0), (0.5, 2.0, 0.0)), (2, 1, (0.5, 3.0, 0.0), (0.5, 4.0, 0.0)),
                   (3, 1, (0.5, 5.0, 0.0), (0.5, 6.0, 0.0)), (4, 0, (0.5, 7.5, 0.0), None)]


def test_obscured_animation_link_is_made_visible_by_straight_path(oracle_backend):  # path_test.rs:602-719
    arch, n, l = link_arch(
        oracle_backend,
        [(0.0, 0.0), (7.0, 0.0), (7.0, 1.0), (6.0, 1.0), (8.0, 3.0), (0.0, 3.0), (0.0, 1.0), (0.0, 5.0),
         (8.0, 5.0), (8.0, 6.0), (0.0, 6.0), (8.0, 7.0), (0.0, 7.0)],
        [[0, 1, 2, 3, 6], [3, 4, 5, 6], [7, 8, 9, 10], [10, 9, 11, 12]],
        [(((0.0, 3.0), (8.0, 3.0)), ((0.0, 5.0), (8.0, 5.0)))])
    portals = [3, _abi.PORTAL_LINK | l[0], 2, _abi.NONE]
    got = collect_steps(oracle_backend, n, portals, (0, (6.5, 0.5, 0.0)), (3, (0.1, 6.9, 0.0)), 10)
    assert got == [(0, 0, (6.0, 1.0, 0.0), None), (2, 1, (6.0, 3.0, 0.0), (6.0, 5.0, 0.0)),
                   (3, 0, (0.10000000149011612, 6.900000095367432, 0.0), None)]


def test_end_point_after_animation_link_is_reported(oracle_backend):  # path_test.rs:720-840
    arch, n, l = link_arch(
        oracle_backend,
        [(0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0), (0.0, 2.0), (1.0, 2.0), (1.0, 3.0), (0.0, 3.0),
         (0.0, 4.0), (1.0, 4.0), (1.0, 5.0), (0.0, 5.0)],
        [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 10, 11]],
        [(((0.0, 1.0), (1.0, 1.0)), ((0.0, 2.0), (1.0, 2.0))), (((0.0, 3.0), (1.0, 3.0)), ((0.0, 4.0), (1.0, 4.0)))])
    portals = [_abi.PORTAL_LINK | l[0], _abi.PORTAL_LINK | l[1], _abi.NONE]
    got = collect_steps(oracle_backend, n, portals, (0, (0.5, 0.5, 0.0)), (1, (0.5, 2.5, 0.0)), 10)
    assert got == [(1, 1, (0.5, 1.0, 0.0), (0.5, 2.0, 0.0)), (1, 0, (0.5, 2.5, 0.0), None)]


def test_query_finds_path_across_islands(make_backend):  # query_test.rs:272-327
    arch = new_archipelago(make_backend(), cs=XY)
    mesh = valid_mesh(dict(vertices=[(0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0)],
                           polygons=[[0, 1, 2, 3]], polygon_type_indices=[0]), cs=XY)
    off = np.array([10.0, 10.0], dtype=np.float32)
    arch.add_island(Island(Transform(tuple(off), 0.0), mesh))
    arch.add_island(Island(Transform(tuple(off + np.float32([1.0, 0.0])), 0.0), mesh))
    arch.add_island(Island(Transform(tuple(off + np.float32([2.0, 0.5])), 0.0), mesh))
    arch.update(1.0)
    start = arch.sample_point(tuple(off + np.float32([0.5, 0.5])), np.float32(1e-5))
    end = arch.sample_point(tuple(off + np.float32([2.5, 1.25])), np.float32(1e-5))
    sn, en = arch.node_id(*start._node_ref), arch.node_id(*end._node_ref)
    sp, ep = XY.to_landmass(start.point()), XY.to_landmass(end.point())
    succ, begin, kind, pts, link = arch.backend.find_straight_paths([sn], [sp], [en], [ep])
    assert succ[0] == 1 and list(kind) == [0, 0, 0]
    got = [tuple(float(x) for x in p[:2]) for p in pts]
    assert got == [(10.5, 10.5), (12.0, 11.0), (12.5, 11.25)]


def _three_island_corridor(oracle_backend):
    """path_test.rs:919-975: corridor [island 1: 3] -link-> [island 2: 2, 1] -link-> [island 3: 0], a fourth
    island that is not on the path.  The reference builds the `Path` by hand; here the islands are real (five
    quads each, far apart) and the two off-mesh links are animation links, so that the flat corridor is one the
    oracle can turn back into segments."""
    from landmass_b200 import AnimationLink

    arch = new_archipelago(oracle_backend, cs=XY)
    strip = dict(vertices=[(float(x), float(y)) for x in range(6) for y in (0, 1)],
                 polygons=[[2 * k, 2 * k + 2, 2 * k + 3, 2 * k + 1] for k in range(5)], polygon_type_indices=[0] * 5)
    islands = [arch.add_island(Island(Transform((0.0, 10.0 * k), 0.0), valid_mesh(strip, cs=XY))) for k in range(4)]
    # island 1 polygon 3 -> island 2 polygon 2, island 2 polygon 1 -> island 3 polygon 0
    keys = [arch.add_animation_link(AnimationLink(start_edge=((3.2, 0.5), (3.8, 0.5)), end_edge=((2.2, 10.5), (2.8, 10.5)),
                                                  cost=1.0, kind=0, bidirectional=False)),
            arch.add_animation_link(AnimationLink(start_edge=((1.2, 10.5), (1.8, 10.5)), end_edge=((0.2, 20.5), (0.8, 20.5)),
                                                  cost=1.0, kind=0, bidirectional=False))]
    arch._sync_navdata()
    f = arch._flat
    link_of = {}
    for l in range(f["n_links"]):
        if f["link_kind"][l] == _abi.LINK_ANIMATION:
            link_of[int(f["link_anim_id"][l])] = l
    node = lambda isl, poly: arch.node_id(islands[isl - 1], poly)  # noqa: E731
    nodes = [node(1, 3), node(2, 2), node(2, 1), node(3, 0)]
    assert f["link_dst_node"][link_of[keys[0].idx]] == node(2, 2) and f["link_dst_node"][link_of[keys[1].idx]] == node(3, 0)
    portals = [_abi.PORTAL_LINK | link_of[keys[0].idx], 0, _abi.PORTAL_LINK | link_of[keys[1].idx], _abi.NONE]
    return nodes, portals, node


@pytest.mark.parametrize("reverse", [False, True])
def test_indices_in_path_are_found(oracle_backend, reverse):  # path_test.rs:919-1003 / 1005-1089
    nodes, portals, node = _three_island_corridor(oracle_backend)
    find = lambda isl, poly: oracle_backend.path_find_index_of_node(nodes, portals, node(isl, poly), reverse)  # noqa: E731
    # PathIndex {segment, portal} as a flat index: {2, 0} = 3, {0, 0} = 0, {1, 1} = 2
    assert find(3, 0) == 3
    assert find(1, 3) == 0
    assert find(2, 1) == 2
    # missing NodeRefs
    assert find(3, 3) is None
    assert find(1, 1) is None
    assert find(4, 4) is None


def test_project_point_to_line_segment():  # geometry_test.rs:547-589
    from oracle.oracle_py import OracleBackend

    proj = OracleBackend.project_point_to_line_segment
    for point, seg, want_p, want_f in [
            ((0.0, 0.0, 0.0), ((1.0, 0.0, 0.0), (0.0, 1.0, 0.0)), (0.5, 0.5, 0.0), 0.5),
            ((2.0, -2.0, 0.0), ((1.0, 0.0, 0.0), (0.0, 1.0, 0.0)), (1.0, 0.0, 0.0), 0.0),
            ((-2.0, 2.0, 0.0), ((1.0, 0.0, 0.0), (0.0, 1.0, 0.0)), (0.0, 1.0, 0.0), 1.0),
            ((-2.0, 2.0, 0.0), ((10.0, 3.0, 0.0), (10.0, 3.0, 0.0)), (10.0, 3.0, 0.0), 0.0)]:
        p, f = proj(point, seg[0], seg[1])
        assert tuple(p) == tuple(np.float32(want_p)) and f == want_f
