"""Animation links -> off-mesh links on the host (landmass_b200/anim_links.py), transcribed from
crates/landmass/src/nav_data_test.rs:1262 onwards.  The reference keeps per-link start/end
`NodePortal`s and patches the off-mesh links incrementally; here every `update` rebuilds them from
scratch, so the tests compare the state AFTER each update: the node portals each link edge samples
(`world_portal_to_node_portals`) and the off-mesh links per node (destination, portals, cost, kind,
owning animation link)."""
import numpy as np
import pytest

from landmass_b200 import XYZ, AnimationLink, Island, NavigationMesh, Transform, _abi
from landmass_b200.anim_links import world_portal_to_node_portals
from landmass_b200.nav_data import NavigationData

f32 = np.float32


def P(*pts):
    return tuple(tuple(float(f32(c)) for c in p) for p in pts)


def rect_mesh(w, h, z):
    return NavigationMesh([(0.0, 0.0, z), (w, 0.0, z), (w, h, z), (0.0, h, z)], [[0, 1, 2, 3]], [0]).validate()


def link_portals(nd, link_id, distance=1.0):
    from landmass_b200.anim_links import effective_link_edges
    start, end = effective_link_edges(nd.get_animation_link(link_id), nd.cs)
    return tuple(sorted(((k, int(p)), (float(iv[0]), float(iv[1])))
                        for k, p, iv in world_portal_to_node_portals(e, nd.islands.items(), nd.cs, distance))
                 for e in (start, end))


def off_mesh_links(nd):
    """{(island key, polygon): [(destination node, destination type, portal, destination portal, cost, kind,
    animation link key index)]} over the animation off-mesh links (node_to_off_mesh_link_ids)."""
    out = {}
    for l in nd.links:
        if l.kind != _abi.LINK_ANIMATION:
            continue
        out.setdefault((l.src_island, l.src_polygon), []).append(
            ((l.dst_island, l.dst_polygon), l.dst_type, P(*l.portal), P(*l.dst_portal), l.cost, l.anim_kind, l.anim_id))
    return {k: sorted(v, key=repr) for k, v in out.items()}


def link(start, end, kind=0, cost=1.0, bidirectional=False):
    return AnimationLink(start_edge=start, end_edge=end, cost=cost, kind=kind, bidirectional=bidirectional)


def update(nd):
    nd.update(1e-5, 1.0)


def test_generates_animation_links():  # nav_data_test.rs:1262-1428
    mesh = rect_mesh(1.0, 1.0, 13.0)
    nd = NavigationData(XYZ)
    i1 = nd.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), mesh))
    i2 = nd.add_island(Island(Transform((2.0, 0.0, 0.0), 0.0), mesh))
    i3 = nd.add_island(Island(Transform((-2.0, 0.0, 0.0), 0.0), mesh))
    i4 = nd.add_island(Island(Transform((0.0, 4.0, 0.0), 0.0), mesh))
    e1 = (((0.1, 0.9, 13.0), (0.9, 0.9, 13.0)), ((0.1, 4.1, 13.0), (0.9, 4.1, 13.0)))
    e2 = (((0.9, 0.1, 13.0), (0.9, 0.9, 13.0)), ((2.1, 0.1, 13.0), (2.1, 0.9, 13.0)))
    e3 = (((-1.1, 0.1, 13.0), (-1.1, 0.9, 13.0)), ((0.1, 0.1, 13.0), (0.1, 0.9, 13.0)))
    l1 = nd.add_animation_link(link(*e1, kind=0))
    l2 = nd.add_animation_link(link(*e2, kind=1))
    l3 = nd.add_animation_link(link(*e3, kind=2))
    update(nd)
    whole = (0.0, 1.0)
    assert link_portals(nd, l1) == ([((i1, 0), whole)], [((i4, 0), whole)])
    assert link_portals(nd, l2) == ([((i1, 0), whole)], [((i2, 0), whole)])
    assert link_portals(nd, l3) == ([((i3, 0), whole)], [((i1, 0), whole)])
    assert off_mesh_links(nd) == {
        (i1, 0): sorted([((i4, 0), 0, P(*e1[0]), P(*e1[1]), 1.0, 0, l1.idx),
                         ((i2, 0), 0, P(*e2[0]), P(*e2[1]), 1.0, 1, l2.idx)], key=repr),
        (i3, 0): [((i1, 0), 0, P(*e3[0]), P(*e3[1]), 1.0, 2, l3.idx)],
    }


E_12 = (((0.9, 0.1, 7.0), (0.9, 0.9, 7.0)), ((2.1, 0.1, 7.0), (2.1, 0.9, 7.0)))
E_21 = (((2.1, 1.1, 7.0), (2.1, 1.9, 7.0)), ((0.9, 1.1, 7.0), (0.9, 1.9, 7.0)))


def test_removing_animation_link_removes_off_mesh_links():  # nav_data_test.rs:1430-1603
    mesh = rect_mesh(1.0, 2.0, 7.0)
    nd = NavigationData(XYZ)
    i1 = nd.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), mesh))
    i2 = nd.add_island(Island(Transform((2.0, 0.0, 0.0), 0.0), mesh))
    update(nd)  # the islands are not brand new any more
    l1 = nd.add_animation_link(link(*E_12, kind=0))
    l2 = nd.add_animation_link(link(*E_21, kind=1))
    update(nd)
    whole = (0.0, 1.0)
    assert link_portals(nd, l1) == ([((i1, 0), whole)], [((i2, 0), whole)])
    assert link_portals(nd, l2) == ([((i2, 0), whole)], [((i1, 0), whole)])
    assert off_mesh_links(nd) == {
        (i1, 0): [((i2, 0), 0, P(*E_12[0]), P(*E_12[1]), 1.0, 0, l1.idx)],
        (i2, 0): [((i1, 0), 0, P(*E_21[0]), P(*E_21[1]), 1.0, 1, l2.idx)],
    }
    nd.remove_animation_link(l1)
    update(nd)
    links = off_mesh_links(nd)
    assert list(links) == [(i2, 0)] and len(links[(i2, 0)]) == 1
    nd.add_animation_link(link(*E_12, kind=0))
    update(nd)
    links = off_mesh_links(nd)
    assert sorted(links, key=repr) == sorted([(i1, 0), (i2, 0)], key=repr)
    assert len(links[(i1, 0)]) == 1 and len(links[(i2, 0)]) == 1


def test_existing_animation_link_is_linked_for_new_island():  # nav_data_test.rs:1605-1769
    mesh = rect_mesh(1.0, 2.0, 7.0)
    nd = NavigationData(XYZ)
    l1 = nd.add_animation_link(link(*E_12, kind=0))
    l2 = nd.add_animation_link(link(*E_21, kind=1))
    update(nd)  # no islands yet: no portals, no off-mesh links
    assert link_portals(nd, l1) == ([], []) and link_portals(nd, l2) == ([], [])
    assert off_mesh_links(nd) == {} and nd.links == []
    i1 = nd.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), mesh))
    update(nd)  # one end each: still no off-mesh links
    whole = (0.0, 1.0)
    assert link_portals(nd, l1) == ([((i1, 0), whole)], [])
    assert link_portals(nd, l2) == ([], [((i1, 0), whole)])
    assert off_mesh_links(nd) == {}
    i2 = nd.add_island(Island(Transform((2.0, 0.0, 0.0), 0.0), mesh))
    update(nd)
    assert link_portals(nd, l1) == ([((i1, 0), whole)], [((i2, 0), whole)])
    assert link_portals(nd, l2) == ([((i2, 0), whole)], [((i1, 0), whole)])
    assert off_mesh_links(nd) == {
        (i1, 0): [((i2, 0), 0, P(*E_12[0]), P(*E_12[1]), 1.0, 0, l1.idx)],
        (i2, 0): [((i1, 0), 0, P(*E_21[0]), P(*E_21[1]), 1.0, 1, l2.idx)],
    }


def test_added_island_mixes_new_and_old_portals():  # nav_data_test.rs:1771-2002
    mesh = rect_mesh(1.0, 1.0, 7.0)
    nd = NavigationData(XYZ)
    i00 = nd.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), mesh))
    i11 = nd.add_island(Island(Transform((2.0, 2.0, 0.0), 0.0), mesh))
    bottom, top = ((0.5, 0.5, 7.0), (2.5, 0.5, 7.0)), ((0.5, 2.5, 7.0), (2.5, 2.5, 7.0))
    l1 = nd.add_animation_link(link(bottom, top, kind=0))
    l2 = nd.add_animation_link(link(top, bottom, kind=1))
    update(nd)
    lo, hi = (0.0, 0.25), (0.75, 1.0)
    assert link_portals(nd, l1) == ([((i00, 0), lo)], [((i11, 0), hi)])
    assert link_portals(nd, l2) == ([((i11, 0), hi)], [((i00, 0), lo)])
    assert off_mesh_links(nd) == {}  # the intervals do not overlap
    i01 = nd.add_island(Island(Transform((0.0, 2.0, 0.0), 0.0), mesh))
    i10 = nd.add_island(Island(Transform((2.0, 0.0, 0.0), 0.0), mesh))
    update(nd)
    by_key = lambda ps: sorted(ps, key=lambda p: p[1])
    assert tuple(map(by_key, link_portals(nd, l1))) == ([((i00, 0), lo), ((i10, 0), hi)], [((i01, 0), lo), ((i11, 0), hi)])
    assert tuple(map(by_key, link_portals(nd, l2))) == ([((i01, 0), lo), ((i11, 0), hi)], [((i00, 0), lo), ((i10, 0), hi)])
    assert off_mesh_links(nd) == {
        (i00, 0): [((i01, 0), 0, P((0.5, 0.5, 7.0), (1.0, 0.5, 7.0)), P((0.5, 2.5, 7.0), (1.0, 2.5, 7.0)), 1.0, 0, l1.idx)],
        (i10, 0): [((i11, 0), 0, P((2.0, 0.5, 7.0), (2.5, 0.5, 7.0)), P((2.0, 2.5, 7.0), (2.5, 2.5, 7.0)), 1.0, 0, l1.idx)],
        (i01, 0): [((i00, 0), 0, P((0.5, 2.5, 7.0), (1.0, 2.5, 7.0)), P((0.5, 0.5, 7.0), (1.0, 0.5, 7.0)), 1.0, 1, l2.idx)],
        (i11, 0): [((i10, 0), 0, P((2.0, 2.5, 7.0), (2.5, 2.5, 7.0)), P((2.0, 0.5, 7.0), (2.5, 0.5, 7.0)), 1.0, 1, l2.idx)],
    }


def test_same_island_animation_link():  # nav_data_test.rs:2004-2076
    from landmass_b200 import XY
    mesh = NavigationMesh([(0.0, 0.0), (1.0, 0.0), (1.0, 2.0), (0.0, 2.0)], [[0, 1, 2, 3]], [0], cs=XY).validate()
    nd = NavigationData(XY)
    i = nd.add_island(Island(Transform((0.0, 0.0), 0.0), mesh))
    l = nd.add_animation_link(link(((0.1, 0.5), (0.9, 0.5)), ((0.1, 1.5), (0.9, 1.5))))
    update(nd)
    assert link_portals(nd, l) == ([((i, 0), (0.0, 1.0))], [((i, 0), (0.0, 1.0))])
    assert off_mesh_links(nd) == {
        (i, 0): [((i, 0), 0, P((0.1, 0.5, 0.0), (0.9, 0.5, 0.0)), P((0.1, 1.5, 0.0), (0.9, 1.5, 0.0)), 1.0, 0, l.idx)]}


TWO_SQUARES = dict(vertices=[(0.0, 0.0, 7.0), (1.0, 0.0, 7.0), (1.0, 1.0, 7.0), (0.0, 1.0, 7.0),
                             (0.0, 3.0, 7.0), (1.0, 3.0, 7.0), (1.0, 4.0, 7.0), (0.0, 4.0, 7.0)],
                   polygons=[[0, 1, 2, 3], [4, 5, 6, 7]], polygon_type_indices=[0, 0])


@pytest.mark.parametrize("links_first", [False, True])
def test_point_animation_links_are_connected(links_first):  # nav_data_test.rs:2078-2203, 2205-2336
    """A point end edge receives the whole start edge; a point start edge collapses the end edge to its
    midpoint.  Second variant: the links exist (and have been updated once) before the island does."""
    mesh = NavigationMesh(**TWO_SQUARES).validate()
    nd = NavigationData(XYZ)
    if not links_first:
        i = nd.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), mesh))
    edge, point = ((0.0, 0.5, 7.0), (1.0, 0.5, 7.0)), ((0.5, 3.5, 7.0), (0.5, 3.5, 7.0))
    l1 = nd.add_animation_link(link(edge, point, kind=0))
    l2 = nd.add_animation_link(link(point, edge, kind=1))
    if links_first:
        update(nd)
        i = nd.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), mesh))
    update(nd)
    whole = (0.0, 1.0)
    assert link_portals(nd, l1) == ([((i, 0), whole)], [((i, 1), whole)])
    assert link_portals(nd, l2) == ([((i, 1), whole)], [((i, 0), whole)])
    mid = ((0.5, 0.5, 7.0), (0.5, 0.5, 7.0))
    assert off_mesh_links(nd) == {
        (i, 0): [((i, 1), 0, P(*edge), P(*point), 1.0, 0, l1.idx)],
        (i, 1): [((i, 0), 0, P(*point), P(*mid), 1.0, 1, l2.idx)],
    }
    assert len(nd.links) == 2


def test_changing_island_does_not_cause_duplicate_off_mesh_links():  # nav_data_test.rs:2338-2543
    mesh = rect_mesh(1.0, 1.0, 7.0)
    nd = NavigationData(XYZ)
    i00 = nd.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), mesh))
    i10 = nd.add_island(Island(Transform((3.5, 0.0, 0.0), 0.0), mesh))
    i11 = nd.add_island(Island(Transform((3.0, 3.0, 0.0), 0.0), mesh))
    i01 = nd.add_island(Island(Transform((-0.5, 3.0, 0.0), 0.0), mesh))
    l = nd.add_animation_link(link(((0.0, 0.5, 7.0), (4.0, 0.5, 7.0)), ((0.0, 3.5, 7.0), (4.0, 3.5, 7.0))))
    update(nd)
    by_iv = lambda ps: sorted(ps, key=lambda p: p[1])
    assert tuple(map(by_iv, link_portals(nd, l))) == (
        [((i00, 0), (0.0, 0.25)), ((i10, 0), (0.875, 1.0))], [((i01, 0), (0.0, 0.125)), ((i11, 0), (0.75, 1.0))])
    assert off_mesh_links(nd) == {
        (i00, 0): [((i01, 0), 0, P((0.0, 0.5, 7.0), (0.5, 0.5, 7.0)), P((0.0, 3.5, 7.0), (0.5, 3.5, 7.0)), 1.0, 0, l.idx)],
        (i10, 0): [((i11, 0), 0, P((3.5, 0.5, 7.0), (4.0, 0.5, 7.0)), P((3.5, 3.5, 7.0), (4.0, 3.5, 7.0)), 1.0, 0, l.idx)],
    }
    assert len(nd.links) == 2
    nd.get_island(i01).set_transform(Transform((0.0, 3.0, 0.0), 0.0))
    nd.get_island(i10).set_transform(Transform((3.0, 0.0, 0.0), 0.0))
    update(nd)
    assert tuple(map(by_iv, link_portals(nd, l))) == (
        [((i00, 0), (0.0, 0.25)), ((i10, 0), (0.75, 1.0))], [((i01, 0), (0.0, 0.25)), ((i11, 0), (0.75, 1.0))])
    assert off_mesh_links(nd) == {
        (i00, 0): [((i01, 0), 0, P((0.0, 0.5, 7.0), (1.0, 0.5, 7.0)), P((0.0, 3.5, 7.0), (1.0, 3.5, 7.0)), 1.0, 0, l.idx)],
        (i10, 0): [((i11, 0), 0, P((3.0, 0.5, 7.0), (4.0, 0.5, 7.0)), P((3.0, 3.5, 7.0), (4.0, 3.5, 7.0)), 1.0, 0, l.idx)],
    }
    assert len(nd.links) == 2


def test_changing_island_does_not_cause_duplicate_off_mesh_links_for_point_links():  # nav_data_test.rs:2545-2765
    mesh = rect_mesh(1.0, 1.0, 7.0)
    nd = NavigationData(XYZ)
    i1 = nd.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), mesh))
    i2 = nd.add_island(Island(Transform((0.0, 2.0, 0.0), 0.0), mesh))
    p1, p2 = (0.25, 0.5, 7.0), (0.75, 0.5, 7.0)
    l1 = nd.add_animation_link(link((p1, p1), ((0.0, 2.5, 7.0), (0.5, 2.5, 7.0)), kind=0))
    l2 = nd.add_animation_link(link(((0.5, 2.5, 7.0), (1.0, 2.5, 7.0)), (p2, p2), kind=1))
    want = {
        (i1, 0): [((i2, 0), 0, P(p1, p1), P((0.25, 2.5, 7.0), (0.25, 2.5, 7.0)), 1.0, 0, l1.idx)],
        (i2, 0): [((i1, 0), 0, P((0.5, 2.5, 7.0), (1.0, 2.5, 7.0)), P(p2, p2), 1.0, 1, l2.idx)],
    }
    whole = (0.0, 1.0)
    for moved in (False, True):
        if moved:
            nd.get_island(i2).set_transform(Transform((0.1, 2.0, 0.0), 0.0))
        update(nd)
        assert link_portals(nd, l1) == ([((i1, 0), whole)], [((i2, 0), whole)])
        assert link_portals(nd, l2) == ([((i2, 0), whole)], [((i1, 0), whole)])
        assert off_mesh_links(nd) == want and len(nd.links) == 2


TWO_FLOORS = [(0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0)]


@pytest.mark.parametrize("use_height_mesh", [False, True])
def test_generates_animation_links_at_correct_height(use_height_mesh):  # nav_data_test.rs:2767-2864, 2866-2991
    """Two floors at z = 5 and z = 10.  Link edges 0.9 above / below a floor are sampled, 1.1 are not.
    Second variant: the polygons lie at z = -100 / +100 and only the height mesh has the real floors."""
    from landmass_b200 import HeightNavigationMesh, HeightPolygon
    floor = lambda z: [(x, y, z) for x, y in TWO_FLOORS]
    hm = None
    zs = (5.0, 10.0)
    if use_height_mesh:
        hm = HeightNavigationMesh(vertices=floor(5.0) + floor(10.0),
                                  triangles=[[0, 1, 2], [2, 3, 0], [0, 1, 2], [2, 3, 0]],
                                  polygons=[HeightPolygon(0, 4, 0, 2), HeightPolygon(4, 4, 2, 2)])
        zs = (-100.0, 100.0)
    mesh = NavigationMesh(floor(zs[0]) + floor(zs[1]), [[0, 1, 2, 3], [4, 5, 6, 7]], [0, 0], height_mesh=hm).validate()
    nd = NavigationData(XYZ)
    i1 = nd.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), mesh))
    i2 = nd.add_island(Island(Transform((2.0, 0.0, 0.0), 0.0), mesh))
    inside = (((0.9, 0.1, 4.1), (0.9, 0.9, 4.1)), ((2.1, 0.1, 10.9), (2.1, 0.9, 10.9)))
    outside = (((0.9, 0.1, 3.9), (0.9, 0.9, 3.9)), ((2.1, 0.1, 11.1), (2.1, 0.9, 11.1)))
    l1 = nd.add_animation_link(link(*inside, kind=0))
    l2 = nd.add_animation_link(link(*outside, kind=0 if use_height_mesh else 1))
    update(nd)
    assert link_portals(nd, l1) == ([((i1, 0), (0.0, 1.0))], [((i2, 1), (0.0, 1.0))])
    assert link_portals(nd, l2) == ([], [])
    assert off_mesh_links(nd) == {(i1, 0): [((i2, 1), 0, P(*inside[0]), P(*inside[1]), 1.0, 0, l1.idx)]}
    assert len(nd.links) == 1


def test_animation_link_along_angled_surface():  # nav_data_test.rs:2993-3186
    """A ramp between two flat squares; the link edge runs along it at half height.  With a vertical
    distance of 0.4 only the middle of the ramp is close enough; with 0.6 all three polygons are."""
    mesh = NavigationMesh(
        [(0.0, 0.0, 0.0), (1.0, 0.0, 0.0), (1.0, 1.0, 0.0), (0.0, 1.0, 0.0), (1.0, 2.0, 1.0), (0.0, 2.0, 1.0),
         (1.0, 3.0, 1.0), (0.0, 3.0, 1.0)], [[0, 1, 2, 3], [3, 2, 4, 5], [5, 4, 6, 7]], [0] * 3).validate()
    nd = NavigationData(XYZ)
    i1 = nd.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), mesh))
    i2 = nd.add_island(Island(Transform((2.0, 0.0, 0.0), 0.0), mesh))
    edges = (((0.5, 0.5, 0.5), (0.5, 2.5, 0.5)), ((2.5, 0.5, 0.5), (2.5, 2.5, 0.5)))
    l = nd.add_animation_link(link(*edges))
    nd.update(0.01, 0.4)
    iv = (float(f32(0.3)), float(f32(0.7)))
    assert link_portals(nd, l, 0.4) == ([((i1, 1), iv)], [((i2, 1), iv)])
    assert off_mesh_links(nd) == {
        (i1, 1): [((i2, 1), 0, P((0.5, 1.1, 0.5), (0.5, 1.9, 0.5)), P((2.5, 1.1, 0.5), (2.5, 1.9, 0.5)), 1.0, 0, l.idx)]}
    nd.remove_animation_link(l)
    l = nd.add_animation_link(link(*edges))
    nd.update(0.01, 0.6)
    ivs = [(0.0, 0.25), (0.25, 0.75), (0.75, 1.0)]
    assert link_portals(nd, l, 0.6) == ([((i1, p), ivs[p]) for p in range(3)], [((i2, p), ivs[p]) for p in range(3)])
    ys = [(0.5, 1.0), (1.0, 2.0), (2.0, 2.5)]
    assert off_mesh_links(nd) == {
        (i1, p): [((i2, p), 0, P((0.5, ys[p][0], 0.5), (0.5, ys[p][1], 0.5)),
                   P((2.5, ys[p][0], 0.5), (2.5, ys[p][1], 0.5)), 1.0, 0, l.idx)] for p in range(3)}


def test_generates_bidirectional_animation_links():  # nav_data_test.rs:3188-3409
    mesh = rect_mesh(1.0, 1.0, 13.0)
    nd = NavigationData(XYZ)
    i1 = nd.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), mesh))
    i2 = nd.add_island(Island(Transform((2.0, 0.0, 0.0), 0.0), mesh))
    i3 = nd.add_island(Island(Transform((-2.0, 0.0, 0.0), 0.0), mesh))
    i4 = nd.add_island(Island(Transform((0.0, 4.0, 0.0), 0.0), mesh))
    e1 = (((0.1, 0.9, 13.0), (0.9, 0.9, 13.0)), ((0.1, 4.1, 13.0), (0.9, 4.1, 13.0)))
    e2 = (((0.9, 0.1, 13.0), (0.9, 0.9, 13.0)), ((2.1, 0.1, 13.0), (2.1, 0.9, 13.0)))
    e3 = (((-1.1, 0.1, 13.0), (-1.1, 0.9, 13.0)), ((0.1, 0.1, 13.0), (0.1, 0.9, 13.0)))
    l1 = nd.add_animation_link(link(*e1, kind=0, bidirectional=True))
    l2 = nd.add_animation_link(link(*e2, kind=1, bidirectional=True))
    l3 = nd.add_animation_link(link(*e3, kind=2, bidirectional=True))
    update(nd)
    whole = (0.0, 1.0)
    assert link_portals(nd, l1) == ([((i1, 0), whole)], [((i4, 0), whole)])
    assert link_portals(nd, l2) == ([((i1, 0), whole)], [((i2, 0), whole)])
    assert link_portals(nd, l3) == ([((i3, 0), whole)], [((i1, 0), whole)])
    assert off_mesh_links(nd) == {
        (i1, 0): sorted([((i4, 0), 0, P(*e1[0]), P(*e1[1]), 1.0, 0, l1.idx),
                         ((i2, 0), 0, P(*e2[0]), P(*e2[1]), 1.0, 1, l2.idx),
                         ((i3, 0), 0, P(*e3[1]), P(*e3[0]), 1.0, 2, l3.idx)], key=repr),
        (i2, 0): [((i1, 0), 0, P(*e2[1]), P(*e2[0]), 1.0, 1, l2.idx)],
        (i3, 0): [((i1, 0), 0, P(*e3[0]), P(*e3[1]), 1.0, 2, l3.idx)],
        (i4, 0): [((i1, 0), 0, P(*e1[1]), P(*e1[0]), 1.0, 0, l1.idx)],
    }


def test_bidirectional_links_collapse_either_end():  # nav_data_test.rs:3411-3644
    mesh = NavigationMesh(**TWO_SQUARES).validate()
    nd = NavigationData(XYZ)
    i = nd.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), mesh))
    a, b = (0.25, 3.5, 7.0), (0.75, 3.5, 7.0)
    l1 = nd.add_animation_link(link(((0.0, 0.5, 7.0), (0.5, 0.5, 7.0)), (a, a), kind=0, bidirectional=True))
    l2 = nd.add_animation_link(link((b, b), ((0.5, 0.5, 7.0), (1.0, 0.5, 7.0)), kind=1, bidirectional=True))
    a0, b0 = (0.25, 0.5, 7.0), (0.75, 0.5, 7.0)
    whole = (0.0, 1.0)
    want = {
        (i, 0): sorted([((i, 1), 0, P(a0, a0), P(a, a), 1.0, 0, l1.idx), ((i, 1), 0, P(b0, b0), P(b, b), 1.0, 1, l2.idx)],
                       key=repr),
        (i, 1): sorted([((i, 0), 0, P(a, a), P(a0, a0), 1.0, 0, l1.idx), ((i, 0), 0, P(b, b), P(b0, b0), 1.0, 1, l2.idx)],
                       key=repr),
    }
    for step in ("first", "moved away", "moved back"):
        if step == "moved away":
            nd.get_island(i).set_transform(Transform((100.0, 100.0, 100.0), 0.0))
        elif step == "moved back":
            nd.get_island(i).set_transform(Transform((0.0, 0.0, 0.0), 0.0))
        update(nd)
        if step == "moved away":
            assert link_portals(nd, l1) == ([], []) and link_portals(nd, l2) == ([], [])
            assert off_mesh_links(nd) == {} and nd.links == []
        else:
            assert link_portals(nd, l1) == ([((i, 0), whole)], [((i, 1), whole)])
            assert link_portals(nd, l2) == ([((i, 1), whole)], [((i, 0), whole)])
            assert off_mesh_links(nd) == want and len(nd.links) == 4


def test_bidirectional_links_use_correct_indices():  # nav_data_test.rs:3646-3748
    from landmass_b200 import XY
    mesh = NavigationMesh([(0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0), (2.0, 0.0), (2.0, 1.0)],
                          [[0, 1, 2, 3], [2, 1, 4, 5]], [0, 1], cs=XY).validate()
    nd = NavigationData(XY)
    i1 = nd.add_island(Island(Transform((0.0, 0.0), 0.0), mesh))
    i2 = nd.add_island(Island(Transform((-3.0, 0.0), 0.0), mesh))
    l = nd.add_animation_link(link(((0.1, 0.1), (0.1, 0.9)), ((-1.1, 0.1), (-1.1, 0.9)), bidirectional=True))
    update(nd)
    assert link_portals(nd, l) == ([((i1, 0), (0.0, 1.0))], [((i2, 1), (0.0, 1.0))])
    s, e = ((0.1, 0.1, 0.0), (0.1, 0.9, 0.0)), ((-1.1, 0.1, 0.0), (-1.1, 0.9, 0.0))
    assert off_mesh_links(nd) == {
        (i1, 0): [((i2, 1), 1, P(*s), P(*e), 1.0, 0, l.idx)],   # destination type index of (island_2, node 1)
        (i2, 1): [((i1, 0), 0, P(*e), P(*s), 1.0, 0, l.idx)],
    }
