"""Island linking known answers (nav_data.rs:827-1011, host side: landmass_b200/linking.py) —
transcribed from crates/landmass/src/nav_data_test.rs:178-505
(`link_edges_between_islands_links_touching_islands`): two overlapping islands under the same
rotated transform; every boundary link's source node, destination node, destination type and
portal (rounded to 1e-5 like the reference), both for edge_link_distance 1e-5 and 1e-3."""
import math

import numpy as np
import pytest

from landmass_b200 import Island, Transform
from landmass_b200.nav_data import NavigationData, transform_apply
from tests.helpers import valid_mesh

MESH_1 = dict(
    vertices=[(1.0, 0.0, 1.0), (1.0, 1.0, 1.0), (-1.0, 1.0, 1.0), (-1.0, 0.0, 1.0), (-1.0, -1.0, 1.0), (1.0, -1.0, 1.0),
              (2.0, 0.0, 1.0), (2.0, 2.0, 1.0), (-2.0, 2.0, 1.0), (-2.0, 0.0, 1.0), (-2.0, -2.0, 1.0), (2.0, -2.0, 1.0)],
    polygons=[[0, 6, 7, 1], [1, 7, 8, 2], [2, 8, 9, 3], [10, 4, 3, 9], [10, 11, 5, 4], [6, 0, 5, 11]],
    polygon_type_indices=[0, 1, 1, 1, 0, 0])
MESH_2 = dict(
    vertices=[(-1.0, -0.5, 1.0), (-0.5, -0.5, 1.0), (0.5, -0.5, 1.0), (1.0, -0.5, 1.0), (-1.0, 0.5, 1.0), (-0.5, 0.5, 1.0),
              (0.5, 0.5, 1.0), (1.0, 0.5, 1.0), (-0.5, 1.0, 1.0), (0.5, 1.0, 1.0), (-0.5, -1.0, 1.0), (0.5, -1.0, 1.0)],
    polygons=[[5, 4, 0, 1], [1, 2, 6, 5], [3, 7, 6, 2], [5, 6, 9, 8], [10, 11, 2, 1]],
    polygon_type_indices=[0, 0, 0, 0, 2])

T = np.array([1.0, 2.0, 3.0], dtype=np.float32)
ROT = math.pi * -0.25

# (source island, polygon) -> [(destination island, polygon, destination type, portal a, portal b)]
EXPECTED = {
    (1, 0): [(2, 2, 0, (1.0, 0.0, 1.0), (1.0, 0.5, 1.0))],
    (1, 1): [(2, 3, 0, (0.5, 1.0, 1.0), (-0.5, 1.0, 1.0))],
    (1, 2): [(2, 0, 0, (-1.0, 0.5, 1.0), (-1.0, 0.0, 1.0))],
    (1, 3): [(2, 0, 0, (-1.0, 0.0, 1.0), (-1.0, -0.5, 1.0))],
    (1, 4): [(2, 4, 2, (-0.5, -1.0, 1.0), (0.5, -1.0, 1.0))],
    (1, 5): [(2, 2, 0, (1.0, -0.5, 1.0), (1.0, 0.0, 1.0))],
    # reverse links
    (2, 0): [(1, 2, 1, (-1.0, 0.0, 1.0), (-1.0, 0.5, 1.0)), (1, 3, 1, (-1.0, -0.5, 1.0), (-1.0, 0.0, 1.0))],
    (2, 2): [(1, 0, 0, (1.0, 0.5, 1.0), (1.0, 0.0, 1.0)), (1, 5, 0, (1.0, 0.0, 1.0), (1.0, -0.5, 1.0))],
    (2, 3): [(1, 1, 1, (-0.5, 1.0, 1.0), (0.5, 1.0, 1.0))],
    (2, 4): [(1, 4, 0, (0.5, -1.0, 1.0), (-0.5, -1.0, 1.0))],
}


def rounded(p):
    w = transform_apply(T, np.float32(ROT), np.array(p, dtype=np.float32)).astype(np.float64)
    return tuple(np.round(w * 1e5) * 1e-5)


@pytest.mark.parametrize("edge_link_distance,order", [(1e-5, (1, 2)), (1e-3, (2, 1))])
def test_link_edges_between_islands_links_touching_islands(edge_link_distance, order):
    meshes = {1: valid_mesh(MESH_1), 2: valid_mesh(MESH_2)}
    nd = NavigationData()
    keys = {}
    for which in order:  # the reference links (1, 2) and then, separately, (2, 1): same links
        keys[which] = nd.add_island(Island(Transform(tuple(T), ROT), meshes[which]))
    nd.update(edge_link_distance, 0.25)
    f = nd.flatten()
    island_of = {f["_island_key_to_index"][k]: which for which, k in keys.items()}
    begin = f["island_poly_begin"]

    def node_ref(n):
        i = int(np.searchsorted(begin, n, side="right") - 1)
        return island_of[i], int(n - begin[i])

    got = {}
    for l in range(f["n_links"]):
        src = node_ref(int(f["link_src_node"][l]))
        dst = node_ref(int(f["link_dst_node"][l]))
        portal = f["link_portal"][l].astype(np.float64)
        a = tuple(np.round(portal[:3] * 1e5) * 1e-5)
        b = tuple(np.round(portal[3:] * 1e5) * 1e-5)
        got.setdefault(src, []).append((dst[0], dst[1], int(f["link_dst_type"][l]), a, b))
    for v in got.values():
        v.sort(key=lambda e: e[0] * 100 + e[1])
    expected = {k: [(d, p, t, rounded(a), rounded(b)) for d, p, t, a, b in v] for k, v in EXPECTED.items()}
    assert set(got) == set(expected)
    for k in expected:
        assert len(got[k]) == len(expected[k]), k
        for g, e in zip(got[k], expected[k]):
            assert g[:3] == e[:3], (k, g, e)
            np.testing.assert_allclose(g[3], e[3], atol=1.5e-5, rtol=0, err_msg=str(k))
            np.testing.assert_allclose(g[4], e[4], atol=1.5e-5, rtol=0, err_msg=str(k))
    # every node that got a link has its boundary re-cut (modified_node_refs_to_update, nav_data_test.rs:456-476)
    modified = {node_ref(int(n)) for n in f["modified_node"]} if f.get("n_modified", 0) else set()
    assert modified == set(expected)


def test_modifies_node_boundaries_for_linked_islands():  # nav_data_test.rs:805-876
    """Three islands touching along parts of their borders: the nodes with boundary links get a
    `ModifiedNode` — the boundary edges that remain once the linked stretch is cut out, with the new
    vertices the cut introduces (vertex index >= the mesh's vertex count = index into new_vertices)."""
    mesh = valid_mesh(dict(
        vertices=[(1.0, 1.0, 1.0), (2.0, 1.0, 1.0), (2.0, 2.0, 1.0), (1.0, 2.0, 1.0), (2.0, 3.0, 1.0), (1.0, 3.0, 1.0)],
        polygons=[[0, 1, 2, 3], [3, 2, 4, 5]], polygon_type_indices=[0, 0]))
    nd = NavigationData()
    k1 = nd.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), mesh))
    k2 = nd.add_island(Island(Transform((1.0, -1.0, 0.0), 0.0), mesh))
    k3 = nd.add_island(Island(Transform((2.0, 3.5, 0.0), math.pi * -0.5), mesh))
    nd.update(1e-6, 1e-6)
    f = nd.flatten()
    idx = {f["_island_key_to_index"][k]: n for n, k in ((1, k1), (2, k2), (3, k3))}
    begin = f["island_poly_begin"]
    n_mesh_vertices = 6
    got = {}
    for m in range(f["n_modified"]):
        n = int(f["modified_node"][m])
        i = int(np.searchsorted(begin, n, side="right") - 1)
        verts = [tuple(np.round(v.astype(np.float64) / 1e-4) * 1e-4)
                 for v in f["modified_verts"][f["modified_vert_begin"][m]:f["modified_vert_begin"][m + 1]]]
        order = sorted(range(len(verts)), key=lambda a: verts[a])          # clone_sort_round_modified_nodes
        rank = {old: new for new, old in enumerate(order)}
        edges = []
        for l, r in f["modified_edges"][f["modified_edge_begin"][m]:f["modified_edge_begin"][m + 1]]:
            l, r = int(l), int(r)
            if l >= n_mesh_vertices:
                l = rank[l - n_mesh_vertices] + n_mesh_vertices
            if r >= n_mesh_vertices:
                r = rank[r - n_mesh_vertices] + n_mesh_vertices
            edges.append((l, r))
        got[(idx[i], int(n - begin[i]))] = (sorted(edges), [verts[a] for a in order])
    assert got == {
        (1, 0): ([(0, 1), (3, 0)], []),
        (2, 1): ([(2, 6), (4, 5)], [(3.0, 1.5)]),
        (3, 0): ([(0, 6), (1, 2), (3, 0)], [(3.0, 2.0)]),
    }


def boundary_links(nd, names, digits=6):
    """{(island name, polygon): [(destination name, polygon, destination type, portal a, portal b)]} over the
    boundary links, portals rounded like clone_sort_round_links (nav_data_test.rs:134-176)."""
    from landmass_b200 import _abi
    rnd = lambda p: tuple(float(x) + 0.0 for x in np.round(np.asarray(p, dtype=np.float64), digits))
    out = {}
    for l in nd.links:
        if l.kind != _abi.LINK_BOUNDARY:
            continue
        out.setdefault((names[l.src_island], l.src_polygon), []).append(
            (names[l.dst_island], l.dst_polygon, l.dst_type, rnd(l.portal[0]), rnd(l.portal[1])))
    return {k: sorted(v) for k, v in out.items()}


def test_update_links_islands_and_unlinks_on_delete():  # nav_data_test.rs:506-733
    mesh = valid_mesh(dict(
        vertices=[(1.0, 0.0, 1.0), (2.0, 0.0, 1.0), (2.0, 2.0, 1.0), (0.0, 2.0, 1.0), (0.0, 1.0, 1.0), (1.0, 1.0, 1.0)],
        polygons=[[0, 1, 2, 5], [4, 5, 2, 3]], polygon_type_indices=[0, 0]))
    nd = NavigationData()
    pi = math.pi
    keys = [nd.add_island(Island(Transform(t, r), mesh)) for t, r in [
        ((0.0, 0.0, 0.0), 0.0), ((0.0, 0.0, 0.0), pi * 0.5), ((0.0, 0.0, 0.0), pi), ((3.0, 0.0, 0.0), pi),
        ((2.0, 3.0, 0.0), pi * -0.5)]]
    names = {k: n + 1 for n, k in enumerate(keys)}
    nd.update(0.01, 0.01)
    assert boundary_links(nd, names) == {
        (1, 0): [(4, 0, 0, (2.0, 0.0, 1.0), (1.0, 0.0, 1.0)), (5, 0, 0, (2.0, 2.0, 1.0), (2.0, 1.0, 1.0))],
        (1, 1): [(2, 0, 0, (0.0, 1.0, 1.0), (0.0, 2.0, 1.0))],
        (2, 0): [(1, 1, 0, (0.0, 2.0, 1.0), (0.0, 1.0, 1.0))],
        (2, 1): [(3, 0, 0, (-1.0, 0.0, 1.0), (-2.0, 0.0, 1.0))],
        (3, 0): [(2, 1, 0, (-2.0, 0.0, 1.0), (-1.0, 0.0, 1.0))],
        (4, 0): [(1, 0, 0, (1.0, 0.0, 1.0), (2.0, 0.0, 1.0))],
        (5, 0): [(1, 0, 0, (2.0, 1.0, 1.0), (2.0, 2.0, 1.0))],
    }
    nd.remove_island(keys[1])
    nd.remove_island(keys[3])
    nd.get_island(keys[4]).set_transform(Transform((0.0, 0.0, 0.0), pi * 0.5))  # island 5 takes island 2's place
    nd.update(0.01, 0.01)
    assert boundary_links(nd, names) == {
        (1, 1): [(5, 0, 0, (0.0, 1.0, 1.0), (0.0, 2.0, 1.0))],
        (3, 0): [(5, 1, 0, (-2.0, 0.0, 1.0), (-1.0, 0.0, 1.0))],
        (5, 0): [(1, 1, 0, (0.0, 2.0, 1.0), (0.0, 1.0, 1.0))],
        (5, 1): [(3, 0, 0, (-1.0, 0.0, 1.0), (-2.0, 0.0, 1.0))],
    }


SQUARE_2D = dict(vertices=[(0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0)], polygons=[[0, 1, 2, 3]],
                 polygon_type_indices=[0])


def test_stale_modified_nodes_are_removed():  # nav_data_test.rs:877-922
    mesh = valid_mesh(dict(
        vertices=[(1.0, 1.0, 1.0), (2.0, 1.0, 1.0), (2.0, 2.0, 1.0), (1.0, 2.0, 1.0), (2.0, 3.0, 1.0), (1.0, 3.0, 1.0)],
        polygons=[[0, 1, 2, 3], [3, 2, 4, 5]], polygon_type_indices=[0, 0]))
    nd = NavigationData()
    nd.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), mesh))
    k2 = nd.add_island(Island(Transform((1.0, -1.0, 0.0), 0.0), mesh))
    nd.update(1e-6, 1e-6)
    assert len(nd.modified_nodes) == 2
    nd.remove_island(k2)
    nd.update(1e-6, 1e-6)
    assert len(nd.modified_nodes) == 0


def test_modified_node_is_removed_for_no_boundary_links():  # nav_data_test.rs:924-987
    from landmass_b200 import XY, AnimationLink
    mesh = valid_mesh(SQUARE_2D, cs=XY)
    nd = NavigationData(XY)
    k1 = nd.add_island(Island(Transform((0.0, 0.0), 0.0), mesh))
    k2 = nd.add_island(Island(Transform((1.0, 0.0), 0.0), mesh))
    nd.add_island(Island(Transform((-2.0, 0.0), 0.0), mesh))
    nd.add_animation_link(AnimationLink(start_edge=((0.1, 0.1), (0.1, 0.9)), end_edge=((-1.1, 0.1), (-1.1, 0.9)),
                                        cost=1.0, kind=0, bidirectional=False))
    nd.update(1e-5, 1e-5)
    assert any(m.island == k1 and m.polygon == 0 for m in nd.modified_nodes)
    nd.remove_island(k2)  # the animation link stays, the boundary links go
    nd.update(1e-5, 1e-5)
    assert not any(m.island == k1 and m.polygon == 0 for m in nd.modified_nodes)


def test_empty_navigation_mesh_is_safe():  # nav_data_test.rs:989-1027
    nd = NavigationData()
    nd.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), valid_mesh(dict(
        vertices=[(0.0, 0.0, 0.0), (1.0, 0.0, 0.0), (1.0, 1.0, 0.0), (0.0, 1.0, 0.0)], polygons=[[0, 1, 2, 3]],
        polygon_type_indices=[0]))))
    nd.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), valid_mesh(dict(vertices=[], polygons=[],
                                                                          polygon_type_indices=[]))))
    nd.update(1e-6, 1e-6)
    nd.flatten()


def test_error_on_set_zero_or_negative_type_index_cost(make_backend):  # nav_data_test.rs:1029-1041
    from landmass_b200 import XY, Archipelago, ArchipelagoOptions
    from landmass_b200.nav_data import SetTypeIndexCostError
    arch = Archipelago(ArchipelagoOptions.from_agent_radius(0.5, cs=XY), cs=XY, backend=make_backend())
    for cost in (0.0, -1.0):
        with pytest.raises(SetTypeIndexCostError):
            arch.set_type_index_cost(0, cost)


def connected(arch, agent_kwargs=None):
    """`are_nodes_connected(island_1 node 0, island_2 node 0)` as one whole update sees it: an agent on
    island 1 with a target on island 2 either gets a path or — unconnected regions — fails without
    exploring a single node (pathfinding.rs:237-245)."""
    from landmass_b200 import XY, Agent, AgentState
    a = Agent.create((0.5, 0.5), (0.0, 0.0), 0.25, 1.0, 1.0, cs=XY)
    a.current_target = (0.5, arch._island_2_y + 0.5)
    for k, v in (agent_kwargs or {}).items():
        setattr(a, k, v)
    aid = arch.add_agent(a)
    arch.update(0.01)
    (result,) = arch.get_pathing_results()
    state = arch.get_agent(aid).state()
    arch.remove_agent(aid)
    if result.success:
        assert state == AgentState.Moving
        return True
    assert state == AgentState.NoPath and result.explored_nodes == 0
    return False


def two_squares_arch(make_backend):
    from landmass_b200 import XY, Archipelago, ArchipelagoOptions
    arch = Archipelago(ArchipelagoOptions.from_agent_radius(0.5, cs=XY), cs=XY, backend=make_backend())
    mesh = valid_mesh(SQUARE_2D, cs=XY)
    arch.add_island(Island(Transform((0.0, 0.0), 0.0), mesh))
    k2 = arch.add_island(Island(Transform((0.0, 2.0), 0.0), mesh))
    arch._island_2_y = 2.0
    return arch, k2


def test_changed_island_rebuilds_region_connectivity(make_backend):  # nav_data_test.rs:1043-1108
    arch, k2 = two_squares_arch(make_backend)
    assert not connected(arch)
    arch.get_island_mut(k2).set_transform(Transform((0.0, 1.0), 0.0))
    arch._island_2_y = 1.0
    assert connected(arch)
    arch.get_island_mut(k2).set_transform(Transform((0.0, 2.0), 0.0))
    arch._island_2_y = 2.0
    assert not connected(arch)


def test_changed_animation_link_rebuilds_region_connectivity(make_backend):  # nav_data_test.rs:1110-1175
    from landmass_b200 import AnimationLink
    arch, _ = two_squares_arch(make_backend)
    assert not connected(arch)
    l = arch.add_animation_link(AnimationLink(start_edge=((0.1, 0.9), (0.9, 0.9)), end_edge=((0.1, 2.1), (0.9, 2.1)),
                                              cost=1.0, kind=0, bidirectional=False))
    assert connected(arch)
    arch.remove_animation_link(l)
    assert not connected(arch)


def test_permitted_animation_link_blocks_region_connectivity(make_backend):  # nav_data_test.rs:1177-1245
    from landmass_b200 import AnimationLink, PermittedAnimationLinks
    arch, _ = two_squares_arch(make_backend)
    for kind in (0, 1):
        arch.add_animation_link(AnimationLink(start_edge=((0.0, 0.9), (1.0, 0.9)), end_edge=((0.0, 2.1), (1.0, 2.1)),
                                              kind=kind, cost=1.0, bidirectional=False))
    assert connected(arch)
    for kinds, want in [({0}, True), ({1}, True), ({2}, False)]:
        assert connected(arch, dict(permitted_animation_links=PermittedAnimationLinks(frozenset(kinds)))) == want
