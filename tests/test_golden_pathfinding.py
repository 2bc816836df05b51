"""Corridor known answers — transcribed from crates/landmass/src/pathfinding_test.rs.
Corridors, portal edge indices and explored_nodes must match exactly."""
import math

import numpy as np
import pytest

from landmass_b200 import XY, XYZ, Island, Transform
from tests.golden import meshes
from tests.helpers import find_path, find_path_between_nodes, new_archipelago, valid_mesh


def seg(island, corridor, portals):
    return (island, corridor, portals)


def test_finds_path_in_archipelago(make_backend):  # pathfinding_test.rs:52-155
    arch = new_archipelago(make_backend())
    i = arch.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), valid_mesh(meshes.ORGANIC)))
    assert find_path_between_nodes(arch, (i, 0), (i, 2))[0] == ([seg(i, [0, 1, 2], [4, 2])], [])
    assert find_path_between_nodes(arch, (i, 2), (i, 0))[0] == ([seg(i, [2, 1, 0], [0, 0])], [])
    assert find_path_between_nodes(arch, (i, 3), (i, 0))[0] == ([seg(i, [3, 1, 0], [0, 0])], [])


def two_islands(make_backend):
    mesh = valid_mesh(meshes.ORGANIC)
    arch = new_archipelago(make_backend())
    i1 = arch.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), mesh))
    i2 = arch.add_island(Island(Transform((6.0, 0.0, 0.0), math.pi * -0.5), mesh))
    return arch, i1, i2


def test_finds_paths_on_two_islands(make_backend):  # pathfinding_test.rs:157-245
    arch, i1, i2 = two_islands(make_backend)
    assert find_path_between_nodes(arch, (i1, 0), (i1, 2))[0] == ([seg(i1, [0, 1, 2], [4, 2])], [])
    assert find_path_between_nodes(arch, (i2, 0), (i2, 2))[0] == ([seg(i2, [0, 1, 2], [4, 2])], [])


def test_no_path_between_disconnected_islands(make_backend):  # pathfinding_test.rs:247-317
    arch, i1, i2 = two_islands(make_backend)
    assert find_path_between_nodes(arch, (i1, 0), (i2, 0))[0] is None
    assert find_path_between_nodes(arch, (i2, 0), (i1, 0))[0] is None


def link_between(arch, a, b):
    """off_mesh_links[&(island_a, island_b)] of the reference tests."""
    f = arch._flat
    for l in range(f["n_links"]):
        if arch.node_ref(int(f["link_src_node"][l]))[0] == a and arch.node_ref(int(f["link_dst_node"][l]))[0] == b:
            return l
    raise KeyError


def test_find_path_across_connected_islands(make_backend):  # pathfinding_test.rs:319-429
    mesh = valid_mesh(meshes.CENTRED_SQUARE)
    arch = new_archipelago(make_backend())
    i1 = arch.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), mesh))
    i2 = arch.add_island(Island(Transform((1.0, 0.0, 0.0), 0.0), mesh))
    arch.add_island(Island(Transform((1.0, -1.0, 0.0), 0.0), mesh))
    i4 = arch.add_island(Island(Transform((1.0, 1.0, 0.0), 0.0), mesh))
    i5 = arch.add_island(Island(Transform((1.0, 2.0, 0.0), 0.0), mesh))
    path, _ = find_path_between_nodes(arch, (i1, 0), (i5, 0))
    assert path == (
        [seg(i1, [0], []), seg(i2, [0], []), seg(i4, [0], []), seg(i5, [0], [])],
        [((i1, 0), (i2, 0), link_between(arch, i1, i2)), ((i2, 0), (i4, 0), link_between(arch, i2, i4)),
         ((i4, 0), (i5, 0), link_between(arch, i4, i5))])


def test_finds_path_across_different_islands(make_backend):  # pathfinding_test.rs:431-523
    arch = new_archipelago(make_backend())
    i1 = arch.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), valid_mesh(meshes.CENTRED_SQUARE)))
    i2 = arch.add_island(Island(Transform((1.0, 0.0, 0.0), 0.0), valid_mesh(meshes.TWO_SQUARES)))
    path, _ = find_path_between_nodes(arch, (i1, 0), (i2, 1))
    assert path == ([seg(i1, [0], []), seg(i2, [0, 1], [1])],
                    [((i1, 0), (i2, 0), link_between(arch, i1, i2))])


def test_aborts_early_for_unconnected_regions(make_backend):  # pathfinding_test.rs:525-590
    mesh = valid_mesh(meshes.TALL_RECT)
    arch = new_archipelago(make_backend())
    i1 = arch.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), mesh))
    i2 = arch.add_island(Island(Transform((2.0, 0.0, 0.0), 0.0), mesh))
    i3 = arch.add_island(Island(Transform((1.5, 2.0, 0.0), math.pi * 0.5), mesh))
    assert find_path_between_nodes(arch, (i1, 0), (i2, 0))[0] is not None
    arch.remove_island(i3)
    path, explored = find_path_between_nodes(arch, (i1, 0), (i2, 0))
    assert path is None and explored == 0


def xy_arch(make_backend, spec, costs=()):
    arch = new_archipelago(make_backend(), cs=XY)
    for t, c in costs:
        arch.set_type_index_cost(t, c)
    i = arch.add_island(Island(Transform(), valid_mesh(spec, cs=XY)))
    return arch, i


def test_detour_for_high_cost_path(make_backend):  # pathfinding_test.rs:592-669
    arch, i = xy_arch(make_backend, meshes.ring(2, [0, 0, 0, 0, 0, 0, 0, 1]), [(1, 10.0)])
    assert find_path_between_nodes(arch, (i, 0), (i, 6))[0] == (
        [seg(i, [0, 1, 2, 3, 4, 5, 6], [1, 2, 3, 2, 3, 2])], [])


def test_detour_for_high_cost_path_across_boundary_links(make_backend):  # pathfinding_test.rs:671-777
    mesh_1 = dict(vertices=[(0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0), (2.0, 0.0), (2.0, 1.0), (3.0, 0.0), (3.0, 1.0)],
                  polygons=[[0, 1, 2, 3], [2, 1, 4, 5], [5, 4, 6, 7]], polygon_type_indices=[0, 0, 0])
    mesh_2 = dict(vertices=[(0.0, 2.0), (1.0, 2.0), (1.0, 3.0), (0.0, 3.0), (2.0, 2.0), (2.0, 3.0), (3.0, 2.0), (3.0, 3.0),
                            (0.0, 1.0), (1.0, 1.0), (2.0, 1.0), (3.0, 1.0)],
                  polygons=[[0, 1, 2, 3], [2, 1, 4, 5], [5, 4, 6, 7], [1, 0, 8, 9], [6, 4, 10, 11]],
                  polygon_type_indices=[0, 0, 0, 1, 0])
    arch = new_archipelago(make_backend(), cs=XY)
    arch.set_type_index_cost(1, 5.1)
    i1 = arch.add_island(Island(Transform(), valid_mesh(mesh_1, cs=XY)))
    i2 = arch.add_island(Island(Transform(), valid_mesh(mesh_2, cs=XY)))
    path, _ = find_path_between_nodes(arch, (i1, 0), (i2, 0))
    # straight up through the expensive node (type 1, cost 5.1) would be shorter; the detour over the
    # far boundary link is cheaper
    segs, links = path
    assert segs == [seg(i1, [0, 1, 2], [1, 2]), seg(i2, [4, 2, 1, 0], [0, 0, 0])]
    assert len(links) == 1 and links[0][0] == (i1, 2) and links[0][1] == (i2, 4)


def test_fast_path_not_ignored_by_heuristic(make_backend):  # pathfinding_test.rs:779-858
    arch, i = xy_arch(make_backend, meshes.ring(11, [1, 0, 0, 0, 0, 1, 1, 1]), [(1, 0.5)])
    assert find_path_between_nodes(arch, (i, 2), (i, 4))[0] == (
        [seg(i, [2, 1, 0, 7, 6, 5, 4], [0, 0, 2, 2, 0, 0])], [])


def test_infinite_or_nan_cost_cannot_find_path_between_nodes(make_backend):  # pathfinding_test.rs:860-911
    arch, i = xy_arch(make_backend, meshes.THREE_IN_ROW, [(1, math.inf)])
    path, explored = find_path_between_nodes(arch, (i, 0), (i, 2))
    assert path is None and explored == 1
    arch.set_type_index_cost(1, math.nan)
    path, explored = find_path_between_nodes(arch, (i, 0), (i, 2))
    assert path is None and explored == 1


def test_detour_for_overridden_high_cost_path(make_backend):  # pathfinding_test.rs:913-990
    arch, i = xy_arch(make_backend, meshes.ring(2, [0, 0, 0, 0, 0, 0, 0, 1]), [(1, 1.0)])
    assert find_path_between_nodes(arch, (i, 0), (i, 6), overrides={1: 10.0})[0] == (
        [seg(i, [0, 1, 2, 3, 4, 5, 6], [1, 2, 3, 2, 3, 2])], [])


def test_big_node_does_not_skew_pathing(make_backend):  # pathfinding_test.rs:992-1076
    arch, i = xy_arch(make_backend, meshes.BIG_NODE)
    assert find_path_between_nodes(arch, (i, 0), (i, 3))[0] == ([seg(i, [0, 7, 3], [2, 2])], [])


def test_start_and_end_point_influences_path(make_backend):  # pathfinding_test.rs:1078-1169
    arch, i = xy_arch(make_backend, meshes.START_END)
    path, _ = find_path(arch, (i, 1), (1.5, 0.5, 0.0), (i, 3), (1.5, 2.5, 0.0))
    assert path == ([seg(i, [1, 0, 4, 2, 3], [7, 2, 2, 1])], [])
