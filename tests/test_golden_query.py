"""Query-API known answers (`Archipelago::sample_point` / `find_path`: lib.rs:235-268,
query.rs:70-94, 185-273) — transcribed from crates/landmass/src/query_test.rs: every test of the
file (error_on_dirty_nav_mesh 16-47, error_on_out_of_range 49-82, samples_point_on_nav_mesh_or_near
84-139, samples_type_indices 141-220, no_path 222-270, finds_path 272-328,
finds_path_with_override_type_index_costs 330-416, find_path_returns_error_on_invalid_node_cost
418-470, start_and_end_in_same_node 472-516, one_animation_link_path 518-599) plus lib_test.rs
583-667.  Exact equality like the reference (`assert_eq!`), through the host mirror of the API and
the batched device entry points behind it."""
import numpy as np
import pytest

from landmass_b200 import (XY, AnimationLink, Archipelago, ArchipelagoOptions, FindPathError, Island, PathStep,
                           PermittedAnimationLinks, SamplePointError, Transform)
from tests.helpers import valid_mesh

ONE_NODE = dict(vertices=[(0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0)], polygons=[[0, 1, 2, 3]],
                polygon_type_indices=[0])


def new_arch(backend):
    return Archipelago(ArchipelagoOptions.from_agent_radius(0.5, cs=XY), cs=XY, backend=backend)


def waypoints(*points):
    return [PathStep.Waypoint(np.float32(p)) for p in points]


def test_error_on_dirty_nav_mesh(make_backend):  # query_test.rs:16-47
    arch = new_arch(make_backend())
    arch.add_island(Island(Transform(), valid_mesh(ONE_NODE, cs=XY)))
    with pytest.raises(SamplePointError) as e:
        arch.sample_point((0.5, 0.5), np.float32(1.0))
    assert e.value == SamplePointError("NavDataDirty")


def test_error_on_out_of_range(make_backend):  # query_test.rs:49-82
    arch = new_arch(make_backend())
    arch.add_island(Island(Transform(), valid_mesh(ONE_NODE, cs=XY)))
    arch.update(1.0)
    with pytest.raises(SamplePointError) as e:
        arch.sample_point((-0.5, 0.5), np.float32(0.1))
    assert e.value == SamplePointError("OutOfRange")


def test_samples_point_on_nav_mesh_or_near_nav_mesh(make_backend):  # query_test.rs:84-139, lib_test.rs:583-624
    arch = new_arch(make_backend())
    off = np.float32([10.0, 10.0])
    island = arch.add_island(Island(Transform(tuple(off), 0.0), valid_mesh(ONE_NODE, cs=XY)))
    arch.update(1.0)
    for point, expected in [((-0.5, 0.5), (0.0, 0.5)), ((0.5, 0.5), (0.5, 0.5)), ((1.2, 1.2), (1.0, 1.0))]:
        got = arch.sample_point(tuple(off + np.float32(point)), np.float32(0.6))
        assert got.island() == island
        assert tuple(got.point()) == tuple(off + np.float32(expected))


def test_samples_type_indices(make_backend):  # query_test.rs:141-220
    arch = new_arch(make_backend())
    for t in (1, 2, 3):
        arch.set_type_index_cost(t, float(t))
    off = np.float32([10.0, 10.0])
    arch.add_island(Island(Transform(tuple(off), 0.0), valid_mesh(dict(
        vertices=[(0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0), (2.0, 0.0), (2.0, 1.0), (3.0, 0.0), (3.0, 1.0),
                  (4.0, 0.0), (4.0, 1.0)],
        polygons=[[0, 1, 2, 3], [2, 1, 4, 5], [5, 4, 6, 7], [7, 6, 8, 9]], polygon_type_indices=[0, 1, 2, 3]), cs=XY)))
    arch.update(1.0)
    for k in range(4):
        assert arch.sample_point(tuple(off + np.float32([k + 0.5, 0.5])), np.float32(0.1)).type_index() == k


def test_no_path(make_backend):  # query_test.rs:222-270
    arch = new_arch(make_backend())
    off = np.float32([10.0, 10.0])
    mesh = valid_mesh(ONE_NODE, cs=XY)
    arch.add_island(Island(Transform(tuple(off), 0.0), mesh))
    arch.add_island(Island(Transform(tuple(off + np.float32([2.0, 0.0])), 0.0), mesh))
    arch.update(1.0)
    start = arch.sample_point(tuple(off + np.float32([0.5, 0.5])), np.float32(1e-5))
    end = arch.sample_point(tuple(off + np.float32([2.5, 0.5])), np.float32(1e-5))
    with pytest.raises(FindPathError) as e:
        arch.find_path(start, end, {}, PermittedAnimationLinks())
    assert e.value == FindPathError("NoPathFound")


def test_finds_path(make_backend):  # query_test.rs:272-328, lib_test.rs:626-667
    arch = new_arch(make_backend())
    mesh = valid_mesh(ONE_NODE, cs=XY)
    off = np.float32([10.0, 10.0])
    arch.add_island(Island(Transform(tuple(off), 0.0), mesh))
    arch.add_island(Island(Transform(tuple(off + np.float32([1.0, 0.0])), 0.0), mesh))
    arch.add_island(Island(Transform(tuple(off + np.float32([2.0, 0.5])), 0.0), mesh))
    arch.update(1.0)
    start = arch.sample_point(tuple(off + np.float32([0.5, 0.5])), np.float32(1e-5))
    end = arch.sample_point(tuple(off + np.float32([2.5, 1.25])), np.float32(1e-5))
    assert arch.find_path(start, end, {}, PermittedAnimationLinks()) == waypoints(
        off + np.float32([0.5, 0.5]), off + np.float32([2.0, 1.0]), off + np.float32([2.5, 1.25]))


def test_finds_path_with_override_type_index_costs(make_backend):  # query_test.rs:330-416
    arch = new_arch(make_backend())
    mesh = dict(
        vertices=[(0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0), (2.0, 0.0), (2.0, 1.0), (3.0, 0.0), (3.0, 1.0),
                  (2.0, 11.0), (3.0, 11.0), (2.0, 12.0), (3.0, 12.0), (1.0, 12.0), (1.0, 11.0), (0.0, 12.0), (0.0, 11.0)],
        polygons=[[0, 1, 2, 3], [2, 1, 4, 5], [5, 4, 6, 7], [5, 7, 9, 8], [8, 9, 11, 10], [8, 10, 12, 13],
                  [13, 12, 14, 15], [3, 2, 13, 15]],
        polygon_type_indices=[0, 0, 0, 0, 0, 0, 0, 1])
    arch.add_island(Island(Transform(), valid_mesh(mesh, cs=XY)))
    arch.update(1.0)
    start = arch.sample_point((0.5, 0.5), np.float32(0.1))
    end = arch.sample_point((0.5, 11.5), np.float32(0.1))
    # the direct way up costs 10x with the override: around the ring instead
    assert arch.find_path(start, end, {1: 10.0}) == waypoints((0.5, 0.5), (2.0, 1.0), (2.0, 11.0), (0.5, 11.5))
    # (not in the reference test) without the override: straight up through the type-1 node
    assert arch.find_path(start, end) == waypoints((0.5, 0.5), (0.5, 11.5))


def test_find_path_returns_error_on_invalid_node_cost(make_backend):  # query_test.rs:418-470
    arch = new_arch(make_backend())
    arch.add_island(Island(Transform(), valid_mesh(ONE_NODE, cs=XY)))
    arch.update(1.0)
    start = arch.sample_point((0.25, 0.25), np.float32(0.1))
    end = arch.sample_point((0.25, 0.25), np.float32(0.1))
    for cost in (0.0, -0.5):
        with pytest.raises(FindPathError) as e:
            arch.find_path(start, end, {0: cost}, PermittedAnimationLinks())
        assert e.value == FindPathError("NonPositiveTypeIndexCost", 0, cost)


def test_start_and_end_in_same_node(make_backend):  # query_test.rs:472-516
    arch = new_arch(make_backend())
    arch.add_island(Island(Transform(), valid_mesh(ONE_NODE, cs=XY)))
    arch.update(1.0)
    start = arch.sample_point((0.25, 0.25), np.float32(0.1))
    end = arch.sample_point((0.75, 0.75), np.float32(0.1))
    assert arch.find_path(start, end) == waypoints((0.25, 0.25), (0.75, 0.75))


def test_one_animation_link_path(make_backend):  # query_test.rs:518-599
    """Start and end on two mesh chunks joined by an animation link: waypoint, link step (start and
    end points on the link's edges straight ahead of the start), waypoint — exact.  (Not in the
    reference test: with the link's kind not permitted there is no path.)"""
    arch = new_arch(make_backend())
    arch.add_island(Island(Transform(), valid_mesh(dict(
        vertices=[(0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0), (0.0, 2.0), (1.0, 2.0), (1.0, 3.0), (0.0, 3.0)],
        polygons=[[0, 1, 2, 3], [4, 5, 6, 7]], polygon_type_indices=[0, 0]), cs=XY)))
    link_key = arch.add_animation_link(AnimationLink(start_edge=((0.0, 1.0), (1.0, 1.0)),
                                                     end_edge=((0.0, 2.0), (1.0, 2.0)), cost=1.0, kind=0,
                                                     bidirectional=False))
    arch.update(1.0)
    start = arch.sample_point((0.25, 0.25), np.float32(0.1))
    end = arch.sample_point((0.75, 2.75), np.float32(0.1))
    want = [PathStep.Waypoint((0.25, 0.25)), PathStep.AnimationLink((0.25, 1.0), (0.25, 2.0), link_key),
            PathStep.Waypoint((0.75, 2.75))]
    assert arch.find_path(start, end, {}, PermittedAnimationLinks()) == want
    assert arch.find_path(start, end, {}, PermittedAnimationLinks(frozenset({0, 3}))) == want
    for kinds in (frozenset(), frozenset({1})):
        with pytest.raises(FindPathError) as e:
            arch.find_path(start, end, {}, PermittedAnimationLinks(kinds))
        assert e.value == FindPathError("NoPathFound")
