"""Query-API known answers (query.rs:70-94, 185-273) — transcribed from
crates/landmass/src/query_test.rs: `samples_point_on_nav_mesh_or_near_nav_mesh` (85-140), `no_path`
(223-270), `finds_path_with_override_type_index_costs` (331-417), `start_and_end_in_same_node`
(473-517), `one_animation_link_path` (519-599).  Exact equality like the reference (`assert_eq!`), through the batched entry points."""
import numpy as np

from landmass_b200 import XY, Archipelago, ArchipelagoOptions, Island, Transform
from tests.helpers import valid_mesh

ONE_NODE = dict(vertices=[(0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0)], polygons=[[0, 1, 2, 3]],
                polygon_type_indices=[0])


def new_arch(backend):
    return Archipelago(ArchipelagoOptions.from_agent_radius(0.5, cs=XY), cs=XY, backend=backend)


def straight_path(arch, start, end, overrides=None):
    """`Archipelago::find_path` (query.rs:185-273): waypoints between two sampled points, or None."""
    sn, en = arch.node_id(*start[1]), arch.node_id(*end[1])
    sp, ep = XY.to_landmass(start[0]), XY.to_landmass(end[0])
    succ, begin, kind, pts, link = arch.backend.find_straight_paths(
        [sn], [sp], [en], [ep], [overrides] if overrides is not None else None)
    if not succ[0]:
        return None
    assert (kind == 0).all()
    return [tuple(float(x) for x in p[:2]) for p in pts]


def test_samples_point_on_nav_mesh_or_near_nav_mesh(make_backend):  # query_test.rs:85-140
    arch = new_arch(make_backend())
    off = np.float32([10.0, 10.0])
    island = arch.add_island(Island(Transform(tuple(off), 0.0), valid_mesh(ONE_NODE, cs=XY)))
    for point, expected in [((-0.5, 0.5), (0.0, 0.5)), ((0.5, 0.5), (0.5, 0.5)), ((1.2, 1.2), (1.0, 1.0))]:
        got = arch.sample_point(tuple(off + np.float32(point)), np.float32(0.6))
        assert got is not None
        assert got[1][0] == island
        assert tuple(np.asarray(got[0], dtype=np.float32)) == tuple(off + np.float32(expected))


def test_no_path(make_backend):  # query_test.rs:223-270
    arch = new_arch(make_backend())
    off = np.float32([10.0, 10.0])
    mesh = valid_mesh(ONE_NODE, cs=XY)
    arch.add_island(Island(Transform(tuple(off), 0.0), mesh))
    arch.add_island(Island(Transform(tuple(off + np.float32([2.0, 0.0])), 0.0), mesh))
    start = arch.sample_point(tuple(off + np.float32([0.5, 0.5])), np.float32(1e-5))
    end = arch.sample_point(tuple(off + np.float32([2.5, 0.5])), np.float32(1e-5))
    assert start is not None and end is not None
    assert straight_path(arch, start, end) is None


def test_finds_path_with_override_type_index_costs(make_backend):  # query_test.rs:331-417
    arch = new_arch(make_backend())
    mesh = dict(
        vertices=[(0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0), (2.0, 0.0), (2.0, 1.0), (3.0, 0.0), (3.0, 1.0),
                  (2.0, 11.0), (3.0, 11.0), (2.0, 12.0), (3.0, 12.0), (1.0, 12.0), (1.0, 11.0), (0.0, 12.0), (0.0, 11.0)],
        polygons=[[0, 1, 2, 3], [2, 1, 4, 5], [5, 4, 6, 7], [5, 7, 9, 8], [8, 9, 11, 10], [8, 10, 12, 13],
                  [13, 12, 14, 15], [3, 2, 13, 15]],
        polygon_type_indices=[0, 0, 0, 0, 0, 0, 0, 1])
    arch.add_island(Island(Transform(), valid_mesh(mesh, cs=XY)))
    start = arch.sample_point((0.5, 0.5), np.float32(0.1))
    end = arch.sample_point((0.5, 11.5), np.float32(0.1))
    assert start is not None and end is not None
    # the direct way up costs 10x with the override: around the ring instead
    assert straight_path(arch, start, end, overrides={1: 10.0}) == [(0.5, 0.5), (2.0, 1.0), (2.0, 11.0), (0.5, 11.5)]
    # (not in the reference test) without the override: straight up through the type-1 node
    assert straight_path(arch, start, end) == [(0.5, 0.5), (0.5, 11.5)]


def test_start_and_end_in_same_node(make_backend):  # query_test.rs:473-517
    arch = new_arch(make_backend())
    arch.add_island(Island(Transform(), valid_mesh(ONE_NODE, cs=XY)))
    start = arch.sample_point((0.25, 0.25), np.float32(0.1))
    end = arch.sample_point((0.75, 0.75), np.float32(0.1))
    assert straight_path(arch, start, end) == [(0.25, 0.25), (0.75, 0.75)]


def test_one_animation_link_path(make_backend):  # query_test.rs:519-599
    """Start and end on two mesh chunks joined by an animation link: waypoint, link step (start and
    end points on the link's edges straight ahead of the start), waypoint — exact."""
    from landmass_b200 import AnimationLink, _abi

    arch = new_arch(make_backend())
    arch.add_island(Island(Transform(), valid_mesh(dict(
        vertices=[(0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0), (0.0, 2.0), (1.0, 2.0), (1.0, 3.0), (0.0, 3.0)],
        polygons=[[0, 1, 2, 3], [4, 5, 6, 7]], polygon_type_indices=[0, 0]), cs=XY)))
    link_key = arch.add_animation_link(AnimationLink(start_edge=((0.0, 1.0), (1.0, 1.0)),
                                                     end_edge=((0.0, 2.0), (1.0, 2.0)), cost=1.0, kind=0,
                                                     bidirectional=False))
    start = arch.sample_point((0.25, 0.25), np.float32(0.1))
    end = arch.sample_point((0.75, 2.75), np.float32(0.1))
    assert start is not None and end is not None
    sn, en = arch.node_id(*start[1]), arch.node_id(*end[1])
    succ, begin, kind, pts, link = arch.backend.find_straight_paths(
        [sn], [XY.to_landmass(start[0])], [en], [XY.to_landmass(end[0])])
    assert succ[0] == 1 and list(kind) == [0, 1, 0]
    assert tuple(pts[0][:2]) == (0.25, 0.25)
    assert tuple(pts[1][:2]) == (0.25, 1.0) and tuple(pts[1][3:5]) == (0.25, 2.0)
    assert tuple(pts[2][:2]) == (0.75, 2.75)
    assert int(link[1]) == link_key.idx  # the step names the animation link (link_id)
