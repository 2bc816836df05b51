"""`draw_archipelago_debug` known answers — transcribed from crates/landmass/src/debug_test.rs:
draws_island_meshes_and_agents (44-393), draws_boundary_links (395-467), draws_animation_links
(469-547), draws_height_mesh (549-704), fails_to_draw_dirty_archipelago (706-751).  The per-agent
lines come from the device read-backs (`lmb_agent_path`, `lmb_agent_path_endpoints`,
`lmb_agent_waypoints`), so each test runs on the oracle and on the CUDA library.  The
feature-gated draws_avoidance_data_when_requested (753-909) reads `lmb_agent_avoidance_constraints`;
it is the one reference golden that pins individual ORCA constraint lines (obstacle lines and a
neighbour line, to 1e-3)."""
import numpy as np
import pytest

from landmass_b200 import (Agent, AnimationLink, ConstraintKind, ConstraintLine, draw_avoidance_data, Archipelago, ArchipelagoOptions, DebugDrawError, HeightNavigationMesh,
                           HeightPolygon, Island, LineType, PointType, Transform, TriangleType,
                           draw_archipelago_debug)
from tests.helpers import valid_mesh


class FakeDrawer:  # debug_test.rs:14-42
    def __init__(self):
        self.points, self.lines, self.triangles = [], [], []

    def add_point(self, point_type, point):
        self.points.append((point_type, point))

    def add_line(self, line_type, line):
        self.lines.append((line_type, line))

    def add_triangle(self, triangle_type, triangle):
        self.triangles.append((triangle_type, triangle))


def rounded(items, offset=(0.0, 0.0, 0.0)):
    """(type, points) -> sortable (type, ((x, y, z), ...)) moved by -offset and rounded to 0.01 like the test."""
    off = np.float32(offset)
    out = []
    for kind, pts in items:
        pts = pts if isinstance(pts, list) else [pts]
        out.append((kind, tuple(tuple(float(x) + 0.0 for x in
                                      np.round((np.asarray(p, dtype=np.float32) - off).astype(np.float64) * 100.0) / 100.0)
                                for p in pts)))
    return sorted(out, key=repr)


def expect(kind, *pts):
    return (kind, tuple(tuple(float(x) for x in p) for p in pts))


def new_arch(make_backend):
    return Archipelago(ArchipelagoOptions.from_agent_radius(0.5), backend=make_backend())


def test_draws_island_meshes_and_agents(make_backend):  # debug_test.rs:44-393
    mesh = valid_mesh(dict(
        vertices=[(1.0, 0.0, 0.0), (2.0, 0.0, 0.0), (4.0, 1.0, 0.0), (4.0, 2.0, 0.0), (2.0, 3.0, 0.0), (1.0, 3.0, 0.0),
                  (0.0, 2.0, 0.0), (0.0, 1.0, 0.0), (1.0, 5.0, 0.0), (2.0, 5.0, 0.0), (2.0, 4.0, 0.0), (3.0, 5.0, 1.0),
                  (3.0, 4.0, 1.0), (3.0, 4.0, -2.0), (3.0, 3.0, -2.0)],
        polygons=[[0, 1, 2, 3, 4, 5, 6, 7], [5, 4, 10, 9, 8], [9, 10, 12, 11], [10, 4, 14, 13]],
        polygon_type_indices=[0] * 4))
    arch = new_arch(make_backend)
    T = (1.0, 1.0, 1.0)
    arch.add_island(Island(Transform(T, 0.0), mesh))
    a = Agent.create((3.9 + 1.0, 1.5 + 1.0, 0.0 + 1.0), (0.0, 0.0, 0.0), 1.0, 1.0, 1.0)
    a.current_target = (1.5 + 1.0, 4.5 + 1.0, 0.0 + 1.0)
    agent = arch.add_agent(a)
    arch.update(1.0)
    d = FakeDrawer()
    draw_archipelago_debug(arch, d)
    assert rounded(d.points, T) == sorted([
        expect(PointType.AgentPosition(agent), (3.9, 1.5, 0.0)),
        expect(PointType.TargetPosition(agent), (1.5, 4.5, 0.0)),
        expect(PointType.Waypoint(agent), (2.0, 3.0, 0.0))], key=repr)
    B, C = LineType.BoundaryEdge, LineType.ConnectivityEdge
    assert rounded(d.lines, T) == sorted([
        expect(B, (0.0, 1.0, 0.0), (1.0, 0.0, 0.0)), expect(B, (0.0, 2.0, 0.0), (0.0, 1.0, 0.0)),
        expect(B, (1.0, 0.0, 0.0), (2.0, 0.0, 0.0)), expect(B, (1.0, 3.0, 0.0), (0.0, 2.0, 0.0)),
        expect(B, (1.0, 5.0, 0.0), (1.0, 3.0, 0.0)), expect(B, (2.0, 0.0, 0.0), (4.0, 1.0, 0.0)),
        expect(B, (2.0, 3.0, 0.0), (3.0, 3.0, -2.0)), expect(B, (2.0, 4.0, 0.0), (3.0, 4.0, 1.0)),
        expect(B, (2.0, 5.0, 0.0), (1.0, 5.0, 0.0)), expect(B, (3.0, 3.0, -2.0), (3.0, 4.0, -2.0)),
        expect(B, (3.0, 4.0, -2.0), (2.0, 4.0, 0.0)), expect(B, (3.0, 4.0, 1.0), (3.0, 5.0, 1.0)),
        expect(B, (3.0, 5.0, 1.0), (2.0, 5.0, 0.0)), expect(B, (4.0, 1.0, 0.0), (4.0, 2.0, 0.0)),
        expect(B, (4.0, 2.0, 0.0), (2.0, 3.0, 0.0)),
        expect(C, (2.0, 3.0, 0.0), (1.0, 3.0, 0.0)), expect(C, (2.0, 3.0, 0.0), (2.0, 4.0, 0.0)),
        expect(C, (2.0, 4.0, 0.0), (2.0, 5.0, 0.0)),
        expect(LineType.AgentCorridor(agent), (3.9, 1.5, 0.0), (1.5, 3.0, 0.0)),
        expect(LineType.AgentCorridor(agent), (1.5, 3.0, 0.0), (1.5, 4.5, 0.0)),
        expect(LineType.Target(agent), (3.9, 1.5, 0.0), (1.5, 4.5, 0.0)),
        expect(LineType.Waypoint(agent), (3.9, 1.5, 0.0), (2.0, 3.0, 0.0))], key=repr)
    N = TriangleType.Node
    c0, c1, c2, c3 = (1.75, 1.5, 0.0), (1.6, 4.0, 0.0), (2.5, 4.5, 0.5), (2.5, 3.5, -1.0)
    assert rounded(d.triangles, T) == sorted([
        expect(N, (0.0, 1.0, 0.0), (1.0, 0.0, 0.0), c0), expect(N, (0.0, 2.0, 0.0), (0.0, 1.0, 0.0), c0),
        expect(N, (1.0, 0.0, 0.0), (2.0, 0.0, 0.0), c0), expect(N, (1.0, 3.0, 0.0), (0.0, 2.0, 0.0), c0),
        expect(N, (1.0, 3.0, 0.0), (2.0, 3.0, 0.0), c1), expect(N, (1.0, 5.0, 0.0), (1.0, 3.0, 0.0), c1),
        expect(N, (2.0, 0.0, 0.0), (4.0, 1.0, 0.0), c0), expect(N, (2.0, 3.0, 0.0), (1.0, 3.0, 0.0), c0),
        expect(N, (2.0, 3.0, 0.0), (2.0, 4.0, 0.0), c1), expect(N, (2.0, 3.0, 0.0), (3.0, 3.0, -2.0), c3),
        expect(N, (2.0, 4.0, 0.0), (2.0, 3.0, 0.0), c3), expect(N, (2.0, 4.0, 0.0), (2.0, 5.0, 0.0), c1),
        expect(N, (2.0, 4.0, 0.0), (3.0, 4.0, 1.0), c2), expect(N, (2.0, 5.0, 0.0), (1.0, 5.0, 0.0), c1),
        expect(N, (2.0, 5.0, 0.0), (2.0, 4.0, 0.0), c2), expect(N, (3.0, 3.0, -2.0), (3.0, 4.0, -2.0), c3),
        expect(N, (3.0, 4.0, -2.0), (2.0, 4.0, 0.0), c3), expect(N, (3.0, 4.0, 1.0), (3.0, 5.0, 1.0), c2),
        expect(N, (3.0, 5.0, 1.0), (2.0, 5.0, 0.0), c2), expect(N, (4.0, 1.0, 0.0), (4.0, 2.0, 0.0), c0),
        expect(N, (4.0, 2.0, 0.0), (2.0, 3.0, 0.0), c0)], key=repr)


SQUARE_111 = dict(vertices=[(1.0, 1.0, 1.0), (2.0, 1.0, 1.0), (2.0, 2.0, 1.0), (1.0, 2.0, 1.0)], polygons=[[0, 1, 2, 3]],
                  polygon_type_indices=[0])


def test_draws_boundary_links(make_backend):  # debug_test.rs:395-467
    mesh = valid_mesh(SQUARE_111)
    arch = new_arch(make_backend)
    arch.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), mesh))
    arch.add_island(Island(Transform((1.0, 0.0, 0.0), 0.0), mesh))
    a = Agent.create((1.5, 1.25, 1.0), (0.0, 0.0, 0.0), 0.5, 1.0, 2.0)
    a.current_target = (2.5, 1.25, 1.0)
    agent = arch.add_agent(a)
    arch.update(1.0)
    d = FakeDrawer()
    draw_archipelago_debug(arch, d)
    lines = rounded(d.lines)
    for want in [
        expect(LineType.BoundaryLink, (2.0, 2.0, 1.0), (2.0, 1.0, 1.0)),
        expect(LineType.Waypoint(agent), (1.5, 1.25, 1.0), (2.5, 1.25, 1.0)),   # straight to the end
        expect(LineType.Target(agent), (1.5, 1.25, 1.0), (2.5, 1.25, 1.0)),
        expect(LineType.AgentCorridor(agent), (1.5, 1.25, 1.0), (2.0, 1.5, 1.0)),  # through the link's midpoint
        expect(LineType.AgentCorridor(agent), (2.0, 1.5, 1.0), (2.5, 1.25, 1.0)),
    ]:
        assert want in lines, want
    assert sum(1 for kind, _ in lines if kind == LineType.BoundaryLink) == 1  # drawn from one side only


def test_draws_animation_links(make_backend):  # debug_test.rs:469-547
    mesh = valid_mesh(dict(vertices=[(0.0, 0.0, 0.0), (1.0, 0.0, 0.0), (1.0, 1.0, 0.0), (0.0, 1.0, 0.0)],
                           polygons=[[0, 1, 2, 3]], polygon_type_indices=[0]))
    arch = new_arch(make_backend)
    arch.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), mesh))
    arch.add_island(Island(Transform((2.0, 0.0, 0.0), 0.0), mesh))
    link = arch.add_animation_link(AnimationLink(start_edge=((0.9, 0.0, 0.0), (0.9, 1.0, 0.0)),
                                                 end_edge=((2.1, 0.0, 0.0), (2.1, 1.0, 0.0)), kind=0, cost=1.0,
                                                 bidirectional=False))
    a = Agent.create((0.5, 0.25, 0.0), (0.0, 0.0, 0.0), 0.5, 1.0, 2.0)
    a.current_target = (2.5, 0.25, 0.0)
    agent = arch.add_agent(a)
    arch.update(1.0)
    d = FakeDrawer()
    draw_archipelago_debug(arch, d)
    lines = rounded(d.lines)
    for want in [
        expect(LineType.AnimationLinkStart(link), (0.9, 0.0, 0.0), (0.9, 1.0, 0.0)),
        expect(LineType.AnimationLinkEnd(link), (2.1, 0.0, 0.0), (2.1, 1.0, 0.0)),
        expect(LineType.AnimationLinkConnection(link), (0.9, 0.5, 0.0), (2.1, 0.5, 0.0)),
        expect(LineType.CorridorAnimationLink(agent, link), (0.9, 0.5, 0.0), (2.1, 0.5, 0.0)),
        expect(LineType.Waypoint(agent), (0.5, 0.25, 0.0), (0.9, 0.25, 0.0)),      # to the link's start point
        expect(LineType.PathAnimationLink(agent, link), (0.9, 0.25, 0.0), (2.1, 0.25, 0.0)),
    ]:
        assert want in lines, want


def test_draws_height_mesh(make_backend):  # debug_test.rs:549-704
    hm = HeightNavigationMesh(
        vertices=[(0.0, 0.0, -2.0), (1.0, 0.0, -2.0), (1.0, 1.0, -2.0), (0.0, 1.0, -2.0), (0.0, 0.5, -2.0),
                  (0.0, 1.0, 3.0), (1.0, 1.0, 3.0), (1.0, 2.0, 3.0), (0.0, 2.0, 3.0)],
        triangles=[[0, 1, 2], [2, 3, 4], [2, 4, 0], [0, 1, 2], [2, 3, 0]],
        polygons=[HeightPolygon(0, 5, 0, 3), HeightPolygon(5, 4, 3, 2)])
    mesh = valid_mesh(dict(
        vertices=[(0.0, 0.0, 0.0), (1.0, 0.0, 0.0), (1.0, 1.0, 0.0), (0.0, 1.0, 0.0), (1.0, 2.0, 0.0), (0.0, 2.0, 0.0)],
        polygons=[[0, 1, 2, 3], [3, 2, 4, 5]], polygon_type_indices=[0, 0], height_mesh=hm))
    arch = new_arch(make_backend)
    arch.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), mesh))
    arch.update(1.0)
    d = FakeDrawer()
    draw_archipelago_debug(arch, d)
    lines = rounded(d.lines)
    B, C, H = LineType.BoundaryEdge, LineType.ConnectivityEdge, LineType.HeightEdge
    for want in [
        expect(B, (0.0, 0.0, 0.0), (1.0, 0.0, 0.0)), expect(B, (1.0, 0.0, 0.0), (1.0, 1.0, 0.0)),
        expect(B, (1.0, 1.0, 0.0), (1.0, 2.0, 0.0)), expect(B, (1.0, 2.0, 0.0), (0.0, 2.0, 0.0)),
        expect(B, (0.0, 2.0, 0.0), (0.0, 1.0, 0.0)), expect(B, (0.0, 1.0, 0.0), (0.0, 0.0, 0.0)),
        expect(C, (1.0, 1.0, 0.0), (0.0, 1.0, 0.0)),
        expect(H, (0.0, 0.0, -2.0), (1.0, 0.0, -2.0)), expect(H, (1.0, 0.0, -2.0), (1.0, 1.0, -2.0)),
        expect(H, (1.0, 1.0, -2.0), (0.0, 0.0, -2.0)), expect(H, (1.0, 1.0, -2.0), (0.0, 1.0, -2.0)),
        expect(H, (0.0, 1.0, -2.0), (0.0, 0.5, -2.0)), expect(H, (0.0, 0.5, -2.0), (1.0, 1.0, -2.0)),
        expect(H, (1.0, 1.0, -2.0), (0.0, 0.5, -2.0)), expect(H, (0.0, 0.5, -2.0), (0.0, 0.0, -2.0)),
        expect(H, (0.0, 0.0, -2.0), (1.0, 1.0, -2.0)),
        expect(H, (0.0, 1.0, 3.0), (1.0, 1.0, 3.0)), expect(H, (1.0, 1.0, 3.0), (1.0, 2.0, 3.0)),
        expect(H, (1.0, 2.0, 3.0), (0.0, 1.0, 3.0)), expect(H, (1.0, 2.0, 3.0), (0.0, 2.0, 3.0)),
        expect(H, (0.0, 2.0, 3.0), (0.0, 1.0, 3.0)), expect(H, (0.0, 1.0, 3.0), (1.0, 2.0, 3.0)),
    ]:
        assert want in lines, want
    assert sum(1 for kind, _ in lines if kind == H) == 15


def test_fails_to_draw_dirty_archipelago(make_backend):  # debug_test.rs:706-751
    mesh = valid_mesh(SQUARE_111)
    d = FakeDrawer()
    arch = new_arch(make_backend)
    draw_archipelago_debug(arch, d)  # a brand new archipelago is clean
    island = arch.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), mesh))
    with pytest.raises(DebugDrawError) as e:
        draw_archipelago_debug(arch, d)
    assert e.value == DebugDrawError("NavDataDirty")
    arch.update(1.0)
    draw_archipelago_debug(arch, d)
    arch.get_island_mut(island).set_nav_mesh(mesh)
    with pytest.raises(DebugDrawError) as e:
        draw_archipelago_debug(arch, d)
    assert e.value == DebugDrawError("NavDataDirty")


def _equiv_line(expected: ConstraintLine, actual: ConstraintLine) -> bool:  # debug_test.rs:842-853
    e, a = expected.normal.astype(np.float64), actual.normal.astype(np.float64)
    angle = np.arctan2(e[0] * a[1] - e[1] * a[0], e[0] * a[0] + e[1] * a[1])  # Vec2::angle_to
    if abs(angle) >= 1e-3:
        return False
    return abs(float(np.dot(expected.point.astype(np.float64) - actual.point.astype(np.float64), a))) < 1e-3


def test_draws_avoidance_data_when_requested(make_backend):  # debug_test.rs:753-909
    mesh = valid_mesh(dict(vertices=[(1.0, 1.0, 1.0), (11.0, 1.0, 1.0), (11.0, 11.0, 1.0), (1.0, 11.0, 1.0)],
                           polygons=[[0, 1, 2, 3]], polygon_type_indices=[0]))
    opts = ArchipelagoOptions.from_agent_radius(0.5)
    opts.neighbourhood = 100.0
    opts.avoidance_time_horizon = 100.0
    opts.obstacle_avoidance_time_horizon = 100.0
    arch = Archipelago(opts, backend=make_backend())
    arch.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), mesh))
    a1 = Agent.create((6.0, 2.0, 1.0), (1.0, 1.0, 0.0), 0.5, 1.0, 1.0)
    a1.current_target = (6.0, 10.0, 1.0)
    a1.keep_avoidance_data = True
    agent_1 = arch.add_agent(a1)
    a2 = Agent.create((6.0, 10.0, 1.0), (0.0, 0.0, 0.0), 0.5, 1.0, 1.0)
    a2.current_target = (6.0, 2.0, 1.0)
    agent_2 = arch.add_agent(a2)
    arch.update(1.0)

    drawn = {}

    class FakeAvoidanceDrawer:
        def add_constraint(self, agent, constraint, kind):
            drawn.setdefault(agent, {}).setdefault(kind, []).append(constraint)

    draw_avoidance_data(arch, FakeAvoidanceDrawer())
    # only one of the agents was rendered, with original constraints only
    assert list(drawn.keys()) == [agent_1]
    assert list(drawn[agent_1].keys()) == [ConstraintKind.Original]
    actual = drawn[agent_1][ConstraintKind.Original]
    nb = np.array([1.007905, 8.0], dtype=np.float64)
    nb /= np.linalg.norm(nb)
    expected = [
        # lines for the edges of the nav mesh
        ConstraintLine((0.05, 0.0), (-1.0, 0.0)), ConstraintLine((0.0, -0.01), (0.0, 1.0)),
        ConstraintLine((-0.05, 0.0), (1.0, 0.0)), ConstraintLine((0.0, 0.09), (0.0, -1.0)),
        # line for the other agent: normal = -normalize(1.007905, 8).perp(), perp(x, y) = (-y, x)
        ConstraintLine((0.0, 0.0), (nb[1], -nb[0])),
    ]
    assert len(actual) == len(expected), actual
    unmatched = list(actual)
    for e in expected:  # unordered_elements_are!
        hit = next((a for a in unmatched if _equiv_line(e, a)), None)
        assert hit is not None, (e, actual)
        unmatched.remove(hit)

    # agent.avoidance_data = keep_avoidance_data.then_some(..) (avoidance.rs:172): asking later starts the
    # export, no longer asking drops the data at the next update that runs avoidance for the agent
    arch.get_agent(agent_2).keep_avoidance_data = True
    arch.get_agent(agent_1).keep_avoidance_data = False
    arch.update(1.0)
    drawn.clear()
    draw_avoidance_data(arch, FakeAvoidanceDrawer())
    assert list(drawn.keys()) == [agent_2]
    assert len(drawn[agent_2][ConstraintKind.Original]) == 5
