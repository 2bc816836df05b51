"""`debug::draw_archipelago_debug` — host-side mirror of crates/landmass/src/debug.rs:14-413.

Cold path: everything about the navigation data is drawn from the host copy (`NavigationData`), the
per-agent parts come from three read-backs of the device state (`lmb_agent_path`,
`lmb_agent_path_endpoints`, `lmb_agent_waypoints`).  Point / line / triangle types are plain tuples
whose first element is the Rust variant name, e.g. `LineType.AgentCorridor(agent_id)` ==
("AgentCorridor", agent_id).  The `debug-avoidance` feature (debug.rs:415-503) is mirrored by
`draw_avoidance_data` over the read-back `lmb_agent_avoidance_constraints`.
"""
from __future__ import annotations

import numpy as np

from . import _abi
from .nav_data import transform_apply

f32 = np.float32


class PointType:
    """debug.rs:15-23."""

    AgentPosition = staticmethod(lambda agent_id: ("AgentPosition", agent_id))
    TargetPosition = staticmethod(lambda agent_id: ("TargetPosition", agent_id))
    Waypoint = staticmethod(lambda agent_id: ("Waypoint", agent_id))


class LineType:
    """debug.rs:26-58."""

    BoundaryEdge = ("BoundaryEdge",)
    ConnectivityEdge = ("ConnectivityEdge",)
    HeightEdge = ("HeightEdge",)
    BoundaryLink = ("BoundaryLink",)
    AnimationLinkStart = staticmethod(lambda link_id: ("AnimationLinkStart", link_id))
    AnimationLinkEnd = staticmethod(lambda link_id: ("AnimationLinkEnd", link_id))
    AnimationLinkConnection = staticmethod(lambda link_id: ("AnimationLinkConnection", link_id))
    AgentCorridor = staticmethod(lambda agent_id: ("AgentCorridor", agent_id))
    CorridorAnimationLink = staticmethod(lambda agent_id, link_id: ("CorridorAnimationLink", agent_id, link_id))
    Target = staticmethod(lambda agent_id: ("Target", agent_id))
    Waypoint = staticmethod(lambda agent_id: ("Waypoint", agent_id))
    PathAnimationLink = staticmethod(lambda agent_id, link_id: ("PathAnimationLink", agent_id, link_id))


class TriangleType:
    """debug.rs:61-65."""

    Node = ("Node",)


class DebugDrawer:
    """debug.rs:69-77: implement these three to visualise an archipelago."""

    def add_point(self, point_type, point):
        raise NotImplementedError

    def add_line(self, line_type, line):
        raise NotImplementedError

    def add_triangle(self, triangle_type, triangle):
        raise NotImplementedError


class DebugDrawError(Exception):
    """debug.rs:80-86: kind is "NavDataDirty"."""

    def __init__(self, kind: str):
        super().__init__(kind)
        self.kind = kind

    def __eq__(self, other):
        return isinstance(other, DebugDrawError) and other.kind == self.kind

    __hash__ = Exception.__hash__


def _midpoint(a, b):
    return ((np.asarray(a, dtype=f32) + np.asarray(b, dtype=f32)) * f32(0.5)).astype(f32)


def draw_archipelago_debug(archipelago, debug_drawer) -> None:
    """debug.rs:89-283.  Raises DebugDrawError("NavDataDirty") when the navigation data was mutated
    since the last `update`."""
    nav = archipelago.nav_data
    cs = archipelago.cs
    if nav.dirty:
        raise DebugDrawError("NavDataDirty")
    links_of_node = {}
    for link in nav.links:  # canonical order = ascending link index per node
        links_of_node.setdefault((link.src_island, int(link.src_polygon)), []).append(link)
    link_keys = {k.idx: k for k in nav.get_animation_link_ids()}

    for island_id in archipelago.get_island_ids():
        island = archipelago.get_island(island_id)
        assert not island.dirty, \
            "Drawing an archipelago while things are dirty is unsafe! Update the archipelago first."
        mesh = island.nav_mesh
        t, r = island.transform.landmass_translation(cs), f32(island.transform.rotation)

        def vertex(index):
            return cs.from_landmass(transform_apply(t, r, mesh.vertices[index]))

        for polygon in range(mesh.n_polys):
            verts = mesh.polygon_vertices(polygon)
            centre = cs.from_landmass(transform_apply(t, r, mesh.poly_center[polygon]))
            for i in range(len(verts)):
                j = (i + 1) % len(verts)
                debug_drawer.add_triangle(TriangleType.Node, [vertex(int(verts[i])), vertex(int(verts[j])), centre])
            if mesh.hpoly is not None:
                base_v, _, base_t, n_t = (int(x) for x in mesh.hpoly[polygon])
                for tri in range(base_t, base_t + n_t):
                    a, b, c = (cs.from_landmass(transform_apply(t, r, mesh.hverts[base_v + int(k)]))
                               for k in mesh.htris[tri])
                    debug_drawer.add_line(LineType.HeightEdge, [a, b])
                    debug_drawer.add_line(LineType.HeightEdge, [b, c])
                    debug_drawer.add_line(LineType.HeightEdge, [c, a])
            for edge, connection in enumerate(mesh.connectivity(polygon)):
                if connection is None:
                    line_type = LineType.BoundaryEdge
                elif polygon > connection[0]:
                    continue  # the neighbour draws the shared edge
                else:
                    line_type = LineType.ConnectivityEdge
                j = (edge + 1) % len(verts)
                debug_drawer.add_line(line_type, [vertex(int(verts[edge])), vertex(int(verts[j]))])
            for link in links_of_node.get((island_id, polygon), ()):
                if link.kind == _abi.LINK_BOUNDARY:
                    if (island_id, polygon) > (link.dst_island, int(link.dst_polygon)):
                        continue  # drawn from the other side
                    debug_drawer.add_line(LineType.BoundaryLink,
                                          [cs.from_landmass(link.portal[0]), cs.from_landmass(link.portal[1])])
                else:
                    key = link.anim_key
                    debug_drawer.add_line(LineType.AnimationLinkStart(key),
                                          [cs.from_landmass(link.portal[0]), cs.from_landmass(link.portal[1])])
                    debug_drawer.add_line(LineType.AnimationLinkEnd(key),
                                          [cs.from_landmass(link.dst_portal[0]), cs.from_landmass(link.dst_portal[1])])
                    debug_drawer.add_line(LineType.AnimationLinkConnection(key),
                                          [cs.from_landmass(_midpoint(*link.portal)),
                                           cs.from_landmass(_midpoint(*link.dst_portal))])

    waypoints = None
    order = {key: i for i, key in enumerate(archipelago._last_step_keys)}
    for agent_id, agent in archipelago.agents.items():
        if agent.paused or agent.is_using_animation_link():
            continue  # paused agents are not drawn
        position = np.asarray(agent.position, dtype=f32)
        debug_drawer.add_point(PointType.AgentPosition(agent_id), position)
        if agent.current_target is not None:
            target = np.asarray(agent.current_target, dtype=f32)
            debug_drawer.add_line(LineType.Target(agent_id), [position, target])
            debug_drawer.add_point(PointType.TargetPosition(agent_id), target)
        path = archipelago.backend.agent_path(agent_id.idx) if agent_id in order else None
        if path is None or len(path[0]) == 0:
            continue
        if waypoints is None:
            waypoints = archipelago.backend.agent_waypoints()
        _draw_path(archipelago, debug_drawer, agent_id, position, path, order[agent_id], waypoints, link_keys)


def _draw_path(archipelago, debug_drawer, agent_id, position, path, index, waypoints, link_keys):
    """debug.rs:287-413."""
    cs = archipelago.cs
    nav = archipelago.nav_data
    nodes, portals = path
    start_point, end_point = archipelago.backend.agent_path_endpoints(agent_id.idx)
    last_point = cs.from_landmass(start_point)
    for node, portal in zip(nodes, portals):
        portal = int(portal)
        if portal == _abi.NONE:  # the last node of the last island segment
            debug_drawer.add_line(LineType.AgentCorridor(agent_id), [last_point, cs.from_landmass(end_point)])
        elif portal & _abi.PORTAL_LINK:
            link = nav.links[portal & 0x7FFFFFFF]
            next_point = cs.from_landmass(_midpoint(*link.portal))
            debug_drawer.add_line(LineType.AgentCorridor(agent_id), [last_point, next_point])
            if link.kind == _abi.LINK_BOUNDARY:
                last_point = next_point
            else:
                destination = cs.from_landmass(_midpoint(*link.dst_portal))
                debug_drawer.add_line(LineType.CorridorAnimationLink(agent_id, link.anim_key), [next_point, destination])
                last_point = destination
        else:
            island_id, polygon = archipelago.node_ref(int(node))
            island = archipelago.get_island(island_id)
            left, right = island.nav_mesh.get_edge_indices(polygon, portal)
            mid = _midpoint(island.nav_mesh.vertices[left], island.nav_mesh.vertices[right])
            next_point = cs.from_landmass(transform_apply(island.transform.landmass_translation(cs),
                                                          f32(island.transform.rotation), mid))
            debug_drawer.add_line(LineType.AgentCorridor(agent_id), [last_point, next_point])
            last_point = next_point
    kind, points, link = waypoints
    if index >= len(kind) or kind[index] == _abi.NONE:
        return
    waypoint = points[index, :3]
    if kind[index] == 1:
        debug_drawer.add_line(LineType.PathAnimationLink(agent_id, link_keys[int(link[index])]),
                              [cs.from_landmass(points[index, :3]), cs.from_landmass(points[index, 3:])])
    debug_drawer.add_line(LineType.Waypoint(agent_id), [position, cs.from_landmass(waypoint)])
    debug_drawer.add_point(PointType.Waypoint(agent_id), cs.from_landmass(waypoint))


# --- feature `debug-avoidance` (debug.rs:415-503) ---------------------------------------------------
class ConstraintLine:
    """debug.rs:415-426: a half-plane of valid velocities; `normal` points towards the valid side."""

    __slots__ = ("point", "normal")

    def __init__(self, point, normal):
        self.point = np.asarray(point, dtype=f32)
        self.normal = np.asarray(normal, dtype=f32)

    @classmethod
    def from_dodgy(cls, line):
        """debug.rs:453-460: line = (point.x, point.y, direction.x, direction.y)."""
        return cls((line[0], line[1]), (-line[3], line[2]))

    def __repr__(self):
        return f"ConstraintLine(point={self.point.tolist()}, normal={self.normal.tolist()})"


class ConstraintKind:
    """debug.rs:428-439."""

    Original = "Original"
    Fallback = "Fallback"


class AvoidanceDrawer:
    """debug.rs:441-451."""

    def add_constraint(self, agent, constraint: ConstraintLine, kind):
        raise NotImplementedError


def draw_avoidance_data(archipelago, avoidance_drawer):
    """debug.rs:462-503: reports the constraints of every agent that kept its avoidance data."""
    for agent_id, agent in archipelago.agents.items():
        data = agent._avoidance_data
        if data is None:
            continue
        original = data[1]
        fallback = data[2] if data[0] == "Fallback" else ()
        for line in original:
            avoidance_drawer.add_constraint(agent_id, ConstraintLine.from_dodgy(line), ConstraintKind.Original)
        for line in fallback:
            avoidance_drawer.add_constraint(agent_id, ConstraintLine.from_dodgy(line), ConstraintKind.Fallback)
