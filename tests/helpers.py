"""Shared helpers for the parity tests."""
import numpy as np

from landmass_b200 import (XY, XYZ, Archipelago, ArchipelagoOptions, Island, NavigationMesh,
                           Transform, _abi)


def valid_mesh(spec, cs=XYZ):
    return NavigationMesh(vertices=list(spec["vertices"]), polygons=[list(p) for p in spec["polygons"]],
                          polygon_type_indices=list(spec["polygon_type_indices"]),
                          height_mesh=spec.get("height_mesh"), cs=cs).validate()


def new_archipelago(backend, cs=XYZ, radius=0.5, **kw):
    return Archipelago(ArchipelagoOptions.from_agent_radius(radius, cs=cs), cs=cs, backend=backend, **kw)


def core_options(h, below, above, ratio):
    o = _abi.Options()
    o.horizontal_distance = h
    o.distance_below = below
    o.distance_above = above
    o.vertical_preference_ratio = ratio
    o.neighbourhood = 5.0
    o.avoidance_time_horizon = 0.5
    o.obstacle_avoidance_time_horizon = 0.25
    o.reached_destination_avoidance_responsibility = 0.1
    return o


def find_path_between_nodes(arch, start, end, overrides=None):
    """pathfinding_test.rs:25-50: start/end points are the polygon centres (island-local!)."""
    arch._sync_navdata()
    (si, sp), (ei, ep) = start, end
    s_center = arch.get_island(si).nav_mesh.poly_center[sp]
    e_center = arch.get_island(ei).nav_mesh.poly_center[ep]
    return find_path(arch, start, s_center, end, e_center, overrides)


def find_path(arch, start, start_point, end, end_point, overrides=None):
    arch._sync_navdata()
    sn = arch.node_id(*start)
    en = arch.node_id(*end)
    succ, expl, begin, nodes, portals = arch.backend.find_paths(
        [sn], [start_point], [en], [end_point], [overrides] if overrides is not None else None)
    path = None
    if succ[0]:
        path = segments(arch, nodes, portals)
    return path, int(expl[0])


def segments(arch, nodes, portals):
    """Flattened (node, portal) list -> reference `Path` shape:
    (island_segments [(island key, corridor, portal_edge_index)], off_mesh_links [(start node, end node, link)])."""
    segs, links = [], []
    cur = None
    for n, p in zip(nodes, portals):
        key, poly = arch.node_ref(int(n))
        if cur is None:
            cur = (key, [], [])
            segs.append(cur)
        cur[1].append(poly)
        p = int(p)
        if p == _abi.NONE:
            pass
        elif p & _abi.PORTAL_LINK:
            l = p & 0x7FFFFFFF
            links.append(((key, poly), arch.node_ref(int(arch._flat["link_dst_node"][l])), l))
            cur = None
        else:
            cur[2].append(p)
    return segs, links


def create_height_mesh(vertices, polygons):
    """nav_mesh_test.rs:843-884 test helper (fan-triangulates sub-polygons)."""
    from landmass_b200 import HeightNavigationMesh, HeightPolygon

    hv, hp, ht = [], [], []
    for polygon in polygons:
        base_v, base_t = len(hv), len(ht)
        for sub in polygon:
            base = len(hv) - base_v
            hv.append(vertices[sub[0]])
            hv.append(vertices[sub[1]])
            for i in range(2, len(sub)):
                hv.append(vertices[sub[i]])
                ht.append([base, base + i - 1, base + i])
        hp.append(HeightPolygon(base_v, len(hv) - base_v, base_t, len(ht) - base_t))
    return HeightNavigationMesh(polygons=hp, vertices=hv, triangles=ht)
