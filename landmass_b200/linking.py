"""Island linking on the host — the cold half of `nav_data.update`
(crates/landmass/src/nav_data.rs:326-514, 827-1011, 1089-1156, 1254-1328;
geometry.rs:3-55).

Produces the read-only link tables the device path consumes:
  * boundary links between touching islands (`link_edges_between_islands`),
  * `ModifiedNode` boundaries (the reference clips boundary edges with the `geo`
    crate — absent here; for portals lying on boundary edges that clip is a 1-D
    interval subtraction along each edge, which is what is done below),
  * region connectivity (`update_regions`).

All geometry is f32 (numpy float32), evaluated in the reference's order.
Canonical orders replace the reference's HashMap/HashSet iteration orders
(SURVEY.md Appendix E): islands by dense order, boundary edges by (polygon, edge).
This module is rebuilt from scratch on every change rather than incrementally;
the resulting tables are the same up to link numbering.
"""
from __future__ import annotations

import numpy as np

from . import _abi

f32 = np.float32


def _world_vertices(nav_data, island):
    """transform.apply(vertex) for every vertex of the island (util.rs:358-361)."""
    from .nav_data import _rotate_z

    m = island.nav_mesh
    t = island.transform.landmass_translation(nav_data.cs)
    h = f32(f32(island.transform.rotation) * f32(0.5))
    s, c = f32(np.sin(h)), f32(np.cos(h))
    v = m.vertices
    if len(v) == 0:
        return v
    # vectorised _rotate_z
    z = f32(0)
    b2 = f32(f32(z * z + z * z) + s * s)
    k1 = f32(c * c - b2)
    k2 = ((v[:, 0] * z + v[:, 1] * z) + v[:, 2] * s) * f32(2)
    k3 = f32(c * f32(2))
    bxv = np.stack([z * v[:, 2] - v[:, 1] * s, s * v[:, 0] - v[:, 2] * z, z * v[:, 1] - v[:, 0] * z], axis=1)
    b = np.array([0, 0, s], dtype=f32)
    out = (v * k1 + b[None, :] * k2[:, None]) + bxv * k3
    return (out + t[None, :]).astype(f32)


def _dot3(a, b):
    return (a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]


def _clamp01(x):
    return np.minimum(np.maximum(x, f32(0)), f32(1))


def edge_intersection(e1a, e1b, e2a, e2b, dist_sq):
    """Vectorised geometry.rs:3-55.  Inputs (k,3) f32.  Returns (valid, p0, p1)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        d1 = e1b - e1a
        l1 = _dot3(d1, d1)
        t2a = _clamp01(_dot3(d1, e2a - e1a) / l1)
        t2b = _clamp01(_dot3(d1, e2b - e1a) / l1)
        ok = ~(t2a <= t2b)
        d2 = e2b - e2a
        l2 = _dot3(d2, d2)
        t1a = _clamp01(_dot3(d2, e1a - e2a) / l2)
        t1b = _clamp01(_dot3(d2, e1b - e2a) / l2)
        ok &= ~(t1a <= t1b)
        p2on1_0 = t2a[:, None] * d1 + e1a
        p2on1_1 = t2b[:, None] * d1 + e1a
        p1on2_0 = t1a[:, None] * d2 + e2a
        p1on2_1 = t1b[:, None] * d2 + e2a
        da = p2on1_1 - p1on2_0
        db = p2on1_0 - p1on2_1
        ok &= ~((_dot3(da, da) > dist_sq) | (_dot3(db, db) > dist_sq))
        p0 = (p2on1_1 + p1on2_0) * f32(0.5)
        p1 = (p2on1_0 + p1on2_1) * f32(0.5)
    return ok, p0.astype(f32), p1.astype(f32)


def _boundary_world_edges(m, wv):
    """World (left, right) points of every boundary edge, `get_edge_points` order
    (nav_mesh.rs:603-609): left = vertex[edge+1], right = vertex[edge]."""
    be = m.boundary_edges.astype(np.int64)
    if len(be) == 0:
        z = np.zeros((0, 3), dtype=f32)
        return z, z
    begin = m.poly_edge_begin.astype(np.int64)
    cnt = begin[be[:, 0] + 1] - begin[be[:, 0]]
    e = be[:, 1]
    nxt = np.where(e == cnt - 1, 0, e + 1)
    left = wv[m.poly_verts[begin[be[:, 0]] + nxt]]
    right = wv[m.poly_verts[begin[be[:, 0]] + e]]
    return left, right


def link_islands(nav_data, edge_link_distance, animation_link_distance=1.0):
    from .anim_links import create_animation_off_mesh_links
    from .nav_data import ModifiedNode, OffMeshLink

    d = f32(edge_link_distance)
    d2 = f32(d * d)
    items = nav_data.islands.items()
    world = [_world_vertices(nav_data, isl) for _, isl in items]
    bedges = [_boundary_world_edges(isl.nav_mesh, wv) for (_, isl), wv in zip(items, world)]
    bounds = []
    for wv in world:
        bounds.append(None if len(wv) == 0 else (wv.min(axis=0), wv.max(axis=0)))

    links = []
    for i1 in range(len(items)):
        for i2 in range(i1 + 1, len(items)):
            if bounds[i1] is None or bounds[i2] is None:
                continue
            if len(bedges[i1][0]) == 0 or len(bedges[i2][0]) == 0:
                continue
            mn1, mx1 = bounds[i1][0] - d, bounds[i1][1] + d
            mn2, mx2 = bounds[i2]
            if not (np.all(mn1 <= mx2) and np.all(mn2 <= mx1)):
                continue
            # link_edges_between_islands (nav_data.rs:1254-1328): island_1 = i1
            l1, r1 = bedges[i1]
            l2, r2 = bedges[i2]
            # prune to edges near the other island
            near1 = np.flatnonzero(np.all(np.minimum(l1, r1) <= mx2 + d, axis=1) & np.all(np.maximum(l1, r1) >= mn2 - d, axis=1))
            near2 = np.flatnonzero(np.all(np.minimum(l2, r2) <= mx1, axis=1) & np.all(np.maximum(l2, r2) >= mn1, axis=1))
            if len(near1) == 0 or len(near2) == 0:
                continue
            # edge bbox overlap (island_2 edge bbox expanded by d), all pairs
            a_min = np.minimum(l1[near1], r1[near1])
            a_max = np.maximum(l1[near1], r1[near1])
            b_min = np.minimum(l2[near2], r2[near2]) - d
            b_max = np.maximum(l2[near2], r2[near2]) + d
            ov = np.all(a_min[None, :, :] <= b_max[:, None, :], axis=2) & np.all(b_min[:, None, :] <= a_max[None, :, :], axis=2)
            j2, j1 = np.nonzero(ov)  # island_2 edge outer loop, island_1 edge inner
            if len(j1) == 0:
                continue
            e1 = near1[j1]
            e2 = near2[j2]
            ok, p0, p1 = edge_intersection(l1[e1], r1[e1], l2[e2], r2[e2], d2)
            dp = p0 - p1
            ok &= ~(_dot3(dp, dp) < d2)
            m1, m2 = items[i1][1].nav_mesh, items[i2][1].nav_mesh
            for k in np.flatnonzero(ok):
                poly1 = int(m1.boundary_edges[e1[k], 0])
                poly2 = int(m2.boundary_edges[e2[k], 0])
                a = OffMeshLink(items[i1][0], poly1, items[i2][0], poly2, int(m2.poly_type[poly2]),
                                np.stack([p0[k], p1[k]]), _abi.LINK_BOUNDARY)
                b = OffMeshLink(items[i2][0], poly2, items[i1][0], poly1, int(m1.poly_type[poly1]),
                                np.stack([p1[k], p0[k]]), _abi.LINK_BOUNDARY)
                links.append((i1, poly1, a, b))

    # canonical link order: by source node, forward link then its reverse in creation order
    dense = {key: i for i, (key, _) in enumerate(items)}
    flat_links = []
    for _, _, a, b in links:
        a._pair = b
        b._pair = a
        flat_links.extend([a, b])
    for l in create_animation_off_mesh_links(nav_data, animation_link_distance):
        l._pair = None
        flat_links.append(l)
    order = sorted(range(len(flat_links)), key=lambda i: (dense[flat_links[i].src_island],
                                                          flat_links[i].src_polygon, i))
    out_links = [flat_links[i] for i in order]
    index_of = {id(l): i for i, l in enumerate(out_links)}
    for l in out_links:
        l.reverse = index_of[id(l._pair)] if l._pair is not None else -1

    modified = _modified_nodes(nav_data, items, world, out_links, d)
    regions = _update_regions(items, out_links)
    return out_links, modified, regions


def _modified_nodes(nav_data, items, world, links, d):
    """update_modified_node (nav_data.rs:827-1011) with the geo clip restated as
    interval subtraction along each boundary edge."""
    from .nav_data import ModifiedNode

    dense = {key: i for i, (key, _) in enumerate(items)}
    by_node = {}
    for l in links:
        if l.kind != _abi.LINK_BOUNDARY:
            continue
        by_node.setdefault((l.src_island, l.src_polygon), []).append(l)
    result = []
    tol = max(float(d) * 4.0, 1e-5)
    for (ikey, poly), node_links in sorted(by_node.items(), key=lambda kv: (dense[kv[0][0]], kv[0][1])):
        ii = dense[ikey]
        m = items[ii][1].nav_mesh
        wv = world[ii]
        pv = m.polygon_vertices(poly).astype(np.int64)
        conn = m.connectivity(poly)
        n = len(pv)
        node = ModifiedNode(ikey, poly)
        nverts = len(m.vertices)
        center = None
        for e in range(n):
            if conn[e] is not None:
                continue
            ia, ib = int(pv[e]), int(pv[(e + 1) % n])
            a, b = wv[ia, :2].astype(np.float64), wv[ib, :2].astype(np.float64)
            ab = b - a
            L2 = float(ab @ ab)
            if L2 == 0.0:
                continue
            L = np.sqrt(L2)
            cuts = []
            for l in node_links:
                p0, p1 = l.portal[0, :2].astype(np.float64), l.portal[1, :2].astype(np.float64)
                # both portal ends must lie on this edge's line (within the clip half-width)
                dist0 = abs(ab[0] * (p0[1] - a[1]) - ab[1] * (p0[0] - a[0])) / L
                dist1 = abs(ab[0] * (p1[1] - a[1]) - ab[1] * (p1[0] - a[0])) / L
                if dist0 > tol or dist1 > tol:
                    continue
                t0 = float((p0 - a) @ ab) / L2
                t1 = float((p1 - a) @ ab) / L2
                lo, hi = max(min(t0, t1), 0.0), min(max(t0, t1), 1.0)
                if hi - lo <= 1e-9:
                    continue
                cuts.append((lo, hi))
            cuts.sort()
            merged = []
            for lo, hi in cuts:
                if merged and lo <= merged[-1][1] + 1e-9:
                    merged[-1] = (merged[-1][0], max(merged[-1][1], hi))
                else:
                    merged.append((lo, hi))
            pieces = []
            cur = 0.0
            for lo, hi in merged:
                if lo > cur:
                    pieces.append((cur, lo))
                cur = max(cur, hi)
            if cur < 1.0:
                pieces.append((cur, 1.0))
            for t0, t1 in pieces:
                s_pt = a + ab * t0
                e_pt = a + ab * t1

                def snap(pt):
                    # nearest original polygon vertex within squared distance 0.01 (nav_data.rs:958-978)
                    best, best_d = None, None
                    for vi in pv:
                        dd = float(((wv[vi, :2].astype(np.float64) - pt) ** 2).sum())
                        if best_d is None or dd < best_d:
                            best, best_d = int(vi), dd
                    return best if best_d is not None and best_d < 0.01 else None

                si, ei = snap(s_pt), snap(e_pt)
                if si is not None and ei is not None and si == ei:
                    continue
                if si is None:
                    node.new_vertices.append((f32(s_pt[0]), f32(s_pt[1])))
                    si = nverts + len(node.new_vertices) - 1
                if ei is None:
                    node.new_vertices.append((f32(e_pt[0]), f32(e_pt[1])))
                    ei = nverts + len(node.new_vertices) - 1
                if center is None:
                    from .nav_data import transform_apply

                    isl = items[ii][1]
                    center = transform_apply(isl.transform.landmass_translation(nav_data.cs),
                                             isl.transform.rotation, m.poly_center[poly])[:2].astype(np.float64)
                sc, ec = s_pt - center, e_pt - center
                if sc[0] * ec[1] - sc[1] * ec[0] < 0.0:
                    si, ei = ei, si
                node.new_boundary.append((si, ei))
        result.append(node)
    return result


def _update_regions(items, links):
    """update_regions (nav_data.rs:1089-1156) in canonical link order."""
    number = {}
    parent = []

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    def num(island_key, polygon):
        mesh = dict_items[island_key].nav_mesh
        rid = (island_key, int(mesh.poly_region[polygon]))
        if rid not in number:
            number[rid] = len(parent)
            parent.append(len(parent))
        return number[rid]

    dict_items = {k: v for k, v in items}
    for l in links:
        if l.kind != _abi.LINK_BOUNDARY:
            continue
        a, b = num(l.src_island, l.src_polygon), num(l.dst_island, l.dst_polygon)
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[max(ra, rb)] = min(ra, rb)
    possible = []
    for l in links:
        if l.kind != _abi.LINK_ANIMATION:
            continue
        a, b = num(l.src_island, l.src_polygon), num(l.dst_island, l.dst_polygon)
        if find(a) == find(b):
            continue
        possible.append((a, b, l.anim_kind))
    possible.sort(key=lambda p: p[0])  # grouped by source, stable
    return {"number": number, "root": [find(i) for i in range(len(parent))], "possible": possible}
