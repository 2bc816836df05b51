"""Animation links on the host — the cold half that turns `AnimationLink`s into off-mesh links
(crates/landmass/src/link.rs:14-91, nav_data.rs:516-825, 1330-1426; nav_mesh.rs:819-923;
geometry.rs:63-170; util.rs:219-311).  f32 arithmetic in the reference's order; bounding-box
hierarchies are replaced by linear scans in ascending island / polygon order (the hierarchy only
prunes, and the reference's result order leaks through hash iteration anyway)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _abi
from .nav_data import transform_apply, transform_apply_inverse

f32 = np.float32


@dataclass
class AnimationLink:
    """link.rs:14-42."""

    start_edge: tuple
    end_edge: tuple
    kind: int = 0
    cost: float = 0.0
    bidirectional: bool = False


def _lerp(a, b, t):  # f32 FloatExt::lerp
    return f32(a + f32(f32(b - a) * t))


def _perp_dot(a, b):
    return f32(f32(a[0] * b[1]) - f32(a[1] * b[0]))


def _dot2(a, b):
    return f32(f32(a[0] * b[0]) + f32(a[1] * b[1]))


def clip_edge_to_triangle(tri, edge, max_vertical_distance):
    """geometry.rs:63-170.  tri: 3 x (3,) f32, edge: 2 x (3,) f32 -> (t0, t1) or None."""
    mv = f32(max_vertical_distance)
    t2 = [p[:2] for p in tri]
    e2 = [p[:2] for p in edge]
    deltas = (t2[1] - t2[0], t2[2] - t2[1], t2[0] - t2[2])
    edge_delta = e2[1] - e2[0]
    offs = (t2[0] - e2[0], t2[1] - e2[0], t2[2] - e2[0])
    v0, v1 = deltas[0], deltas[2]
    with np.errstate(divide="ignore", invalid="ignore"):
        denominator = _perp_dot(v0, v1)

        def estimate_height(point):
            v2 = point - t2[0]
            v = f32(_perp_dot(v2, v1) / denominator)
            w = f32(_perp_dot(v2, v0) / denominator)
            u = f32(f32(f32(1.0) - v) - w)
            return f32(f32(f32(tri[0][2] * u) + f32(tri[1][2] * v)) + f32(tri[2][2] * w))

        h_at = (estimate_height(e2[0]), estimate_height(e2[1]))

        def intersection_t(d1, d2, off):
            det = f32(-_perp_dot(d1, d2))
            if det == 0.0:
                return f32(np.inf)
            return f32(_dot2(np.array([-d2[1], d2[0]], dtype=f32), off) / det)

        ts = [intersection_t(edge_delta, deltas[k], offs[k]) for k in range(3)]
    tmin, tmax = f32(0.0), f32(1.0)
    for k in range(3):
        if np.isinf(ts[k]):
            if _perp_dot(offs[k], deltas[k]) < 0.0:
                return None
        elif _perp_dot(edge_delta, deltas[k]) < 0.0:
            tmin = max(tmin, ts[k])
        else:
            tmax = min(tmax, ts[k])
    if tmin >= 1.0 or tmax <= 0.0:
        return None
    th_min, th_max = _lerp(h_at[0], h_at[1], tmin), _lerp(h_at[0], h_at[1], tmax)
    eh_min, eh_max = _lerp(edge[0][2], edge[1][2], tmin), _lerp(edge[0][2], edge[1][2], tmax)
    slope = f32(f32(eh_max - eh_min) - f32(th_max - th_min))
    if slope == 0.0:
        return (tmin, tmax) if abs(f32(th_min - eh_min)) < mv else None
    t0 = f32(f32(f32(mv - eh_min) + th_min) / slope)
    t1 = f32(f32(f32(-mv - eh_min) + th_min) / slope)
    t0, t1 = min(t0, t1), max(t0, t1)
    if t0 >= 1.0 or t1 <= 0.0:
        return None
    t0, t1 = max(t0, f32(0.0)), min(t1, f32(1.0))
    return _lerp(tmin, tmax, t0), _lerp(tmin, tmax, t1)


def intersects_ray_segment(bmin, bmax, p0, p1):
    """BoundingBox::intersects_ray_segment (util.rs:219-262) with RaySegment::new (util.rs:491-502)."""
    d = (p1 - p0).astype(f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = (f32(1.0) / d).astype(f32)
        sign = [int(d[k] < 0.0) for k in range(3)]
        b = [[bmin[k], bmax[k]] for k in range(3)]
        tmin = f32(f32(b[0][sign[0]] - p0[0]) * inv[0])
        tmax = f32(f32(b[0][1 - sign[0]] - p0[0]) * inv[0])
        if tmin >= 1.0 or tmax <= 0.0:
            return False
        tymin = f32(f32(b[1][sign[1]] - p0[1]) * inv[1])
        tymax = f32(f32(b[1][1 - sign[1]] - p0[1]) * inv[1])
        if tmin > tymax or tymin > tmax:
            return False
        tmin, tmax = max(tmin, tymin), min(tmax, tymax)   # f32::max/min ignore NaN
        if np.isnan(tmin):
            tmin = tymin
        if tmin >= 1.0 or tmax <= 0.0:
            return False
        tzmin = f32(f32(b[2][sign[2]] - p0[2]) * inv[2])
        tzmax = f32(f32(b[2][1 - sign[2]] - p0[2]) * inv[2])
        if tmin > tzmax or tzmin > tmax:
            return False
        tmin = np.fmax(tmin, tzmin)
        tmax = np.fmin(tmax, tzmax)
        return bool(tmin < 1.0 and tmax > 0.0)


def _triangles(mesh, node):
    if mesh.hpoly is not None:
        bv, _, bt, nt = (int(x) for x in mesh.hpoly[node])
        for t in range(bt, bt + nt):
            a, b, c = (int(x) for x in mesh.htris[t])
            yield mesh.hverts[bv + a], mesh.hverts[bv + b], mesh.hverts[bv + c]
    else:
        pv = mesh.polygon_vertices(node)
        for i in range(2, len(pv)):
            yield mesh.vertices[pv[0]], mesh.vertices[pv[i - 1]], mesh.vertices[pv[i]]


def sample_edge(mesh, edge, max_vertical_distance):
    """ValidNavigationMesh::sample_edge (nav_mesh.rs:819-923): [(node, (t0, t1))]."""
    mv = f32(max_vertical_distance)
    out = []
    for node in range(mesh.n_polys):
        b = mesh.poly_bounds[node]
        if not b[0] <= b[3]:
            continue
        bmin = np.array([b[0], b[1], f32(b[2] - mv)], dtype=f32)
        bmax = np.array([b[3], b[4], f32(b[5] + mv)], dtype=f32)
        if not (bmax - bmin >= 0).all():
            continue
        if not intersects_ray_segment(bmin, bmax, edge[0], edge[1]):
            continue
        intervals = []
        for tri in _triangles(mesh, node):
            iv = clip_edge_to_triangle(tri, edge, mv)
            if iv is not None:
                intervals.append(iv)
        if not intervals:
            continue
        intervals.sort(key=lambda iv: (np.isnan(iv[0]), iv[0]))  # FloatOrd: NaN largest; stable
        cur = intervals[0]
        for iv in intervals[1:]:
            if f32(iv[0] - cur[1]) <= f32(1e-5):
                cur = (cur[0], iv[1])
            else:
                out.append((node, cur))
                cur = iv
        out.append((node, cur))
    return out


def _project_to_triangle(t0, t1, t2, p):
    """nav_mesh.rs:634-665 (f32)."""
    d = (t1 - t0, t2 - t1, t0 - t2)
    ts = (t0, t1, t2)
    for k in range(3):
        f = d[k][:2]
        rel = (p[:2] - ts[k][:2]).astype(f32)
        if _perp_dot(f, rel) < 0.0:
            s = f32(_dot2(f, rel) / _dot2(f, f))
            s = min(max(s, f32(0.0)), f32(1.0))
            return (d[k] * s + ts[k]).astype(f32)
    a, b = d[0], d[2]
    cr = np.array([f32(a[1] * b[2]) - f32(b[1] * a[2]), f32(a[2] * b[0]) - f32(b[2] * a[0]),
                   f32(a[0] * b[1]) - f32(b[0] * a[1])], dtype=f32)
    ln = f32(np.sqrt(f32(f32(f32(cr[0] * cr[0]) + f32(cr[1] * cr[1])) + f32(cr[2] * cr[2]))))
    normal = -(cr * f32(f32(1.0) / ln)).astype(f32)
    rel = (p - t0).astype(f32)
    dotv = f32(f32(f32(normal[0] * rel[0]) + f32(normal[1] * rel[1])) + f32(normal[2] * rel[2]))
    height = f32(dotv / normal[2])
    return np.array([p[0], p[1], f32(p[2] - height)], dtype=f32)


def sample_point_mesh(mesh, point, horizontal, above, below, ratio):
    """ValidNavigationMesh::sample_point (nav_mesh.rs:614-737) on the host (cold path only)."""
    h, above, below, ratio = f32(horizontal), f32(above), f32(below), f32(ratio)
    smin = point + np.array([-h, -h, -below], dtype=f32)
    smax = point + np.array([h, h, above], dtype=f32)
    best = None
    for node in range(mesh.n_polys):
        b = mesh.poly_bounds[node]
        if not b[0] <= b[3]:
            continue
        if not (smin[0] <= b[3] and b[0] <= smax[0] and smin[1] <= b[4] and b[1] <= smax[1]
                and smin[2] <= b[5] and b[2] <= smax[2]):
            continue
        for tri in _triangles(mesh, node):
            q = _project_to_triangle(tri[0], tri[1], tri[2], point)
            dxy = (point[:2] - q[:2]).astype(f32)
            dh = f32(np.sqrt(_dot2(dxy, dxy)))
            dv = f32(q[2] - point[2])
            if dh <= h and -below <= dv <= above:
                dist = f32(f32(dh * ratio) + abs(dv))
                if best is None or dist < best[2]:
                    best = (node, q, dist)
    return None if best is None else (best[1], best[0])


def transformed_bounds(mesh, translation, rotation):
    """BoundingBox::transform (util.rs:266-311)."""
    if mesh.mesh_bounds is None:
        return None
    mn, mx = mesh.mesh_bounds[:3], mesh.mesh_bounds[3:]
    pts = [transform_apply(translation, rotation, np.array(p, dtype=f32))
           for p in ((mn[0], mn[1], mn[2]), (mx[0], mx[1], mn[2]), (mn[0], mx[1], mn[2]), (mx[0], mn[1], mn[2]))]
    xs = [p[0] for p in pts]
    ys = [p[1] for p in pts]
    z0 = pts[0][2]
    return (np.array([min(xs), min(ys), z0], dtype=f32),
            np.array([max(xs), max(ys), f32(z0 + f32(mx[2] - mn[2]))], dtype=f32))


def _boxes_intersect(amin, amax, bmin, bmax):
    return bool(np.all(amin <= bmax) and np.all(bmin <= amax))


def world_portal_to_node_portals(portal, islands, cs, max_vertical_distance):
    """nav_data.rs:1330-1426: [(island key, polygon, (t0, t1))]."""
    mv = f32(max_vertical_distance)
    p0, p1 = portal
    out = []
    if (p0 == p1).all():
        # sample_animation_link_point
        qmin = p0 - np.array([0, 0, mv], dtype=f32)
        qmax = p0 + np.array([0, 0, mv], dtype=f32)
        best = None
        for key, island in islands:
            t = island.transform.landmass_translation(cs)
            tb = transformed_bounds(island.nav_mesh, t, island.transform.rotation)
            if tb is None or not _boxes_intersect(tb[0], tb[1], qmin, qmax):
                continue
            rel = transform_apply_inverse(t, island.transform.rotation, p0)
            s = sample_point_mesh(island.nav_mesh, rel, 0.0, mv, mv, 1.0)
            if s is None:
                continue
            dd = (rel - s[0]).astype(f32)
            dist = f32(f32(f32(dd[0] * dd[0]) + f32(dd[1] * dd[1])) + f32(dd[2] * dd[2]))
            if best is not None and dist >= best[0]:
                continue
            best = (dist, key, s[1])
        if best is not None:
            out.append((best[1], best[2], (f32(0.0), f32(1.0))))
        return out
    qmin = np.minimum(p0, p1) - np.array([0, 0, mv], dtype=f32)
    qmax = np.maximum(p0, p1) + np.array([0, 0, mv], dtype=f32)
    for key, island in islands:
        m = island.nav_mesh
        if m.n_polys == 0:
            continue
        t = island.transform.landmass_translation(cs)
        tb = transformed_bounds(m, t, island.transform.rotation)
        if tb is None or not _boxes_intersect(tb[0], tb[1], qmin, qmax):
            continue
        local = (transform_apply_inverse(t, island.transform.rotation, p0),
                 transform_apply_inverse(t, island.transform.rotation, p1))
        for node, iv in sample_edge(m, local, mv):
            out.append((key, node, iv))
    return out


def _vec_lerp(a, b, s):  # glam Vec3::lerp
    s = f32(s)
    return (a * f32(f32(1.0) - s) + b * s).astype(f32)


def effective_link_edges(link, cs):
    """The (start, end) world portals a link is sampled with: a point start edge turns the end edge
    into its midpoint, and for bidirectional links a point end edge does the same to the start."""
    start = (cs.to_landmass(link.start_edge[0]), cs.to_landmass(link.start_edge[1]))
    end = (cs.to_landmass(link.end_edge[0]), cs.to_landmass(link.end_edge[1]))
    if (start[0] == start[1]).all():
        mid = ((end[0] + end[1]) * f32(0.5)).astype(f32)
        end = (mid, mid)
    elif link.bidirectional and (end[0] == end[1]).all():
        mid = ((start[0] + start[1]) * f32(0.5)).astype(f32)
        start = (mid, mid)
    return start, end


def create_animation_off_mesh_links(nav_data, max_vertical_distance):
    """create_off_mesh_links_for_animation_links (nav_data.rs:516-825), rebuilt from scratch:
    for every animation link, every (start portal, end portal) pair with overlapping intervals."""
    from .nav_data import OffMeshLink

    cs = nav_data.cs
    islands = nav_data.islands.items()
    key_to_island = dict(islands)
    links = []
    for link_key, link in nav_data.animation_links.items():
        start, end = effective_link_edges(link, cs)
        start_portals = world_portal_to_node_portals(start, islands, cs, max_vertical_distance)
        if not start_portals:
            continue
        end_portals = world_portal_to_node_portals(end, islands, cs, max_vertical_distance)
        if not end_portals:
            continue
        for sk, sp, siv in start_portals:
            for ek, ep, eiv in end_portals:
                lo, hi = max(siv[0], eiv[0]), min(siv[1], eiv[1])
                if hi <= lo:
                    continue
                s_edge = np.stack([_vec_lerp(start[0], start[1], lo), _vec_lerp(start[0], start[1], hi)])
                e_edge = np.stack([_vec_lerp(end[0], end[1], lo), _vec_lerp(end[0], end[1], hi)])
                links.append(OffMeshLink(sk, sp, ek, ep, int(key_to_island[ek].nav_mesh.poly_type[ep]), s_edge,
                                         _abi.LINK_ANIMATION, -1, e_edge, float(link.cost), int(link.kind),
                                         int(link_key.idx), link_key))
                if link.bidirectional:
                    links.append(OffMeshLink(ek, ep, sk, sp, int(key_to_island[sk].nav_mesh.poly_type[sp]),
                                             e_edge, _abi.LINK_ANIMATION, -1, s_edge, float(link.cost),
                                             int(link.kind), int(link_key.idx), link_key))
    return links
