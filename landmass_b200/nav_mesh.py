"""`NavigationMesh::validate` -> `ValidNavigationMesh` on the host.

Mirror of crates/landmass/src/nav_mesh.rs:19-35, 131-185, 187-479, 483-583.
Validation is a one-off, host-side input producer for the device path (its
OUTPUT layout is the kernels' input), so it is vectorised numpy, not a kernel.
All arithmetic that the reference does in f32 is done in numpy float32.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .coords import XYZ

NONE = 0xFFFFFFFF


class ValidationError(Exception):
    """nav_mesh.rs:131-185.  `kind` is the Rust variant name, `args` its payload."""

    def __init__(self, kind: str, *args):
        super().__init__(f"{kind}{args}")
        self.kind = kind
        self.args_ = args

    def __eq__(self, other):
        return (isinstance(other, ValidationError) and self.kind == other.kind
                and self.args_ == other.args_)

    def __hash__(self):
        return hash((self.kind, self.args_))


@dataclass
class HeightPolygon:
    """nav_mesh.rs:74-90."""

    base_vertex_index: int
    vertex_count: int
    base_triangle_index: int
    triangle_count: int


@dataclass
class HeightNavigationMesh:
    """nav_mesh.rs:52-68."""

    polygons: list
    vertices: list
    triangles: list


@dataclass
class ValidNavigationMesh:
    """nav_mesh.rs:483-583, as SoA numpy arrays (island-local, landmass coordinates)."""

    vertices: np.ndarray          # (V,3) f32
    poly_edge_begin: np.ndarray   # (P+1,) u32
    poly_verts: np.ndarray        # (E,) u32 vertex index (local)
    edge_neighbor: np.ndarray     # (E,) u32 polygon index or NONE
    edge_reverse: np.ndarray      # (E,) u32
    poly_type: np.ndarray         # (P,) u32
    poly_region: np.ndarray       # (P,) u32
    poly_bounds: np.ndarray       # (P,6) f32
    poly_center: np.ndarray       # (P,3) f32
    mesh_bounds: np.ndarray | None  # (6,) f32 or None (Empty)
    boundary_edges: np.ndarray    # (B,2) u32 (polygon, edge), sorted
    # height mesh (optional)
    hpoly: np.ndarray | None = None     # (P,4) u32
    hverts: np.ndarray | None = None    # (HV,3) f32
    htris: np.ndarray | None = None     # (HT,3) u8
    cs: type = XYZ

    @property
    def n_polys(self) -> int:
        return len(self.poly_type)

    def polygon_vertices(self, p: int) -> np.ndarray:
        return self.poly_verts[self.poly_edge_begin[p]:self.poly_edge_begin[p + 1]]

    def connectivity(self, p: int):
        b, e = self.poly_edge_begin[p], self.poly_edge_begin[p + 1]
        return [None if n == NONE else (int(n), int(r))
                for n, r in zip(self.edge_neighbor[b:e], self.edge_reverse[b:e])]

    def get_bounds(self):
        return self.mesh_bounds

    def get_edge_indices(self, p: int, edge: int):
        """ValidPolygon::get_edge_indices (nav_mesh.rs:945-950): (left, right)."""
        v = self.polygon_vertices(p)
        return int(v[0 if edge == len(v) - 1 else edge + 1]), int(v[edge])


@dataclass
class NavigationMesh:
    """nav_mesh.rs:19-35."""

    vertices: list
    polygons: list
    polygon_type_indices: list
    height_mesh: HeightNavigationMesh | None = None
    cs: type = XYZ

    def validate(self) -> ValidNavigationMesh:
        cs = self.cs
        n_polys = len(self.polygons)
        if n_polys != len(self.polygon_type_indices):
            raise ValidationError("TypeIndicesHaveWrongLength", n_polys,
                                  len(self.polygon_type_indices))

        hm = None
        if self.height_mesh is not None:
            hm = _validate_height_mesh(self.height_mesh, n_polys, cs)

        if isinstance(self.vertices, np.ndarray) and self.vertices.ndim == 2 and self.vertices.shape[1] == 3 \
                and cs is XYZ:
            vertices = np.ascontiguousarray(self.vertices, dtype=np.float32)
        elif len(self.vertices):
            vertices = np.stack([cs.to_landmass(v) for v in self.vertices]).astype(np.float32)
        else:
            vertices = np.zeros((0, 3), dtype=np.float32)
        n_verts = len(vertices)

        uniform = isinstance(self.polygons, np.ndarray) and self.polygons.ndim == 2
        if uniform:
            counts = np.full(n_polys, self.polygons.shape[1], dtype=np.int64)
        else:
            counts = np.array([len(p) for p in self.polygons], dtype=np.int64)
        if n_polys and counts.min() < 3:
            raise ValidationError("NotEnoughVerticesInPolygon", int(np.argmax(counts < 3)))
        begin = np.zeros(n_polys + 1, dtype=np.int64)
        np.cumsum(counts, out=begin[1:])
        n_edges = int(begin[-1])
        if uniform:
            arr = self.polygons[:, ::-1] if cs.FLIP_POLYGONS else self.polygons
            flat = np.ascontiguousarray(arr, dtype=np.int64).ravel()
        elif cs.FLIP_POLYGONS:
            flat = np.fromiter((v for p in self.polygons for v in reversed(p)), dtype=np.int64,
                               count=n_edges)
        else:
            flat = np.fromiter((v for p in self.polygons for v in p), dtype=np.int64, count=n_edges)
        edge_poly = np.repeat(np.arange(n_polys, dtype=np.int64), counts)
        edge_idx = np.arange(n_edges, dtype=np.int64) - begin[edge_poly]

        # Errors are reported in the reference's discovery order: polygon by polygon,
        # first the out-of-range check, then per edge: degenerate, doubly connected, concave.
        errors = []  # (polygon, stage, edge, ValidationError)
        bad = (flat >= n_verts) | (flat < 0)
        if bad.any():
            p = int(edge_poly[np.argmax(bad)])
            errors.append((p, 0, 0, ValidationError("InvalidVertexIndexInPolygon", p)))
            # later checks index vertices; restrict them to polygons before p
            limit = int(begin[p])
        else:
            limit = n_edges

        nxt = np.arange(n_edges, dtype=np.int64) + 1
        last = edge_idx == counts[edge_poly] - 1
        nxt[last] = begin[edge_poly[last]]
        prv = np.arange(n_edges, dtype=np.int64) - 1
        first = edge_idx == 0
        prv[first] = begin[edge_poly[first] + 1] - 1

        a = flat[:limit]
        b = flat[nxt[:limit]]
        degenerate = a == b
        if degenerate.any():
            i = int(np.argmax(degenerate))
            errors.append((int(edge_poly[i]), 1, int(edge_idx[i]),
                           ValidationError("DegenerateEdgeInPolygon", int(edge_poly[i]))))

        lo = np.minimum(a, b)
        hi = np.maximum(a, b)
        key = lo * (n_verts + 1) + hi
        order = np.argsort(key, kind="stable")  # stable: ties keep polygon order
        sk = key[order]
        grp_start = np.ones(len(sk), dtype=bool)
        grp_start[1:] = sk[1:] != sk[:-1]
        grp_id = np.cumsum(grp_start) - 1
        grp_size = np.bincount(grp_id) if len(sk) else np.zeros(0, dtype=np.int64)
        rank = np.arange(len(sk)) - np.flatnonzero(grp_start)[grp_id] if len(sk) else np.zeros(0, dtype=np.int64)
        third = rank == 2
        if third.any():
            i = int(order[third].min())  # first edge (in polygon order) that is a third user
            errors.append((int(edge_poly[i]), 2, int(edge_idx[i]),
                           ValidationError("DoublyConnectedEdge", int(lo[i]), int(hi[i]))))

        # concavity (nav_mesh.rs:300-319), f32
        if limit:
            lv = vertices[flat[prv[:limit]], :2]
            cv = vertices[a, :2]
            rv = vertices[b, :2]
            left_edge = lv - cv
            right_edge = rv - cv
            pd = right_edge[:, 0] * left_edge[:, 1] - right_edge[:, 1] * left_edge[:, 0]
            dt = right_edge[:, 0] * left_edge[:, 0] + right_edge[:, 1] * left_edge[:, 1]
            ok = (pd > 0) | ((pd == 0) & (dt < 0))
            if not ok.all():
                i = int(np.argmax(~ok))
                errors.append((int(edge_poly[i]), 3, int(edge_idx[i]),
                               ValidationError("ConcavePolygon", int(edge_poly[i]))))
        if errors:
            # order: polygon, then (stage 0 before the edge loop), then edge, then stage
            errors.sort(key=lambda e: (e[0], 0 if e[1] == 0 else 1, e[2], e[1]))
            raise errors[0][3]

        # connectivity
        edge_neighbor = np.full(n_edges, NONE, dtype=np.uint32)
        edge_reverse = np.zeros(n_edges, dtype=np.uint32)
        pair_first = np.flatnonzero(grp_start & (grp_size[grp_id] == 2)) if len(sk) else np.zeros(0, dtype=np.int64)
        e1 = order[pair_first]
        e2 = order[pair_first + 1]
        edge_neighbor[e1] = edge_poly[e2]
        edge_reverse[e1] = edge_idx[e2]
        edge_neighbor[e2] = edge_poly[e1]
        edge_reverse[e2] = edge_idx[e1]

        # regions: connected components, normalised by first appearance (nav_mesh.rs:354-364)
        if n_polys:
            from scipy.sparse import coo_matrix
            from scipy.sparse.csgraph import connected_components

            g = coo_matrix((np.ones(len(e1), dtype=np.int8), (edge_poly[e1], edge_poly[e2])),
                           shape=(n_polys, n_polys))
            _, labels = connected_components(g, directed=False)
            _, first_idx = np.unique(labels, return_index=True)
            remap = np.empty(len(first_idx), dtype=np.int64)
            remap[np.argsort(first_idx)] = np.arange(len(first_idx))
            region = remap[labels].astype(np.uint32)
        else:
            region = np.zeros(0, dtype=np.uint32)

        # bounds (from the height mesh when present: nav_mesh.rs:334-345)
        if hm is not None:
            hpoly, hverts, _ = hm
            poly_bounds = np.zeros((n_polys, 6), dtype=np.float32)
            for p in range(n_polys):
                vs = hverts[hpoly[p, 0]:hpoly[p, 0] + hpoly[p, 1]]
                if len(vs) == 0:
                    poly_bounds[p] = (1, 1, 1, 0, 0, 0)  # Empty
                else:
                    poly_bounds[p, :3] = vs.min(axis=0)
                    poly_bounds[p, 3:] = vs.max(axis=0)
            bsrc = hverts
        else:
            pv = vertices[flat]
            poly_bounds = np.zeros((n_polys, 6), dtype=np.float32)
            if n_polys:
                poly_bounds[:, :3] = np.minimum.reduceat(pv, begin[:-1], axis=0)
                poly_bounds[:, 3:] = np.maximum.reduceat(pv, begin[:-1], axis=0)
            bsrc = vertices
        mesh_bounds = None
        if len(bsrc):
            mesh_bounds = np.concatenate([bsrc.min(axis=0), bsrc.max(axis=0)]).astype(np.float32)

        # center = sum(vertices in order, starting from 0) / n  (f32, sequential adds)
        center = np.zeros((n_polys, 3), dtype=np.float32)
        maxc = int(counts.max()) if n_polys else 0
        for k in range(maxc):
            m = counts > k
            center[m] = center[m] + vertices[flat[begin[:-1][m] + k]]
        if n_polys:
            center = (center / counts.astype(np.float32)[:, None]).astype(np.float32)

        boundary = np.flatnonzero(edge_neighbor == NONE)
        boundary_edges = np.stack([edge_poly[boundary], edge_idx[boundary]], axis=1).astype(np.uint32) \
            if len(boundary) else np.zeros((0, 2), dtype=np.uint32)

        return ValidNavigationMesh(
            vertices=vertices,
            poly_edge_begin=begin.astype(np.uint32),
            poly_verts=flat.astype(np.uint32),
            edge_neighbor=edge_neighbor,
            edge_reverse=edge_reverse,
            poly_type=np.asarray(self.polygon_type_indices, dtype=np.uint32),
            poly_region=region,
            poly_bounds=poly_bounds,
            poly_center=center,
            mesh_bounds=mesh_bounds,
            boundary_edges=boundary_edges,
            hpoly=hm[0] if hm is not None else None,
            hverts=hm[1] if hm is not None else None,
            htris=hm[2] if hm is not None else None,
            cs=cs,
        )


def _validate_height_mesh(hm: HeightNavigationMesh, expected_polygons: int, cs):
    """HeightNavigationMesh::validate (nav_mesh.rs:410-479)."""
    if len(hm.polygons) != expected_polygons:
        raise ValidationError("IncorrectNumberOfPolygons", expected_polygons, len(hm.polygons))
    verts = (np.stack([cs.to_landmass(v) for v in hm.vertices]).astype(np.float32)
             if len(hm.vertices) else np.zeros((0, 3), dtype=np.float32))
    tris = np.asarray(hm.triangles, dtype=np.int64).reshape(-1, 3)
    if cs.FLIP_POLYGONS:
        tris = tris[:, [0, 2, 1]]
    hpoly = np.zeros((expected_polygons, 4), dtype=np.uint32)
    for pi, poly in enumerate(hm.polygons):
        last = poly.base_triangle_index + poly.triangle_count
        if last > len(tris):
            raise ValidationError("InvalidPolygonIndices", pi)
        for ti in range(poly.base_triangle_index, last):
            real = []
            for idx in tris[ti]:
                r = poly.base_vertex_index + int(idx)
                if r >= len(verts) or idx >= poly.vertex_count:
                    raise ValidationError("InvalidIndexInTriangle", ti, r)
                real.append(r)
            a, b, c = (verts[r, :2] for r in real)
            ab = b - a
            ac = c - a
            if np.float32(ab[0] * ac[1]) - np.float32(ab[1] * ac[0]) < 0:
                raise ValidationError("ClockwiseTriangle", pi)
        hpoly[pi] = (poly.base_vertex_index, poly.vertex_count, poly.base_triangle_index,
                     poly.triangle_count)
    return hpoly, verts, tris.astype(np.uint8)
