"""Host-side `NavigationData` (crates/landmass/src/nav_data.rs:27-62): islands,
type costs, off-mesh links, and the flattening into the SoA description that
crosses the C ABI (include/landmass_b200.h `lmb_navdata_desc`).

The per-frame path never touches this module: `flatten()` runs only when
`update()` reports a change (reference nav_data.rs:1158-1180, cold path).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .coords import XYZ
from .linking import link_islands
from .nav_mesh import NONE, ValidNavigationMesh
from .slotmap import DenseSlotMap, Key


@dataclass
class Transform:
    """util.rs:315-368: translation + rotation about the up axis."""

    translation: object = None
    rotation: float = 0.0

    def landmass_translation(self, cs) -> np.ndarray:
        if self.translation is None:
            return np.zeros(3, dtype=np.float32)
        return cs.to_landmass(self.translation)


def _rotate_z(s, c, v):
    """Quat::from_rotation_z * v, glam SSE2 order; f32."""
    f = np.float32
    v = np.asarray(v, dtype=np.float32)
    b2 = f(f(f(0) * f(0) + f(0) * f(0)) + s * s)
    k1 = f(c * c - b2)
    k2 = f(f(f(v[0] * f(0) + v[1] * f(0)) + v[2] * s) * f(2))
    k3 = f(c * f(2))
    bxv = np.array([f(f(0) * v[2]) - f(v[1] * s), f(s * v[0]) - f(v[2] * f(0)),
                    f(f(0) * v[1]) - f(v[0] * f(0))], dtype=np.float32)
    b = np.array([0, 0, s], dtype=np.float32)
    return (v * k1 + b * k2) + bxv * k3


def transform_apply(translation, rotation, p):
    h = np.float32(np.float32(rotation) * np.float32(0.5))
    return _rotate_z(np.float32(np.sin(h)), np.float32(np.cos(h)), p) + translation


def transform_apply_inverse(translation, rotation, p):
    h = np.float32(np.float32(-np.float32(rotation)) * np.float32(0.5))
    return _rotate_z(np.float32(np.sin(h)), np.float32(np.cos(h)),
                     np.asarray(p, dtype=np.float32) - translation)


class Island:
    """island.rs:16-69."""

    def __init__(self, transform: Transform, nav_mesh: ValidNavigationMesh):
        self.transform = transform
        self.nav_mesh = nav_mesh
        self.dirty = True

    def get_transform(self):
        return self.transform

    def set_transform(self, transform: Transform):
        self.transform = transform
        self.dirty = True

    def get_nav_mesh(self):
        return self.nav_mesh

    def set_nav_mesh(self, nav_mesh: ValidNavigationMesh):
        self.nav_mesh = nav_mesh
        self.dirty = True


class SetTypeIndexCostError(Exception):
    """nav_data.rs:1184-1190."""


@dataclass
class OffMeshLink:
    """nav_data.rs:78-117 (ids are island keys + polygon indices until flattening)."""

    src_island: Key
    src_polygon: int
    dst_island: Key
    dst_polygon: int
    dst_type: int
    portal: np.ndarray            # (2,3) world
    kind: int                     # _abi.LINK_*
    reverse: int = -1             # index into NavigationData.links
    dst_portal: np.ndarray | None = None
    cost: float = 0.0
    anim_kind: int = 0
    anim_id: int = 0
    anim_key: Key | None = None   # slot-map key of the AnimationLink (anim_id is its index)

    def identity(self):
        """What makes this the same link after a rebuild (the reference keeps the OffMeshLinkId of a
        link whose islands and animation link did not change: nav_data.rs:352-403)."""
        return (self.src_island, int(self.src_polygon), self.dst_island, int(self.dst_polygon), int(self.kind),
                self.anim_key, np.asarray(self.portal, dtype=np.float32).tobytes(),
                None if self.dst_portal is None else np.asarray(self.dst_portal, dtype=np.float32).tobytes())


@dataclass
class ModifiedNode:
    """nav_data.rs:119-132."""

    island: Key
    polygon: int
    new_boundary: list = field(default_factory=list)   # [(left, right)]
    new_vertices: list = field(default_factory=list)   # [(x, y)] world


class NavigationData:
    def __init__(self, cs=XYZ):
        self.cs = cs
        self.islands = DenseSlotMap()
        self.animation_links = DenseSlotMap()
        self.type_index_to_cost: dict[int, float] = {}
        self.dirty = False
        self.links: list[OffMeshLink] = []
        self.modified_nodes: list[ModifiedNode] = []
        self.generation = 0  # bumped whenever the flattened data must be re-uploaded
        # `changed_islands` of nav_data.rs:326-344 (deleted + dirty), accumulated until a consumer
        # (Archipelago._sync_navdata) takes them to decide which stored paths survive
        self._changed_islands: set = set()

    # --- mutation (all mark dirty) ---
    def add_island(self, island: Island) -> Key:
        self.dirty = True
        return self.islands.insert(island)

    def remove_island(self, island_id: Key):
        self.dirty = True
        if self.islands.remove(island_id) is None:
            raise KeyError("Island should be present in the Archipelago")
        self._changed_islands.add(island_id)

    def get_island(self, island_id: Key):
        return self.islands.get(island_id)

    def get_island_ids(self):
        return self.islands.keys()

    def add_animation_link(self, link) -> Key:
        self.dirty = True
        return self.animation_links.insert(link)

    def remove_animation_link(self, link_id: Key):
        self.dirty = True
        self.animation_links.remove(link_id)

    def get_animation_link(self, link_id: Key):
        return self.animation_links.get(link_id)

    def get_animation_link_ids(self):
        return self.animation_links.keys()

    def set_type_index_cost(self, type_index: int, cost: float):
        if cost <= 0.0:
            raise SetTypeIndexCostError(f"NonPositiveCost({cost})")
        self.type_index_to_cost[type_index] = float(cost)
        self.generation += 1

    def get_type_index_cost(self, type_index: int):
        return self.type_index_to_cost.get(type_index)

    # --- update (nav_data.rs:1158-1180) ---
    def update(self, edge_link_distance: float, animation_link_distance: float) -> bool:
        """Returns True when navigation data changed (device copy must be replaced)."""
        if not self.dirty and not any(i.dirty for i in self.islands.values()):
            return False
        self.dirty = False
        for key, island in self.islands.items():
            if island.dirty:
                self._changed_islands.add(key)
            island.dirty = False
        self.links, self.modified_nodes, self._regions = link_islands(self, edge_link_distance,
                                                                      animation_link_distance)
        self.generation += 1
        return True

    def take_changed_islands(self) -> set:
        """Keys of the islands deleted or dirty since the last call (`invalidated_islands`, lib.rs:274)."""
        out, self._changed_islands = self._changed_islands, set()
        return out

    # --- flattening ---
    def flatten(self) -> dict:
        cs = self.cs
        islands = self.islands.items()
        n_islands = len(islands)
        f = {}
        f["n_islands"] = n_islands
        trans = np.zeros((n_islands, 3), dtype=np.float32)
        rot = np.zeros(n_islands, dtype=np.float32)
        bounds = np.zeros((n_islands, 6), dtype=np.float32)
        poly_begin = np.zeros(n_islands + 1, dtype=np.uint32)
        vert_begin = np.zeros(n_islands + 1, dtype=np.uint32)
        has_height = np.zeros(n_islands, dtype=np.uint8)
        verts, pe_counts, edge_vertex, edge_nb, edge_rev = [], [], [], [], []
        ptype, pregion, pbounds, pcenter = [], [], [], []
        hpoly, hverts, htris = [], [], []
        any_height = False
        key_to_index = {}
        pb = vb = hvb = htb = 0
        for i, (key, island) in enumerate(islands):
            m = island.nav_mesh
            key_to_index[key] = i
            trans[i] = island.transform.landmass_translation(cs)
            rot[i] = island.transform.rotation
            bounds[i] = m.mesh_bounds if m.mesh_bounds is not None else (1, 1, 1, 0, 0, 0)
            poly_begin[i] = pb
            vert_begin[i] = vb
            verts.append(m.vertices)
            pe_counts.append(np.diff(m.poly_edge_begin.astype(np.int64)))
            edge_vertex.append(m.poly_verts.astype(np.int64) + vb)
            nb = m.edge_neighbor.astype(np.int64)
            edge_nb.append(np.where(nb == NONE, NONE, nb + pb))
            edge_rev.append(m.edge_reverse)
            ptype.append(m.poly_type)
            pregion.append(m.poly_region)
            pbounds.append(m.poly_bounds)
            pcenter.append(m.poly_center)
            if m.hpoly is not None:
                any_height = True
                has_height[i] = 1
                hp = m.hpoly.astype(np.int64).copy()
                hp[:, 0] += hvb
                hp[:, 2] += htb
                hpoly.append(hp)
                hverts.append(m.hverts)
                htris.append(m.htris)
                hvb += len(m.hverts)
                htb += len(m.htris)
            else:
                hpoly.append(np.zeros((m.n_polys, 4), dtype=np.int64))
            pb += m.n_polys
            vb += len(m.vertices)
        poly_begin[n_islands] = pb
        vert_begin[n_islands] = vb

        def cat(lst, dtype, shape_tail=()):
            if lst:
                return np.concatenate(lst).astype(dtype)
            return np.zeros((0,) + shape_tail, dtype=dtype)

        f["island_translation"] = trans
        f["island_rotation"] = rot
        f["island_bounds"] = bounds
        f["island_poly_begin"] = poly_begin
        f["island_vert_begin"] = vert_begin
        f["n_vertices"] = vb
        f["vertices"] = cat(verts, np.float32, (3,))
        f["n_polys"] = pb
        counts = cat(pe_counts, np.int64)
        peb = np.zeros(pb + 1, dtype=np.uint32)
        np.cumsum(counts, out=peb[1:])
        f["poly_edge_begin"] = peb
        f["n_edges"] = int(peb[-1])
        f["edge_vertex"] = cat(edge_vertex, np.uint32)
        f["edge_neighbor"] = cat(edge_nb, np.uint32)
        f["edge_reverse"] = cat(edge_rev, np.uint32)
        f["poly_type"] = cat(ptype, np.uint32)
        f["poly_region"] = cat(pregion, np.uint32)
        f["poly_bounds"] = cat(pbounds, np.float32, (6,))
        f["poly_center"] = cat(pcenter, np.float32, (3,))
        if any_height:
            f["island_has_height"] = has_height
            f["hpoly"] = cat(hpoly, np.uint32, (4,))
            f["n_hverts"] = hvb
            f["hverts"] = cat(hverts, np.float32, (3,))
            f["n_htris"] = htb
            f["htris"] = cat(htris, np.uint8, (3,))
        costs = sorted(self.type_index_to_cost.items())
        f["n_type_costs"] = len(costs)
        f["type_cost_index"] = np.array([c[0] for c in costs], dtype=np.uint32)
        f["type_cost_value"] = np.array([c[1] for c in costs], dtype=np.float32)

        # off-mesh links
        links = self.links
        nl = len(links)
        f["n_links"] = nl

        def node_id(island_key, polygon):
            return int(poly_begin[key_to_index[island_key]]) + int(polygon)

        src = np.array([node_id(l.src_island, l.src_polygon) for l in links], dtype=np.uint32)
        f["link_src_node"] = src
        f["link_dst_node"] = np.array([node_id(l.dst_island, l.dst_polygon) for l in links], dtype=np.uint32)
        f["link_dst_type"] = np.array([l.dst_type for l in links], dtype=np.uint32)
        f["link_portal"] = (np.stack([l.portal for l in links]).astype(np.float32).reshape(nl, 6)
                            if nl else np.zeros((0, 6), dtype=np.float32))
        f["link_kind"] = np.array([l.kind for l in links], dtype=np.uint32)
        f["link_reverse"] = np.array([l.reverse if l.reverse >= 0 else NONE for l in links], dtype=np.uint32)
        f["link_dst_portal"] = (np.stack([l.dst_portal if l.dst_portal is not None else l.portal
                                          for l in links]).astype(np.float32).reshape(nl, 6)
                                if nl else np.zeros((0, 6), dtype=np.float32))
        f["link_cost"] = np.array([l.cost for l in links], dtype=np.float32)
        f["link_anim_kind"] = np.array([l.anim_kind for l in links], dtype=np.uint32)
        f["link_anim_id"] = np.array([l.anim_id for l in links], dtype=np.uint32)
        nlb = np.zeros(pb + 1, dtype=np.uint32)
        if nl:
            order = np.argsort(src, kind="stable")  # ascending link index within a node
            np.cumsum(np.bincount(src, minlength=pb), out=nlb[1:])
            f["node_links"] = order.astype(np.uint32)
        else:
            f["node_links"] = np.zeros(0, dtype=np.uint32)
        f["node_link_begin"] = nlb

        # modified nodes, ascending node id
        mods = sorted(self.modified_nodes, key=lambda m: node_id(m.island, m.polygon))
        f["n_modified"] = len(mods)
        f["modified_node"] = np.array([node_id(m.island, m.polygon) for m in mods], dtype=np.uint32)
        meb = np.zeros(len(mods) + 1, dtype=np.uint32)
        mvb = np.zeros(len(mods) + 1, dtype=np.uint32)
        medges, mverts = [], []
        for i, m in enumerate(mods):
            medges.extend(m.new_boundary)
            mverts.extend(m.new_vertices)
            meb[i + 1] = len(medges)
            mvb[i + 1] = len(mverts)
        f["modified_edge_begin"] = meb
        f["modified_edges"] = np.array(medges, dtype=np.uint32).reshape(-1, 2)
        f["modified_vert_begin"] = mvb
        f["modified_verts"] = np.array(mverts, dtype=np.float32).reshape(-1, 2)

        # regions (nav_data.rs:1089-1156)
        regions = getattr(self, "_regions", None)
        nrn = np.full(pb, NONE, dtype=np.uint32)
        if regions is not None and regions["number"]:
            for (island_key, region), number in regions["number"].items():
                i = key_to_index[island_key]
                sl = slice(int(poly_begin[i]), int(poly_begin[i + 1]))
                nrn[sl] = np.where(f["poly_region"][sl] == region, number, nrn[sl])
            f["n_region_numbers"] = len(regions["root"])
            f["region_root"] = np.array(regions["root"], dtype=np.uint32)
            pl = regions["possible"]
            f["n_possible_links"] = len(pl)
            f["possible_link_src"] = np.array([p[0] for p in pl], dtype=np.uint32)
            f["possible_link_dst"] = np.array([p[1] for p in pl], dtype=np.uint32)
            f["possible_link_kind"] = np.array([p[2] for p in pl], dtype=np.uint32)
        else:
            f["n_region_numbers"] = 0
            f["region_root"] = np.zeros(0, dtype=np.uint32)
            f["n_possible_links"] = 0
            f["possible_link_src"] = np.zeros(0, dtype=np.uint32)
            f["possible_link_dst"] = np.zeros(0, dtype=np.uint32)
            f["possible_link_kind"] = np.zeros(0, dtype=np.uint32)
        f["node_region_number"] = nrn
        f["_island_key_to_index"] = key_to_index
        f["_island_keys"] = [key for key, _ in islands]
        f["_link_identity"] = [l.identity() for l in links]
        return f
