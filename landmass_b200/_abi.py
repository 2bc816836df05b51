"""ctypes mirror of include/landmass_b200.h (the C ABI of the agent step).

Only plain structs, constants and array marshalling live here; no compute.
The same marshalling is reused by the test-only oracle binding so both sides
are fed byte-identical inputs.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

ABI_VERSION = 4
NONE = 0xFFFFFFFF
PORTAL_LINK = 0x80000000

# lmb_status
OK = 0
ERR_INVALID_ARGUMENT = -1
ERR_CUDA = -2
ERR_NO_NAVDATA = -3
ERR_CAPACITY = -4
ERR_UNSUPPORTED = -5
ERR_INTERNAL = -6

# agent flags
AGENT_HAS_TARGET = 1
AGENT_PAUSED = 2
AGENT_USING_ANIMATION_LINK = 4
AGENT_RESET = 8
AGENT_PERMIT_ALL_LINKS = 16
AGENT_KEEP_AVOIDANCE_DATA = 32

REACH_DISTANCE = 0
REACH_VISIBLE_AT_DISTANCE = 1
REACH_STRAIGHT_PATH_DISTANCE = 2

LINK_BOUNDARY = 0
LINK_ANIMATION = 1

CONFIG_NO_AVOIDANCE = 1
CONFIG_DEBUG_LOG = 2
CONFIG_ASTAR_COMPACT = 4
CONFIG_ASTAR_DENSE = 8
CONFIG_ASTAR_HASHED = 16
CONFIG_ASTAR_LARGE_SHAPES = 32
CONFIG_ASTAR_THREAD_ONLY = 64
CONFIG_ASTAR_WARP_ONLY = 128
CONFIG_ASTAR_NO_FOREST = 256
CONFIG_NO_L2_WINDOW = 512
CONFIG_ASTAR_SMALL_HEAP = 1024
CONFIG_ASTAR_NO_PAIR = 2048
CONFIG_ORDER_ALWAYS = 4096

f32p = C.POINTER(C.c_float)
u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)


class Options(C.Structure):
    _fields_ = [
        ("horizontal_distance", C.c_float),
        ("distance_above", C.c_float),
        ("distance_below", C.c_float),
        ("vertical_preference_ratio", C.c_float),
        ("neighbourhood", C.c_float),
        ("avoidance_time_horizon", C.c_float),
        ("obstacle_avoidance_time_horizon", C.c_float),
        ("reached_destination_avoidance_responsibility", C.c_float),
    ]


class NavDataDesc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("n_islands", C.c_uint32),
        ("island_translation", f32p),
        ("island_rotation", f32p),
        ("island_bounds", f32p),
        ("island_poly_begin", u32p),
        ("island_vert_begin", u32p),
        ("n_vertices", C.c_uint32),
        ("vertices", f32p),
        ("n_polys", C.c_uint32),
        ("poly_edge_begin", u32p),
        ("n_edges", C.c_uint32),
        ("edge_vertex", u32p),
        ("edge_neighbor", u32p),
        ("edge_reverse", u32p),
        ("poly_type", u32p),
        ("poly_region", u32p),
        ("poly_bounds", f32p),
        ("poly_center", f32p),
        ("island_has_height", u8p),
        ("hpoly", u32p),
        ("n_hverts", C.c_uint32),
        ("hverts", f32p),
        ("n_htris", C.c_uint32),
        ("htris", u8p),
        ("n_type_costs", C.c_uint32),
        ("type_cost_index", u32p),
        ("type_cost_value", f32p),
        ("n_links", C.c_uint32),
        ("link_src_node", u32p),
        ("link_dst_node", u32p),
        ("link_dst_type", u32p),
        ("link_portal", f32p),
        ("link_kind", u32p),
        ("link_reverse", u32p),
        ("link_dst_portal", f32p),
        ("link_cost", f32p),
        ("link_anim_kind", u32p),
        ("link_anim_id", u32p),
        ("node_link_begin", u32p),
        ("node_links", u32p),
        ("n_modified", C.c_uint32),
        ("modified_node", u32p),
        ("modified_edge_begin", u32p),
        ("modified_edges", u32p),
        ("modified_vert_begin", u32p),
        ("modified_verts", f32p),
        ("node_region_number", u32p),
        ("n_region_numbers", C.c_uint32),
        ("region_root", u32p),
        ("n_possible_links", C.c_uint32),
        ("possible_link_src", u32p),
        ("possible_link_dst", u32p),
        ("possible_link_kind", u32p),
    ]


class AgentsIn(C.Structure):
    _fields_ = [
        ("n", C.c_uint32),
        ("slot", u32p),
        ("flags", u32p),
        ("position", f32p),
        ("velocity", f32p),
        ("target", f32p),
        ("radius", f32p),
        ("desired_speed", f32p),
        ("max_speed", f32p),
        ("reach_condition", u32p),
        ("reach_distance", f32p),
        ("anim_link_reach_distance", f32p),
        ("override_begin", u32p),
        ("override_type", u32p),
        ("override_cost", f32p),
        ("permitted_begin", u32p),
        ("permitted_kind", u32p),
    ]


class CharactersIn(C.Structure):
    _fields_ = [
        ("n", C.c_uint32),
        ("position", f32p),
        ("velocity", f32p),
        ("radius", f32p),
    ]


class AgentsOut(C.Structure):
    _fields_ = [
        ("desired_velocity", f32p),
        ("state", u32p),
        ("reached_link_id", u32p),
        ("reached_link_start", f32p),
        ("reached_link_end", f32p),
    ]


class NavRemap(C.Structure):
    """lmb_nav_remap: the previous upload's islands / links -> the new flattening (NONE = invalidated)."""

    _fields_ = [
        ("struct_size", C.c_uint32),
        ("n_old_islands", C.c_uint32),
        ("island_map", u32p),
        ("n_old_links", C.c_uint32),
        ("link_map", u32p),
    ]


def make_nav_remap(remap: dict):
    """remap: {"island_map": uint32[n_old_islands], "link_map": uint32[n_old_links]} -> (NavRemap, keepalive)."""
    im, lm = as_u32(remap["island_map"]), as_u32(remap["link_map"])
    r = NavRemap()
    r.struct_size = C.sizeof(NavRemap)
    r.n_old_islands = len(im)
    r.island_map = im.ctypes.data_as(u32p)
    r.n_old_links = len(lm)
    r.link_map = lm.ctypes.data_as(u32p)
    return r, (im, lm)


class PathingResult(C.Structure):
    _fields_ = [
        ("agent", C.c_uint32),
        ("success", C.c_uint32),
        ("explored_nodes", C.c_uint32),
    ]


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("device", C.c_int32),
        ("max_agents", C.c_uint32),
        ("astar_slots", C.c_uint32),
        ("scratch_bytes", C.c_uint64),
        ("flags", C.c_uint32),
        ("astar_shape", C.c_uint32),
    ]


class StepTimings(C.Structure):
    _fields_ = [
        ("h2d_ms", C.c_float),
        ("sample_ms", C.c_float),
        ("repath_decide_ms", C.c_float),
        ("astar_ms", C.c_float),
        ("follow_ms", C.c_float),
        ("avoidance_ms", C.c_float),
        ("d2h_ms", C.c_float),
        ("total_ms", C.c_float),
        ("n_repath", C.c_uint32),
        ("kernel_launches", C.c_uint32),
        ("explored_nodes", C.c_uint64),
        ("path_nodes", C.c_uint64),
        ("allgather_ms", C.c_float),
        ("astar_launch_kind", C.c_uint32),
    ]


def as_f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


def as_u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def ptr(a, ty):
    """Pointer to a contiguous numpy array (or NULL for None)."""
    if a is None:
        return C.cast(None, ty)
    return a.ctypes.data_as(ty)


_NAV_F32 = {
    "island_translation", "island_rotation", "island_bounds", "vertices", "poly_bounds",
    "poly_center", "hverts", "type_cost_value", "link_portal", "link_dst_portal", "link_cost",
    "modified_verts",
}
_NAV_U8 = {"island_has_height", "htris"}


def make_navdata_desc(flat: dict):
    """Builds a NavDataDesc from a dict of arrays (see nav_data.FlatNavData).

    Returns (desc, keepalive) — keepalive holds the numpy arrays the pointers
    refer to and must outlive the call that consumes `desc`.
    """
    desc = NavDataDesc()
    keep = {}
    desc.struct_size = C.sizeof(NavDataDesc)
    for name, ty in NavDataDesc._fields_:
        if name == "struct_size":
            continue
        if ty is C.c_uint32:
            setattr(desc, name, int(flat.get(name, 0)))
            continue
        arr = flat.get(name)
        if arr is None:
            setattr(desc, name, C.cast(None, ty))
            continue
        if name in _NAV_F32:
            arr = np.ascontiguousarray(arr, dtype=np.float32)
        elif name in _NAV_U8:
            arr = np.ascontiguousarray(arr, dtype=np.uint8)
        else:
            arr = np.ascontiguousarray(arr, dtype=np.uint32)
        if arr.size == 0:
            # keep a 1-element dummy so the pointer is valid
            arr = np.zeros(1, dtype=arr.dtype)
        keep[name] = arr
        setattr(desc, name, arr.ctypes.data_as(ty))
    return desc, keep


def make_options(opts) -> Options:
    o = Options()
    for name, _ in Options._fields_:
        setattr(o, name, float(getattr(opts, name) if not isinstance(opts, dict) else opts[name]))
    return o


_AG_F32 = {"position", "velocity", "target", "radius", "desired_speed", "max_speed",
           "reach_distance", "anim_link_reach_distance", "override_cost"}


def make_agents_in(ag: dict):
    s = AgentsIn()
    keep = {}
    s.n = int(ag["n"])
    for name, ty in AgentsIn._fields_:
        if name == "n":
            continue
        arr = ag.get(name)
        if arr is None:
            setattr(s, name, C.cast(None, ty))
            continue
        arr = np.ascontiguousarray(arr, dtype=np.float32 if name in _AG_F32 else np.uint32)
        if arr.size == 0:
            arr = np.zeros(1, dtype=arr.dtype)
        keep[name] = arr
        setattr(s, name, arr.ctypes.data_as(ty))
    return s, keep


def make_characters_in(ch: dict | None):
    s = CharactersIn()
    keep = {}
    if ch is None:
        s.n = 0
        return s, keep
    s.n = int(ch["n"])
    for name in ("position", "velocity", "radius"):
        arr = np.ascontiguousarray(ch[name], dtype=np.float32)
        if arr.size == 0:
            arr = np.zeros(1, dtype=np.float32)
        keep[name] = arr
        setattr(s, name, arr.ctypes.data_as(f32p))
    return s, keep


def make_agents_out(n: int):
    out = {
        "desired_velocity": np.zeros((n, 3), dtype=np.float32),
        "state": np.zeros(n, dtype=np.uint32),
        "reached_link_id": np.full(n, NONE, dtype=np.uint32),
        "reached_link_start": np.zeros((n, 3), dtype=np.float32),
        "reached_link_end": np.zeros((n, 3), dtype=np.float32),
    }
    s = AgentsOut()
    # pointers must be valid even for n == 0
    keep = {k: (v if v.size else np.zeros(1, dtype=v.dtype)) for k, v in out.items()}
    s.desired_velocity = keep["desired_velocity"].ctypes.data_as(f32p)
    s.state = keep["state"].ctypes.data_as(u32p)
    s.reached_link_id = keep["reached_link_id"].ctypes.data_as(u32p)
    s.reached_link_start = keep["reached_link_start"].ctypes.data_as(f32p)
    s.reached_link_end = keep["reached_link_end"].ctypes.data_as(f32p)
    if n:
        out = keep
    return s, out
