"""TEST INFRASTRUCTURE — ctypes binding of the CPU oracle (oracle/liblandmass_oracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  It exposes the same Python-facing backend interface as
landmass_b200._native.NativeBackend so identical inputs can be fed to both.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from landmass_b200 import _abi
from landmass_b200._native import CBackend, declare_prototypes

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_DIR, "liblandmass_oracle.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(_DIR, f) for f in ("landmass_oracle.cpp", "avoidance_oracle.cpp",
                                            "landmass_oracle.h", "orc_internal.h", "orc_math.h")]
    srcs.append(os.path.join(_DIR, "..", "include", "landmass_b200.h"))
    if (not force and os.path.exists(_LIB)
            and all(os.path.getmtime(_LIB) >= os.path.getmtime(s) for s in srcs)):
        return _LIB
    subprocess.run(["make", "-C", _DIR, "-B", "liblandmass_oracle.so"], check=True,
                   stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        lib = C.CDLL(_LIB)
        declare_prototypes(lib, "orc_")
        lib.orc_create.restype = C.c_void_p
        lib.orc_create.argtypes = [C.POINTER(_abi.NavDataDesc), C.c_uint32]
        lib.orc_destroy.restype = None
        lib.orc_destroy.argtypes = [C.c_void_p]
        lib.orc_step_begin.restype = C.c_int32
        lib.orc_step_begin.argtypes = [C.c_void_p, C.POINTER(_abi.AgentsIn), C.POINTER(_abi.CharactersIn),
                                       C.POINTER(_abi.Options), C.c_float, C.c_uint32, C.c_uint32]
        lib.orc_halo_buffer.restype = C.c_int32
        lib.orc_halo_buffer.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64),
                                        C.POINTER(C.c_uint64)]
        lib.orc_step_end.restype = C.c_int32
        lib.orc_step_end.argtypes = [C.c_void_p, C.POINTER(_abi.Options), C.c_float, C.POINTER(_abi.AgentsOut)]
        lib.orc_astar_adjacency.restype = C.c_int32
        lib.orc_next_straight_point.restype = C.c_int32
        lib.orc_borders_to_obstacles.restype = C.c_int32
        lib.orc_compute_avoiding_velocity.restype = C.c_int32
        _lib = lib
    return _lib


class OracleBackend(CBackend):
    prefix = "orc_"

    def __init__(self, flags: int = 0, **_ignored):
        super().__init__()
        self.lib = load()
        self.flags = flags
        self.ctx = None

    def navdata_upload(self, flat: dict, remap: dict | None = None):
        desc, keep = _abi.make_navdata_desc(flat)
        if self.ctx is None:
            self.ctx = C.c_void_p(self.lib.orc_create(C.byref(desc), self.flags))
            if not self.ctx:
                raise RuntimeError("orc_create failed")
        elif remap is None:
            self._check(self.lib.orc_navdata_upload(self.ctx, C.byref(desc)))
        else:
            r, keep_r = _abi.make_nav_remap(remap)
            self._check(self.lib.orc_navdata_update(self.ctx, C.byref(desc), C.byref(r)))
            del keep_r
        del keep

    # ---- multi-rank form (same contract as NativeBackend.step_begin / halo_buffer / step_end) ----
    def step_begin_host(self, ag: dict, ch, options, dt: float, n_total: int, first: int):
        a, self._ka = _abi.make_agents_in(ag)
        c, self._kc = _abi.make_characters_in(ch)
        self._n_local = a.n
        self._check(self.lib.orc_step_begin(self.ctx, C.byref(a), C.byref(c), C.byref(options),
                                            C.c_float(dt), C.c_uint32(n_total), C.c_uint32(first)))

    def halo_numpy(self):
        """The halo buffer as a writable (n_total, 8) float32 view of the oracle's host memory."""
        p, rb, nr = C.c_void_p(), C.c_uint64(), C.c_uint64()
        self._check(self.lib.orc_halo_buffer(self.ctx, C.byref(p), C.byref(rb), C.byref(nr)))
        buf = (C.c_float * (nr.value * rb.value // 4)).from_address(p.value)
        return np.ctypeslib.as_array(buf).reshape(nr.value, rb.value // 4)

    def step_end_host(self, options, dt: float) -> dict:
        o, out = _abi.make_agents_out(self._n_local)
        self._check(self.lib.orc_step_end(self.ctx, C.byref(options), C.c_float(dt), C.byref(o)))
        return out

    def next_straight_point(self, nodes, portals, start_index, start_point, end_index, end_point):
        nodes = _abi.as_u32(nodes)
        portals = _abi.as_u32(portals)
        sp = _abi.as_f32(start_point)
        ep = _abi.as_f32(end_point)
        oi = C.c_uint32()
        ok = C.c_uint32()
        op = np.zeros(6, dtype=np.float32)
        self._check(self.lib.orc_next_straight_point(
            self.ctx, C.c_uint32(len(nodes)), _abi.ptr(nodes, _abi.u32p), _abi.ptr(portals, _abi.u32p),
            C.c_uint32(start_index), _abi.ptr(sp, _abi.f32p), C.c_uint32(end_index),
            _abi.ptr(ep, _abi.f32p), C.byref(oi), C.byref(ok), _abi.ptr(op, _abi.f32p)))
        return oi.value, ok.value, op

    def path_find_index_of_node(self, nodes, portals, node, reverse=False):
        """Path::find_index_of_node[_rev] (path.rs:402-446) on a hand-made corridor: flat index or None."""
        nodes = _abi.as_u32(nodes)
        portals = _abi.as_u32(portals)
        found, idx = C.c_uint32(), C.c_uint32()
        self.lib.orc_path_find_index_of_node.restype = C.c_int32
        self._check(self.lib.orc_path_find_index_of_node(
            self.ctx, C.c_uint32(len(nodes)), _abi.ptr(nodes, _abi.u32p), _abi.ptr(portals, _abi.u32p),
            C.c_uint32(node), C.c_uint32(1 if reverse else 0), C.byref(found), C.byref(idx)))
        return idx.value if found.value else None

    @staticmethod
    def project_point_to_line_segment(point, start, end):
        """geometry.rs:176-189 -> (point (3,), fraction)."""
        lib = load()
        lib.orc_project_point_to_line_segment.restype = C.c_int32
        p, a, b = _abi.as_f32(point), _abi.as_f32(start), _abi.as_f32(end)
        out, frac = np.zeros(3, dtype=np.float32), C.c_float()
        lib.orc_project_point_to_line_segment(_abi.ptr(p, _abi.f32p), _abi.ptr(a, _abi.f32p), _abi.ptr(b, _abi.f32p),
                                              _abi.ptr(out, _abi.f32p), C.byref(frac))
        return out, frac.value

    def borders_to_obstacles(self, point, node, distance_limit):
        p = _abi.as_f32(point)
        cap_o, cap_v = 256, 8192
        n = C.c_uint32()
        vb = np.zeros(cap_o + 1, dtype=np.uint32)
        closed = np.zeros(cap_o, dtype=np.uint32)
        verts = np.zeros((cap_v, 2), dtype=np.float32)
        self._check(self.lib.orc_borders_to_obstacles(
            self.ctx, _abi.ptr(p, _abi.f32p), C.c_uint32(node), C.c_float(distance_limit), C.byref(n),
            _abi.ptr(vb, _abi.u32p), _abi.ptr(closed, _abi.u32p), _abi.ptr(verts, _abi.f32p),
            C.c_uint32(cap_o), C.c_uint32(cap_v)))
        return [(bool(closed[i]), verts[vb[i]:vb[i + 1]].copy()) for i in range(n.value)]

    def close(self):
        if self.ctx:
            self.lib.orc_destroy(self.ctx)
            self.ctx = None


def astar_adjacency(states, start, end):
    """states: list of (adjacency [(cost, action, state)], heuristic). astar_test.rs harness."""
    lib = load()
    begin = np.zeros(len(states) + 1, dtype=np.uint32)
    cost, action, state, h = [], [], [], []
    for i, (adj, heur) in enumerate(states):
        for c, a, s in adj:
            cost.append(c)
            action.append(a)
            state.append(s)
        begin[i + 1] = len(cost)
        h.append(heur)
    cost = np.array(cost + [0], dtype=np.float32)
    action = np.array(action + [0], dtype=np.int32)
    state = np.array(state + [0], dtype=np.uint32)
    h = np.array(h + [0], dtype=np.float32)
    out = np.zeros(4096, dtype=np.int32)
    ln, found, expl = C.c_uint32(), C.c_uint32(), C.c_uint32()
    st = lib.orc_astar_adjacency(
        C.c_uint32(len(states)), _abi.ptr(begin, _abi.u32p), _abi.ptr(cost, _abi.f32p),
        action.ctypes.data_as(C.POINTER(C.c_int32)), _abi.ptr(state, _abi.u32p),
        _abi.ptr(h, _abi.f32p), C.c_uint32(start), C.c_uint32(end),
        out.ctypes.data_as(C.POINTER(C.c_int32)), C.c_uint32(len(out)), C.byref(ln), C.byref(found),
        C.byref(expl))
    assert st == 0
    return (list(out[:ln.value]) if found.value else None), expl.value


def compute_avoiding_velocity(agent, neighbours, obstacles, preferred, max_speed, time_step,
                              obstacle_margin, time_horizon, obstacle_time_horizon):
    """agent/neighbours: (px,py,vx,vy,radius,responsibility); obstacles: [(closed, [(x,y)...])]."""
    lib = load()
    a = _abi.as_f32(agent)
    nb = _abi.as_f32(np.array(neighbours, dtype=np.float32).reshape(-1, 6))
    vb = np.zeros(len(obstacles) + 1, dtype=np.uint32)
    closed = np.zeros(max(len(obstacles), 1), dtype=np.uint32)
    verts = []
    for i, (c, vs) in enumerate(obstacles):
        closed[i] = 1 if c else 0
        verts.extend(vs)
        vb[i + 1] = len(verts)
    verts = np.array(verts + [(0, 0)], dtype=np.float32)
    pref = _abi.as_f32(preferred)
    out = np.zeros(2, dtype=np.float32)
    if nb.size == 0:
        nb = np.zeros((1, 6), dtype=np.float32)
        n_nb = 0
    else:
        n_nb = len(nb)
    st = lib.orc_compute_avoiding_velocity(
        _abi.ptr(a, _abi.f32p), C.c_uint32(n_nb), _abi.ptr(nb, _abi.f32p), C.c_uint32(len(obstacles)),
        _abi.ptr(vb, _abi.u32p), _abi.ptr(closed, _abi.u32p), _abi.ptr(verts, _abi.f32p),
        _abi.ptr(pref, _abi.f32p), C.c_float(max_speed), C.c_float(time_step), C.c_float(obstacle_margin),
        C.c_float(time_horizon), C.c_float(obstacle_time_horizon), _abi.ptr(out, _abi.f32p))
    assert st == 0
    return out
