"""ctypes binding of the CUDA library (landmass_b200/csrc -> liblandmass_b200.so).

This is the ONLY compute backend of the package.  If the shared library is
missing or cannot be loaded (e.g. no CUDA driver), construction raises — there
is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _abi

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "liblandmass_b200.so")

EXPORTED_SYMBOLS = [
    "lmb_abi_version", "lmb_create", "lmb_destroy", "lmb_last_error", "lmb_navdata_upload",
    "lmb_navdata_update", "lmb_agent_path_endpoints", "lmb_agent_waypoints",
    "lmb_step", "lmb_agents_upload", "lmb_step_resident", "lmb_agents_download",
    "lmb_pathing_results", "lmb_agent_path", "lmb_sample_points", "lmb_find_paths",
    "lmb_get_step_timings", "lmb_step_begin", "lmb_halo_buffer", "lmb_step_end",
    "lmb_find_straight_paths", "lmb_agent_avoidance_constraints", "lmb_get_query_timings",
    "lmb_comm_unique_id", "lmb_comm_init", "lmb_step_sharded", "lmb_agents_integrate", "lmb_agents_positions",
]


class NativeError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"landmass_b200 native error {status}: {message}")
        self.status = status


def load_library(path: str | None = None) -> C.CDLL:
    path = path or _LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError(
            f"landmass_b200: CUDA library not built ({path}); run `python -c 'import "
            "__graft_entry__ as g; g.build()'` — there is no CPU fallback")
    lib = C.CDLL(path)
    declare_prototypes(lib, "lmb_")
    return lib


def declare_prototypes(lib, prefix: str):
    """Declares argument/return types shared by the lmb_* and (test-only) orc_* entry points."""
    u32, i32, f32 = C.c_uint32, C.c_int32, C.c_float
    vp = C.c_void_p

    def fn(name, restype, argtypes):
        f = getattr(lib, prefix + name, None)
        if f is not None:
            f.restype = restype
            f.argtypes = argtypes

    fn("navdata_upload", i32, [vp, C.POINTER(_abi.NavDataDesc)])
    fn("navdata_update", i32, [vp, C.POINTER(_abi.NavDataDesc), C.POINTER(_abi.NavRemap)])
    fn("step", i32, [vp, C.POINTER(_abi.AgentsIn), C.POINTER(_abi.CharactersIn),
                     C.POINTER(_abi.Options), f32, C.POINTER(_abi.AgentsOut)])
    fn("agents_upload", i32, [vp, C.POINTER(_abi.AgentsIn), C.POINTER(_abi.CharactersIn)])
    fn("step_resident", i32, [vp, C.POINTER(_abi.Options), f32])
    fn("agents_download", i32, [vp, C.POINTER(_abi.AgentsOut)])
    fn("pathing_results", i32, [vp, C.POINTER(_abi.PathingResult), u32, C.POINTER(u32)])
    fn("agent_path", i32, [vp, u32, _abi.u32p, _abi.u32p, u32, C.POINTER(u32)])
    fn("sample_points", i32, [vp, u32, _abi.f32p, C.POINTER(_abi.Options), _abi.f32p, _abi.u32p])
    fn("find_paths", i32, [vp, u32, _abi.u32p, _abi.f32p, _abi.u32p, _abi.f32p, _abi.u32p,
                           _abi.u32p, _abi.f32p, _abi.u32p, _abi.u32p, _abi.u32p, _abi.u32p,
                           _abi.u32p, u32, _abi.u32p, _abi.u32p, _abi.u32p])
    fn("find_straight_paths", i32, [vp, u32, _abi.u32p, _abi.f32p, _abi.u32p, _abi.f32p, _abi.u32p,
                                    _abi.u32p, _abi.f32p, _abi.u32p, _abi.u32p, _abi.u32p, _abi.f32p,
                                    _abi.u32p, u32, _abi.u32p, _abi.u32p, _abi.u32p])
    fn("agent_path_endpoints", i32, [vp, u32, _abi.f32p, _abi.f32p])
    fn("agent_waypoints", i32, [vp, u32, _abi.u32p, _abi.f32p, _abi.u32p, C.POINTER(u32)])
    fn("agent_avoidance_constraints", i32, [vp, u32, u32, _abi.f32p, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)])
    fn("get_step_timings", i32, [vp, C.POINTER(_abi.StepTimings)])
    fn("step_begin", i32, [vp, C.POINTER(_abi.Options), f32, u32, u32])
    fn("halo_buffer", i32, [vp, C.POINTER(vp), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)])
    fn("step_end", i32, [vp, C.POINTER(_abi.Options), f32])
    fn("get_query_timings", i32, [vp, C.POINTER(_abi.StepTimings)])
    fn("comm_unique_id", i32, [vp, C.POINTER(C.c_uint8)])
    fn("comm_init", i32, [vp, C.POINTER(C.c_uint8), u32, u32])
    fn("step_sharded", i32, [vp, C.POINTER(_abi.Options), f32, u32])
    fn("agents_integrate", i32, [vp, f32])
    fn("agents_positions", i32, [vp, _abi.f32p, u32])
    if prefix == "lmb_":
        lib.lmb_abi_version.restype = u32
        lib.lmb_abi_version.argtypes = []
        lib.lmb_create.restype = i32
        lib.lmb_create.argtypes = [C.POINTER(_abi.Config), C.POINTER(vp)]
        lib.lmb_destroy.restype = None
        lib.lmb_destroy.argtypes = [vp]
        lib.lmb_last_error.restype = C.c_char_p
        lib.lmb_last_error.argtypes = [vp]


class CBackend:
    """Calls shared by the CUDA library and the test-only oracle (same signatures)."""

    prefix = "lmb_"

    def __init__(self):
        self.lib = None
        self.ctx = None

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    def _check(self, status: int):
        if status != 0:
            msg = ""
            if self.prefix == "lmb_":
                raw = self.lib.lmb_last_error(self.ctx)
                msg = raw.decode() if raw else ""
            raise NativeError(status, msg)

    def navdata_upload(self, flat: dict, remap: dict | None = None):
        """remap (see _abi.make_nav_remap) keeps the stored paths the reference's `Path::is_valid`
        keeps; without it every path is dropped."""
        desc, keep = _abi.make_navdata_desc(flat)
        if remap is None:
            self._check(self._f("navdata_upload")(self.ctx, C.byref(desc)))
        else:
            r, keep_r = _abi.make_nav_remap(remap)
            self._check(self._f("navdata_update")(self.ctx, C.byref(desc), C.byref(r)))
            del keep_r
        self._n_polys = int(flat["n_polys"])
        del keep

    def step(self, ag: dict, ch: dict | None, options: _abi.Options, dt: float) -> dict:
        a, ka = _abi.make_agents_in(ag)
        c, kc = _abi.make_characters_in(ch)
        o, out = _abi.make_agents_out(a.n)
        self._check(self._f("step")(self.ctx, C.byref(a), C.byref(c), C.byref(options),
                                    C.c_float(dt), C.byref(o)))
        del ka, kc
        return out

    def pathing_results(self):
        cap = 1024
        while True:
            buf = (_abi.PathingResult * cap)()
            cnt = C.c_uint32(0)
            st = self._f("pathing_results")(self.ctx, buf, cap, C.byref(cnt))
            if st == _abi.ERR_CAPACITY:
                cap = max(cnt.value, cap * 2)
                continue
            self._check(st)
            return [(buf[i].agent, buf[i].success, buf[i].explored_nodes) for i in range(cnt.value)]

    def pathing_results_array(self):
        """Same as pathing_results but as an (n,3) uint32 array (for large batches)."""
        cap = 1 << 16
        while True:
            buf = np.zeros((cap, 3), dtype=np.uint32)
            cnt = C.c_uint32(0)
            st = self._f("pathing_results")(
                self.ctx, C.cast(buf.ctypes.data, C.POINTER(_abi.PathingResult)), cap, C.byref(cnt))
            if st == _abi.ERR_CAPACITY:
                cap = max(cnt.value, cap * 2)
                continue
            self._check(st)
            return buf[:cnt.value]

    def agent_path(self, slot: int):
        cap = 256
        while True:
            nodes = np.zeros(cap, dtype=np.uint32)
            portals = np.zeros(cap, dtype=np.uint32)
            ln = C.c_uint32(0)
            st = self._f("agent_path")(self.ctx, slot, _abi.ptr(nodes, _abi.u32p),
                                       _abi.ptr(portals, _abi.u32p), cap, C.byref(ln))
            if st == _abi.ERR_CAPACITY:
                cap = max(ln.value, cap * 2)
                continue
            self._check(st)
            return nodes[:ln.value], portals[:ln.value]

    def agent_path_endpoints(self, slot: int):
        """(Path.start_point, Path.end_point) of the path stored for `slot` (see lmb_agent_path_endpoints)."""
        a, b = np.zeros(3, dtype=np.float32), np.zeros(3, dtype=np.float32)
        self._check(self._f("agent_path_endpoints")(self.ctx, slot, _abi.ptr(a, _abi.f32p), _abi.ptr(b, _abi.f32p)))
        return a, b

    def agent_waypoints(self):
        """Per agent of the last step: (kind, points (n,6), link) — see lmb_agent_waypoints."""
        cap = 1024
        while True:
            kind = np.zeros(cap, dtype=np.uint32)
            pts = np.zeros((cap, 6), dtype=np.float32)
            link = np.zeros(cap, dtype=np.uint32)
            cnt = C.c_uint32(0)
            st = self._f("agent_waypoints")(self.ctx, cap, _abi.ptr(kind, _abi.u32p), _abi.ptr(pts, _abi.f32p),
                                            _abi.ptr(link, _abi.u32p), C.byref(cnt))
            if st == _abi.ERR_CAPACITY:
                cap = cnt.value
                continue
            self._check(st)
            return kind[:cnt.value], pts[:cnt.value], link[:cnt.value]

    def agent_avoidance_constraints(self, agent_index: int):
        """(ran, original (n,4), fallback (m,4), fell_back) of the agent at `agent_index` of the last step; lines
        are {point.xy, direction.xy} — see lmb_agent_avoidance_constraints."""
        cap = 64
        while True:
            lines = np.zeros((cap, 4), dtype=np.float32)
            n_o, n_f, ran = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
            st = self._f("agent_avoidance_constraints")(self.ctx, agent_index, cap, _abi.ptr(lines, _abi.f32p),
                                                        C.byref(n_o), C.byref(n_f), C.byref(ran))
            if st == _abi.ERR_CAPACITY:
                cap = n_o.value + n_f.value
                continue
            self._check(st)
            return (ran.value != 0, lines[:n_o.value].copy(), lines[n_o.value:n_o.value + n_f.value].copy(),
                    ran.value == 2)

    def sample_points(self, points, options: _abi.Options):
        pts = _abi.as_f32(points).reshape(-1, 3)
        n = len(pts)
        out_p = np.zeros((max(n, 1), 3), dtype=np.float32)
        out_n = np.zeros(max(n, 1), dtype=np.uint32)
        self._check(self._f("sample_points")(self.ctx, n, _abi.ptr(pts, _abi.f32p), C.byref(options),
                                             _abi.ptr(out_p, _abi.f32p), _abi.ptr(out_n, _abi.u32p)))
        return out_p[:n], out_n[:n]

    def find_paths(self, start_node, start_point, end_node, end_point, overrides=None,
                   path_capacity: int | None = None, permitted=None):
        """overrides: list (per query) of dict{type: cost} or None.
        permitted: list (per query) of None (= All) or an iterable of permitted link kinds; None = All."""
        sn = _abi.as_u32(start_node)
        en = _abi.as_u32(end_node)
        sp = _abi.as_f32(start_point).reshape(-1, 3)
        ep = _abi.as_f32(end_point).reshape(-1, 3)
        n = len(sn)
        ob, ot, oc = self._overrides_csr(overrides, n)
        pa, pb, pk = self._permitted_csr(permitted, n)
        cap = path_capacity or max(1024, 64 * n)
        while True:
            succ = np.zeros(max(n, 1), dtype=np.uint32)
            expl = np.zeros(max(n, 1), dtype=np.uint32)
            begin = np.zeros(n + 1, dtype=np.uint32)
            nodes = np.zeros(cap, dtype=np.uint32)
            portals = np.zeros(cap, dtype=np.uint32)
            st = self._f("find_paths")(
                self.ctx, n, _abi.ptr(sn, _abi.u32p), _abi.ptr(sp, _abi.f32p), _abi.ptr(en, _abi.u32p),
                _abi.ptr(ep, _abi.f32p), _abi.ptr(ob, _abi.u32p), _abi.ptr(ot, _abi.u32p),
                _abi.ptr(oc, _abi.f32p), _abi.ptr(succ, _abi.u32p), _abi.ptr(expl, _abi.u32p),
                _abi.ptr(begin, _abi.u32p), _abi.ptr(nodes, _abi.u32p), _abi.ptr(portals, _abi.u32p), cap,
                _abi.ptr(pa, _abi.u32p), _abi.ptr(pb, _abi.u32p), _abi.ptr(pk, _abi.u32p))
            if st == _abi.ERR_CAPACITY:
                cap = max(int(begin[n]), cap * 2)
                continue
            self._check(st)
            return succ[:n], expl[:n], begin, nodes[:begin[n]], portals[:begin[n]]


    def _overrides_csr(self, overrides, n):
        if overrides is None:
            return None, None, None
        ob = np.zeros(n + 1, dtype=np.uint32)
        tl, cl = [], []
        for i, ov in enumerate(overrides):
            for t, c in sorted((ov or {}).items()):
                tl.append(t)
                cl.append(c)
            ob[i + 1] = len(tl)
        return ob, np.array(tl + [0], dtype=np.uint32), np.array(cl + [0], dtype=np.float32)

    def _permitted_csr(self, permitted, n):
        if permitted is None:
            return None, None, None
        pa = np.array([1 if p is None else 0 for p in permitted], dtype=np.uint32)
        pb = np.zeros(n + 1, dtype=np.uint32)
        kinds = []
        for i, p in enumerate(permitted):
            kinds.extend(sorted(p or ()))
            pb[i + 1] = len(kinds)
        return pa, pb, np.array(kinds + [0], dtype=np.uint32)

    def find_straight_paths(self, start_node, start_point, end_node, end_point, overrides=None, permitted=None):
        """Batched query-API find_path: returns (success, step_begin, kind, points (k,6), link)."""
        sn = _abi.as_u32(start_node)
        en = _abi.as_u32(end_node)
        sp = _abi.as_f32(start_point).reshape(-1, 3)
        ep = _abi.as_f32(end_point).reshape(-1, 3)
        n = len(sn)
        ob, ot, oc = self._overrides_csr(overrides, n)
        pa, pb, pk = self._permitted_csr(permitted, n)
        cap = max(1024, 64 * n)
        while True:
            succ = np.zeros(max(n, 1), dtype=np.uint32)
            begin = np.zeros(n + 1, dtype=np.uint32)
            kind = np.zeros(cap, dtype=np.uint32)
            pts = np.zeros((cap, 6), dtype=np.float32)
            link = np.zeros(cap, dtype=np.uint32)
            st = self._f("find_straight_paths")(
                self.ctx, n, _abi.ptr(sn, _abi.u32p), _abi.ptr(sp, _abi.f32p), _abi.ptr(en, _abi.u32p),
                _abi.ptr(ep, _abi.f32p), _abi.ptr(ob, _abi.u32p), _abi.ptr(ot, _abi.u32p),
                _abi.ptr(oc, _abi.f32p), _abi.ptr(succ, _abi.u32p), _abi.ptr(begin, _abi.u32p),
                _abi.ptr(kind, _abi.u32p), _abi.ptr(pts, _abi.f32p), _abi.ptr(link, _abi.u32p), cap,
                _abi.ptr(pa, _abi.u32p), _abi.ptr(pb, _abi.u32p), _abi.ptr(pk, _abi.u32p))
            if st == _abi.ERR_CAPACITY:
                cap = max(int(begin[n]), cap * 2)
                continue
            self._check(st)
            k = int(begin[n])
            return succ[:n], begin, kind[:k], pts[:k], link[:k]


_nccl_preloaded = False


def _preload_nccl():
    """The library dlopens "libnccl.so.2" and gets whichever copy the process already holds under that soname,
    else the loader's search path.  In a Python process that later imports torch that must be the copy torch was
    built against (the `nvidia-nccl` wheel: a system libnccl of another version under the same soname makes
    `import torch` fail with an undefined symbol), so the wheel's copy — if there is one — is loaded first."""
    global _nccl_preloaded
    if _nccl_preloaded:
        return
    _nccl_preloaded = True
    import importlib.util
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
    except (ImportError, ValueError):
        spec = None
    for d in (spec.submodule_search_locations if spec and spec.submodule_search_locations else []):
        path = os.path.join(d, "lib", "libnccl.so.2")
        if os.path.exists(path):
            C.CDLL(path, mode=C.RTLD_GLOBAL)
            return


class NativeBackend(CBackend):
    prefix = "lmb_"

    def __init__(self, max_agents: int = 1024, device: int = 0, astar_slots: int = 0,
                 scratch_bytes: int = 0, flags: int = 0, lib_path: str | None = None, astar_shape: int = 0):
        super().__init__()
        self.lib = load_library(lib_path)
        if self.lib.lmb_abi_version() != _abi.ABI_VERSION:
            raise RuntimeError("landmass_b200: ABI version mismatch between header mirror and library")
        cfg = _abi.Config()
        cfg.struct_size = C.sizeof(_abi.Config)
        cfg.device = device
        cfg.max_agents = max_agents
        cfg.astar_slots = astar_slots
        cfg.scratch_bytes = scratch_bytes
        cfg.flags = flags
        cfg.astar_shape = astar_shape
        ctx = C.c_void_p()
        st = self.lib.lmb_create(C.byref(cfg), C.byref(ctx))
        if st != 0:
            raw = self.lib.lmb_last_error(None)
            raise NativeError(st, raw.decode() if raw else "lmb_create failed")
        self.ctx = ctx

    # resident-input variants (bench `value` path)
    def agents_upload(self, ag: dict, ch: dict | None):
        a, ka = _abi.make_agents_in(ag)
        c, kc = _abi.make_characters_in(ch)
        self._check(self.lib.lmb_agents_upload(self.ctx, C.byref(a), C.byref(c)))
        self._n_agents = a.n

    def step_resident(self, options: _abi.Options, dt: float):
        self._check(self.lib.lmb_step_resident(self.ctx, C.byref(options), C.c_float(dt)))

    def step_begin(self, options: _abi.Options, dt: float, n_total: int, first: int):
        self._check(self.lib.lmb_step_begin(self.ctx, C.byref(options), C.c_float(dt), n_total, first))

    def halo_buffer(self):
        """(device pointer, record bytes, n records) of the avoidance-record halo buffer."""
        p, rb, nr = C.c_void_p(), C.c_uint64(), C.c_uint64()
        self._check(self.lib.lmb_halo_buffer(self.ctx, C.byref(p), C.byref(rb), C.byref(nr)))
        return p.value, rb.value, nr.value

    def step_end(self, options: _abi.Options, dt: float):
        self._check(self.lib.lmb_step_end(self.ctx, C.byref(options), C.c_float(dt)))

    # library-owned halo exchange (lmb_comm.cu): NCCL opened by the library itself
    def comm_unique_id(self) -> bytes:
        _preload_nccl()
        buf = (C.c_uint8 * 128)()
        self._check(self.lib.lmb_comm_unique_id(self.ctx, buf))
        return bytes(buf)

    def comm_init(self, unique_id: bytes, n_ranks: int, rank: int):
        _preload_nccl()
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._check(self.lib.lmb_comm_init(self.ctx, buf, n_ranks, rank))

    def step_sharded(self, options: _abi.Options, dt: float, agents_per_rank: int):
        self._check(self.lib.lmb_step_sharded(self.ctx, C.byref(options), C.c_float(dt), agents_per_rank))

    def agents_integrate(self, dt: float):
        """position += desired_velocity * dt, velocity = desired_velocity, RESET cleared — on the device."""
        self._check(self.lib.lmb_agents_integrate(self.ctx, C.c_float(dt)))

    def resident_positions(self) -> np.ndarray:
        out = np.zeros((max(self._n_agents, 1), 3), dtype=np.float32)
        self._check(self.lib.lmb_agents_positions(self.ctx, _abi.ptr(out, _abi.f32p), self._n_agents))
        return out[:self._n_agents]

    def query_timings(self) -> _abi.StepTimings:
        """Timings of the last find_paths / find_straight_paths batch (astar_ms, explored_nodes ...)."""
        t = _abi.StepTimings()
        self._check(self.lib.lmb_get_query_timings(self.ctx, C.byref(t)))
        return t

    def agents_download(self) -> dict:
        o, out = _abi.make_agents_out(self._n_agents)
        self._check(self.lib.lmb_agents_download(self.ctx, C.byref(o)))
        return out

    def step_timings(self) -> _abi.StepTimings:
        t = _abi.StepTimings()
        self._check(self.lib.lmb_get_step_timings(self.ctx, C.byref(t)))
        return t

    def close(self):
        if self.ctx is not None and self.lib is not None:
            self.lib.lmb_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
