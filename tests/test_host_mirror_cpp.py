"""include/landmass_b200.hpp — the C++ host mirror of the per-frame API (Archipelago / Agent / Character /
ArchipelagoOptions / AgentState / DenseSlotMap over the C ABI) — against the Python mirror: the same scripted
game loop (agents added and removed, a re-used slot, a paused agent, a cost override, a character), frame by
frame, on the same kind of backend.  The harness (tests/cpp/host_mirror_harness.cpp) is compiled here with g++
and gets the backend's entry points as function pointers: the CPU oracle on CPU, the CUDA library on a GPU box.
Everything must be identical: both mirrors pack the same arrays for the same step function."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from landmass_b200 import (Agent, Archipelago, ArchipelagoOptions, Character, Island, TargetReachedCondition,
                           Transform, _abi)
from tests.helpers import valid_mesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FRAMES, DT, MAX_AGENTS = 40, 0.1, 16


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hm") / "libhost_mirror_harness.so")
    subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-Wall", "-Wextra", "-Werror", "-shared", "-fPIC",
                    "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "host_mirror_harness.cpp"),
                    "-o", out], check=True)
    lib = C.CDLL(out)
    lib.hm_run_script.restype = C.c_int32
    return lib


def python_trace(backend):
    """The script of host_mirror_harness.cpp on the Python mirror."""
    arch = Archipelago(ArchipelagoOptions.from_agent_radius(0.5), backend=backend)
    arch.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), valid_mesh(dict(
        vertices=[(0.0, 0.0, 0.0), (15.0, 0.0, 0.0), (15.0, 15.0, 0.0), (0.0, 15.0, 0.0)],
        polygons=[[0, 1, 2, 3]], polygon_type_indices=[0]))))

    def make(pos, target, radius):
        a = Agent.create(np.array(pos, dtype=np.float32), np.zeros(3, dtype=np.float32), radius, 1.0, 2.0)
        a.current_target = target
        a.target_reached_condition = TargetReachedCondition.Distance(0.01)
        return a

    ids = [arch.add_agent(make((1.0, 1.0, 0.0), (11.0, 1.1, 0.0), 1.0)),
           arch.add_agent(make((11.0, 1.1, 0.0), (1.0, 1.0, 0.0), 1.0)), None, None]
    c = make((6.0, 8.0, 0.0), (6.0, 2.0, 0.0), 0.5)
    c.target_reached_condition = TargetReachedCondition.StraightPathDistance(0.25)
    assert c.override_type_index_cost(0, 3.0)
    ids[2] = arch.add_agent(c)
    alive = [True, True, True, False]
    ch = arch.add_character(Character(np.array((6.0, 4.0, 0.0), dtype=np.float32),
                                      np.array((0.2, 0.0, 0.0), dtype=np.float32), 0.5))
    dt = np.float32(DT)
    trace_f = np.zeros((FRAMES, 4, 6), dtype=np.float32)
    trace_u = np.zeros((FRAMES, 14), dtype=np.uint32)
    for f in range(FRAMES):
        if f == 10:
            arch.remove_agent(ids[0])
            alive[0] = False
        if f == 12:
            ids[3] = arch.add_agent(make((2.0, 12.0, 0.0), (12.0, 12.0, 0.0), 0.5))
            alive[3] = True
        if f == 20:
            arch.get_agent_mut(ids[1]).paused = True
        if f == 25:
            arch.get_agent_mut(ids[1]).paused = False
        arch.update(float(dt))
        for k in range(4):
            if not alive[k]:
                trace_f[f, k] = -1000.0
                trace_u[f, k] = 99
                continue
            a = arch.get_agent_mut(ids[k])
            a.velocity = np.asarray(a.get_desired_velocity(), dtype=np.float32).copy()
            a.position = (np.asarray(a.position, dtype=np.float32) + a.velocity * dt).astype(np.float32)
            trace_f[f, k, :3] = a.position
            trace_f[f, k, 3:] = a.velocity
            trace_u[f, k] = int(a.state())
        chr_ = arch.get_character_mut(ch) if hasattr(arch, "get_character_mut") else arch.characters.get(ch)
        chr_.position = (np.asarray(chr_.position, dtype=np.float32)
                         + np.asarray(chr_.velocity, dtype=np.float32) * dt).astype(np.float32)
        pr = arch.get_pathing_results()
        trace_u[f, 4] = len(pr)
        trace_u[f, 5] = sum(r.explored_nodes for r in pr)
        for k, r in enumerate(pr[:4]):
            trace_u[f, 6 + 2 * k] = r.agent.idx
            trace_u[f, 7 + 2 * k] = r.agent.version
    arch._sync_navdata()
    return trace_f, trace_u, arch._flat


def cpp_trace(harness, ctx, fns, desc):
    trace_f = np.zeros((FRAMES, 4, 6), dtype=np.float32)
    trace_u = np.zeros((FRAMES, 14), dtype=np.uint32)
    err = C.create_string_buffer(512)
    st = harness.hm_run_script(ctx, *[C.cast(f, C.c_void_p) for f in fns], C.byref(desc), C.c_uint32(MAX_AGENTS),
                               C.c_uint32(FRAMES), C.c_float(DT), trace_f.ctypes.data_as(C.c_void_p),
                               trace_u.ctypes.data_as(C.c_void_p), err, C.c_uint32(512))
    assert st == 0, err.value.decode()
    return trace_f, trace_u


def check(py, cpp):
    assert np.array_equal(py[1], cpp[1]), "states / pathing results differ"
    assert np.array_equal(py[0].view(np.uint32), cpp[0].view(np.uint32)), "positions / velocities differ"
    states = py[1][:, :4]
    assert (states[:10, 0] != 99).all() and (states[10:, 0] == 99).all()      # A removed at frame 10
    assert (states[:12, 3] == 99).all() and (states[12:, 3] != 99).all()      # D added at frame 12 ...
    assert py[1][12, 4] >= 1 and (py[1][12, 6:8] == py[1][0, 6:8] + [0, 2]).all()  # ... in A's slot, next version
    assert (states[20:25, 1] == 8).all() and (states[25:, 1] != 8).all()      # B paused for five frames


def test_cpp_host_mirror_matches_python_mirror_on_the_oracle(harness):
    from oracle import oracle_py
    from oracle.oracle_py import OracleBackend

    py = python_trace(OracleBackend())
    desc, keep = _abi.make_navdata_desc(py[2])
    lib = oracle_py.load()
    lib.orc_create.restype = C.c_void_p
    ctx = C.c_void_p(lib.orc_create(C.byref(desc), C.c_uint32(0)))
    assert ctx
    cpp = cpp_trace(harness, ctx, (lib.orc_navdata_upload, lib.orc_step, lib.orc_pathing_results), desc)
    check(py, cpp)
    del keep


@pytest.mark.gpu
def test_cpp_host_mirror_matches_python_mirror_on_cuda(harness):
    from landmass_b200._native import NativeBackend

    py = python_trace(NativeBackend(max_agents=MAX_AGENTS))
    desc, keep = _abi.make_navdata_desc(py[2])
    gpu = NativeBackend(max_agents=MAX_AGENTS)   # a second context, driven by the C++ mirror
    lib = gpu.lib
    cpp = cpp_trace(harness, gpu.ctx, (lib.lmb_navdata_upload, lib.lmb_step, lib.lmb_pathing_results), desc)
    check(py, cpp)
    gpu.close()
    del keep
