"""The C-ABI library loads and exports every symbol include/landmass_b200.h declares, the ctypes
mirror has the same struct layouts as the C header, and the product refuses to run without a GPU
(no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

from landmass_b200 import _abi, _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "landmass_b200.h")


def header_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lmb_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_native._LIB_PATH), "run __graft_entry__.build() first"
    lib = C.CDLL(_native._LIB_PATH)
    declared = header_functions()
    assert len(declared) >= 17
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(_native.EXPORTED_SYMBOLS) == set(declared)
    lib.lmb_abi_version.restype = C.c_uint32
    assert lib.lmb_abi_version() == _abi.ABI_VERSION


def test_ctypes_layout_matches_header(tmp_path):
    structs = {"lmb_options": _abi.Options, "lmb_navdata_desc": _abi.NavDataDesc,
               "lmb_agents_in": _abi.AgentsIn, "lmb_characters_in": _abi.CharactersIn,
               "lmb_agents_out": _abi.AgentsOut, "lmb_pathing_result": _abi.PathingResult,
               "lmb_config": _abi.Config, "lmb_step_timings": _abi.StepTimings, "lmb_nav_remap": _abi.NavRemap}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void) {"]
    for cname, py in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in py._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-o", str(exe), str(src)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, py in structs.items():
        assert int(out[cname]) == C.sizeof(py), cname
        for fname, _ in py._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(py, fname).offset, f"{cname}.{fname}"


def test_constants_match_header():
    text = open(HEADER).read()
    for name, value in [("LMB_ABI_VERSION", _abi.ABI_VERSION), ("LMB_AGENT_HAS_TARGET", _abi.AGENT_HAS_TARGET),
                        ("LMB_AGENT_PAUSED", _abi.AGENT_PAUSED), ("LMB_AGENT_RESET", _abi.AGENT_RESET),
                        ("LMB_AGENT_PERMIT_ALL_LINKS", _abi.AGENT_PERMIT_ALL_LINKS),
                        ("LMB_AGENT_KEEP_AVOIDANCE_DATA", _abi.AGENT_KEEP_AVOIDANCE_DATA),
                        ("LMB_CONFIG_NO_AVOIDANCE", _abi.CONFIG_NO_AVOIDANCE),
                        ("LMB_CONFIG_DEBUG_LOG", _abi.CONFIG_DEBUG_LOG),
                        ("LMB_CONFIG_ASTAR_COMPACT", _abi.CONFIG_ASTAR_COMPACT),
                        ("LMB_CONFIG_ASTAR_DENSE", _abi.CONFIG_ASTAR_DENSE),
                        ("LMB_CONFIG_ASTAR_HASHED", _abi.CONFIG_ASTAR_HASHED),
                        ("LMB_CONFIG_ASTAR_LARGE_SHAPES", _abi.CONFIG_ASTAR_LARGE_SHAPES),
                        ("LMB_CONFIG_ASTAR_THREAD_ONLY", _abi.CONFIG_ASTAR_THREAD_ONLY),
                        ("LMB_CONFIG_ASTAR_WARP_ONLY", _abi.CONFIG_ASTAR_WARP_ONLY),
                        ("LMB_CONFIG_ASTAR_NO_FOREST", _abi.CONFIG_ASTAR_NO_FOREST),
                        ("LMB_CONFIG_NO_L2_WINDOW", _abi.CONFIG_NO_L2_WINDOW),
                        ("LMB_CONFIG_ASTAR_SMALL_HEAP", _abi.CONFIG_ASTAR_SMALL_HEAP),
                        ("LMB_CONFIG_ASTAR_NO_PAIR", _abi.CONFIG_ASTAR_NO_PAIR),
                        ("LMB_CONFIG_ORDER_ALWAYS", _abi.CONFIG_ORDER_ALWAYS)]:
        m = re.search(rf"#define {name} (\d+)u", text)
        assert m and int(m.group(1)) == value, name


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_native.NativeError) as e:
        _native.NativeBackend()
    assert "no CPU fallback" in str(e.value)


def test_package_does_not_import_oracle():
    code = "import sys; import landmass_b200, landmass_b200._native, landmass_b200.sharded, landmass_b200.synth; " \
           "assert not any(m.startswith('oracle') for m in sys.modules), 'oracle imported by the product'"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)


@pytest.mark.gpu
def test_error_paths_of_the_newer_entry_points():
    """lmb_navdata_update refuses a remap that does not describe the previous upload; the read-backs
    refuse bad slots; every failure leaves a message in lmb_last_error."""
    import numpy as np

    from landmass_b200 import synth

    mesh = synth.grid_mesh(4, 4)
    flat = synth.single_island_flat(mesh)
    gpu = _native.NativeBackend(max_agents=8)
    # before any upload a remap is ignored (there is nothing to carry over)
    gpu.navdata_upload(flat, {"island_map": np.zeros(0, dtype=np.uint32), "link_map": np.zeros(0, dtype=np.uint32)})
    for remap, what in [
        ({"island_map": np.zeros(3, dtype=np.uint32), "link_map": np.zeros(0, dtype=np.uint32)}, "previous upload"),
        ({"island_map": np.array([5], dtype=np.uint32), "link_map": np.zeros(0, dtype=np.uint32)}, "not the same island"),
    ]:
        with pytest.raises(_native.NativeError) as e:
            gpu.navdata_upload(flat, remap)
        assert e.value.status == _abi.ERR_INVALID_ARGUMENT and what in str(e.value)
    gpu.navdata_upload(flat, {"island_map": np.array([0], dtype=np.uint32), "link_map": np.zeros(0, dtype=np.uint32)})
    with pytest.raises(_native.NativeError):
        gpu.agent_path_endpoints(8)
    kind, pts, link = gpu.agent_waypoints()
    assert len(kind) == 0
    ag = synth.make_agents(mesh, 4)
    gpu.step(ag, None, synth.default_options(), 0.1)
    kind, pts, link = gpu.agent_waypoints()
    assert len(kind) == 4 and set(kind.tolist()) <= {0, _abi.NONE}
    # lmb_agent_avoidance_constraints: nobody asked -> nothing; out of range -> error; asked but no step
    # since the upload -> "did not run"; after a step -> the lines
    assert gpu.agent_avoidance_constraints(0)[0] is False
    with pytest.raises(_native.NativeError):
        gpu.agent_avoidance_constraints(4)
    ag["flags"] = ag["flags"] | np.uint32(_abi.AGENT_KEEP_AVOIDANCE_DATA)
    gpu.agents_upload(ag, None)
    assert gpu.agent_avoidance_constraints(1)[0] is False
    gpu.step_resident(synth.default_options(), 0.1)
    ran, original, fallback, fell_back = gpu.agent_avoidance_constraints(1)
    assert ran and original.shape[1] == 4 and len(original) >= 3
    gpu.close()


@pytest.mark.gpu
def test_error_paths_of_the_round_2_entry_points():
    """lmb_step_sharded without a communicator, with too small a share, lmb_comm_init with a bad rank,
    lmb_agents_integrate before a step, a step range outside the halo, null CSR arrays."""
    import numpy as np

    from landmass_b200 import synth

    mesh = synth.grid_mesh(4, 4)
    flat = synth.single_island_flat(mesh)
    opts = synth.default_options()
    gpu = _native.NativeBackend(max_agents=8)
    gpu.navdata_upload(flat)
    ag = synth.make_agents(mesh, 4)
    gpu.agents_upload(ag, None)
    with pytest.raises(_native.NativeError) as e:
        gpu.agents_integrate(0.1)  # no step has produced velocities for these agents
    assert e.value.status == _abi.ERR_INVALID_ARGUMENT
    with pytest.raises(_native.NativeError) as e:
        gpu.step_sharded(opts, 0.1, 4)
    assert e.value.status == _abi.ERR_INVALID_ARGUMENT and "lmb_comm_init" in str(e.value)
    with pytest.raises(_native.NativeError) as e:
        gpu.comm_init(b"\0" * 128, 2, 2)  # rank out of range: refused before NCCL is touched
    assert e.value.status == _abi.ERR_INVALID_ARGUMENT
    gpu.comm_init(gpu.comm_unique_id(), 1, 0)
    with pytest.raises(_native.NativeError) as e:
        gpu.step_sharded(opts, 0.1, 3)  # this rank holds 4 agents
    assert e.value.status == _abi.ERR_INVALID_ARGUMENT and "agents_per_rank" in str(e.value)
    with pytest.raises(_native.NativeError) as e:
        gpu.step_begin(opts, 0.1, 6, 3)  # [3, 7) does not fit 6 records: refused before anything is stepped
    assert e.value.status == _abi.ERR_INVALID_ARGUMENT
    out = None
    gpu.step_sharded(opts, 0.1, 4)
    gpu.agents_integrate(0.1)
    moved = gpu.resident_positions()
    out = gpu.agents_download()
    np.testing.assert_array_equal(moved, (ag["position"] + out["desired_velocity"] * np.float32(0.1)).astype(np.float32))
    t = gpu.step_timings()
    assert 1 <= t.astar_launch_kind <= 7 and t.allgather_ms >= 0.0
    gpu.close()
