"""Where a small frame's time goes: 100 agents on the 64x64 grid (BASELINE config 1), steady frames.
Prints device time per phase (CUDA events on the step stream) next to the wall time of lmb_step."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from landmass_b200 import _abi, synth
from landmass_b200._native import NativeBackend

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
lib = sys.argv[2] if len(sys.argv) > 2 else None  # an alternative liblandmass_b200.so (A/B of kernel variants)
mesh = synth.grid_mesh(64, 64)
flat = synth.single_island_flat(mesh)
ag = synth.make_agents(mesh, n)
opts = synth.default_options()
for flags, name in ((0, "full step"), (_abi.CONFIG_NO_AVOIDANCE, "without avoidance")):
    gpu = NativeBackend(max_agents=n, flags=flags, lib_path=lib)
    gpu.navdata_upload(flat)
    a = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in ag.items()}
    rows = []
    for f in range(40):
        t0 = time.perf_counter()
        out = gpu.step(a, None, opts, 1 / 60)
        wall = (time.perf_counter() - t0) * 1e3
        T = gpu.step_timings()
        rows.append((wall, T.total_ms, T.sample_ms, T.repath_decide_ms, T.astar_ms, T.follow_ms, T.avoidance_ms, T.h2d_ms, T.d2h_ms, T.n_repath))
        a["flags"] = a["flags"] & ~np.uint32(_abi.AGENT_RESET)
        a["velocity"] = out["desired_velocity"].copy()
        a["position"] = a["position"] + out["desired_velocity"] * np.float32(1 / 60)
    r = np.array(rows)
    print(name, "frame0: wall %.2f dev %.2f (astar %.2f, %d repaths)" % (r[0, 0], r[0, 1], r[0, 4], r[0, 9]))
    m = np.median(r[5:], axis=0)
    print(name, "steady median: wall %.3f  device total %.3f = sample %.3f decide %.3f astar %.3f follow %.3f avoidance %.3f | h2d %.3f d2h %.3f"
          % tuple(m[:9]), flush=True)
    gpu.close()
