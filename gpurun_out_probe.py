import sys, time
sys.path.insert(0, '.')
import numpy as np
from landmass_b200 import _abi, synth
from landmass_b200._native import NativeBackend
rooms = int(sys.argv[1]); n = int(sys.argv[2])
t=time.time(); mesh = synth.maze_mesh(rooms); flat = synth.single_island_flat(mesh); print("mesh", mesh.n_polys, time.time()-t)
gpu = NativeBackend(max_agents=n, flags=_abi.CONFIG_NO_AVOIDANCE)
t=time.time(); gpu.navdata_upload(flat); print("upload", time.time()-t)
ag = synth.make_agents(mesh, n)
opts = synth.default_options()
for f in range(3):
    t=time.time(); out = gpu.step(ag, None, opts, 1/60); dt=time.time()-t
    T = gpu.step_timings()
    print(f"frame {f}: wall {dt*1e3:.1f} ms total {T.total_ms:.1f} sample {T.sample_ms:.2f} decide {T.repath_decide_ms:.2f} astar {T.astar_ms:.1f} follow {T.follow_ms:.2f} n_repath {T.n_repath} explored {T.explored_nodes} path_nodes {T.path_nodes} launches {T.kernel_launches}")
    ag["flags"] &= ~np.uint32(_abi.AGENT_RESET)
    ag["velocity"] = out["desired_velocity"].copy(); ag["position"] += out["desired_velocity"]*np.float32(1/60)
print(np.bincount(out["state"], minlength=9))
