"""Experiment: effect of query order on the A* makespan (longest-first vs random)."""
import sys
sys.path.insert(0, ".")
import numpy as np
from landmass_b200 import _abi, synth
from landmass_b200._native import NativeBackend

rooms = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
mesh = synth.maze_mesh(rooms)
flat = synth.single_island_flat(mesh)
gpu = NativeBackend(max_agents=n)
gpu.navdata_upload(flat)
ag = synth.make_agents(mesh, n)
opts = synth.default_options()

def run(ag, label):
    gpu.agents_upload(ag, None)
    for step in range(2):
        gpu.step_resident(opts, 1.0 / 60.0)
    T = gpu.step_timings()
    print(f"{label}: astar {T.astar_ms:.1f} ms explored {T.explored_nodes}", flush=True)
    return gpu.pathing_results_array()

res = run(ag, "random order")
expl = np.zeros(n, dtype=np.int64)
expl[res[:, 0]] = res[:, 2]
def reorder(ag, order):
    out = dict(ag)
    for k, v in ag.items():
        if isinstance(v, np.ndarray) and len(v) == n and k != "slot":
            out[k] = np.ascontiguousarray(v[order])
    return out
o1 = np.argsort(-expl, kind="stable")
run(reorder(ag, o1), "explored-descending (oracle order)")
d = np.linalg.norm(ag["position"] - ag["target"], axis=1)
o2 = np.argsort(-d, kind="stable")
run(reorder(ag, o2), "euclidean-distance-descending")
gpu.close()
