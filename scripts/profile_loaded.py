"""Fully loaded A* launch for `ncu --set full`: many more queries than slots, modest arena
(128x128-room maze, slots capped) so kernel replay's save/restore stays affordable."""
import sys
sys.path.insert(0, ".")
from landmass_b200 import synth
from landmass_b200._native import NativeBackend

rooms = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n = int(sys.argv[2]) if len(sys.argv) > 2 else 150000
slots = int(sys.argv[3]) if len(sys.argv) > 3 else 18944
mesh = synth.maze_mesh(rooms)
flat = synth.single_island_flat(mesh)
gpu = NativeBackend(max_agents=n, astar_slots=slots)
gpu.navdata_upload(flat)
ag = synth.make_agents(mesh, n)
opts = synth.default_options()
gpu.agents_upload(ag, None)
for step in range(2):
    gpu.step_resident(opts, 1.0 / 60.0)
    T = gpu.step_timings()
    print(f"step {step}: astar {T.astar_ms:.2f} ms n_repath {T.n_repath} explored {T.explored_nodes} "
          f"path_nodes {T.path_nodes} launches {T.kernel_launches}", flush=True)
gpu.close()
