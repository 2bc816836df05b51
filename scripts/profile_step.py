"""Small, deterministic step for `ncu --set full` captures (kernel replay must save/restore
every byte the kernel writes, so the A* arenas are kept modest: 64x64-room maze)."""
import sys

sys.path.insert(0, ".")
import numpy as np

from landmass_b200 import _abi, synth
from landmass_b200._native import NativeBackend

rooms = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
mesh = synth.maze_mesh(rooms)
flat = synth.single_island_flat(mesh)
gpu = NativeBackend(max_agents=n)
gpu.navdata_upload(flat)
ag = synth.make_agents(mesh, n)
opts = synth.default_options()
gpu.agents_upload(ag, None)
for step in range(2):
    gpu.step_resident(opts, 1.0 / 60.0)
    T = gpu.step_timings()
    print(f"step {step}: total {T.total_ms:.2f} ms sample {T.sample_ms:.3f} decide {T.repath_decide_ms:.3f} "
          f"astar {T.astar_ms:.2f} follow {T.follow_ms:.3f} avoid {T.avoidance_ms:.3f} n_repath {T.n_repath} "
          f"explored {T.explored_nodes} path_nodes {T.path_nodes} launches {T.kernel_launches}")
gpu.close()
