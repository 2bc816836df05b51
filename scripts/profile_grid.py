"""Open-field A* timing: 256x256 quads of 8 m, antipodal targets (config 3's frame 0, fewer agents)."""
import sys
sys.path.insert(0, ".")
import numpy as np
from landmass_b200 import synth
from landmass_b200._native import NativeBackend

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
mesh = synth.grid_mesh(256, 256, cell=8.0)
flat = synth.single_island_flat(mesh)
rng = np.random.Generator(np.random.PCG64(synth.SEED))
pos, _ = synth.random_points_on_mesh(mesh, n, rng)
ag = synth.make_agents(mesh, n)
ag["position"] = pos
ag["target"] = np.stack([2048.0 - pos[:, 0], 2048.0 - pos[:, 1], pos[:, 2]], axis=1).astype(np.float32)
opts = synth.default_options()
import os
gpu = NativeBackend(max_agents=n, astar_slots=int(os.environ.get("SLOTS", "0")))
gpu.navdata_upload(flat)
gpu.agents_upload(ag, None)
for step in range(2):
    gpu.step_resident(opts, 1.0 / 60.0)
    T = gpu.step_timings()
    print(f"step {step}: astar {T.astar_ms:.2f} ms n_repath {T.n_repath} explored {T.explored_nodes} "
          f"path_nodes {T.path_nodes} launches {T.kernel_launches} -> {T.explored_nodes / T.astar_ms / 1e6:.3f} G explored/s", flush=True)
gpu.close()
