"""One batch of long A* queries through one kernel family, for `ncu --set full --import-source on`:
  python scripts/profile_single.py <maze|field> <pair|warp|thread> <n_queries> [rooms]
The first call sizes the arena; profile the second launch (ncu --launch-skip 1 -k regex:k_astar)."""
import sys
sys.path.insert(0, ".")
import numpy as np
from landmass_b200 import _abi, synth
from landmass_b200._native import NativeBackend

kind, kernel, nq = sys.argv[1], sys.argv[2], int(sys.argv[3])
rooms = int(sys.argv[4]) if len(sys.argv) > 4 else 256
if kind == "maze":
    mesh = synth.maze_mesh(rooms)
else:
    mesh = synth.grid_mesh(256, 256, cell=8.0)
flat = synth.single_island_flat(mesh)
flags = {"pair": _abi.CONFIG_ASTAR_WARP_ONLY, "warp": _abi.CONFIG_ASTAR_WARP_ONLY | _abi.CONFIG_ASTAR_NO_PAIR,
         "thread": _abi.CONFIG_ASTAR_THREAD_ONLY}[kernel]
gpu = NativeBackend(max_agents=1024, flags=flags)
gpu.navdata_upload(flat)
rng = np.random.Generator(np.random.PCG64(1000 + nq))
sp, sn = synth.random_points_on_mesh(mesh, nq, rng)
if kind == "field":
    ep = np.stack([2048.0 - sp[:, 0], 2048.0 - sp[:, 1], sp[:, 2]], axis=1).astype(np.float32)
    en = ((ep[:, 1] // 8.0).astype(np.int64) * 256 + (ep[:, 0] // 8.0).astype(np.int64)).astype(np.uint32)
else:
    ep, en = synth.random_points_on_mesh(mesh, nq, rng)
for rep in range(2):
    succ, expl, begin, nodes, portals = gpu.find_paths(sn, sp, en, ep, path_capacity=1 << 24)
    T = gpu.query_timings()
    print(f"rep {rep}: astar {T.astar_ms:.2f} ms explored {int(expl.sum())} max {int(expl.max())} "
          f"-> {1e3 * T.astar_ms / max(int(expl.max()), 1):.3f} us per node of the longest", flush=True)
gpu.close()
