"""Single-search latency and small-batch throughput of the A* kernel families (k_astar_warp.cu: a search warp +
a heap warp per query up to 4 queries per SM, one warp per query; k_astar.cu: one query per thread) on a forest (256-room maze) and an open field
(256 x 256 quads of 8 m, antipodal queries).  Prints one JSON line per (mesh, kernel, batch).
Usage: python scripts/astar_latency.py [maze] [field] [tiled]"""
import json
import sys
import time

sys.path.insert(0, ".")
import numpy as np

from landmass_b200 import _abi, synth
from landmass_b200._native import NativeBackend


def queries(kind, mesh, nq, rng):
    sp, sn = synth.random_points_on_mesh(mesh, nq, rng)
    if kind == "field":
        ep = np.stack([2048.0 - sp[:, 0], 2048.0 - sp[:, 1], sp[:, 2]], axis=1).astype(np.float32)
        cell = 8.0
        en = ((ep[:, 1] // cell).astype(np.int64) * 256 + (ep[:, 0] // cell).astype(np.int64)).astype(np.uint32)
    else:
        ep, en = synth.random_points_on_mesh(mesh, nq, rng)
    return sn, sp, en, ep


def run(kind):
    if kind == "maze":
        mesh = synth.maze_mesh(256)
        flat = synth.single_island_flat(mesh)
    elif kind == "field":
        mesh = synth.grid_mesh(256, 256, cell=8.0)
        flat = synth.single_island_flat(mesh)
    else:
        mesh = None
        flat = synth.tiled_archipelago_flat(8, 128)
    kernels = (("pair", _abi.CONFIG_ASTAR_WARP_ONLY), ("warp", _abi.CONFIG_ASTAR_WARP_ONLY | _abi.CONFIG_ASTAR_NO_PAIR),
               ("thread", _abi.CONFIG_ASTAR_THREAD_ONLY))
    for name, flags in kernels:
        gpu = NativeBackend(max_agents=1024, flags=flags)
        gpu.navdata_upload(flat)
        opts = synth.default_options()
        for nq in (1, 16, 148, 592, 1184) if name == "pair" else (1, 16, 148, 592, 1184, 2368):
            rng = np.random.Generator(np.random.PCG64(1000 + nq))
            if mesh is not None:
                sn, sp, en, ep = queries(kind, mesh, nq, rng)
            else:
                pts = np.zeros((2 * nq, 3), dtype=np.float32)
                pts[:, :2] = rng.uniform(0.5, 1023.5, size=(2 * nq, 2))
                p, nodes = gpu.sample_points(pts, opts)
                sn, sp, en, ep = nodes[:nq], p[:nq], nodes[nq:], p[nq:]
            best = None
            for rep in range(3):
                succ, expl, begin, nodes_, portals = gpu.find_paths(sn, sp, en, ep, path_capacity=1 << 24)
                T = gpu.query_timings()
                if best is None or T.astar_ms < best:
                    best = T.astar_ms
            assert succ.all()
            print(json.dumps({"mesh": kind, "kernel": name, "queries": nq, "astar_ms": best,
                              "explored_total": int(expl.sum()), "explored_max": int(expl.max()),
                              "us_per_node_longest": 1e3 * best / max(int(expl.max()), 1),
                              "Mnodes_per_s": expl.sum() / best / 1e3,
                              "checksum": int(expl.astype(np.uint64).sum() * 31 + begin[-1])}), flush=True)
        gpu.close()


if __name__ == "__main__":
    for k in sys.argv[1:] or ["maze", "field"]:
        run(k)
