"""Dense against hashed best_estimates for small batches on a 1 M-polygon world (8 x 8 islands of 128 x 128 quads):
microseconds per explored node of the longest search, warp-pair kernel.  python scripts/layout_latency.py"""
import json
import sys

sys.path.insert(0, ".")
import numpy as np

from landmass_b200 import _abi, synth
from landmass_b200._native import NativeBackend

flat = synth.tiled_archipelago_flat(8, 128)
opts = synth.default_options()
for name, flags in (("hashed", _abi.CONFIG_ASTAR_WARP_ONLY), ("dense", _abi.CONFIG_ASTAR_WARP_ONLY | _abi.CONFIG_ASTAR_DENSE)):
    gpu = NativeBackend(max_agents=1024, flags=flags | _abi.CONFIG_DEBUG_LOG, astar_slots=592)
    gpu.navdata_upload(flat)
    for nq in (1, 148):
        rng = np.random.Generator(np.random.PCG64(1000 + nq))
        pts = np.zeros((2 * nq, 3), dtype=np.float32)
        pts[:, :2] = rng.uniform(0.5, 1023.5, size=(2 * nq, 2))
        p, nodes = gpu.sample_points(pts, opts)
        sn, sp, en, ep = nodes[:nq], p[:nq], nodes[nq:], p[nq:]
        best = None
        for rep in range(2):
            succ, expl, begin, nodes_, portals = gpu.find_paths(sn, sp, en, ep, path_capacity=1 << 24)
            T = gpu.query_timings()
            best = T.astar_ms if best is None else min(best, T.astar_ms)
        print(json.dumps({"layout": name, "queries": nq, "astar_ms": best, "explored_max": int(expl.max()),
                          "us_per_node_longest": 1e3 * best / max(int(expl.max()), 1),
                          "checksum": int(expl.astype(np.uint64).sum() * 31 + begin[-1])}), flush=True)
    gpu.close()
