"""Host-side logic of the multi-GPU step: contiguous agent ranges per rank,
replicated navigation data, and ONE exchange per frame — the all-gather of the 32-byte
avoidance records between phase C and phase D.  No other collective exists on this path.
"""
from __future__ import annotations

import numpy as np

_PER_AGENT = ("slot", "flags", "position", "velocity", "target", "radius", "desired_speed", "max_speed",
              "reach_condition", "reach_distance", "anim_link_reach_distance")


def shard_range(n_total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous split of the global agent order (keeps PathingResult order by rank)."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def slice_agents(ag: dict, lo: int, hi: int, slot_base: int | None = None) -> dict:
    """This rank's slice of a global agent SoA (CSR arrays are re-based)."""
    out = {"n": hi - lo}
    for k in _PER_AGENT:
        if ag.get(k) is not None:
            out[k] = np.ascontiguousarray(ag[k][lo:hi])
    if slot_base is not None:
        out["slot"] = (out["slot"] - np.uint32(slot_base)).astype(np.uint32)
    for begin, *vals in (("override_begin", "override_type", "override_cost"),
                         ("permitted_begin", "permitted_kind")):
        if ag.get(begin) is not None:
            b = np.asarray(ag[begin], dtype=np.int64)
            out[begin] = (b[lo:hi + 1] - b[lo]).astype(np.uint32)
            for v in vals:
                out[v] = np.ascontiguousarray(ag[v][b[lo]:b[hi]])
    return out


def gather_pathing_results(local_results, lo: int):
    """Re-bases this rank's (agent, success, explored) tuples to global agent indices."""
    return [(a + lo, s, e) for (a, s, e) in local_results]
