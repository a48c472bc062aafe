"""Host-side helpers for the column-sharded pricing pass (one process per GPU).

``torch.distributed`` is used for plumbing only: shipping NCCL's unique id from rank 0 to
the other ranks.  The per-iteration exchange (one all-gather of the tableau-row shards) is
issued from C++ on the engine's own CUDA stream (clp_b200/csrc/capi.cu, engine.cu).
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int):
    """Contiguous column block of `rank`; must match Engine::enqueueIteration (engine.cu)."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def broadcast_unique_id(uid: np.ndarray, src: int = 0) -> np.ndarray:
    import torch
    import torch.distributed as dist

    t = torch.from_numpy(np.ascontiguousarray(uid, dtype=np.uint8).copy())
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=src)
    return t.cpu().numpy()
