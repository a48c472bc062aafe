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


def round_up8(v: int) -> int:
    return (v + 7) // 8 * 8


def factor_row_range(k: int, rank: int, world: int):
    """Rows of the nucleus inverse rank `rank` streams (gemv_rows_kernel in solve.cu): contiguous blocks
    of per_k = roundUp8(ceil(k / world)) rows; k is the CURRENT nucleus size."""
    per_k = round_up8((k + world - 1) // world)
    lo = min(k, rank * per_k)
    return lo, min(k, lo + per_k), per_k


def gather_slot(i: int, c: int, k: int, world: int, nrhs: int, per_max: int) -> int:
    """Index of GEMV result (row i, right-hand side c) in the gathered buffer [rank][rhs][per_max]
    (gemv_result / btran_scatter_kernel in solve.cu); per_max = roundUp8(ceil(m / world)) is fixed for the
    whole solve so that the all-gather size does not depend on the nucleus size."""
    per_k = round_up8((k + world - 1) // world)
    rk = i // per_k
    return (rk * nrhs + c) * per_max + (i - rk * per_k)
