"""Coalition sharding across the GPUs of one box (one process per GPU, torch.distributed).

The reference is single-process (no collective call sites; only
``nn.DataParallel`` at start.py:283).  Coalitions are independent units, so the path shards
with no data-path collective: every rank holds the same stacked deltas / W0 / validation set
(one NCCL broadcast at set-up, ``broadcast_``), evaluates a contiguous slice of the pending
coalition list, and ONE all-gather of the per-coalition (correct:int64, loss_sum:fp64) pairs
gives every rank the same memo.  Each pair is computed by exactly one rank with a fixed
reduction order, so results are bit-identical for any world size.
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Sequence, Tuple

import torch


def _td():
    import torch.distributed as td

    return td


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    td = _td()
    if td.is_available() and td.is_initialized():
        return td.get_rank(), td.get_world_size()
    return 0, 1


def default_device() -> str:
    return f"cuda:{int(os.environ.get('LOCAL_RANK', '0'))}"


def _comm_device() -> torch.device:
    td = _td()
    if td.get_backend() == "nccl":
        return torch.device(default_device())
    return torch.device("cpu")


def shard_bounds(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous balanced slice [lo, hi) of n_items for ``rank``."""
    per = (n_items + world_size - 1) // world_size
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)


def _rows_digest(rows: Sequence[Sequence[float]]) -> int:
    """63-bit digest of the dense ratio rows: every rank must be about to evaluate the same list."""
    import hashlib

    import numpy as np

    h = hashlib.blake2b(np.asarray(rows, dtype=np.float64).tobytes(), digest_size=8).digest()
    return int.from_bytes(h, "little") >> 1


def sharded_evaluate(evaluate: Callable[..., Tuple[List[int], List[float]]],
                     rows: Sequence[Sequence[float]], n_val: int = 0,
                     evaluate_into: Optional[Callable[..., None]] = None) -> Tuple[List[int], List[float]]:
    """Evaluate ``rows`` (dense ratio rows) across ranks; every rank returns all results.

    ONE collective per call: an all-gather of packed 16-byte records
    (correct:int64, loss_sum:fp64).  With ``evaluate_into`` (the engine's device path) a rank's K5 launches write
    its records straight into its slot of the send buffer -- no host round trip between the kernels and the
    collective; the gathered buffer is read back once.  Record 0 of every slot is a header (number of rows, digest
    of the row list): ranks that disagree on the pending list (e.g. differently seeded estimators) raise instead of
    silently memoising results under wrong keys.

    Two axes:
    * at least one row per rank: contiguous slices of the rows -- each record is produced by one rank, so the
      result does not depend on the world size;
    * fewer rows than ranks (late truncation waves, tiny games) and ``n_val`` given: every rank evaluates ALL rows
      on its slice of the validation images (``image_range=(lo, hi)``); the gathered partial records are summed on
      the host in rank order -- integer counts stay exact, the fp64 loss sums differ from the single-rank value only
      by the order of the final additions (and are identical on every rank)."""
    rank, ws = world()
    if ws == 1:
        return evaluate(rows)
    td = _td()
    n = len(rows)
    dev = _comm_device()
    split_images = 0 < n < ws and n_val >= ws
    per = n if split_images else (n + ws - 1) // ws
    send = torch.zeros((1 + per, 2), dtype=torch.int64, device=dev)
    send[0, 0], send[0, 1] = n, _rows_digest(rows)
    if split_images:
        mine, kw = rows, {"image_range": shard_bounds(n_val, rank, ws)}
    else:
        lo, hi = shard_bounds(n, rank, ws)
        mine, kw = rows[lo:hi], {}
    if len(mine):
        if evaluate_into is not None and dev.type == "cuda":
            evaluate_into(mine, send[1:1 + len(mine)], **kw)
        else:
            c, l = evaluate(mine, **kw)
            send[1:1 + len(mine), 0] = torch.tensor(c, dtype=torch.int64)
            send[1:1 + len(mine), 1] = torch.tensor(l, dtype=torch.float64).view(torch.int64)
    recv = torch.empty((ws, 1 + per, 2), dtype=torch.int64, device=dev)
    td.all_gather_into_tensor(recv.view(-1), send.view(-1))
    host = recv.cpu()
    if not bool((host[:, 0] == host[0, 0]).all()):
        raise RuntimeError("ranks disagree on the pending coalition list (different estimator seeds?): "
                           f"(n, digest) per rank = {host[:, 0].tolist()}")
    correct, loss = host[:, 1:, 0], host[:, 1:, 1].contiguous().view(torch.float64)
    if split_images:
        tc, tl = correct[0].clone(), loss[0].clone()
        for r in range(1, ws):                       # fixed rank order: the same sums on every rank
            tc += correct[r]
            tl += loss[r]
        return tc.tolist(), tl.tolist()
    return correct.reshape(-1)[:n].tolist(), loss.reshape(-1)[:n].tolist()


def shared_seed(seed: Optional[int]) -> Optional[int]:
    """The seed every rank must use for a stochastic estimator.  Single process: ``seed`` as given (None keeps the
    reference's OS-entropy behaviour).  world_size > 1: a given seed is returned as is (and checked to be the same
    everywhere); None is replaced by one entropy draw of rank 0, broadcast -- otherwise every rank would sample its
    own coalitions and the sharded evaluation would mix them up."""
    rank, ws = world()
    if ws == 1:
        return seed
    td = _td()
    dev = _comm_device()
    if seed is None:
        t = torch.zeros(1, dtype=torch.int64, device=dev)
        if rank == 0:
            t[0] = int.from_bytes(os.urandom(4), "little")
        td.broadcast(t, src=0)
        return int(t.item())
    t = torch.tensor([int(seed)], dtype=torch.int64, device=dev)
    lo, hi = t.clone(), t.clone()
    td.all_reduce(lo, op=td.ReduceOp.MIN)
    td.all_reduce(hi, op=td.ReduceOp.MAX)
    if int(lo.item()) != int(hi.item()):
        raise RuntimeError(f"estimator seed differs across ranks ({int(lo.item())} .. {int(hi.item())})")
    return int(seed)


def broadcast_(t: torch.Tensor, src: int = 0) -> torch.Tensor:
    """In-place broadcast of a tensor from ``src`` (no-op when not distributed)."""
    _, ws = world()
    if ws > 1:
        _td().broadcast(t, src=src)
    return t


def barrier() -> None:
    _, ws = world()
    if ws > 1:
        td = _td()
        if td.get_backend() == "nccl":
            td.barrier(device_ids=[int(os.environ.get("LOCAL_RANK", "0"))])
        else:
            td.barrier()
