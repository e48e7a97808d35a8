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
from typing import Callable, List, Sequence, Tuple

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


def sharded_evaluate(evaluate: Callable[..., Tuple[List[int], List[float]]],
                     rows: Sequence[Sequence[float]], n_val: int = 0) -> Tuple[List[int], List[float]]:
    """Evaluate ``rows`` (dense ratio rows) across ranks; every rank returns all results.

    Two axes:
    * at least one row per rank: contiguous slices of the rows, one all-gather of the (correct, loss_sum)
      pairs -- each pair is produced by one rank, so the result does not depend on the world size;
    * fewer rows than ranks (late truncation waves, tiny games) and ``n_val`` given: every rank evaluates
      ALL rows on its slice of the validation images (``evaluate(rows, image_range=(lo, hi))``) and one
      all-reduce sums the pairs -- integer counts stay exact, the fp64 loss sums differ from the
      single-rank value only by the order of the final additions."""
    rank, ws = world()
    if ws == 1:
        return evaluate(rows)
    td = _td()
    n = len(rows)
    dev = _comm_device()
    if 0 < n < ws and n_val >= ws:
        lo, hi = shard_bounds(n_val, rank, ws)
        c, l = evaluate(rows, image_range=(lo, hi))
        tc = torch.tensor(c, dtype=torch.int64, device=dev)
        tl = torch.tensor(l, dtype=torch.float64, device=dev)
        td.all_reduce(tc)
        td.all_reduce(tl)
        return tc.cpu().tolist(), tl.cpu().tolist()
    per = (n + ws - 1) // ws
    lo, hi = shard_bounds(n, rank, ws)
    c, l = evaluate(rows[lo:hi]) if hi > lo else ([], [])
    mine_c = torch.zeros(per, dtype=torch.int64, device=dev)
    mine_l = torch.zeros(per, dtype=torch.float64, device=dev)
    if hi > lo:
        mine_c[:hi - lo] = torch.tensor(c, dtype=torch.int64)
        mine_l[:hi - lo] = torch.tensor(l, dtype=torch.float64)
    all_c = torch.empty(ws * per, dtype=torch.int64, device=dev)
    all_l = torch.empty(ws * per, dtype=torch.float64, device=dev)
    td.all_gather_into_tensor(all_c, mine_c)
    td.all_gather_into_tensor(all_l, mine_l)
    return all_c[:n].cpu().tolist(), all_l[:n].cpu().tolist()


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
