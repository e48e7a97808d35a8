"""CoalitionEngine: device-resident state of one utility game and the batched evaluation of
coalitions through libsvit.

What the reference does per coalition (``Game.eval_utility``, reference
fed_client_contribution/game.py:88-107): build FedAvg ratios, aggregate the member deltas,
add W0, ``load_state_dict``, then run ``evaluation`` over the whole validation loader with
an H2D copy per batch and two host syncs per batch.  Here, per BATCH of coalitions:

    ratios [C, N] (host, fp64 -> fp32)  --kernel parameters-->
    svit_aggregate (vec region, fp32)  +  svit_aggregate (mat region, operand dtype)
    for each chunk of validation images:  svit_forward_batched -> logits [C, n_val, n_cls]
    svit_score -> correct [C] int64, loss_sum [C] fp64   --D2H-->

The stacked client deltas, W0, the patchified validation set and the labels are uploaded
once and stay in HBM.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib, ops
from .layout import PlanLayout, VitConfig, pack_state_dict, plan_layout


class ValidationSet:
    """The held-out validation set, resident in HBM: images are uploaded once and turned into
    the patch matrix [n * n_patches, patch_dim] (the A operand of the patch-embedding GEMM,
    shared by every coalition); labels as int64.  The reference instead re-uploads every batch
    for every coalition (federated_learning/utils.py:880-882)."""

    def __init__(self, cfg: VitConfig, images: torch.Tensor, labels: torch.Tensor, precision: int,
                 device: str | torch.device = "cuda:0"):
        self.cfg, self.precision, self.device = cfg, precision, torch.device(device)
        self.n = int(images.shape[0])
        if tuple(images.shape[1:]) != (cfg.channels, cfg.image, cfg.image):
            raise ValueError(f"images must be [n, {cfg.channels}, {cfg.image}, {cfg.image}], got {tuple(images.shape)}")
        with torch.cuda.device(self.device):
            helper = ops.Plan(cfg, precision, 1, 1, self.device)
            self.operand_dtype = helper.operand_dtype
            # the patch matrix in the plan's operand format (split precisions: planes, written by patchify itself)
            self.patches = helper.operand_array((self.n * cfg.n_patches, cfg.patch_dim))
            self.upload(images, labels, helper)
            helper.close()

    UPLOAD_STEP = 1024   # images per host -> device copy + patchify launch

    def upload(self, images: torch.Tensor, labels: torch.Tensor, plan=None, events: Optional[list] = None) -> int:
        """(Re-)upload host images/labels; returns the bytes copied host -> device.  ``events``: a list that receives
        one CUDA event per UPLOAD_STEP images (recorded on the current stream after their patch rows are written), so a
        consumer on another stream can start on the first images while the rest is still in flight
        (``CoalitionEngine.chunk_ready``)."""
        cfg = self.cfg
        own = plan is None
        with torch.cuda.device(self.device):
            if own:
                plan = ops.Plan(cfg, self.precision, 1, 1, self.device)
            step = self.UPLOAD_STEP
            self.labels = labels.to(self.device, dtype=torch.int64, non_blocking=True).contiguous()
            for s in range(0, self.n, step):
                img = images[s:s + step].to(self.device, dtype=torch.float32, non_blocking=True)
                plan.patchify(img, out=self.patches, row0=s * cfg.n_patches)
                if events is not None:
                    ev = torch.cuda.Event()
                    ev.record()
                    events.append(ev)
            if own:
                torch.cuda.current_stream().synchronize()
                plan.close()
        return images.numel() * 4 + labels.numel() * 8

    @staticmethod
    def from_loader(cfg: VitConfig, loader, precision: int, device="cuda:0") -> "ValidationSet":
        """Drain a DataLoader of the reference's dict samples {'image','label',...}
        (federated_learning/utils.py:880) or of (x, y) tuples."""
        imgs, labs = [], []
        for sample in loader:
            if isinstance(sample, dict):
                x, y = sample["image"], sample["label"]
            else:
                x, y = sample[0], sample[1]
            imgs.append(torch.as_tensor(x, dtype=torch.float32))
            labs.append(torch.as_tensor(y).long().reshape(-1))
        return ValidationSet(cfg, torch.cat(imgs), torch.cat(labs), precision, device)


class CoalitionEngine:
    def __init__(self, cfg: VitConfig, w0, deltas, images, labels: Optional[torch.Tensor] = None,
                 precision: str = _lib.DEFAULT_PRECISION, coalition_batch: int = 8, image_chunk: int = 128,
                 device: str | torch.device = "cuda:0", keep_logits: bool = False, n_clients: Optional[int] = None):
        """``images`` is either a host tensor [n, C, H, W] (with ``labels``) or a ValidationSet.
        ``deltas=None`` (with ``n_clients``): no single-round stack is resident -- the engine serves the multi-round
        mode (``set_round_deltas`` / ``evaluate_rounds``) and, for all-zero ratio rows, W_0 itself."""
        if not torch.cuda.is_available():
            raise RuntimeError("CoalitionEngine needs a CUDA device (sm_100a); there is no CPU path")
        self.cfg = cfg
        self.device = torch.device(device)
        self.lay: PlanLayout = plan_layout(cfg)
        self.precision = _lib.PRECISIONS[precision] if isinstance(precision, str) else int(precision)
        if deltas is None:
            if n_clients is None:
                raise ValueError("deltas=None needs n_clients")
            self.n_clients = int(n_clients)
        else:
            self.n_clients = int(deltas.shape[0]) if isinstance(deltas, torch.Tensor) else len(deltas)
        if not 1 <= self.n_clients <= 64:
            raise ValueError("1..64 clients supported")
        self.n_val = images.n if isinstance(images, ValidationSet) else int(images.shape[0])
        self.coalition_batch = max(1, int(coalition_batch))
        self.image_chunk = max(1, min(int(image_chunk), self.n_val))
        self.keep_logits = keep_logits
        self.last_logits: Optional[torch.Tensor] = None
        self.kernel_launches = 0
        self.profile = False                      # bench.py: CUDA-event pairs around the K1 launches
        self.chunk_ready = None                   # optional callable(lo, hi) run before the forward of images [lo, hi)
        self.agg_spans: List[Tuple[torch.cuda.Event, torch.cuda.Event]] = []
        self.round_deltas: List[torch.Tensor] = []   # multi-round mode (set_round_deltas)
        self._partial: Optional[torch.Tensor] = None
        with torch.cuda.device(self.device):
            self.plan = ops.Plan(cfg, self.precision, self.coalition_batch, self.image_chunk, self.device)
            total = self.lay.total
            # stacked deltas [N, total] and W0 [total], plan layout, fp32, resident.  Already
            # packed tensors (e.g. received by NCCL broadcast) are taken as they are.
            self._zero_row: Optional[torch.Tensor] = None
            if deltas is None:
                self.deltas = None
            elif isinstance(deltas, torch.Tensor):
                if deltas.shape != (self.n_clients, total) or deltas.dtype != torch.float32:
                    raise ValueError(f"packed deltas must be fp32 [N, {total}]")
                self.deltas = deltas.to(self.device).contiguous()
            else:
                host = torch.empty((self.n_clients, total), dtype=torch.float32, pin_memory=True)
                for j, d in enumerate(deltas):
                    pack_state_dict(self.lay, d, out=host[j])
                self.deltas = host.to(self.device, non_blocking=True)
            if isinstance(w0, torch.Tensor):
                self.w0 = w0.to(self.device, dtype=torch.float32).contiguous()
            elif w0 is not None:
                self.w0 = pack_state_dict(self.lay, w0).to(self.device)
            else:
                self.w0 = None
            self.val = images if isinstance(images, ValidationSet) else ValidationSet(
                cfg, images, labels, self.precision, self.device)
            if self.val.precision != self.precision or self.val.device != self.device:
                raise ValueError("ValidationSet was built for another precision/device")
            cb = self.coalition_batch
            self.wvec = torch.empty((cb, self.lay.vec_size), dtype=torch.float32, device=self.device)
            self.wmat = self.plan.operand_array((cb, self.lay.mat_size))   # aggregated weight matrices, operand format
            self.logits = torch.empty((cb, self.n_val, cfg.n_cls), dtype=torch.float32, device=self.device)
            torch.cuda.synchronize(self.device)

    # ------------------------------------------------------------------ #
    @property
    def patches(self) -> torch.Tensor:
        return self.val.patches

    @property
    def labels(self) -> torch.Tensor:
        return self.val.labels

    def upload_bytes_per_batch(self) -> int:
        return self.coalition_batch * self.n_clients * 4

    # ------------------------------------------------------------------ #
    def _aggregate_batch(self, ratios: torch.Tensor, Cn: int) -> None:
        """K1: W_S = W_0 + sum_j r_j Delta_j for the batch's coalitions into wvec / wmat (both regions)."""
        V, Mz = self.lay.vec_size, self.lay.mat_size
        w0v = self.w0[:V] if self.w0 is not None else None
        w0m = self.w0[V:] if self.w0 is not None else None
        deltas = self.deltas
        if deltas is None:
            # no single-round stack: only W_0 itself can be asked for (all-zero rows = no member).  K1 skips
            # zero-ratio clients, so one zero placeholder row serves any N.
            if bool((ratios != 0).any()):
                raise ValueError("this engine holds no single-round deltas (deltas=None): use evaluate_rounds")
            if self._zero_row is None:
                self._zero_row = torch.zeros((1, self.lay.total), dtype=torch.float32, device=self.device)
            deltas, ratios = self._zero_row, torch.zeros((Cn, 1), dtype=torch.float32)
        ops.aggregate(deltas[:, :V], w0v, ratios, out=self.wvec[:Cn], P=V)
        ops.aggregate(deltas[:, V:], w0m, ratios, out=self.wmat, P=Mz)

    def _forward(self, Cn: int, row0: int, n_images: int, logits: torch.Tensor, image_offset: int) -> None:
        """One svit_forward_batched launch sequence (overridden by the shared-weight LoRA engine)."""
        self.plan.forward(self.wvec[:Cn], self.wmat, self.patches, row0, n_images, logits, image_offset=image_offset)

    def _run_batch(self, ratio_rows: Sequence[Sequence[float]], image_range: Optional[Tuple[int, int]] = None,
                   records: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """One batch of <= coalition_batch coalitions given their dense ratio rows, scored on the validation
        images [lo, hi) (default: all of them).  With ``records`` (int64 [Cn, 2] device view, e.g. a slice of the
        all-gather send buffer) K5 writes the packed (correct, loss_sum) records there and nothing is returned."""
        Cn = len(ratio_rows)
        lay, cfg = self.lay, self.cfg
        # FedAvg ratios: fp64 on the host (as the reference computes them), rounded once to fp32;
        # svit_aggregate copies them into its kernel parameters (no separate H2D copy)
        ratios = torch.as_tensor(ratio_rows, dtype=torch.float64).to(torch.float32)
        if self.profile:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        self._aggregate_batch(ratios, Cn)
        if self.profile:
            e1.record()
            self.agg_spans.append((e0, e1))
        logits = self.logits[:Cn]
        npch = cfg.n_patches
        lo, hi = image_range if image_range is not None else (0, self.n_val)
        for s in range(lo, hi, self.image_chunk):
            b = min(self.image_chunk, hi - s)
            if self.chunk_ready is not None:
                self.chunk_ready(s, s + b)        # e.g. wait for the upload events covering these images
            self._forward(Cn, s * npch, b, logits, s)
        self.kernel_launches += 2 + ((hi - lo + self.image_chunk - 1) // self.image_chunk) * (3 + 7 * cfg.layers) + 1
        if records is not None:
            if hi > lo:
                ops.score_records(logits[:, lo:hi], self.labels[lo:hi], records)
            else:
                records.zero_()
            if self.keep_logits:
                self.last_logits = logits.clone()
            return None, None
        if (lo, hi) == (0, self.n_val):
            correct, loss = ops.score(logits, self.labels)
        elif hi > lo:
            correct, loss = ops.score(logits[:, lo:hi], self.labels[lo:hi])
        else:
            correct = torch.zeros(Cn, dtype=torch.int64, device=self.device)
            loss = torch.zeros(Cn, dtype=torch.float64, device=self.device)
        if self.keep_logits:
            self.last_logits = logits.clone()
        return correct, loss

    # ------------------------------------------------------------------ #
    # multi-round "lazy" reconstruction (reference utils_fed_shapley.py:146-196 + server2.py:121-127):
    #   W_S = W_0 + agg_{S, t0} + agg_{S, t1} + ...   one FedAvg aggregate per FL round, added in round order
    def set_round_deltas(self, round_deltas) -> None:
        """``round_deltas[t]``: the clients' deltas of FL round t -- a packed fp32 tensor [N, total] or a list of N
        state-dict-like deltas (None = client absent in that round: a zero row, never selected)."""
        packed = []
        with torch.cuda.device(self.device):
            for rd in round_deltas:
                if isinstance(rd, torch.Tensor):
                    if rd.shape != (self.n_clients, self.lay.total) or rd.dtype != torch.float32:
                        raise ValueError(f"packed round deltas must be fp32 [N, {self.lay.total}]")
                    packed.append(rd.to(self.device).contiguous())
                else:
                    host = torch.zeros((self.n_clients, self.lay.total), dtype=torch.float32)
                    for j, d in enumerate(rd):
                        if d is not None:
                            pack_state_dict(self.lay, d, out=host[j])
                    packed.append(host.to(self.device))
        self.round_deltas = packed
        self._partial = None

    def _run_batch_rounds(self, rows_per_round: Sequence[Sequence[Sequence[float]]]):
        """One batch of coalitions; ``rows_per_round[t][c]`` = dense FedAvg ratio row of coalition c in round t
        (all zeros when none of its members was selected in that round)."""
        R = len(rows_per_round)
        if R != len(self.round_deltas):
            raise ValueError("one ratio table per round expected (set_round_deltas first)")
        if R == 0:
            raise ValueError("no rounds: use evaluate() with all-zero rows (the model is W_0)")
        Cn = len(rows_per_round[0])
        lay, cfg = self.lay, self.cfg
        V, Mz = lay.vec_size, lay.mat_size
        if self._partial is None:   # fp32 partial models [coalition_batch, total], W_0 + the rounds so far
            self._partial = torch.empty((self.coalition_batch, lay.total), dtype=torch.float32, device=self.device)
        part = self._partial[:Cn]
        if self.w0 is not None:
            part.copy_(self.w0.unsqueeze(0).expand(Cn, -1))
        else:
            part.zero_()
        for t in range(R):
            ratios = torch.as_tensor(rows_per_round[t], dtype=torch.float64).to(torch.float32)
            D = self.round_deltas[t]
            if t < R - 1:
                ops.aggregate_onto(D, part, ratios, part)                       # fp32, in place, two-rounding exact
            else:                                                                # last round: straight into the plan buffers
                ops.aggregate_onto(D[:, :V], part[:, :V], ratios, self.wvec[:Cn], P=V)
                ops.aggregate_onto(D[:, V:], part[:, V:], ratios, self.wmat, P=Mz)
        logits = self.logits[:Cn]
        npch = cfg.n_patches
        for s in range(0, self.n_val, self.image_chunk):
            b = min(self.image_chunk, self.n_val - s)
            self.plan.forward(self.wvec[:Cn], self.wmat, self.patches, s * npch, b, logits, image_offset=s)
        correct, loss = ops.score(logits, self.labels)
        if self.keep_logits:
            self.last_logits = logits.clone()
        return correct, loss

    def evaluate_rounds(self, rows_per_round: Sequence[Sequence[Sequence[float]]],
                        n_coalitions: Optional[int] = None) -> Tuple[List[int], List[float]]:
        """Like ``evaluate`` for models reconstructed from several FL rounds: ``rows_per_round[t][c]``.
        No round at all (``include_from_round > current_round`` in the reference's loop,
        utils_fed_shapley.py:166-176): every model is W_0; ``n_coalitions`` says how many."""
        if len(rows_per_round) == 0:
            return self.evaluate([[0.0] * self.n_clients] * int(n_coalitions or 0))
        n = len(rows_per_round[0])
        correct: List[int] = []
        loss: List[float] = []
        with torch.cuda.device(self.device):
            pending = []
            for s in range(0, n, self.coalition_batch):
                pending.append(self._run_batch_rounds([rt[s:s + self.coalition_batch] for rt in rows_per_round]))
            for c, l in pending:
                correct += c.cpu().tolist()
                loss += l.cpu().tolist()
        return correct, loss

    def aggregated_rows_rounds(self, rows_per_round) -> torch.Tensor:
        """Reconstructed multi-round models (plan layout, [C, total] fp32) for parity checks."""
        with torch.cuda.device(self.device):
            Cn = len(rows_per_round[0])
            part = (self.w0.unsqueeze(0).expand(Cn, -1).clone() if self.w0 is not None
                    else torch.zeros((Cn, self.lay.total), dtype=torch.float32, device=self.device))
            for t, rows in enumerate(rows_per_round):
                ops.aggregate_onto(self.round_deltas[t], part, torch.as_tensor(rows, dtype=torch.float64).to(torch.float32), part)
            return part

    def ratio_row(self, members: Sequence[int], ratios: Sequence[float]) -> List[float]:
        row = [0.0] * self.n_clients
        for j, r in zip(members, ratios):
            row[int(j)] = float(r)
        return row

    def evaluate(self, ratio_rows: Sequence[Sequence[float]],
                 image_range: Optional[Tuple[int, int]] = None) -> Tuple[List[int], List[float]]:
        """Evaluate coalitions given as dense FedAvg ratio rows [n_clients] (0 = non-member).
        Returns per coalition (#correct, sum of cross-entropy) over the validation set, or over its
        images [lo, hi) when ``image_range`` is given (the validation-split axis of dist.sharded_evaluate)."""
        correct: List[int] = []
        loss: List[float] = []
        with torch.cuda.device(self.device):
            pending = []
            for s in range(0, len(ratio_rows), self.coalition_batch):
                c, l = self._run_batch(ratio_rows[s:s + self.coalition_batch], image_range)
                pending.append((c, l))
            for c, l in pending:       # single host sync at the end
                correct += c.cpu().tolist()
                loss += l.cpu().tolist()
        return correct, loss

    def evaluate_into(self, ratio_rows: Sequence[Sequence[float]], records: torch.Tensor,
                      image_range: Optional[Tuple[int, int]] = None) -> None:
        """``evaluate`` without the host read-back: the packed per-coalition records land in ``records``
        (int64 [len(ratio_rows), 2] on this device), asynchronously on the current stream."""
        with torch.cuda.device(self.device):
            for s in range(0, len(ratio_rows), self.coalition_batch):
                self._run_batch(ratio_rows[s:s + self.coalition_batch], image_range, records=records[s:s + self.coalition_batch])

    def evaluate_state_dict(self, sd: Dict[str, torch.Tensor]) -> Tuple[int, float]:
        """Score one explicit model (the reference's plain ``evaluation(args, net, loader)``)."""
        with torch.cuda.device(self.device):
            row = pack_state_dict(self.lay, sd).to(self.device).unsqueeze(0)
            one = torch.ones((1, 1), dtype=torch.float32)
            V = self.lay.vec_size
            ops.aggregate(row[:, :V], None, one, out=self.wvec[:1], P=V)
            ops.aggregate(row[:, V:], None, one, out=self.wmat, P=self.lay.mat_size)
            logits = self.logits[:1]
            npch = self.cfg.n_patches
            for s in range(0, self.n_val, self.image_chunk):
                b = min(self.image_chunk, self.n_val - s)
                self.plan.forward(self.wvec[:1], self.wmat, self.patches, s * npch, b, logits, image_offset=s)
            correct, loss = ops.score(logits, self.labels)
            if self.keep_logits:
                self.last_logits = logits.clone()
            return int(correct.item()), float(loss.item())

    def aggregated_rows(self, ratio_rows: Sequence[Sequence[float]], dtype=torch.float32) -> torch.Tensor:
        """Aggregated models (plan layout, [C, total]) for parity checks of K1."""
        with torch.cuda.device(self.device):
            r = torch.as_tensor(ratio_rows, dtype=torch.float64).to(torch.float32)
            return ops.aggregate(self.deltas, self.w0, r, out_dtype=dtype)
