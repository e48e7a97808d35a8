"""LoRA-aware coalition path: the reference author's real setup.

``start.py:274-283`` wraps the HF ViT with PEFT LoRA (r = 16, alpha = 8) on the ``query`` and
``value`` projections and trains the classifier fully (``modules_to_save``); every other tensor is
frozen, so between clients only ~0.6 M of 86 M entries differ.  The reference still FedAvg-averages
the whole state_dict entry by entry -- in particular **A and B separately**:

    A_S = A_0 + sum_j r_j dA_j        B_S = B_0 + sum_j r_j dB_j        (get_aggregated_model, utils.py:781-792)

and the model it scores computes, per wrapped projection (PEFT ``lora.Linear.forward``, eval mode),

    y = x W^T + b + (alpha / r) * (x A_S^T) B_S^T

which is the dense projection with  W_S = W + (alpha / r) * B_S A_S  -- not linear in the coalition.

Here: K1 aggregates the packed LoRA rows (A stored transposed, B pre-scaled by alpha / r) for the
whole coalition batch, ONE grouped fp32 GEMM over (coalition, layer, target) forms
``W + (alpha/r) B_S A_S`` for every wrapped projection, the blocks are written into the weight
matrix region in the operand dtype, and the usual batched forward runs.  When the base weights are
frozen (all base deltas zero, the author's case) the 343 MB matrix region is never re-aggregated:
its rows are filled with W_0 once and only the query / value blocks change per batch.

PEFT is not installed in the build container, so the oracle for this row is a restatement of the
published algorithm (``oracle/restate.py::vit_forward(..., lora=...)``): parity unpinned.
Key conventions accepted (PEFT 0.5 - 0.13): optional ``module.`` / ``base_model.model.`` prefixes,
``<proj>.base_layer.{weight,bias}`` or ``<proj>.{weight,bias}``, ``<proj>.lora_{A,B}.<adapter>.weight``,
``classifier.modules_to_save.<adapter>.*`` (used) and ``classifier.original_module.*`` (ignored).
"""
from __future__ import annotations

import re
from collections import OrderedDict
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _lib, ops
from .engine import CoalitionEngine
from .fl import _clean_keys
from .layout import K_WQ, K_WV, VitConfig

TARGETS = ("query", "value")
_LORA_KEY = re.compile(r"^vit\.encoder\.layer\.(\d+)\.attention\.attention\.(query|key|value)\.lora_([AB])\.[^.]+\.weight$")
LoraDict = Dict[Tuple[int, str, str], torch.Tensor]   # (layer, target, 'A' | 'B') -> tensor


def is_lora_state_dict(sd) -> bool:
    return any(".lora_A." in k for k in sd.keys())


def split_state_dict(sd) -> Tuple["OrderedDict[str, torch.Tensor]", LoraDict]:
    """PEFT-wrapped ViT state_dict -> (plain HF-keyed state_dict of the base model, LoRA factors)."""
    hf: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    lora: LoraDict = {}
    for k, v in _clean_keys(sd).items():
        m = _LORA_KEY.match(k)
        if m:
            if m.group(2) not in TARGETS:
                raise ValueError(f"LoRA on '{m.group(2)}' is not supported (query / value only, start.py:275)")
            lora[(int(m.group(1)), m.group(2), m.group(3))] = v
            continue
        if ".lora_" in k:
            raise ValueError(f"unsupported LoRA entry {k}")
        if k.startswith("classifier.original_module."):
            continue
        k = re.sub(r"^classifier\.modules_to_save\.[^.]+\.", "classifier.", k.replace(".base_layer.", "."))
        hf[k] = v
    return hf, lora


def lora_rank(lora: LoraDict) -> int:
    ranks = {v.shape[0] for (_, _, ab), v in lora.items() if ab == "A"}
    if len(ranks) != 1:
        raise ValueError(f"one LoRA rank expected, got {sorted(ranks)}")
    return ranks.pop()


def pack_lora(cfg: VitConfig, lora: LoraDict, r: int, scaling: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One fp32 row [layers, 2 targets, 2, h, r]: A^T and scaling * B per wrapped projection (zeros if absent)."""
    h = cfg.hidden
    if out is None:
        out = torch.zeros(cfg.layers * 2 * 2 * h * r, dtype=torch.float32)
    else:
        out.zero_()
    view = out.view(cfg.layers, 2, 2, h, r)
    for (layer, target, ab), v in lora.items():
        t = TARGETS.index(target)
        v = v.detach().to(torch.float32)
        if ab == "A":
            if tuple(v.shape) != (r, h):
                raise ValueError(f"lora_A of layer {layer} {target}: expected {(r, h)}, got {tuple(v.shape)}")
            view[layer, t, 0].copy_(v.t())
        else:
            if tuple(v.shape) != (h, r):
                raise ValueError(f"lora_B of layer {layer} {target}: expected {(h, r)}, got {tuple(v.shape)}")
            view[layer, t, 1].copy_(v * scaling)
    return out


EXT_K = 64   # kGemmExtK: columns of the K-extension block (query factors in 0 .. r-1, value factors in 32 .. 32+r-1)


def pack_lora_ext(cfg: VitConfig, lora: LoraDict, r: int, scaling: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One fp32 row in the layout svit_forward_lora_batched reads (include/svit.h): per layer Acat [64, h] (rows 0 .. r-1 =
    A_q, rows 32 .. 32+r-1 = A_v) then Bext [3h, 64] (query rows: columns 0 .. r-1 = scaling * B_q; value rows: columns
    32 .. = scaling * B_v).  The zero padding is part of the row, so K1 on these rows (entry-wise FedAvg of A and of B,
    as the reference aggregates the PEFT state_dict) yields the operand arrays directly."""
    h = cfg.hidden
    if r > 32:
        raise ValueError("LoRA rank <= 32 supported by the K-extension path")
    per = EXT_K * h * 4
    if out is None:
        out = torch.zeros(cfg.layers * per, dtype=torch.float32)
    else:
        out.zero_()
    for (layer, target, ab), v in lora.items():
        v = v.detach().to(torch.float32)
        c0 = 0 if target == "query" else 32
        base = layer * per
        if ab == "A":
            if tuple(v.shape) != (r, h):
                raise ValueError(f"lora_A of layer {layer} {target}: expected {(r, h)}, got {tuple(v.shape)}")
            out[base:base + EXT_K * h].view(EXT_K, h)[c0:c0 + r].copy_(v)
        else:
            if tuple(v.shape) != (h, r):
                raise ValueError(f"lora_B of layer {layer} {target}: expected {(h, r)}, got {tuple(v.shape)}")
            rows0 = 0 if target == "query" else 2 * h
            out[base + EXT_K * h:base + per].view(3 * h, EXT_K)[rows0:rows0 + h, c0:c0 + r].copy_(v * scaling)
    return out


def merged_state_dict(sd, lora_alpha: float = 8.0) -> "OrderedDict[str, torch.Tensor]":
    """Plain HF-keyed state_dict of ONE PEFT-LoRA model with W + (alpha / r) B A folded into the wrapped
    projections (host fp32): what scoring a single explicit model needs (``fl.evaluation``)."""
    hf, lo = split_state_dict(sd)
    if not lo:
        return hf
    scaling = float(lora_alpha) / lora_rank(lo)
    out = OrderedDict((k, v) for k, v in hf.items())
    for (layer, target, ab), a in lo.items():
        if ab != "A":
            continue
        b = lo[(layer, target, "B")]
        key = f"vit.encoder.layer.{layer}.attention.attention.{target}.weight"
        out[key] = hf[key].to(torch.float32) + scaling * (b.to(torch.float32) @ a.to(torch.float32))
    return out


class LoraCoalitionEngine(CoalitionEngine):
    """CoalitionEngine for PEFT-LoRA client models: same evaluate()/Game interface, state_dicts with LoRA keys."""

    def __init__(self, cfg: VitConfig, w0_sd, delta_sds: Sequence[dict], images, labels=None, lora_alpha: float = 8.0,
                 shared_base: Optional[bool] = None, **kw):
        """Two forwards for frozen-base LoRA clients (both need rank <= 32 / hidden % 64 == 0 for the first):
        * shared (svit_forward_lora_batched): ONE weight region for every coalition, the
          aggregated factors as a K-extension of the QKV GEMM; K1 touches ~2 % of a dense model's bytes and the
          per-coalition weights (343 MB each for ViT-B) are never formed;
        * dense: W + (alpha/r) B_S A_S merged per coalition into the ordinary grouped forward (also the general path
          when the clients trained the base).
        *Measured* (ViT-B, rank 16, 8 coalitions x 2 048 images, f16c8, one session): dense 6.69 evals/s, shared 6.33,
        the plain dense model 6.50 -- the grouped GEMMs are tensor-bound, not weight-bandwidth-bound, so sharing B buys
        nothing there while the extension adds 1/12 of a k-loop plus the rank-64 projection to every QKV GEMM; the
        shared path wins where the merge dominates (short validation sets) or memory is tight.
        ``shared_base``: True / False force a path; None picks shared for <= 256 validation images, dense above."""
        hf0, l0 = split_state_dict(w0_sd)
        parts = [split_state_dict(d) for d in delta_sds]
        self.r = lora_rank(l0)
        self.scaling = float(lora_alpha) / self.r
        super().__init__(cfg, hf0, [p[0] for p in parts], images, labels, **kw)
        h, L, r = cfg.hidden, cfg.layers, self.r
        self.n_proj = L * 2
        with torch.cuda.device(self.device):
            n = self.n_clients
            host = torch.empty((n + 1, L * 2 * 2 * h * r), dtype=torch.float32, pin_memory=True)
            pack_lora(cfg, l0, r, self.scaling, out=host[n])
            for j, p in enumerate(parts):
                pack_lora(cfg, p[1], r, self.scaling, out=host[j])
            dev = host.to(self.device)
            self.lora_deltas, self.lora_w0 = dev[:n], dev[n]
            cb, V = self.coalition_batch, self.lay.vec_size
            self.lora_s = torch.empty((cb, host.shape[1]), dtype=torch.float32, device=self.device)
            self.qv32 = torch.empty((cb, self.n_proj, h * h), dtype=torch.float32, device=self.device)
            # element offsets of the wrapped projections inside the matrix region, (layer, target) order
            self.qv_off = [self.lay.find(kind, l).offset for l in range(L) for kind in (K_WQ, K_WV)]
            # the author's case: nothing but LoRA (and the vec region: classifier, biases) differs between clients
            self.base_frozen = not bool(self.deltas[:, V:].any().item())
            self._eye = torch.eye(cb, dtype=torch.float32)               # K1 as a row-wise cast: out[c] = 1 * rows[c]
            want_shared = shared_base if shared_base is not None else self.n_val <= 256
            self.shared = bool(want_shared) and self.base_frozen and self.r <= 32 and cfg.hidden % 64 == 0
            if self.shared:
                # N1: the K / out-proj / MLP weights (and W_q, W_v under the LoRA term) are the same
                # for every coalition: ONE mat-region row, shared B operands; the factors ride the QKV GEMM as a K-extension
                ext = torch.empty((n + 1, L * EXT_K * h * 4), dtype=torch.float32, pin_memory=True)
                pack_lora_ext(cfg, l0, r, self.scaling, out=ext[n])
                for j, p in enumerate(parts):
                    pack_lora_ext(cfg, p[1], r, self.scaling, out=ext[j])
                edev = ext.to(self.device)
                self.ext_deltas, self.ext_w0 = edev[:n], edev[n]
                self.lora_rows = self.plan.operand_array((cb, ext.shape[1]))
                self.wmat_shared = self.plan.operand_array((1, self.lay.mat_size))
                ops.aggregate(self.w0[V:].unsqueeze(0), None, torch.ones((1, 1)), out=self.wmat_shared, P=self.lay.mat_size)
            elif self.base_frozen:
                ops.aggregate(self.w0[V:].unsqueeze(0), None, torch.ones((cb, 1)), out=self.wmat, P=self.lay.mat_size)
                self.qv_base = torch.stack([self.w0[V + o:V + o + h * h] for o in self.qv_off])   # [n_proj, h*h] fp32
            torch.cuda.synchronize(self.device)

    def _aggregate_batch(self, ratios: torch.Tensor, Cn: int) -> None:
        cfg, lay = self.cfg, self.lay
        h, r, V, Mz = cfg.hidden, self.r, lay.vec_size, lay.mat_size
        hh = h * h
        ops.aggregate(self.deltas[:, :V], self.w0[:V], ratios, out=self.wvec[:Cn], P=V)
        if self.shared:
            # entry-wise FedAvg of the factors (A and B separately, as the reference averages the PEFT state_dict),
            # written straight as the operand rows of the K-extension: one K1 launch, ~2 % of a dense model's bytes
            ops.aggregate(self.ext_deltas, self.ext_w0, ratios, out=self.lora_rows)
            self.kernel_launches += 2
            return
        qv = self.qv32[:Cn]
        if self.base_frozen:
            qv.copy_(self.qv_base)                                   # W_0 blocks, shared by every coalition
        else:
            ops.aggregate(self.deltas[:, V:], self.w0[V:], ratios, out=self.wmat, P=Mz)
            for i, o in enumerate(self.qv_off):                       # the wrapped projections again, in fp32
                ops.aggregate(self.deltas[:, V + o:V + o + hh], self.w0[V + o:V + o + hh], ratios, out=qv[:, i], P=hh)
        # A_S and B_S, averaged separately (K1 on the packed LoRA rows), then W + (alpha/r) B_S A_S per projection
        ops.aggregate(self.lora_deltas, self.lora_w0, ratios, out=self.lora_s[:Cn])
        f = self.lora_s[:Cn].view(Cn * self.n_proj, 2, h, r)
        w = qv.view(Cn * self.n_proj, h, h)
        ops.gemm(_lib.PREC_F32, f[:, 1], f[:, 0], residual=w, out=w, out_dtype=torch.float32)
        eye = self._eye[:Cn, :Cn]
        for i, o in enumerate(self.qv_off):                            # fp32 blocks -> operand dtype, in place in wmat
            ops.aggregate(qv[:, i], None, eye, out=self.wmat, P=hh, col0=o)
        self.kernel_launches += 2 + self.n_proj + (0 if self.base_frozen else 1 + self.n_proj)

    def _forward(self, Cn: int, row0: int, n_images: int, logits: torch.Tensor, image_offset: int) -> None:
        if self.shared:
            self.plan.forward_lora(self.wvec[:Cn], self.wmat_shared, self.lora_rows, self.patches, row0, n_images, logits,
                                   image_offset=image_offset)
        else:
            super()._forward(Cn, row0, n_images, logits, image_offset)

    def evaluate_state_dict(self, sd) -> Tuple[int, float]:
        """Score one explicit model; PEFT-keyed state_dicts are merged on the host first."""
        shared, self.shared = self.shared, False      # an explicit dense model goes through the dense forward
        try:
            out = super().evaluate_state_dict(merged_state_dict(sd, self.scaling * self.r) if is_lora_state_dict(sd) else sd)
        finally:
            self.shared = shared
        if self.base_frozen and not self.shared:   # the call went through wmat[0]: restore the W_0 rows the frozen-base path relies on
            V = self.lay.vec_size
            ops.aggregate(self.w0[V:].unsqueeze(0), None, torch.ones((1, 1)), out=self.wmat, P=self.lay.mat_size)
        return out

    def merged_rows(self, ratio_rows) -> torch.Tensor:
        """fp32 [C, n_proj, h, h]: the merged query / value weights of each coalition (tests)."""
        ratios = torch.as_tensor(ratio_rows, dtype=torch.float64).to(torch.float32)
        self._aggregate_batch(ratios, len(ratio_rows))
        h = self.cfg.hidden
        return self.qv32[:len(ratio_rows)].view(len(ratio_rows), self.n_proj, h, h).clone()
