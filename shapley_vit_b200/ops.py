"""torch-tensor front end of the C ABI: every function here takes CUDA tensors, passes
their ``data_ptr()`` and the current CUDA stream to ``libsvit.so`` and returns tensors.
PyTorch is used for device memory and streams only -- no torch operator computes anything
on these paths, and there is no fallback if the library is missing."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import BF16, F16, F32, PREC_F16X3, EpilogueC, check

TORCH_DTYPE = {F32: torch.float32, BF16: torch.bfloat16, F16: torch.float16}
SVIT_DTYPE = {v: k for k, v in TORCH_DTYPE.items()}


def _stream(t: torch.Tensor) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _cuda(t: torch.Tensor, name: str, dtype=None) -> None:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (libsvit has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")


def aggregate(deltas: torch.Tensor, w0: Optional[torch.Tensor], ratios: torch.Tensor,
              out_dtype: torch.dtype = torch.float32, out: Optional[torch.Tensor] = None,
              P: Optional[int] = None) -> torch.Tensor:
    """K1.  deltas [N, stride] fp32 (row-contiguous), w0 [>=P] fp32 or None, ratios [C, N] fp32
    on the HOST (0 = non-member; a CUDA tensor is copied back, which synchronises).
    Returns out [C, out_stride] with out[c, :P] = w0 + sum_j ratios[c, j] * deltas[j]."""
    _cuda(deltas, "deltas", torch.float32)
    ratios = ratios.detach().to("cpu", torch.float32)
    if deltas.dim() != 2 or deltas.stride(1) != 1:
        raise ValueError("deltas must be [N, P] with unit inner stride")
    N, width = deltas.shape
    P = width if P is None else int(P)
    ratios = ratios.contiguous()
    Cn = ratios.shape[0]
    if ratios.shape != (Cn, N):
        raise ValueError(f"ratios must be [C, {N}]")
    if w0 is not None:
        _cuda(w0, "w0", torch.float32)
    if out is None:
        stride = (P + 7) // 8 * 8
        out = torch.empty((Cn, stride), dtype=out_dtype, device=deltas.device)
    _cuda(out, "out")
    if out.dim() != 2 or out.stride(1) != 1 or out.shape[0] != Cn:
        raise ValueError("out must be [C, >=P] with unit inner stride")
    check(_lib.load().svit_aggregate(_ptr(deltas), deltas.stride(0), _ptr(w0), _ptr(ratios), _ptr(out), out.stride(0),
                                     SVIT_DTYPE[out.dtype], P, N, Cn, _stream(deltas)))
    return out


def aggregate_onto(deltas: torch.Tensor, base: torch.Tensor, ratios: torch.Tensor, out: torch.Tensor,
                   P: Optional[int] = None) -> torch.Tensor:
    """One more FL round folded onto per-coalition partial models (svit_aggregate_onto):
    out[c, :P] = cast(base[c, :P] + sum_j ratios[c, j] * deltas[j]).  base [C, stride] fp32 (may be ``out``
    itself when ``out`` is fp32), ratios [C, N] on the host."""
    _cuda(deltas, "deltas", torch.float32)
    _cuda(base, "base", torch.float32)
    _cuda(out, "out")
    ratios = ratios.detach().to("cpu", torch.float32).contiguous()
    N, width = deltas.shape
    P = width if P is None else int(P)
    Cn = ratios.shape[0]
    if ratios.shape != (Cn, N) or base.shape[0] != Cn or out.shape[0] != Cn:
        raise ValueError("ratios must be [C, N]; base and out must have C rows")
    if deltas.stride(1) != 1 or base.stride(1) != 1 or out.stride(1) != 1:
        raise ValueError("deltas / base / out must have unit inner stride")
    check(_lib.load().svit_aggregate_onto(_ptr(deltas), deltas.stride(0), _ptr(base), base.stride(0), _ptr(ratios), _ptr(out),
                                          out.stride(0), SVIT_DTYPE[out.dtype], P, N, Cn, _stream(deltas)))
    return out


def score(logits: torch.Tensor, labels: torch.Tensor, correct: Optional[torch.Tensor] = None,
          loss_sum: Optional[torch.Tensor] = None, accumulate: bool = False, want_pred: bool = False):
    """K5.  logits [C, n, n_cls] fp32, labels [n] int64 -> (correct int64 [C], loss_sum fp64 [C][, pred int32 [C, n]])."""
    _cuda(logits, "logits", torch.float32)
    _cuda(labels, "labels", torch.int64)
    if logits.dim() != 3 or not logits[0].is_contiguous():
        raise ValueError("logits must be [C, n, n_cls] with contiguous [n, n_cls] slabs")
    Cn, n, n_cls = logits.shape
    if labels.shape != (n,):
        raise ValueError("labels must be [n]")
    labels = labels.contiguous()
    if correct is None:
        correct = torch.zeros(Cn, dtype=torch.int64, device=logits.device)
        loss_sum = torch.zeros(Cn, dtype=torch.float64, device=logits.device)
    pred = torch.empty((Cn, n), dtype=torch.int32, device=logits.device) if want_pred else None
    check(_lib.load().svit_score(_ptr(logits), logits.stride(0) if Cn > 1 else n * n_cls, _ptr(labels), Cn, n, n_cls,
                                 _ptr(correct), _ptr(loss_sum), _ptr(pred), n, int(accumulate), _stream(logits)))
    return (correct, loss_sum, pred) if want_pred else (correct, loss_sum)


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float,
              out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """x [G, rows, h] fp32, gamma/beta [G, h] fp32 -> y [G, rows, h]."""
    _cuda(x, "x", torch.float32)
    G, rows, h = x.shape
    x, gamma, beta = x.contiguous(), gamma.contiguous(), beta.contiguous()
    y = torch.empty((G, rows, h), dtype=out_dtype, device=x.device)
    check(_lib.load().svit_layernorm(_ptr(x), rows * h, h, _ptr(gamma), _ptr(beta), h, _ptr(y), rows * h, h,
                                     SVIT_DTYPE[out_dtype], G, rows, h, float(eps), _stream(x)))
    return y


def attention(qkv: torch.Tensor, heads: int) -> torch.Tensor:
    """qkv [n_seq, T, 3h] -> ctx [n_seq, T, h] (same dtype)."""
    _cuda(qkv, "qkv")
    qkv = qkv.contiguous()
    n_seq, T, h3 = qkv.shape
    h = h3 // 3
    ctx = torch.empty((n_seq, T, h), dtype=qkv.dtype, device=qkv.device)
    check(_lib.load().svit_attention(_ptr(qkv), _ptr(ctx), SVIT_DTYPE[qkv.dtype], n_seq, T, heads, h // heads,
                                     _stream(qkv)))
    return ctx


def attention_f16x3(qkv: torch.Tensor, heads: int) -> torch.Tensor:
    """fp32 qkv [n_seq, T, 3h] -> fp32 ctx: the split-precision tensor-core attention of PREC_F16X3 (head_dim 64)."""
    _cuda(qkv, "qkv")
    qkv = qkv.contiguous()
    n_seq, T, h3 = qkv.shape
    h = h3 // 3
    if qkv.dtype != torch.float32 or h // heads != 64:
        raise ValueError("attention_f16x3 takes fp32 qkv with head_dim 64")
    ctx = torch.empty((n_seq, T, h), dtype=torch.float32, device=qkv.device)
    check(_lib.load().svit_attention_f16x3(_ptr(qkv), _ptr(ctx), n_seq, T, heads, _stream(qkv)))
    return ctx


def split_f16(x: torch.Tensor) -> torch.Tensor:
    """fp32 [G, rows, K] -> fp16 [G, rows, 2K] = [hi | lo] per row, the operand format of PREC_F16X3."""
    _cuda(x, "x")
    x = x.contiguous()
    G, rows, K = x.shape
    out = torch.empty((G, rows, 2 * K), dtype=torch.float16, device=x.device)
    check(_lib.load().svit_split_f16(_ptr(x), rows * K, _ptr(out), G, rows, K, _stream(x)))
    return out


def gemm(precision: int, A: torch.Tensor, B: torch.Tensor, bias: Optional[torch.Tensor] = None,
         residual: Optional[torch.Tensor] = None, gelu: bool = False, out_dtype: Optional[torch.dtype] = None,
         rowvec: Optional[torch.Tensor] = None, rows_in: int = 0, rows_out: int = 0, row_shift: int = 0,
         out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[g] = epilogue(A[g] @ B[g].T).  A [G|1, M, K], B [G, N, K] in the operand dtype of
    ``precision``; bias [G, N] fp32; residual fp32 [G, M_out, N]; rowvec fp32 [G, rows_out, N]."""
    _cuda(A, "A")
    _cuda(B, "B")
    A, B = A.contiguous(), B.contiguous()
    G, N, K = B.shape
    M = A.shape[1]
    out_dtype = out_dtype or A.dtype
    kp = K                      # row length of the operands in memory
    if precision == PREC_F16X3:  # fp32 in, split here; the library takes the [hi | lo] rows
        A, B, kp = split_f16(A), split_f16(B), 2 * K
    a_gs = 0 if (A.shape[0] == 1 and G > 1) else M * kp
    m_out = M if rows_in <= 0 else (M // rows_in) * rows_out
    if out is None:
        out = torch.zeros((G, m_out, N), dtype=out_dtype, device=A.device)
    epi = EpilogueC()
    keep = []
    if bias is not None:
        bias = bias.contiguous(); keep.append(bias)
        epi.bias, epi.bias_gs = bias.data_ptr(), N
    if rowvec is not None:
        rowvec = rowvec.contiguous(); keep.append(rowvec)
        epi.rowvec, epi.rowvec_gs = rowvec.data_ptr(), rows_out * N
    if residual is not None:
        epi.residual, epi.residual_gs = residual.data_ptr(), m_out * N
    epi.gelu, epi.rows_in, epi.rows_out, epi.row_shift = int(gelu), rows_in, rows_out, row_shift
    check(_lib.load().svit_gemm(precision, _ptr(A), a_gs, _ptr(B), N * kp, _ptr(out), m_out * N, SVIT_DTYPE[out.dtype],
                                G, M, N, K, C.byref(epi), _stream(A)))
    return out


class Plan:
    """Owner of an ``svit_plan`` plus its torch-allocated workspace."""

    def __init__(self, cfg, precision: int, max_coalitions: int, max_images: int, device):
        self.cfg, self.precision = cfg, precision
        self.max_coalitions, self.max_images = max_coalitions, max_images
        self.device = torch.device(device)
        self._h = C.c_void_p()
        c = _lib.cfg_struct(cfg)
        check(_lib.load().svit_plan_create(C.byref(c), precision, max_coalitions, max_images, C.byref(self._h)))
        self.operand_dtype = TORCH_DTYPE[_lib.load().svit_plan_operand_dtype(self._h)]
        self.workspace_bytes = _lib.load().svit_plan_workspace_bytes(self._h)
        self.workspace = torch.empty(self.workspace_bytes, dtype=torch.uint8, device=self.device)

    def close(self):
        if self._h:
            _lib.load().svit_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def timing_begin(self) -> None:
        check(_lib.load().svit_plan_timing_begin(self._h))

    def timing_end(self):
        """{class: (ms, work, launches)} accumulated since timing_begin (synchronises)."""
        t = _lib.TimingC()
        check(_lib.load().svit_plan_timing_end(self._h, C.byref(t)))
        return {name: (t.ms[i], t.work[i], t.launches[i]) for i, name in enumerate(_lib.KERNEL_CLASSES)}

    def patchify(self, images: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        _cuda(images, "images", torch.float32)
        images = images.contiguous()
        n = images.shape[0]
        if out is None:
            out = torch.empty((n * self.cfg.n_patches, self.cfg.patch_dim), dtype=self.operand_dtype,
                              device=images.device)
        check(_lib.load().svit_patchify(self._h, _ptr(images), _ptr(out), n, _stream(images)))
        return out

    def forward(self, wvec: torch.Tensor, wmat: torch.Tensor, patches: torch.Tensor, n_images: int,
                logits: torch.Tensor, image_offset: int = 0) -> torch.Tensor:
        """logits [C, n_total, n_cls] fp32; writes rows image_offset .. image_offset + n_images."""
        _cuda(wvec, "wvec", torch.float32)
        _cuda(wmat, "wmat", self.operand_dtype)
        _cuda(patches, "patches", self.operand_dtype)
        Cn = wvec.shape[0]
        n_cls = self.cfg.n_cls
        lptr = logits.data_ptr() + image_offset * n_cls * 4
        check(_lib.load().svit_forward_batched(
            self._h, _ptr(wvec), wvec.stride(0), _ptr(wmat), wmat.stride(0), _ptr(patches), C.c_void_p(lptr),
            logits.stride(0), Cn, n_images, _ptr(self.workspace), self.workspace_bytes, _stream(wvec)))
        return logits
