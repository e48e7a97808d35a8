"""torch-tensor front end of the C ABI: every function here takes CUDA tensors, passes
their ``data_ptr()`` and the current CUDA stream to ``libsvit.so`` and returns tensors.
PyTorch is used for device memory and streams only -- no torch operator computes anything
on these paths, and there is no fallback if the library is missing."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import BF16, F16, F32, FMT_C8, FMT_PLAIN, FMT_X3, OPERAND_FORMAT, EpilogueC, check

TORCH_DTYPE = {F32: torch.float32, BF16: torch.bfloat16, F16: torch.float16}
SVIT_DTYPE = {v: k for k, v in TORCH_DTYPE.items()}


class OperandArray:
    """An operand array of ``shape`` elements in an operand format of the library (include/svit.h):
    FMT_PLAIN -- a torch tensor of ``dtype``; FMT_X3 / FMT_C8 -- ONE uint8 allocation holding the planes of a
    packed operand array of ``alloc`` elements (the plane pitch): fp16 hi at byte 0, then fp16 lo (X3) or the
    e4m3 planes hi8 | lo8 (C8); 4 bytes per element.  Element (i, j, ...) has the same index in every plane."""

    def __init__(self, shape, dtype: torch.dtype, fmt: int, device):
        self.shape = tuple(int(x) for x in shape)
        self.fmt, self.device = int(fmt), torch.device(device)
        self.elems = 1
        for x in self.shape:
            self.elems *= x
        self.alloc = (self.elems + 15) // 16 * 16
        if self.fmt == FMT_PLAIN:
            self.dtype = dtype
            self.buf = torch.empty(self.shape, dtype=dtype, device=self.device)
        else:
            self.dtype = torch.float16
            self.buf = torch.empty(4 * max(self.alloc, 16), dtype=torch.uint8, device=self.device)

    @property
    def ptr(self) -> int:
        return self.buf.data_ptr()

    @property
    def nbytes(self) -> int:
        return self.buf.numel() * self.buf.element_size()

    def plane(self, k: int) -> torch.Tensor:
        """Plane k as a tensor of ``shape``: 0 = main (fp16 hi, or the plain array), 1 / 2 = auxiliary planes."""
        a, n = self.alloc, self.elems
        if self.fmt == FMT_PLAIN:
            return self.buf
        if k == 0:
            return self.buf[:2 * a].view(torch.float16)[:n].view(self.shape)
        if self.fmt == FMT_X3:
            return self.buf[2 * a:4 * a].view(torch.float16)[:n].view(self.shape)
        lo = (2 if k == 1 else 3) * a
        return self.buf[lo:lo + a][:n].view(torch.float8_e4m3fn).view(self.shape)

    def to_float(self) -> torch.Tensor:
        """fp32 value the planes stand for: hi + lo (X3), hi + lo8 / 2^13 (C8)."""
        if self.fmt == FMT_PLAIN:
            return self.buf.float()
        if self.fmt == FMT_X3:
            return self.plane(0).float() + self.plane(1).float()
        return self.plane(0).float() + self.plane(2).float() / _lib.C8_LO_SCALE

    @staticmethod
    def from_float(x: torch.Tensor, fmt: int, dtype: torch.dtype = torch.float32) -> "OperandArray":
        """fp32 CUDA tensor -> operand array (svit_split_operand for the split formats)."""
        _cuda(x, "x")
        out = OperandArray(x.shape, dtype, fmt, x.device)
        if fmt == FMT_PLAIN:
            out.buf.copy_(x)
            return out
        x = x.contiguous().to(torch.float32)
        if out.elems % 4:
            raise ValueError("split operand arrays need a multiple of 4 elements")
        check(_lib.load().svit_split_operand(_ptr(x), C.c_void_p(out.ptr), out.alloc, fmt, out.elems, _stream(x)))
        return out


def operand_like(precision: int, shape, device) -> OperandArray:
    """Empty operand array in the format the GEMMs of ``precision`` read."""
    return OperandArray(shape, TORCH_DTYPE[_lib.OPERAND_DTYPE[precision]], OPERAND_FORMAT[precision], device)


def _stream(t: torch.Tensor) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _cuda(t: torch.Tensor, name: str, dtype=None) -> None:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (libsvit has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")


def _out_args(out, Cn: int, P: int, col0: int):
    """(is_split, ptr, element offset, row stride, alloc, fmt / dtype code) of an aggregation output: a torch tensor
    [C, >= P] (plain) or an OperandArray [rows >= C, width] written at columns col0 .. col0 + P."""
    if isinstance(out, OperandArray):
        if len(out.shape) != 2 or out.shape[0] < Cn or col0 + P > out.shape[1]:
            raise ValueError("out must be an OperandArray [rows >= C, >= col0 + P]")
        if out.fmt != FMT_PLAIN:
            return True, out.ptr, col0, out.shape[1], out.alloc, out.fmt
        out = out.buf[:Cn, col0:col0 + P] if (col0 or P != out.shape[1]) else out.buf[:Cn]
    _cuda(out, "out")
    if out.dim() != 2 or out.stride(1) != 1 or out.shape[0] != Cn:
        raise ValueError("out must be [C, >=P] with unit inner stride")
    return False, out.data_ptr(), 0, out.stride(0), 0, SVIT_DTYPE[out.dtype]


def aggregate(deltas: torch.Tensor, w0: Optional[torch.Tensor], ratios: torch.Tensor,
              out_dtype: torch.dtype = torch.float32, out=None, P: Optional[int] = None, col0: int = 0):
    """K1.  deltas [N, stride] fp32 (row-contiguous), w0 [>=P] fp32 or None, ratios [C, N] fp32
    on the HOST (0 = non-member; a CUDA tensor is copied back, which synchronises).
    Returns out with out[c, col0 : col0 + P] = w0 + sum_j ratios[c, j] * deltas[j]; ``out`` is a torch tensor
    [C, >= P] or an OperandArray (split formats: svit_aggregate_split writes the planes directly)."""
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
    split, optr, off, ostride, alloc, code = _out_args(out, Cn, P, col0)
    if split:
        check(_lib.load().svit_aggregate_split(_ptr(deltas), deltas.stride(0), _ptr(w0), None, 0, _ptr(ratios),
                                               C.c_void_p(optr), off, ostride, alloc, code, P, N, Cn, _stream(deltas)))
    else:
        check(_lib.load().svit_aggregate(_ptr(deltas), deltas.stride(0), _ptr(w0), _ptr(ratios), C.c_void_p(optr), ostride,
                                         code, P, N, Cn, _stream(deltas)))
    return out


def aggregate_onto(deltas: torch.Tensor, base: torch.Tensor, ratios: torch.Tensor, out,
                   P: Optional[int] = None, col0: int = 0):
    """One more FL round folded onto per-coalition partial models (svit_aggregate_onto / _split):
    out[c, col0 : col0 + P] = cast(base[c, :P] + sum_j ratios[c, j] * deltas[j]).  base [C, stride] fp32 (may be
    ``out`` itself when ``out`` is fp32), ratios [C, N] on the host; ``out`` a tensor or an OperandArray."""
    _cuda(deltas, "deltas", torch.float32)
    _cuda(base, "base", torch.float32)
    ratios = ratios.detach().to("cpu", torch.float32).contiguous()
    N, width = deltas.shape
    P = width if P is None else int(P)
    Cn = ratios.shape[0]
    if ratios.shape != (Cn, N) or base.shape[0] != Cn:
        raise ValueError("ratios must be [C, N]; base must have C rows")
    if deltas.stride(1) != 1 or base.stride(1) != 1:
        raise ValueError("deltas / base must have unit inner stride")
    split, optr, off, ostride, alloc, code = _out_args(out, Cn, P, col0)
    if split:
        check(_lib.load().svit_aggregate_split(_ptr(deltas), deltas.stride(0), None, _ptr(base), base.stride(0), _ptr(ratios),
                                               C.c_void_p(optr), off, ostride, alloc, code, P, N, Cn, _stream(deltas)))
    else:
        check(_lib.load().svit_aggregate_onto(_ptr(deltas), deltas.stride(0), _ptr(base), base.stride(0), _ptr(ratios),
                                              C.c_void_p(optr), ostride, code, P, N, Cn, _stream(deltas)))
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


def score_records(logits: torch.Tensor, labels: torch.Tensor, records: torch.Tensor, accumulate: bool = False) -> torch.Tensor:
    """K5 into packed records: ``records`` is an int64 [C, 2] device tensor (a slice of the all-gather send buffer);
    row c receives (correct, bit pattern of the fp64 loss_sum) -- the 16-byte svit_record of include/svit.h."""
    _cuda(logits, "logits", torch.float32)
    _cuda(labels, "labels", torch.int64)
    _cuda(records, "records", torch.int64)
    if logits.dim() != 3 or not logits[0].is_contiguous():
        raise ValueError("logits must be [C, n, n_cls] with contiguous [n, n_cls] slabs")
    Cn, n, n_cls = logits.shape
    if records.shape != (Cn, 2) or not records.is_contiguous():
        raise ValueError("records must be a contiguous int64 [C, 2] tensor")
    check(_lib.load().svit_score_records(_ptr(logits), logits.stride(0) if Cn > 1 else n * n_cls, _ptr(labels.contiguous()), Cn, n,
                                         n_cls, _ptr(records), int(accumulate), _stream(logits)))
    return records


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float,
              out_dtype: torch.dtype = torch.float32, fmt: int = FMT_PLAIN):
    """x [G, rows, h] fp32, gamma/beta [G, h] fp32 -> y [G, rows, h] (a tensor, or an OperandArray for a split fmt)."""
    _cuda(x, "x", torch.float32)
    G, rows, h = x.shape
    x, gamma, beta = x.contiguous(), gamma.contiguous(), beta.contiguous()
    y = OperandArray((G, rows, h), out_dtype, fmt, x.device)
    check(_lib.load().svit_layernorm(_ptr(x), rows * h, h, _ptr(gamma), _ptr(beta), h, C.c_void_p(y.ptr), rows * h, h,
                                     SVIT_DTYPE[y.dtype], fmt, y.alloc, G, rows, h, float(eps), _stream(x)))
    return y.buf if fmt == FMT_PLAIN else y


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


def attention_split(qkv: torch.Tensor, heads: int, out_fmt: int = FMT_X3) -> OperandArray:
    """fp32 qkv [n_seq, T, 3h] -> ctx [n_seq, T, h] as an operand array of ``out_fmt``: the attention of the split
    precisions (qkv is split into X3 planes here; in the forward the QKV GEMM's epilogue emits them)."""
    _cuda(qkv, "qkv", torch.float32)
    n_seq, T, h3 = qkv.shape
    h = h3 // 3
    q = OperandArray.from_float(qkv, FMT_X3)
    ctx = OperandArray((n_seq, T, h), torch.float16, out_fmt, qkv.device)
    check(_lib.load().svit_attention_split(C.c_void_p(q.ptr), C.c_void_p(ctx.ptr), out_fmt, n_seq, T, heads, h // heads,
                                           _stream(qkv)))
    return ctx


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
    fmt = OPERAND_FORMAT[precision]
    a_gs = 0 if (A.shape[0] == 1 and G > 1) else M * K
    a_ptr, b_ptr = _ptr(A), _ptr(B)
    if fmt != FMT_PLAIN:         # fp32 in, split here (in the forward the producers emit the planes)
        A, B = OperandArray.from_float(A, fmt), OperandArray.from_float(B, fmt)
        keep_ab = (A, B)
        a_ptr, b_ptr = C.c_void_p(A.ptr), C.c_void_p(B.ptr)
    m_out = M if rows_in <= 0 else (M // rows_in) * rows_out
    if isinstance(out, OperandArray):
        out_arr, out = out, out.plane(0)
    elif out is None and fmt != FMT_PLAIN and out_dtype == torch.float16:
        out_arr = OperandArray((G, m_out, N), torch.float16, fmt, B.device)
        out = out_arr.plane(0)
    else:
        out_arr = None
    if out is None:
        out = torch.zeros((G, m_out, N), dtype=out_dtype, device=B.device)
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
    check(_lib.load().svit_gemm(precision, a_ptr, a_gs, b_ptr, N * K, _ptr(out), m_out * N, SVIT_DTYPE[out.dtype],
                                G, M, N, K, C.byref(epi), _stream(out)))
    return out_arr if out_arr is not None else out


def gemm_ext(precision: int, A: torch.Tensor, B: torch.Tensor, Ae: torch.Tensor, Be: torch.Tensor,
             bias: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.float32):
    """out[g] = A[g] @ B[g or 0].T + Ae[g] @ Be[g].T (+ bias): svit_gemm_ext.  A [G, M, K], B [G | 1, N, K], Ae [G, M, 64],
    Be [G, N, 64] fp32 CUDA tensors (converted to the precision's operand format here); fp32 or operand-format output."""
    for t in (A, B, Ae, Be):
        _cuda(t, "operand")
    G, M, K = A.shape
    N = B.shape[1]
    fmt = OPERAND_FORMAT[precision]
    odt = TORCH_DTYPE[_lib.OPERAND_DTYPE[precision]]
    conv = lambda t: OperandArray.from_float(t.contiguous().float(), fmt, odt)
    a, b, ae, be = conv(A), conv(B), conv(Ae), conv(Be)
    if out_dtype == torch.float32:
        out_arr, out = None, torch.zeros((G, M, N), dtype=torch.float32, device=A.device)
        optr = _ptr(out)
    else:
        out_arr = OperandArray((G, M, N), odt, fmt, A.device)
        out, optr = out_arr.plane(0), C.c_void_p(out_arr.ptr)
    epi = EpilogueC()
    if bias is not None:
        bias = bias.contiguous()
        epi.bias, epi.bias_gs = bias.data_ptr(), N
    check(_lib.load().svit_gemm_ext(precision, C.c_void_p(a.ptr), M * K, C.c_void_p(b.ptr), N * K if B.shape[0] == G else 0,
                                    C.c_void_p(ae.ptr), C.c_void_p(be.ptr), optr, M * N, SVIT_DTYPE[out.dtype], G, M, N, K,
                                    C.byref(epi), _stream(A)))
    return out_arr if out_arr is not None else out


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
        self.operand_format = _lib.load().svit_plan_operand_format(self._h)
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

    def operand_array(self, shape) -> OperandArray:
        """Empty operand array in this plan's operand format (weights, patch matrix)."""
        return OperandArray(shape, self.operand_dtype, self.operand_format, self.device)

    def patchify(self, images: torch.Tensor, out: Optional[OperandArray] = None, row0: int = 0) -> OperandArray:
        """images [n, C, H, W] fp32 -> rows [row0, row0 + n * n_patches) of the patch matrix ``out``."""
        _cuda(images, "images", torch.float32)
        images = images.contiguous()
        n = images.shape[0]
        if out is None:
            out = self.operand_array((n * self.cfg.n_patches, self.cfg.patch_dim))
        check(_lib.load().svit_patchify(self._h, _ptr(images), C.c_void_p(out.ptr), out.alloc, row0, n, _stream(images)))
        return out

    def forward(self, wvec: torch.Tensor, wmat: OperandArray, patches: OperandArray, row0: int, n_images: int,
                logits: torch.Tensor, image_offset: int = 0) -> torch.Tensor:
        """logits [C, n_total, n_cls] fp32; writes rows image_offset .. image_offset + n_images from the images whose
        patch rows start at ``row0`` of ``patches``."""
        _cuda(wvec, "wvec", torch.float32)
        if wmat.fmt != self.operand_format or patches.fmt != self.operand_format:
            raise ValueError("wmat / patches are not in this plan's operand format")
        Cn = wvec.shape[0]
        n_cls = self.cfg.n_cls
        lptr = logits.data_ptr() + image_offset * n_cls * 4
        check(_lib.load().svit_forward_batched(
            self._h, _ptr(wvec), wvec.stride(0), C.c_void_p(wmat.ptr), wmat.shape[1], wmat.alloc, C.c_void_p(patches.ptr),
            patches.alloc, row0, C.c_void_p(lptr), logits.stride(0), Cn, n_images, _ptr(self.workspace),
            self.workspace_bytes, _stream(wvec)))
        return logits

    def forward_lora(self, wvec: torch.Tensor, wmat_shared: OperandArray, lora: OperandArray, patches: OperandArray, row0: int,
                     n_images: int, logits: torch.Tensor, image_offset: int = 0) -> torch.Tensor:
        """``forward`` for PEFT-LoRA coalition models over a frozen base: ``wmat_shared`` is ONE mat-region row shared by
        every coalition, ``lora`` the per-coalition [C, layers * 256 * hidden] rows (Acat | Bext per layer, include/svit.h)."""
        _cuda(wvec, "wvec", torch.float32)
        if wmat_shared.fmt != self.operand_format or patches.fmt != self.operand_format or lora.fmt != self.operand_format:
            raise ValueError("wmat / lora / patches are not in this plan's operand format")
        Cn = wvec.shape[0]
        lptr = logits.data_ptr() + image_offset * self.cfg.n_cls * 4
        check(_lib.load().svit_forward_lora_batched(
            self._h, _ptr(wvec), wvec.stride(0), C.c_void_p(wmat_shared.ptr), wmat_shared.alloc, C.c_void_p(lora.ptr), lora.shape[1],
            lora.alloc, C.c_void_p(patches.ptr), patches.alloc, row0, C.c_void_p(lptr), logits.stride(0), Cn, n_images,
            _ptr(self.workspace), self.workspace_bytes, _stream(wvec)))
        return logits
