"""ctypes binding of ``libsvit.so`` (the C ABI declared in include/svit.h).

There is no fallback: if the shared library is missing it is built with nvcc, and if
that is impossible the import fails loudly.  Every wrapper raises ``SvitError`` with
the library's own message on a non-zero status.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

from . import build as _build

# enums (mirror include/svit.h)
F32, BF16, F16 = 0, 1, 2
PREC_F32, PREC_TF32, PREC_BF16, PREC_F16, PREC_F16X3, PREC_F16C8 = 0, 1, 2, 3, 4, 5
PRECISIONS = {"f32": PREC_F32, "fp32": PREC_F32, "tf32": PREC_TF32, "bf16": PREC_BF16, "f16": PREC_F16,
              "fp16": PREC_F16, "f16x3": PREC_F16X3, "f16c8": PREC_F16C8}
PRECISION_NAMES = {PREC_F32: "f32", PREC_TF32: "tf32", PREC_BF16: "bf16", PREC_F16: "f16", PREC_F16X3: "f16x3",
                   PREC_F16C8: "f16c8"}
# The precision a caller gets when it names none: the fastest mode that meets the parity gates of the path
# (>= 99.9 % top-1 agreement with the fp32 reference, utilities within one sample): fp16 main pass + e4m3
# compensation passes.  f16 / bf16 / tf32 are explicit opt-in throughput modes outside that tolerance.
DEFAULT_PRECISION = "f16c8"
OPERAND_DTYPE = {PREC_F32: F32, PREC_TF32: F32, PREC_BF16: BF16, PREC_F16: F16, PREC_F16X3: F16, PREC_F16C8: F16}
FMT_PLAIN, FMT_X3, FMT_C8 = 0, 1, 2     # svit_operand_format
OPERAND_FORMAT = {PREC_F32: FMT_PLAIN, PREC_TF32: FMT_PLAIN, PREC_BF16: FMT_PLAIN, PREC_F16: FMT_PLAIN,
                  PREC_F16X3: FMT_X3, PREC_F16C8: FMT_C8}
C8_HI_SCALE, C8_LO_SCALE = 4.0, 8192.0  # SVIT_C8_HI_SCALE / SVIT_C8_LO_SCALE (csrc/common.cuh)

EXPORTS = [
    "svit_version", "svit_last_error", "svit_device_info", "svit_layout_sizes", "svit_layout_segment",
    "svit_aggregate", "svit_aggregate_onto", "svit_aggregate_split", "svit_plan_create", "svit_plan_destroy",
    "svit_plan_workspace_bytes", "svit_plan_operand_dtype", "svit_plan_operand_format", "svit_patchify",
    "svit_forward_batched", "svit_forward_lora_batched", "svit_gemm_ext", "svit_score", "svit_score_records", "svit_gemm", "svit_split_operand", "svit_layernorm", "svit_attention",
    "svit_attention_split", "svit_plan_timing_begin", "svit_plan_timing_end",
]
KERNEL_CLASSES = ("gemm", "attention", "layernorm", "forward")


class SvitError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libsvit error {code}: {message}")
        self.code = code


class VitCfgC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("hidden", "layers", "heads", "ff", "image", "patch", "channels", "n_cls")]
    _fields_.append(("ln_eps", C.c_float))


class SegmentC(C.Structure):
    _fields_ = [("kind", C.c_int32), ("layer", C.c_int32), ("region", C.c_int32), ("reserved", C.c_int32),
                ("offset", C.c_int64), ("size", C.c_int64), ("rows", C.c_int64), ("cols", C.c_int64)]


class TimingC(C.Structure):
    _fields_ = [("ms", C.c_double * 4), ("work", C.c_double * 4), ("launches", C.c_int64 * 4)]


class EpilogueC(C.Structure):
    _fields_ = [("bias", C.c_void_p), ("bias_gs", C.c_int64), ("rowvec", C.c_void_p), ("rowvec_gs", C.c_int64),
                ("residual", C.c_void_p), ("residual_gs", C.c_int64), ("gelu", C.c_int32), ("rows_in", C.c_int32),
                ("rows_out", C.c_int32), ("row_shift", C.c_int32)]


_lib: Optional[C.CDLL] = None


def lib_path() -> str:
    return _build.LIB


def load() -> C.CDLL:
    """Load (building first if needed) libsvit.so.  Raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("SVIT_LIB")  # A/B runs against another build of the same ABI
    if path:
        if not os.path.exists(path):
            raise RuntimeError(f"SVIT_LIB={path} does not exist")
    else:
        path = _build.LIB
    if path == _build.LIB and not _build.is_current():
        try:
            _build.build()
        except Exception as e:  # a stale-but-present .so on a box without nvcc is still usable
            if not os.path.exists(path):
                raise RuntimeError(
                    "libsvit.so is not built and cannot be built here (nvcc missing?). "
                    "shapley_vit_b200 has no CPU or PyTorch fallback.") from e
    lib = C.CDLL(path)
    vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float
    sig = {
        "svit_version": (C.c_char_p, []),
        "svit_last_error": (C.c_char_p, []),
        "svit_device_info": (i32, [C.POINTER(i32)] * 3),
        "svit_layout_sizes": (i32, [C.POINTER(VitCfgC), C.POINTER(i64), C.POINTER(i64), C.POINTER(C.c_int32)]),
        "svit_layout_segment": (i32, [C.POINTER(VitCfgC), C.c_int32, C.POINTER(SegmentC)]),
        "svit_aggregate": (i32, [vp, i64, vp, vp, vp, i64, i32, i64, i32, i32, vp]),
        "svit_aggregate_onto": (i32, [vp, i64, vp, i64, vp, vp, i64, i32, i64, i32, i32, vp]),
        "svit_aggregate_split": (i32, [vp, i64, vp, vp, i64, vp, vp, i64, i64, i64, i32, i64, i32, i32, vp]),
        "svit_plan_create": (i32, [C.POINTER(VitCfgC), i32, i32, i32, C.POINTER(vp)]),
        "svit_plan_destroy": (i32, [vp]),
        "svit_plan_workspace_bytes": (i64, [vp]),
        "svit_plan_operand_dtype": (i32, [vp]),
        "svit_plan_operand_format": (i32, [vp]),
        "svit_patchify": (i32, [vp, vp, vp, i64, i64, i64, vp]),
        "svit_forward_batched": (i32, [vp, vp, i64, vp, i64, i64, vp, i64, i64, vp, i64, i32, i32, vp, C.c_size_t, vp]),
        "svit_forward_lora_batched": (i32, [vp, vp, i64, vp, i64, vp, i64, i64, vp, i64, i64, vp, i64, i32, i32, vp, C.c_size_t, vp]),
        "svit_gemm_ext": (i32, [i32, vp, i64, vp, i64, vp, vp, vp, i64, i32, i32, i32, i32, i32, C.POINTER(EpilogueC), vp]),
        "svit_score": (i32, [vp, i64, vp, i32, i64, i32, vp, vp, vp, i64, i32, vp]),
        "svit_score_records": (i32, [vp, i64, vp, i32, i64, i32, vp, i32, vp]),
        "svit_gemm": (i32, [i32, vp, i64, vp, i64, vp, i64, i32, i32, i32, i32, i32, C.POINTER(EpilogueC), vp]),
        "svit_split_operand": (i32, [vp, vp, i64, i32, i64, vp]),
        "svit_layernorm": (i32, [vp, i64, i64, vp, vp, i64, vp, i64, i64, i32, i32, i64, i32, i64, i32, f32, vp]),
        "svit_attention": (i32, [vp, vp, i32, i64, i32, i32, i32, vp]),
        "svit_attention_split": (i32, [vp, vp, i32, i64, i32, i32, i32, vp]),
        "svit_plan_timing_begin": (i32, [vp]),
        "svit_plan_timing_end": (i32, [vp, C.POINTER(TimingC)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError here = ABI mismatch; let it propagate
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise SvitError(rc, load().svit_last_error().decode("utf-8", "replace"))


def cfg_struct(cfg) -> VitCfgC:
    return VitCfgC(cfg.hidden, cfg.layers, cfg.heads, cfg.ff, cfg.image, cfg.patch, cfg.channels, cfg.n_cls,
                   cfg.ln_eps)


def version() -> str:
    return load().svit_version().decode()


def device_info():
    sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
    check(load().svit_device_info(C.byref(sm), C.byref(ma), C.byref(mi)))
    return sm.value, ma.value, mi.value


def layout_segments(cfg):
    """The C library's plan layout table: (vec_size, mat_size, [SegmentC...])."""
    lib, c = load(), cfg_struct(cfg)
    vs, ms, ns = C.c_int64(), C.c_int64(), C.c_int32()
    check(lib.svit_layout_sizes(C.byref(c), C.byref(vs), C.byref(ms), C.byref(ns)))
    segs = []
    for i in range(ns.value):
        s = SegmentC()
        check(lib.svit_layout_segment(C.byref(c), i, C.byref(s)))
        segs.append(s)
    return vs.value, ms.value, segs
