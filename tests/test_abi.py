"""The C-ABI library loads, exports every symbol include/svit.h declares, and validates
arguments without touching a device (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

from shapley_vit_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "svit.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(svit_[a-z_0-9]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/svit.h but not exported"
    assert set(names) == set(_lib.EXPORTS)
    assert b"sm_100a" in lib.svit_version()


def test_argument_validation_reports_errors_without_a_device():
    lib = _lib.load()
    rc = lib.svit_aggregate(None, 0, None, None, None, 0, 0, 16, 1, 1, None)
    assert rc == -1 and b"null" in lib.svit_last_error()
    buf = (C.c_float * 64)()
    p = C.cast(buf, C.c_void_p)
    rc = lib.svit_aggregate(p, 64, None, p, p, 64, 0, 16, 0, 1, None)      # N = 0
    assert rc == -1
    rc = lib.svit_aggregate(p, 60, None, p, p, 64, 0, 16, 1, 1, None)      # stride % 8 != 0
    assert rc == -2
    with pytest.raises(_lib.SvitError):
        _lib.check(rc)


def test_plan_create_rejects_bad_geometry():
    from shapley_vit_b200.layout import VitConfig

    lib = _lib.load()
    h = C.c_void_p()
    bad = _lib.VitCfgC(100, 2, 3, 768, 32, 16, 3, 10, 1e-12)
    assert lib.svit_plan_create(C.byref(bad), 0, 1, 1, C.byref(h)) == -1
    good = _lib.cfg_struct(VitConfig(192, 2, 3, 768, 32, 10))
    assert lib.svit_plan_create(C.byref(good), 3, 2, 4, C.byref(h)) == 0
    assert lib.svit_plan_workspace_bytes(h) > 0 and lib.svit_plan_operand_dtype(h) == _lib.F16
    assert lib.svit_plan_destroy(h) == 0


def test_product_has_no_cpu_path():
    import torch

    from shapley_vit_b200 import ops

    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(ValueError, match="CUDA"):
        ops.aggregate(torch.zeros(1, 8), None, torch.ones(1, 1))


def test_split_precision_plans_use_packed_fp16_planes():
    """The split precisions keep every GEMM operand as planes of a packed operand array (4 bytes / element, main plane
    fp16): the same workspace as an fp32-storage mode, no split scratch."""
    from shapley_vit_b200.layout import VitConfig

    lib = _lib.load()
    cfg = _lib.cfg_struct(VitConfig(192, 2, 3, 768, 32, 10))
    sizes = {}
    for name in ("f16", "tf32", "f16x3", "f16c8"):
        h = C.c_void_p()
        assert lib.svit_plan_create(C.byref(cfg), _lib.PRECISIONS[name], 2, 4, C.byref(h)) == 0
        sizes[name] = (lib.svit_plan_operand_dtype(h), lib.svit_plan_operand_format(h), lib.svit_plan_workspace_bytes(h))
        assert lib.svit_plan_destroy(h) == 0
    assert sizes["f16x3"][:2] == (_lib.F16, _lib.FMT_X3) and sizes["f16c8"][:2] == (_lib.F16, _lib.FMT_C8)
    assert sizes["tf32"][:2] == (_lib.F32, _lib.FMT_PLAIN) and sizes["f16"][:2] == (_lib.F16, _lib.FMT_PLAIN)
    assert sizes["f16x3"][2] == sizes["f16c8"][2] == sizes["tf32"][2] > sizes["f16"][2]
    h = C.c_void_p()
    assert lib.svit_plan_create(C.byref(cfg), 7, 2, 4, C.byref(h)) == -1          # unknown precision
    assert _lib.DEFAULT_PRECISION in ("f16c8", "f16x3", "f32")                   # the default must be a parity mode


def test_lib_override_must_exist(monkeypatch):
    """SVIT_LIB (A/B against another build of the same ABI) fails loudly on a bad path; there is no fallback."""
    monkeypatch.setenv("SVIT_LIB", "/nonexistent/libsvit.so")
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(RuntimeError):
        _lib.load()
