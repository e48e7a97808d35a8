#!/usr/bin/env python3
"""Per-kernel timing on one B200 (CUDA events, warm, inputs larger than L2 where the real step's are).

    python scripts/microbench.py gemm        # the ViT-B layer's GEMM shapes, C coalitions x B images
    python scripts/microbench.py agg         # K1 at BASELINE config 2 / config 5 points
    python scripts/microbench.py attn ln     # attention, LayerNorm

Prints one line per case: time, achieved TFLOP/s or GB/s, fraction of the measured peak.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shapley_vit_b200 import _lib, ops  # noqa: E402

PEAKS = {"hbm_gbs": 6555.5, "bf16_tflops": 1650.1, "bf16_tflops_sustained": 1388.5}
try:
    PEAKS.update(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))))
except Exception:
    pass


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def bench_gemm(C=8, B=128, prec="f16", T=197, h=768, ff=3072):
    P = _lib.PRECISIONS[prec]
    dt = ops.TORCH_DTYPE[_lib.OPERAND_DTYPE[P]]
    M = B * T
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)

    def rnd(*s, dtype=dt, scale=0.05):
        return (torch.randn(*s, device=dev, generator=g) * scale).to(dtype)

    X = rnd(C, M, h, dtype=torch.float32, scale=1.0)
    cases = [
        ("qkv   (N=2304,K=768, bias, 16-bit out)", rnd(C, M, h), rnd(C, 3 * h, h), dict(bias=rnd(C, 3 * h, dtype=torch.float32))),
        ("proj  (N=768, K=768, bias+residual, f32)", rnd(C, M, h), rnd(C, h, h),
         dict(bias=rnd(C, h, dtype=torch.float32), residual=X, out=X, out_dtype=torch.float32)),
        ("mlp-up(N=3072,K=768, bias+GELU, 16-bit)", rnd(C, M, h), rnd(C, ff, h), dict(bias=rnd(C, ff, dtype=torch.float32), gelu=True)),
        ("mlp-dn(N=768, K=3072,bias+residual, f32)", rnd(C, M, ff), rnd(C, h, ff),
         dict(bias=rnd(C, h, dtype=torch.float32), residual=X, out=X, out_dtype=torch.float32)),
        ("plain (N=2304,K=768, no epilogue)", rnd(C, M, h), rnd(C, 3 * h, h), dict()),
    ]
    tot_ms, tot_fl = 0.0, 0.0
    for name, A, W, kw in cases:
        N, K = W.shape[1], W.shape[2]
        if "out" not in kw:
            kw["out"] = torch.empty(C, M, N, dtype=kw.get("out_dtype", dt), device=dev)
        ms = timeit(lambda: ops.gemm(P, A, W, **kw))
        fl = 2.0 * C * M * N * K
        if not name.startswith("plain"):
            tot_ms += ms
            tot_fl += fl
        print(f"gemm[{prec}] C={C} B={B} {name}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s  "
              f"{fl/ms/1e9/PEAKS['bf16_tflops_sustained']:.3f} of sustained cuBLAS peak", flush=True)
    print(f"gemm[{prec}] layer total: {tot_ms*1e3:8.1f} us  {tot_fl/tot_ms/1e9:7.1f} TFLOP/s", flush=True)


def bench_agg():
    dev = "cuda"
    torch.manual_seed(0)
    sizes = {"S": 21_669_514, "B": 85_806_346, "L": 303_311_882}
    cases = [("B", 8, 8, torch.float16), ("B", 8, 8, torch.float32), ("B", 8, 32, torch.float16), ("B", 8, 1, torch.float16),
             ("B", 4, 8, torch.float16), ("B", 16, 8, torch.float16), ("B", 16, 32, torch.float16), ("S", 32, 8, torch.float16),
             ("S", 64, 8, torch.float16), ("S", 64, 64, torch.float16), ("L", 10, 32, torch.bfloat16), ("B", 8, 128, torch.float16)]
    for model, N, Cn, dt in cases:
        P = sizes[model]
        stride = (P + 63) // 64 * 64
        need = (N + 1) * stride * 4 + Cn * stride * (4 if dt == torch.float32 else 2)
        if need > 150e9:
            print(f"agg {model} N={N} C={Cn}: skipped ({need/1e9:.0f} GB)")
            continue
        deltas = torch.randn(N, stride, device=dev) * 0.02
        w0 = torch.randn(stride, device=dev) * 0.02
        masks = torch.rand(Cn, N) < 0.5
        masks[:, 0] |= ~masks.any(dim=1)
        n = torch.arange(1, N + 1, dtype=torch.float64) * 1000
        ratios = (masks * n / (masks * n).sum(dim=1, keepdim=True)).float()
        out = torch.empty(Cn, stride, dtype=dt, device=dev)
        ms = timeit(lambda: ops.aggregate(deltas, w0, ratios, out=out, P=P), iters=5, warm=2)
        es = 4 if dt == torch.float32 else 2
        by = 4.0 * P * (N + 1) + es * P * Cn
        print(f"agg ViT-{model} N={N:2d} C={Cn:3d} out={str(dt)[6:]:8s}: {ms:8.3f} ms  {by/ms/1e6:7.1f} GB/s  "
              f"{by/ms/1e6/PEAKS['hbm_gbs']:.3f} of measured copy peak", flush=True)
        del deltas, w0, out


def bench_attn(C=8, B=128, T=197, h=768, heads=12, dts=(torch.float16, torch.bfloat16)):
    for dt in dts:
        qkv = (torch.randn(C * B, T, 3 * h, device="cuda") * 0.5).to(dt)
        ms = timeit(lambda: ops.attention(qkv, heads))
        fl = 4.0 * C * B * T * T * h
        print(f"attention {str(dt)[6:]} n_seq={C*B} T={T}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s", flush=True)


def bench_ln(C=8, B=128, T=197, h=768):
    x = torch.randn(C, B * T, h, device="cuda")
    g, b = torch.randn(C, h, device="cuda"), torch.randn(C, h, device="cuda")
    for dt in (torch.float16, torch.float32):
        ms = timeit(lambda: ops.layernorm(x, g, b, 1e-12, out_dtype=dt))
        by = C * B * T * h * (4.0 + (2 if dt == torch.float16 else 4))
        print(f"layernorm out={str(dt)[6:]}: {ms*1e3:8.1f} us  {by/ms/1e6:7.1f} GB/s  {by/ms/1e6/PEAKS['hbm_gbs']:.3f} of copy peak",
              flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["gemm", "agg", "attn", "ln"]
    print(_lib.version(), torch.cuda.get_device_name(0), flush=True)
    if "gemm" in what:
        bench_gemm(prec="f16")
        bench_gemm(prec="bf16", C=4)
    if "gemmtf32" in what:
        bench_gemm(prec="tf32", C=4, B=64)
    if "agg" in what:
        bench_agg()
    if "attn" in what:
        bench_attn()
    if "attn32" in what:   # the register-tiled fp32 kernel of the f32 / tf32 modes, the split-precision one of f16x3
        bench_attn(C=2, dts=(torch.float32,))
        qkv = torch.randn(1024, 197, 2304, device="cuda") * 0.5
        ms = timeit(lambda: ops.attention_split(qkv, 12))
        print(f"attention f16x3 n_seq=1024 T=197: {ms*1e3:8.1f} us  {4.0*1024*197*197*768/ms/1e9:7.1f} TFLOP/s", flush=True)
    if "ln" in what:
        bench_ln()
