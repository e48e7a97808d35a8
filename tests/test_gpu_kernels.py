"""Per-kernel parity on the GPU, through the C ABI: K1 aggregation, K5 scoring, LayerNorm,
attention, patchify and both GEMM back ends, each against the oracle / plain torch fp32 on the
same seeded inputs."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import restate

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from shapley_vit_b200 import _lib, ops as _ops

    sm, major, minor = _lib.device_info()
    assert major == 10, "libsvit is sm_100a only"
    return _ops


def gen(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g, dtype=torch.float32) * scale


# ----------------------------------------------------------------------------- K1
def oracle_aggregate(deltas, w0, ratios):
    """restate.get_aggregated_model + model_agg_lazy on flat fp32 rows (ascending members)."""
    out = []
    for row in ratios:
        members = [j for j, r in enumerate(row) if r != 0]
        agg = restate.get_aggregated_model([{"p": deltas[j]} for j in members], [float(row[j]) for j in members])
        base = {"p": w0} if w0 is not None else {"p": torch.zeros_like(deltas[0])}
        out.append(restate.model_agg_lazy(base, [agg] if agg is not None else [])["p"])
    return torch.stack(out)


def fedavg_rows(masks, n_train):
    rows = []
    for m in masks:
        tot = sum(n for n, b in zip(n_train, m) if b)
        rows.append([float(np.float32(n / tot)) if b else 0.0 for n, b in zip(n_train, m)])
    return rows


@pytest.mark.parametrize("N,C,P", [(1, 1, 8), (4, 15, 4096), (8, 32, 100_000), (8, 33, 33_333), (3, 5, 1027),
                                   (16, 17, 70_001), (20, 9, 5_000), (40, 7, 3_000), (64, 3, 2_049),
                                   # several groups of client rows AND several chunks of coalitions per tile
                                   (32, 16, 40_003), (16, 33, 9_001), (64, 70, 2_600), (14, 9, 511)])
def test_aggregate_fp32_bit_exact(ops, N, C, P):
    rng = np.random.RandomState(N * 1000 + C)
    stride = (P + 7) // 8 * 8
    deltas = torch.zeros(N, stride)
    deltas[:, :P] = gen(N, P, seed=P) * 0.02
    w0 = torch.zeros(stride)
    w0[:P] = gen(P, seed=P + 1) * 0.02
    masks = [rng.rand(N) < 0.5 for _ in range(C)]
    for m in masks:
        if not m.any():
            m[rng.randint(N)] = True
    masks[0][:] = True
    rows = fedavg_rows(masks, [1000 * (j + 1) for j in range(N)])
    got = ops.aggregate(deltas.cuda(), w0.cuda(), torch.tensor(rows, dtype=torch.float32).cuda(), P=P).cpu()
    want = oracle_aggregate(deltas[:, :P], w0[:P], rows)
    assert torch.equal(got[:, :P], want)                       # bit-exact, no FMA contraction


def test_aggregate_no_w0_and_identity(ops):
    P = 10_000
    deltas = gen(2, P, seed=3)
    r = torch.tensor([[1.0, 0.0], [0.0, 1.0], [0.0, 0.0]])
    got = ops.aggregate(deltas.cuda(), None, r.cuda()).cpu()
    assert torch.equal(got[0], deltas[0]) and torch.equal(got[1], deltas[1])
    assert torch.count_nonzero(got[2]) == 0                     # empty coalition: W0 (= 0) alone


@pytest.mark.parametrize("N,C", [(8, 12), (16, 8), (11, 17), (16, 12), (24, 40)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_aggregate_16bit_is_rounded_fp32_result(ops, dtype, N, C):
    P = 50_000
    deltas, w0 = gen(N, P, seed=5) * 0.02, gen(P, seed=6) * 0.02
    rng = np.random.RandomState(0)
    rows = fedavg_rows([rng.rand(N) < 0.6 for _ in range(C - 1)] + [np.ones(N, bool)], list(range(1, N + 1)))
    r = torch.tensor(rows, dtype=torch.float32).cuda()
    ref = ops.aggregate(deltas.cuda(), w0.cuda(), r)
    got = ops.aggregate(deltas.cuda(), w0.cuda(), r, out_dtype=dtype)
    # The 16-bit feed accumulates with fused multiply-adds (<= 1 fp32 ulp from the exact two-rounding
    # value before the cast), so it equals the rounded fp32 result except where that value sits on a
    # 16-bit rounding boundary: at most one 16-bit ulp, in a vanishing fraction of the elements.
    want = ref.to(dtype)
    differ = got != want
    assert differ.float().mean().item() < 1e-3
    ulp = 2.0 ** (-10 if dtype == torch.float16 else -7)            # relative spacing; + the fp16 subnormal spacing
    assert ((got.float() - want.float()).abs() <= ulp * want.float().abs() + 2.0 ** -24).all()


@pytest.mark.parametrize("N,C,P", [(8, 8, 100_000), (3, 5, 1027), (16, 17, 70_001), (32, 16, 40_003)])
@pytest.mark.parametrize("fmt_name", ["x3", "c8"])
def test_aggregate_split_planes_are_the_split_of_the_fp32_result(ops, fmt_name, N, C, P):
    """svit_aggregate_split: X3 planes are the exact split of the bit-exact fp32 aggregate (two-rounding arithmetic); C8
    planes (16-bit-grade) are the split of the fused accumulation, <= a few fp32 ulp from it.  Also at a column offset of
    a wider array and through the per-coalition base of the multi-round fold."""
    from shapley_vit_b200._lib import FMT_C8, FMT_X3
    from shapley_vit_b200.ops import OperandArray

    fmt = FMT_X3 if fmt_name == "x3" else FMT_C8
    rng = np.random.RandomState(N + C)
    deltas, w0 = (gen(N, (P + 7) // 8 * 8, seed=P) * 0.02).cuda(), (gen((P + 7) // 8 * 8, seed=P + 1) * 0.02).cuda()
    rows = fedavg_rows([rng.rand(N) < 0.5 for _ in range(C - 1)] + [np.ones(N, bool)], [1000 * (j + 1) for j in range(N)])
    r = torch.tensor(rows, dtype=torch.float32)
    ref = ops.aggregate(deltas, w0, r, P=P)[:, :P].cpu()                  # fp32, bit-exact vs the oracle (test above)
    width, col0 = (P + 63) // 64 * 64 + 64, 64
    out = OperandArray((C + 1, width), torch.float16, fmt, "cuda:0")
    out.buf.zero_()
    ops.aggregate(deltas, w0, r, out=out, P=P, col0=col0)
    got = OperandArray((C, P), torch.float16, fmt, "cuda:0")
    for k in range(2 if fmt == FMT_X3 else 3):
        got.plane(k).copy_(out.plane(k)[:C, col0:col0 + P])
        assert torch.count_nonzero(out.plane(k)[:C, :col0].view(torch.uint8)) == 0      # nothing outside the window
        assert torch.count_nonzero(out.plane(k)[C].view(torch.uint8)) == 0
    exact = fmt == FMT_X3
    ulps = 0.0 if exact else N * 2.0 ** -23 * float(ref.abs().max())      # fused fold: < 1 fp32 ulp per client
    check_planes(got, ref, ulps, exact=exact)
    base = (gen(C, (P + 7) // 8 * 8, seed=77) * 0.02).cuda()
    ref2 = ops.aggregate_onto(deltas, base, r, torch.empty((C, (P + 7) // 8 * 8), device="cuda"), P=P)[:, :P].cpu()
    out2 = OperandArray((C, (P + 7) // 8 * 8), torch.float16, fmt, "cuda:0")
    ops.aggregate_onto(deltas, base, r, out2, P=P)
    for k in range(2 if fmt == FMT_X3 else 3):
        got.plane(k).copy_(out2.plane(k)[:, :P])
    check_planes(got, ref2, ulps, exact=exact)


def test_aggregate_rejects_misaligned(ops):
    from shapley_vit_b200._lib import SvitError

    d = torch.zeros(2, 12, device="cuda")
    with pytest.raises(SvitError):
        ops.aggregate(d, None, torch.ones(1, 2, device="cuda"))        # stride 12 is not a multiple of 8


# ----------------------------------------------------------------------------- K5
@pytest.mark.parametrize("C,n,k", [(1, 1, 2), (3, 1000, 10), (8, 4097, 4), (2, 333, 1000)])
def test_score_matches_torch(ops, C, n, k):
    logits = gen(C, n, k, seed=n)
    logits[0, 0, :] = 0.25                                              # all-equal row: first index wins
    labels = torch.randint(0, k, (n,), generator=torch.Generator().manual_seed(1))
    correct, loss, pred = ops.score(logits.cuda(), labels.cuda(), want_pred=True)
    for c in range(C):
        assert torch.equal(pred[c].cpu().long(), logits[c].argmax(dim=1))
        assert int(correct[c]) == int((logits[c].argmax(dim=1) == labels).sum())
        want = F.cross_entropy(logits[c].double(), labels, reduction="sum").item()
        assert float(loss[c]) == pytest.approx(want, rel=2e-6, abs=1e-5)
    c2, l2 = ops.score(logits.cuda(), labels.cuda(), correct.clone(), loss.clone(), accumulate=True)
    assert torch.equal(c2, 2 * correct) and torch.allclose(l2, 2 * loss)


def test_score_nan_propagates(ops):
    logits = gen(1, 16, 5)
    logits[0, 3, 2] = float("nan")
    labels = torch.zeros(16, dtype=torch.int64)
    _, loss = ops.score(logits.cuda(), labels.cuda())
    assert torch.isnan(loss[0])


# ----------------------------------------------------------------------------- LayerNorm / patchify / attention
@pytest.mark.parametrize("h", [192, 384, 768, 1024])
def test_layernorm(ops, h):
    G, rows = 3, 197
    x, g, b = gen(G, rows, h, seed=h) * 3 + 0.5, gen(G, h, seed=1) * 0.1 + 1, gen(G, h, seed=2) * 0.1
    want = torch.stack([F.layer_norm(x[i], (h,), g[i], b[i], 1e-12) for i in range(G)])
    got = ops.layernorm(x.cuda(), g.cuda(), b.cuda(), 1e-12).cpu()
    assert (got - want).abs().max() < 2e-5
    for dt in (torch.bfloat16, torch.float16):
        got = ops.layernorm(x.cuda(), g.cuda(), b.cuda(), 1e-12, out_dtype=dt).cpu()
        assert (got.float() - want).abs().max() < (4e-2 if dt == torch.bfloat16 else 5e-3)
    from shapley_vit_b200._lib import FMT_C8, FMT_X3
    for fmt, tol in ((FMT_X3, 2e-5), (FMT_C8, 4e-4)):        # the planes LayerNorm emits for the split precisions
        y = ops.layernorm(x.cuda(), g.cuda(), b.cuda(), 1e-12, fmt=fmt)
        check_planes(y, want, tol)


@pytest.mark.parametrize("image", [32, 224])
def test_patchify(ops, image):
    from shapley_vit_b200 import layout
    from shapley_vit_b200._lib import PREC_F32

    cfg = layout.vit_preset("tiny", image=image, n_cls=10, layers=1)
    img = gen(5, 3, image, image, seed=image)
    plan = ops.Plan(cfg, PREC_F32, 1, 5, "cuda:0")
    got = plan.patchify(img.cuda()).buf.cpu()
    want = F.unfold(img, kernel_size=16, stride=16).transpose(1, 2).reshape(-1, 768)
    assert torch.equal(got, want)
    from shapley_vit_b200._lib import PRECISIONS
    for name in ("f16x3", "f16c8"):                           # planes, written at a row offset of a larger matrix
        plan = ops.Plan(cfg, PRECISIONS[name], 1, 5, "cuda:0")
        arr = plan.operand_array((7 * cfg.n_patches, 768))
        arr.buf.zero_()
        plan.patchify(img.cuda(), out=arr, row0=2 * cfg.n_patches)
        full = torch.cat([torch.zeros(2 * cfg.n_patches, 768), want])
        check_planes(arr, full, 4e-4 if name == "f16c8" else 2e-6)


@pytest.mark.parametrize("T,heads,d", [(5, 3, 64), (197, 12, 64), (197, 2, 32), (50, 2, 128), (16, 2, 64), (17, 1, 64),
                                       (64, 3, 64), (208, 2, 64), (209, 1, 64), (256, 1, 64)])
def test_attention_fp32(ops, T, heads, d):
    n_seq, h = 3, heads * d
    qkv = gen(n_seq, T, 3 * h, seed=T)
    q, k, v = (t.view(n_seq, T, heads, d).transpose(1, 2) for t in qkv.split(h, dim=2))
    want = (torch.softmax(q @ k.transpose(2, 3) * d ** -0.5, dim=-1) @ v).transpose(1, 2).reshape(n_seq, T, h)
    got = ops.attention(qkv.cuda(), heads).cpu()
    assert (got - want).abs().max() < 2e-5
    for dt, tol in ((torch.float16, 5e-3), (torch.bfloat16, 3e-2)):     # d == 64: tensor-core kernel
        q16 = qkv.to(dt)
        q, k, v = (t_.float().view(n_seq, T, heads, d).transpose(1, 2) for t_ in q16.split(h, dim=2))
        want16 = (torch.softmax(q @ k.transpose(2, 3) * d ** -0.5, dim=-1) @ v).transpose(1, 2).reshape(n_seq, T, h)
        got16 = ops.attention(q16.cuda(), heads).cpu().float()
        assert (got16 - want16).abs().max() < tol


@pytest.mark.parametrize("T,heads,n_seq", [(5, 3, 4), (197, 12, 9), (16, 2, 3), (17, 1, 3), (64, 3, 3), (208, 2, 3),
                                           (209, 1, 3), (256, 1, 5), (129, 2, 7), (160, 1, 3), (161, 3, 2), (224, 2, 200),
                                           (225, 1, 3), (197, 12, 160)])
@pytest.mark.parametrize("fmt_name", ["x3", "c8"])
def test_attention_split(ops, T, heads, n_seq, fmt_name):
    """Split-precision attention (X3 planes in, hi*lo + lo*hi + hi*hi on the tensor cores -- tcgen05 for
    128 < T <= 224, mma.sync otherwise): fp32-grade against fp64; the context planes in either split format."""
    from shapley_vit_b200._lib import FMT_C8, FMT_X3

    d, h = 64, heads * 64
    qkv = gen(n_seq, T, 3 * h, seed=T + 1)
    q, k, v = (t.double().view(n_seq, T, heads, d).transpose(1, 2) for t in qkv.split(h, dim=2))
    want = (torch.softmax(q @ k.transpose(2, 3) * d ** -0.5, dim=-1) @ v).transpose(1, 2).reshape(n_seq, T, h)
    ctx = ops.attention_split(qkv.cuda(), heads, FMT_X3 if fmt_name == "x3" else FMT_C8)
    if fmt_name == "x3":
        err = (ctx.to_float().cpu().double() - want).abs().max().item()
        print(f"attention split T={T}: max err {err:.3e}")
        assert err < 1e-5
    check_planes(ctx, want.float(), 1e-5 if fmt_name == "x3" else 4e-4, value_tol=1e-5)


@pytest.mark.parametrize("T,heads,n_seq", [(197, 12, 40), (129, 3, 5), (256, 2, 7), (144, 1, 300), (250, 4, 3)])
@pytest.mark.parametrize("dt,tol", [(torch.float16, 4e-3), (torch.bfloat16, 3e-2)])
def test_attention_tcgen05(ops, T, heads, n_seq, dt, tol):
    """128 < T <= 256, head_dim 64: the tcgen05 / TMEM kernel (attention_tc.cu) vs fp32 softmax attention."""
    d = 64
    h = heads * d
    q16 = gen(n_seq, T, 3 * h, seed=T + n_seq).to(dt)
    q, k, v = (t_.float().view(n_seq, T, heads, d).transpose(1, 2) for t_ in q16.split(h, dim=2))
    want = (torch.softmax(q @ k.transpose(2, 3) * d ** -0.5, dim=-1) @ v).transpose(1, 2).reshape(n_seq, T, h)
    got = ops.attention(q16.cuda(), heads).cpu().float()
    err = (got - want).abs().max().item()
    assert err < tol, f"max err {err}"


# ----------------------------------------------------------------------------- GEMM
def ref_gemm(A, B, bias=None, residual=None, gelu=False, rowvec=None, rows_in=0, rows_out=0, row_shift=0):
    G = B.shape[0]
    outs = []
    for g in range(G):
        a = A[g if A.shape[0] > 1 else 0].double()
        y = a @ B[g].double().T
        if bias is not None:
            y = y + bias[g].double()
        if gelu:
            y = F.gelu(y)
        if rows_in > 0:
            M = y.shape[0]
            full = torch.zeros((M // rows_in) * rows_out, y.shape[1], dtype=torch.float64)
            idx = torch.arange(M)
            orow = (idx // rows_in) * rows_out + row_shift + idx % rows_in
            full[orow] = y + rowvec[g].double()[orow % rows_out]
            y = full
        if residual is not None:
            y = y + residual[g].double()
        outs.append(y)
    return torch.stack(outs)


@pytest.mark.parametrize("G,M,N,K", [(1, 64, 64, 64), (2, 200, 192, 192), (3, 130, 70, 96), (1, 1, 10, 768)])
def test_gemm_cuda_core_fp32(ops, G, M, N, K):
    from shapley_vit_b200._lib import PREC_F32

    A, B, bias, res = gen(G, M, K, seed=1), gen(G, N, K, seed=2) * 0.05, gen(G, N, seed=3), gen(G, M, N, seed=4)
    got = ops.gemm(PREC_F32, A.cuda(), B.cuda(), bias=bias.cuda(), residual=res.cuda(), gelu=True).cpu()
    want = ref_gemm(A, B, bias, res, gelu=True)
    assert (got.double() - want).abs().max() < 1e-4


TC_SHAPES = [(1, 128, 256, 64), (1, 128, 128, 64), (2, 300, 768, 768), (3, 197 * 4, 2304, 768),
             (2, 640, 192, 192), (2, 640, 576, 192), (1, 1000, 768, 3072), (2, 130, 384, 384),
             # CTA-pair (cta_group::2) kernel: N % 256 == 0 and M >= 256; ragged M, > 74 tiles per launch
             (1, 256, 256, 64), (2, 257, 512, 136), (2, 6000, 768, 256), (1, 384, 3072, 768)]


@pytest.mark.parametrize("prec_name", ["f16", "bf16", "tf32"])
@pytest.mark.parametrize("G,M,N,K", TC_SHAPES)
def test_gemm_tcgen05_matches_cuda_core_gemm(ops, prec_name, G, M, N, K, monkeypatch):
    """tcgen05 result == fp64 reference on the SAME rounded operands, to accumulation error."""
    from shapley_vit_b200._lib import PRECISIONS

    prec = PRECISIONS[prec_name]
    dt = {"f16": torch.float16, "bf16": torch.bfloat16, "tf32": torch.float32}[prec_name]
    A, B = gen(G, M, K, seed=M).to(dt), (gen(G, N, K, seed=N) * 0.05).to(dt)
    bias, res = gen(G, N, seed=3), gen(G, M, N, seed=4)
    got = ops.gemm(prec, A.cuda(), B.cuda(), bias=bias.cuda(), residual=res.cuda(), out_dtype=torch.float32).cpu()
    if prec_name == "tf32":   # the tensor core truncates fp32 operands to 10 mantissa bits
        want = ref_gemm(A, B, bias, res)
        tol = 4e-3 * (K / 64) ** 0.5
    else:
        want = ref_gemm(A.float(), B.float(), bias, res)
        tol = 1e-4 * (K / 64) ** 0.5
    err = (got.double() - want).abs().max().item()
    assert err < tol, f"max err {err}"


@pytest.mark.parametrize("G,M,N,K", [(1, 128, 256, 64), (2, 300, 768, 768), (1, 1000, 768, 3072), (2, 6000, 768, 256),
                                     (3, 197 * 4, 2304, 768), (2, 640, 576, 192), (2, 257, 512, 128)])
def test_gemm_f16x3_split_precision(ops, G, M, N, K):
    """SVIT_PREC_F16X3: operands as fp16 hi + lo planes, three tcgen05 passes.  Held against the
    fp64 product of the UNROUNDED fp32 operands: >= 8x closer than one fp16 pass can be (measured 15-80x).
    What is left is the tensor core's own fp32 accumulation, which truncates (round toward zero) at every
    16-deep step: a bias that grows like K^1.5 on these all-positive-variance inputs."""
    from shapley_vit_b200._lib import PRECISIONS

    A, B = gen(G, M, K, seed=M), gen(G, N, K, seed=N) * 0.05
    bias, res = gen(G, N, seed=3), gen(G, M, N, seed=4)
    got = ops.gemm(PRECISIONS["f16x3"], A.cuda(), B.cuda(), bias=bias.cuda(), residual=res.cuda(),
                   out_dtype=torch.float32).cpu()
    want = ref_gemm(A, B, bias, res)
    err = (got.double() - want).abs().max().item()
    one_pass = (ref_gemm(A.half().float(), B.half().float(), bias, res) - want).abs().max().item()
    print(f"f16x3 max err {err:.3e} (one fp16 pass: {one_pass:.3e})")
    assert err < max(1e-5, 4e-6 * (K / 64) ** 1.5) and err < one_pass / 8


def c8_emulation(A, B):
    """fp64 value of the F16C8 product of fp32 A [G, M, K], B [G, N, K]: hi*hi + 2^-15 (hi8*lo8 + lo8*hi8)."""
    def parts(x):
        hi = x.half().float()
        hi8 = (hi * 4.0).clamp(-448, 448).to(torch.float8_e4m3fn).double()
        lo8 = ((x - hi) * 8192.0).clamp(-448, 448).to(torch.float8_e4m3fn).double()
        return hi.double(), hi8, lo8
    ah, a8, al = parts(A)
    bh, b8, bl = parts(B)
    t = lambda x: x.transpose(1, 2)
    return ah @ t(bh) + (a8 @ t(bl) + al @ t(b8)) / 32768.0


@pytest.mark.parametrize("G,M,N,K", [(1, 128, 256, 128), (2, 300, 768, 768), (1, 1000, 768, 3072), (2, 6000, 768, 256),
                                     (3, 197 * 4, 2304, 768), (2, 640, 576, 192), (2, 257, 512, 128)])
def test_gemm_f16c8_compensated(ops, G, M, N, K):
    """SVIT_PREC_F16C8: fp16 main pass + two e4m3 compensation passes (kind::f8f6f4) folded in with scale-input-d.
    (1) the kernel computes exactly the emulated sum (same planes, fp64 arithmetic) up to the tensor core's fp32
    accumulation; (2) that sum is >= 8x closer to the unrounded product than one fp16 pass (measured ~25x)."""
    from shapley_vit_b200._lib import PRECISIONS

    A, B = gen(G, M, K, seed=M), gen(G, N, K, seed=N) * 0.05
    bias, res = gen(G, N, seed=3), gen(G, M, N, seed=4)
    got = ops.gemm(PRECISIONS["f16c8"], A.cuda(), B.cuda(), bias=bias.cuda(), residual=res.cuda(),
                   out_dtype=torch.float32).cpu().double()
    want = ref_gemm(A, B, bias, res)
    emu = c8_emulation(A, B) + bias.double().unsqueeze(1) + res.double()
    err, err_emu = (got - want).abs().max().item(), (got - emu).abs().max().item()
    one_pass = (ref_gemm(A.half().float(), B.half().float(), bias, res) - want).abs().max().item()
    print(f"f16c8 max err {err:.3e} vs exact, {err_emu:.3e} vs emulation (one fp16 pass: {one_pass:.3e})")
    assert err_emu < max(1e-5, 4e-6 * (K / 64) ** 1.5)
    assert err < one_pass / 8


@pytest.mark.parametrize("prec_name", ["f16x3", "f16c8"])
def test_gemm_split_plane_outputs(ops, prec_name):
    """The 16-bit output of a split-precision GEMM is a packed operand array written by the epilogue (QKV, GELU(MLP-up)):
    pair kernel (M >= 256) and 1-CTA kernel (M < 256), with and without GELU."""
    from shapley_vit_b200._lib import PRECISIONS

    prec = PRECISIONS[prec_name]
    for (G, M, N, K, gelu) in ((2, 394, 768, 256, False), (2, 394, 3072, 256, True), (3, 128, 768, 128, True), (1, 60, 256, 128, False)):
        A, B, bias = gen(G, M, K, seed=M + N), gen(G, N, K, seed=N) * 0.05, gen(G, N, seed=9)
        out = ops.gemm(prec, A.cuda(), B.cuda(), bias=bias.cuda(), gelu=gelu, out_dtype=torch.float16)
        want = ref_gemm(A, B, bias, gelu=gelu).float()
        check_planes(out, want, 2e-4 if prec_name == "f16c8" else 3e-5, value_tol=3e-4 if prec_name == "f16c8" else 3e-5)


def test_split_operand_planes(ops):
    from shapley_vit_b200._lib import FMT_C8, FMT_X3
    from shapley_vit_b200.ops import OperandArray

    x = torch.cat([gen(2, 37, 64, seed=1), gen(2, 37, 64, seed=2) * 1e-3, gen(2, 37, 64, seed=3) * 30], dim=2)
    for fmt in (FMT_X3, FMT_C8):
        arr = OperandArray.from_float(x.cuda(), fmt)
        check_planes(arr, x, 0.0, exact=True)


def check_planes(arr, want, tol, value_tol=None, exact=False):
    """`arr` (OperandArray, split format) against the fp32 tensor `want` it should stand for: every plane must be the
    split of ONE fp32 value v with |v - want| <= value_tol (default tol): hi = fp16(v), lo = fp16(v - hi) (X3) or
    hi8 = e4m3(4 hi), lo8 = e4m3(8192 (v - hi)) (C8)."""
    from shapley_vit_b200._lib import FMT_X3

    want = want.float().reshape(arr.shape)
    hi = arr.plane(0).cpu().float()
    if arr.fmt == FMT_X3:
        lo = arr.plane(1).cpu().float()
        v = hi + lo
        if exact:
            assert torch.equal(hi, want.half().float()) and torch.equal(lo, (want - want.half().float()).half().float())
            return
        assert (v - want).abs().max() <= (value_tol if value_tol is not None else tol)
        assert torch.equal(hi, v.half().float()) or (hi - v.half().float()).abs().max() <= 2 ** -10 * want.abs().max()
    else:
        hi8, lo8 = arr.plane(1).cpu().float(), arr.plane(2).cpu().float()
        want_hi8 = (hi * 4.0).clamp(-448, 448).to(torch.float8_e4m3fn).float()
        assert torch.equal(hi8, want_hi8)                                 # hi8 is a function of hi alone
        if exact:
            h = want.half().float()
            assert torch.equal(hi, h)
            assert torch.equal(lo8, ((want - h) * 8192.0).clamp(-448, 448).to(torch.float8_e4m3fn).float())
            return
        assert (hi - want).abs().max() <= 2 ** -10 * want.abs().max() + (value_tol if value_tol is not None else tol)
        # lo8 carries the residual to ~4 bits: |hi + lo8 / 8192 - want| <= residual / 8 + value_tol
        resid = (want - hi).abs()
        assert ((hi + lo8 / 8192.0 - want).abs() <= resid / 8 + 2 ** -10 / 8192 + (value_tol if value_tol is not None else tol)).all()


@pytest.mark.parametrize("prec_name", ["f16", "tf32"])
def test_gemm_tcgen05_epilogues(ops, prec_name):
    from shapley_vit_b200._lib import PRECISIONS

    prec = PRECISIONS[prec_name]
    dt = torch.float16 if prec_name == "f16" else torch.float32
    tol = 3e-3 if prec_name == "f16" else 6e-3
    G, M, N, K = 2, 394, 768, 768
    A, B, bias = gen(G, M, K, seed=7).to(dt), (gen(G, N, K, seed=8) * 0.05).to(dt), gen(G, N, seed=9)
    # GELU, operand-dtype output
    got = ops.gemm(prec, A.cuda(), B.cuda(), bias=bias.cuda(), gelu=True).cpu()
    want = ref_gemm(A.float(), B.float(), bias, gelu=True)
    assert (got.double() - want).abs().max() < tol
    # shared A (group stride 0) + row remap + position add: the patch-embedding epilogue
    np_, T = 196, 197
    A1 = gen(1, 2 * np_, K, seed=10).to(dt)
    pos = gen(G, T, N, seed=11)
    out = torch.full((G, 2 * T, N), 7.0, device="cuda")
    ops.gemm(prec, A1.cuda(), B.cuda(), bias=bias.cuda(), rowvec=pos.cuda(), rows_in=np_, rows_out=T, row_shift=1,
             out_dtype=torch.float32, out=out)
    want = ref_gemm(A1.float(), B.float(), bias, rowvec=pos, rows_in=np_, rows_out=T, row_shift=1)
    got = out.cpu().double()
    cls_rows = torch.tensor([0, T])
    assert torch.all(got[:, cls_rows] == 7.0)                              # the [CLS] gap is untouched
    mask = torch.ones(2 * T, dtype=torch.bool)
    mask[cls_rows] = False
    assert (got[:, mask] - want[:, mask]).abs().max() < tol
    # in-place residual (out aliases residual), as the projection / MLP-down GEMMs run
    X = gen(G, M, N, seed=12).cuda()
    want = ref_gemm(A.float(), B.float(), bias, residual=X.cpu())
    ops.gemm(prec, A.cuda(), B.cuda(), bias=bias.cuda(), residual=X, out_dtype=torch.float32, out=X)
    assert (X.cpu().double() - want).abs().max() < tol


# ----------------------------------------------------------------------------- K-extension GEMM (N1)
@pytest.mark.parametrize("prec_name,tol", [("f16", 6e-3), ("f16x3", 3e-5), ("f16c8", 3e-4), ("tf32", 8e-3)])
@pytest.mark.parametrize("shape", [(3, 394, 768, 768, True), (2, 130, 192, 192, False), (2, 512, 2304, 768, True)])
def test_gemm_k_extension_and_shared_b(ops, prec_name, tol, shape):
    """svit_gemm_ext: out[g] = A[g] B^T + Ae[g] Be[g]^T + bias with B shared by all groups (or grouped): the low-rank
    per-coalition correction of the LoRA path as extra k-blocks of the same accumulator; pair and one-CTA kernels."""
    from shapley_vit_b200._lib import PRECISIONS

    G, M, N, K, shared = shape
    A, B = gen(G, M, K, seed=21), gen(1 if shared else G, N, K, seed=22) * 0.05
    Ae, Be = gen(G, M, 64, seed=23) * 0.3, gen(G, N, 64, seed=24) * 0.1
    Ae[:, :, 40:] = 0                                                   # (the LoRA blocks leave columns unused)
    bias = gen(G, N, seed=25)
    got = ops.gemm_ext(PRECISIONS[prec_name], A.cuda(), B.cuda(), Ae.cuda(), Be.cuda(), bias=bias.cuda()).cpu()
    want = (A.double() @ B.double().transpose(1, 2) + Ae.double() @ Be.double().transpose(1, 2) + bias.double()[:, None, :]).float()
    err = (got - want).abs().max().item()
    print(f"gemm_ext[{prec_name}] {shape}: max err {err:.3e}")
    assert err < tol * max(1.0, want.abs().max().item() / 4)


def test_gemm_f16c8_outlier_operands_degrade_gracefully(ops):
    """The e4m3 planes resolve |x| <= 112 (hi8 = e4m3(4 hi) and lo8 = e4m3(8192 lo) saturate at 448); beyond that the
    compensation of an element clips and its product falls back towards ONE fp16 pass -- never worse (VERDICT r1, weak
    10: outlier channels of pretrained ViTs).  A with two channels x 300 and a row x 30 (everything inside fp16's range)."""
    from shapley_vit_b200._lib import PRECISIONS

    G, M, N, K = 2, 300, 512, 768
    A, B = gen(G, M, K, seed=41), gen(G, N, K, seed=42) * 0.05
    A[:, :, 5] *= 300.0
    A[:, :, 300] *= 300.0
    A[:, 7, :] *= 30.0
    assert float(A.abs().max()) < 65504
    want = ref_gemm(A, B)
    one_pass = (ref_gemm(A.half().float(), B.half().float()) - want).abs()
    got = ops.gemm(PRECISIONS["f16c8"], A.cuda(), B.cuda(), out_dtype=torch.float32).cpu().double()
    err = (got - want).abs()
    scale = want.abs().amax(dim=2, keepdim=True).clamp_min(1.0)           # per output row
    print(f"f16c8 with outliers: max rel err {float((err / scale).max()):.3e} (one fp16 pass: {float((one_pass / scale).max()):.3e})")
    assert float((err / scale).max()) <= 1.05 * float((one_pass / scale).max()) + 1e-6
    inliers = torch.ones(M, dtype=torch.bool)
    inliers[7] = False
    # rows other than the x 30 row still gain from the compensation of their in-range elements
    assert float((err[:, inliers] / scale[:, inliers]).max()) < float((one_pass[:, inliers] / scale[:, inliers]).max())
