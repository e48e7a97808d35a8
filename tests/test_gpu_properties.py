"""Size-independent properties at BASELINE config 2's full parameter size (ViT-B/16, P = 85.8 M,
8 clients, 32 coalitions) where the CPU oracle is too slow to be the checker."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_aggregate_full_size_properties():
    from shapley_vit_b200 import layout, ops

    lay = layout.plan_layout(layout.vit_preset("base", image=224, n_cls=10))
    P, N, C = lay.total, 8, 32
    g = torch.Generator(device="cuda").manual_seed(0)
    deltas = torch.randn((N, P), generator=g, device="cuda") * 0.02
    w0 = torch.randn(P, generator=g, device="cuda") * 0.02
    ratios = torch.zeros((C, N), device="cuda")
    for c in range(N):
        ratios[c, c] = 1.0                                   # singletons with ratio 1: W0 + delta_c exactly
    ratios[8, :] = 0.125                                     # uniform grand coalition (power-of-two ratio)
    ratios[9, [1, 5]] = torch.tensor([0.25, 0.75], device="cuda")
    ratios[10:, :] = torch.rand((C - 10, N), generator=g, device="cuda")
    out = ops.aggregate(deltas, w0, ratios)
    for c in range(N):
        assert torch.equal(out[c], w0 + deltas[c])
    # linearity with exactly representable ratios: sequential fp32 sum in ascending order
    acc = 0.125 * deltas[0]
    for j in range(1, N):
        acc = acc + 0.125 * deltas[j]
    assert torch.equal(out[8], w0 + acc)
    assert torch.equal(out[9], w0 + (0.25 * deltas[1] + 0.75 * deltas[5]))
    # coalition order inside a batch does not matter; nor does batch composition
    perm = torch.randperm(C, generator=torch.Generator().manual_seed(1)).cuda()
    out_p = ops.aggregate(deltas, w0, ratios[perm])
    assert torch.equal(out_p, out[perm])
    assert torch.equal(ops.aggregate(deltas, w0, ratios[10:13]), out[10:13])
    # 16-bit outputs are the rounded fp32 outputs
    assert torch.equal(ops.aggregate(deltas, w0, ratios[:4], out_dtype=torch.float16), out[:4].half())


def test_score_split_sum_property():
    from shapley_vit_b200 import ops

    C, n, k = 4, 10_000, 10
    g = torch.Generator(device="cuda").manual_seed(0)
    logits = torch.randn((C, n, k), generator=g, device="cuda")
    labels = torch.randint(0, k, (n,), generator=g, device="cuda")
    c_all, l_all = ops.score(logits, labels)
    c_sum = torch.zeros_like(c_all)
    for s in range(0, n, 2500):
        c_part, _ = ops.score(logits[:, s:s + 2500].contiguous(), labels[s:s + 2500])
        c_sum += c_part
    assert torch.equal(c_sum, c_all)
    want = (logits.argmax(2) == labels).sum(1)
    assert torch.equal(c_all, want)
