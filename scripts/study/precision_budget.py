#!/usr/bin/env python
"""Which contractions of the ViT forward need compensated (split) operands?  CPU emulation.

Runs a random-init ViT-B/16 (the bench's synthetic weights) on a few images in fp32 and, per op class,
rounds that class's matmul OPERANDS to fp16 (one tensor-core pass) while every other class stays fp32.
Prints max / rms logit error per class: the error budget that decides which ops get extra passes.
    python scripts/study/precision_budget.py [n_images]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from shapley_vit_b200 import layout, synth  # noqa: E402

CLASSES = ["patch", "qkv", "scores", "pv", "proj", "fc1", "fc2"]


def h(x, on):
    return x.half().float() if on else x


def forward(sd, cfg, images, lowp=()):
    lp = set(lowp)
    B = images.shape[0]
    p = cfg.patch
    x = images.unfold(2, p, p).unfold(3, p, p).permute(0, 2, 3, 1, 4, 5).reshape(B, -1, cfg.channels * p * p)
    w = sd["vit.embeddings.patch_embeddings.projection.weight"].reshape(cfg.hidden, -1)
    x = h(x, "patch" in lp) @ h(w, "patch" in lp).T + sd["vit.embeddings.patch_embeddings.projection.bias"]
    x = torch.cat([sd["vit.embeddings.cls_token"].expand(B, -1, -1), x], 1) + sd["vit.embeddings.position_embeddings"]
    H, d = cfg.heads, cfg.hidden // cfg.heads
    for l in range(cfg.layers):
        k = f"vit.encoder.layer.{l}."
        y = torch.nn.functional.layer_norm(x, (cfg.hidden,), sd[k + "layernorm_before.weight"], sd[k + "layernorm_before.bias"], cfg.ln_eps)
        on = "qkv" in lp
        q, kk, v = (h(y, on) @ h(sd[k + f"attention.attention.{n}.weight"], on).T + sd[k + f"attention.attention.{n}.bias"]
                    for n in ("query", "key", "value"))
        sp = lambda t: t.reshape(B, -1, H, d).transpose(1, 2)
        q, kk, v = sp(q), sp(kk), sp(v)
        on = "scores" in lp
        s = (h(q, on or "q" in lp) @ h(kk, on or "k" in lp).transpose(-1, -2)) / d ** 0.5
        pr = s.softmax(-1)
        on = "pv" in lp
        ctx = (h(pr, on or "p" in lp) @ h(v, on or "v" in lp)).transpose(1, 2).reshape(B, -1, cfg.hidden)
        on = "proj" in lp
        x = x + (h(ctx, on) @ h(sd[k + "attention.output.dense.weight"], on).T + sd[k + "attention.output.dense.bias"])
        y = torch.nn.functional.layer_norm(x, (cfg.hidden,), sd[k + "layernorm_after.weight"], sd[k + "layernorm_after.bias"], cfg.ln_eps)
        on = "fc1" in lp
        y = torch.nn.functional.gelu(h(y, on) @ h(sd[k + "intermediate.dense.weight"], on).T + sd[k + "intermediate.dense.bias"])
        on = "fc2" in lp
        x = x + (h(y, on) @ h(sd[k + "output.dense.weight"], on).T + sd[k + "output.dense.bias"])
    x = torch.nn.functional.layer_norm(x[:, 0], (cfg.hidden,), sd["vit.layernorm.weight"], sd["vit.layernorm.bias"], cfg.ln_eps)
    return x @ sd["classifier.weight"].T + sd["classifier.bias"]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    torch.set_num_threads(os.cpu_count())
    cfg = layout.vit_preset("base", image=224, n_cls=10)
    w0 = synth.make_state_dict(cfg, 0)
    sd = synth.make_client_state_dict(w0, 0, 0)
    images, _ = synth.make_val_set(cfg, n, 0)
    sd64 = {k: v.double() for k, v in sd.items()}
    with torch.no_grad():
        ref = forward(sd64, cfg, images.double()).float()
        base = forward(sd, cfg, images)
        print(f"fp32 vs fp64: max {float((base - ref).abs().max()):.2e}")
        margins = ref.topk(2, dim=1).values
        print(f"top-1 margins: min {float((margins[:, 0] - margins[:, 1]).min()):.4f}  median {float((margins[:, 0] - margins[:, 1]).median()):.4f}")
        tot = 0.0
        for c in CLASSES + [tuple(CLASSES), ("scores", "pv"), ("q",), ("k",), ("p",), ("v",), ("q", "p"), ("k", "v")]:
            lp = (c,) if isinstance(c, str) else c
            out = forward(sd, cfg, images, lp)
            e = (out - ref)
            print(f"{'+'.join(lp):40s} max {float(e.abs().max()):.2e}  rms {float(e.pow(2).mean().sqrt()):.2e}")


if __name__ == "__main__":
    main()
