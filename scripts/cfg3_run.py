#!/usr/bin/env python3
"""BASELINE config 3 through the public API, at the real geometry with a reduced budget: 16-client ViT-B/16 @ 224,
seeded Monte-Carlo permutation Shapley (reference utils_shapley.py:248-269) and the truncated GTG estimator
(compared_methods.py:251-346), every utility query through Game.eval_utilities (de-duplicated, batched waves).

    python scripts/cfg3_run.py --perms 200 --val 500 > profiles/r1_cfg3_mc.json

The full configuration (2 000 permutations, 10 000 images) is ~25 000 distinct coalitions x 0.4 s on one GPU;
this run bounds both and reports the throughput of the whole estimator pipeline (host planning + K1 + forward +
score), which is what a user of the reference's entry points sees."""
import argparse
import json
import os
import random
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shapley_vit_b200 import compared, estimators, layout, synth  # noqa: E402
from shapley_vit_b200.engine import CoalitionEngine  # noqa: E402
from shapley_vit_b200.fl import ClientBase, ServerBase  # noqa: E402
from shapley_vit_b200.game import Game  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clients", type=int, default=16)
    ap.add_argument("--perms", type=int, default=200)
    ap.add_argument("--val", type=int, default=500)
    ap.add_argument("--precision", default="f16c8")
    ap.add_argument("--coalition-batch", type=int, default=8)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    cfg = layout.vit_preset("base", image=224, n_cls=10)
    lay = layout.plan_layout(cfg)
    N = a.clients
    deltas = torch.empty((N, lay.total), dtype=torch.float32, device=dev)
    w0_sd = synth.make_state_dict(cfg, a.seed)
    w0 = layout.pack_state_dict(lay, w0_sd).to(dev)
    row = torch.empty(lay.total, dtype=torch.float32).pin_memory()
    for j in range(N):
        cj = synth.make_client_state_dict(w0_sd, j, a.seed)
        layout.pack_state_dict(lay, {k: cj[k] - w0_sd[k] for k in cj}, out=row)
        deltas[j].copy_(row)
    images, labels = synth.make_val_set(cfg, a.val, a.seed)
    eng = CoalitionEngine(cfg, w0, deltas, images, labels, precision=a.precision, coalition_batch=a.coalition_batch,
                          image_chunk=min(128, a.val), device=dev)
    n_train = synth.client_sizes(N)
    clients = [ClientBase(i, {}, None, synth.SizedStub(n)) for i, n in enumerate(n_train)]
    server = ServerBase({}, None, clients, None, eng.val, None)
    c0, l0 = eng.evaluate_state_dict(w0_sd)                      # previous utility = the initial model's
    prev = [c0 / a.val, l0 / a.val]

    def fresh_game():
        g = Game(clients, server, None, [None] * N, [True] * N, prev, 2, {"precision": a.precision})
        g._engine = eng
        return g

    out = {"config": f"BASELINE config 3 geometry: {N}-client ViT-B/16 @224, {a.val} validation images, {a.precision} operands",
           "full_config": "2 000 permutations x 10 000 images", "coalition_batch": a.coalition_batch}
    # --- seeded Monte-Carlo permutations ---
    game = fresh_game()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sv = estimators.shapley_monte_carlo(game, m=a.perms, seed=a.seed)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    distinct = game.n_evaluated
    out["monte_carlo"] = {"permutations": a.perms, "utility_queries": a.perms * N, "distinct_coalitions_evaluated": distinct,
                          "wall_s": wall, "coalition_evals_per_s": distinct / wall,
                          "full_val_equivalent_evals_per_s": distinct / wall * a.val / 10000.0,
                          "shapley_accuracy": [sv[0][c] for c in range(N)], "sum_phi_acc": sum(sv[0].values())}
    # --- GTG: guided permutations with within-permutation truncation (reuses nothing: fresh memo) ---
    game = fresh_game()
    random.seed(a.seed)
    np.random.seed(a.seed)
    t0 = time.perf_counter()
    g = compared.GTG(utility_index=0)
    phi = g.compute_shapley_value(game, 0)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    out["gtg_truncated"] = {"distinct_coalitions_evaluated": game.n_evaluated, "wall_s": wall,
                            "coalition_evals_per_s": game.n_evaluated / wall if wall > 0 else None,
                            "shapley_accuracy": [phi[c] for c in range(N)]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
