"""Scaled-down runs of BASELINE configs 3 and 4 through the CUDA engine (parity-test cases, not bench lines):
cfg 3 = 16 clients, seeded Monte-Carlo permutations + GTG truncation; cfg 4 = 10-client ViT-L geometry in bf16 with
coalition batches of 32 models per GEMM group.  The checker is the oracle (CPU) on the same seeded inputs."""
import random

import numpy as np
import pytest

from helpers import synthetic_game
from oracle import restate
from oracle_evaluator import OracleEvaluator
from shapley_vit_b200 import compared, estimators
from shapley_vit_b200.fl import ClientBase, ServerBase
from shapley_vit_b200.game import Game
from shapley_vit_b200.synth import SizedStub

pytestmark = pytest.mark.gpu


def _games(n_clients, n_val, prec):
    """The same game twice: CUDA engine (precision `prec`) and oracle evaluator."""
    from shapley_vit_b200.engine import CoalitionEngine

    cfg, w0, _, deltas, n_train, images, labels = synthetic_game(n_clients=n_clients, n_val=n_val, layers=1, seed=3)
    acc0, loss0 = restate.evaluation(w0, cfg, images, labels)
    out = []
    for kind in ("gpu", "oracle"):
        clients = [ClientBase(i, {}, None, SizedStub(n)) for i, n in enumerate(n_train)]
        server = ServerBase({}, w0, clients, None, None, None)
        game = Game(clients, server, w0, deltas, [True] * n_clients, [acc0, loss0], 2, {})
        if kind == "gpu":
            game._engine = CoalitionEngine(cfg, w0, deltas, images, labels, precision=prec, coalition_batch=32,
                                           image_chunk=64, device="cuda:0")
        else:
            game._evaluator = OracleEvaluator(cfg, w0, deltas, images, labels)
        out.append(game)
    return out


def test_cfg3_sixteen_clients_monte_carlo_and_gtg_truncation():
    """16 clients (K1 walks two groups of 8 client rows per tile), fp32 mode: identical permutation streams, identical utilities
    (integer correct counts), hence identical Shapley vectors up to fp32 loss noise."""
    gpu, ora = _games(16, 64, "f32")
    sv = []
    for game in (gpu, ora):
        sv.append(estimators.shapley_monte_carlo(game, m=6, seed=5))
    for d in range(2):
        got, want = [sv[0][d][c] for c in range(16)], [sv[1][d][c] for c in range(16)]
        assert got == pytest.approx(want, abs=1e-5 if d else 1e-12)
    # every coalition both sides evaluated carries the same correct count
    common = set(gpu.counts) & set(ora.counts)
    assert len(common) > 60 and all(gpu.counts[k][0] == ora.counts[k][0] for k in common)
    res = []
    for game in (gpu, ora):
        random.seed(7)
        np.random.seed(7)
        g = compared.GTG(utility_index=0)
        res.append(g.compute_shapley_value(game, 0))
    assert [res[0][c] for c in range(16)] == pytest.approx([res[1][c] for c in range(16)], abs=1e-12)


@pytest.mark.parametrize("prec,tol", [("f32", 2e-4), ("bf16", 8e-2)])
def test_cfg4_vit_large_batches_of_32_models(prec, tol):
    """ViT-L/16 @ 224 geometry (2 layers), 10 clients, one batch of 32 coalitions per GEMM group."""
    from shapley_vit_b200.engine import CoalitionEngine

    cfg, w0, _, deltas, n_train, images, labels = synthetic_game("large", 224, 10, 10, 4, 9, layers=2)
    rng = np.random.RandomState(1)
    coalitions = [tuple(sorted(rng.choice(10, size=rng.randint(1, 11), replace=False).tolist())) for _ in range(32)]
    rows = []
    for S in coalitions:
        r = restate.get_agg_ratio([n_train[j] for j in S])
        row = [0.0] * 10
        for j, v in zip(S, r):
            row[j] = v
        rows.append(row)
    eng = CoalitionEngine(cfg, w0, deltas, images, labels, precision=prec, coalition_batch=32, image_chunk=4,
                          device="cuda:0", keep_logits=True)
    eng.evaluate(rows)
    logits = eng.last_logits.cpu()
    worst = 0.0
    for ci in (0, 7, 19, 31):
        want = restate.vit_forward(restate.coalition_state_dict(w0, deltas, n_train, list(coalitions[ci])), cfg, images)
        worst = max(worst, (logits[ci] - want).abs().max().item())
    print(f"[{prec}] ViT-L, 32 models per group: max |dlogit| = {worst:.3e}")
    assert worst < tol


def test_group_testing_on_the_engine_matches_the_oracle_driven_run():
    """Fed_SV (group testing, compared_methods.py:121-243 with the SciPy LP for the Wolfram solve) through the CUDA engine:
    same global-RNG draw sequence, same sampled coalitions, the same difference matrix UD (integer counts -> identical
    accuracy utilities) and the same feasible point as the oracle-driven run.  5 clients, fp32 mode."""
    gpu, ora = _games(5, 64, "f32")
    res = []
    for game in (gpu, ora):
        np.random.seed(11)
        fed = compared.Fed_SV(utility_index=0)
        fed.CONVERGE_MIN_K = 40                      # 40 samples instead of 200: a parity case, not an estimate
        sv = fed.compute_shapley_value(game, 0)
        res.append((sv, fed.UD.copy(), fed.iterations, sorted(fed.Ut[0])))
    (sv_g, ud_g, it_g, keys_g), (sv_o, ud_o, it_o, keys_o) = res
    assert it_g == it_o and keys_g == keys_o
    assert np.array_equal(ud_g, ud_o)
    assert [sv_g[c + 1] for c in range(5)] == pytest.approx([sv_o[c + 1] for c in range(5)], abs=1e-9)
    assert sum(sv_g.values()) == pytest.approx(gpu.eval_utility(range(5))[0], abs=1e-9)


def test_cfg3_default_precision_keeps_the_monte_carlo_vector():
    """The same 16-client Monte-Carlo run in the default precision (f16c8): utilities within one sample of the oracle's,
    Shapley vector within north_star's 1e-3."""
    gpu, ora = _games(16, 64, "f16c8")
    sv = [estimators.shapley_monte_carlo(game, m=4, seed=9) for game in (gpu, ora)]
    for d in range(2):
        got, want = [sv[0][d][c] for c in range(16)], [sv[1][d][c] for c in range(16)]
        assert max(abs(x - y) for x, y in zip(got, want)) < 1e-3
    common = set(gpu.counts) & set(ora.counts)
    assert all(abs(gpu.counts[k][0] - ora.counts[k][0]) <= 1 for k in common)
