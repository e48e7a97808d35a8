"""Pins oracle/restate.py (the CPU restatement) against fixtures produced by the REAL reference
(oracle/make_golden.py): utilities, predictions, logits and aggregated-weight fingerprints of
BASELINE config 1, and ViT-B/16@224 logits."""
import hashlib

import numpy as np
import pytest
import torch

from helpers import load_golden, synthetic_game
from oracle import restate
from shapley_vit_b200 import layout


@pytest.fixture(scope="module")
def cfg1():
    meta, arr = load_golden("cfg1_tiny")
    cfg, w0, clients, deltas, n_train, images, labels = synthetic_game(
        meta["vit"], meta["image"], meta["n_cls"], meta["n_clients"], meta["n_val"], meta["seed"])
    assert n_train == meta["n_train"]
    assert np.array_equal(labels.numpy().astype(np.int8), arr["labels"])
    game = restate.OracleGame(cfg, w0, deltas, n_train, images, labels)
    return meta, arr, cfg, w0, deltas, n_train, game


def test_init_model_utility(cfg1):
    meta, _, _, _, _, _, game = cfg1
    assert game.previous_utility[0] == meta["acc0"]
    assert abs(game.previous_utility[1] - meta["loss0"]) < 1e-6


def test_all_coalition_utilities_and_predictions(cfg1):
    meta, arr, cfg, w0, deltas, n_train, game = cfg1
    for ci, S in enumerate(meta["coalitions"]):
        u = game.eval_utility(S)
        assert u[0] == pytest.approx(arr["utility"][ci, 0], abs=1e-12)      # accuracy: exact counts
        assert u[1] == pytest.approx(arr["utility"][ci, 1], abs=2e-6)       # mean CE
        det = game.details[frozenset(S)]
        assert np.array_equal(det["pred"].numpy().astype(np.int8), arr["pred"][ci])
    for row, ci in enumerate(arr["logits_rows"]):
        mine = game.details[frozenset(meta["coalitions"][ci])]["logits"].numpy()
        assert np.abs(mine - arr["logits"][row]).max() < 1e-5
    assert game.eval_utility([]) == [0, 0]


def test_aggregated_weights_bit_exact(cfg1):
    """W_S from the restatement hashes to the reference's W_S when summed in the reference's
    frozenset order (game.py:90-91)."""
    meta, arr, cfg, w0, deltas, n_train, _ = cfg1
    keys = [k for k, _ in layout.state_dict_spec(cfg)]
    stride = meta["sample_stride"]
    for ci, S in enumerate(meta["coalitions"]):
        order = restate.reference_member_order(S)
        assert order == meta["frozenset_order"][ci]
        sd = restate.coalition_state_dict(w0, deltas, n_train, order)
        flat = torch.cat([sd[k].reshape(-1) for k in keys]).contiguous()
        assert hashlib.sha256(flat.numpy().tobytes()).hexdigest() == meta["agg_sha256"][ci]
        assert np.array_equal(flat[::stride].numpy(), arr["agg_sample"][ci])


def test_exact_shapley_matches_reference(cfg1):
    meta, _, _, _, _, _, game = cfg1
    sv = restate.shapley_exact(game)
    ref = meta["estimators"]["exact"]
    for dim in range(2):
        for c in range(meta["n_clients"]):
            assert sv[dim][c] == pytest.approx(ref[dim][c], abs=2e-6)
    # efficiency axiom: sum phi = v(N)
    vN = game.eval_utility(range(meta["n_clients"]))
    assert sum(sv[0].values()) == pytest.approx(vN[0], abs=1e-12)


def test_vit_base_geometry_logits():
    meta, arr = load_golden("base_probe")
    cfg, w0, clients, deltas, n_train, images, labels = synthetic_game(
        meta["vit"], meta["image"], meta["n_cls"], meta["n_clients"], meta["n_val"], meta["seed"])
    for row, S in enumerate(meta["coalitions"]):
        sd = restate.coalition_state_dict(w0, deltas, n_train, restate.reference_member_order(S))
        mine = restate.vit_forward(sd, cfg, images).numpy()
        assert np.abs(mine - arr["logits"][row]).max() < 2e-5


def test_nan_loss_raises():
    cfg, w0, clients, deltas, n_train, images, labels = synthetic_game(n_clients=1, n_val=4, layers=1)
    bad = {k: v.clone() for k, v in w0.items()}
    bad["classifier.bias"][0] = float("nan")
    with pytest.raises(ValueError, match="loss is nan"):
        restate.evaluation(bad, cfg, images, labels)
