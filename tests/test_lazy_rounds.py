"""Multi-round "lazy" utilities: oracle restatement and the product's host
logic against the fixture generated from the reference's own compute_utilities_lazy
(fed_client_contribution/utils_fed_shapley.py:146-196, driven by oracle/make_golden.py)."""
import types

import pytest
import torch

from helpers import load_golden
from oracle import restate
from shapley_vit_b200 import synth
from shapley_vit_b200.estimators import powerset
from shapley_vit_b200.fl import ClientBase, ServerBase


def _inputs():
    meta, _ = load_golden("lazy_rounds")
    cfg, w0, round_sds, selection, n_train, images, labels = synth.lazy_rounds_inputs(
        meta["seed"], meta["n_clients"], meta["n_rounds"], meta["n_val"])
    assert selection == meta["selection"] and list(n_train) == meta["n_train"]
    round_deltas = [[restate.get_difference_between_network_weights(sd, w0) for sd in sds] for sds in round_sds]
    return meta, cfg, w0, round_deltas, selection, n_train, images, labels


def test_oracle_restatement_matches_reference_fixture():
    meta, cfg, w0, round_deltas, selection, n_train, images, labels = _inputs()
    acc0, loss0 = restate.evaluation(w0, cfg, images, labels)
    assert acc0 == pytest.approx(meta["previous_utility"][0], abs=1e-12)
    assert loss0 == pytest.approx(meta["previous_utility"][1], rel=1e-6)
    for case in meta["cases"]:
        util, subsets = restate.compute_utilities_lazy(w0, round_deltas, selection, n_train, cfg, images, labels,
                                                       meta["previous_utility"], case["current_round"],
                                                       case["include_from_round"])
        assert [list(s) for s in subsets] == meta["subsets"]
        assert util[0] == pytest.approx(case["acc"], abs=1e-12)          # integer correct counts / n
        assert util[1] == pytest.approx(case["loss"], abs=2e-6)


class OracleRoundsEvaluator:
    """Test double with the CoalitionEngine.evaluate_rounds interface, backed by oracle/restate.py."""

    def __init__(self, cfg, w0, round_deltas, images, labels):
        self.cfg, self.w0, self.round_deltas, self.images, self.labels = cfg, w0, round_deltas, images, labels
        self.n_val = images.shape[0]

    def evaluate_rounds(self, rows_per_round):
        correct, loss = [], []
        for c in range(len(rows_per_round[0])):
            per_round = []
            for t, rows in enumerate(rows_per_round):
                members = [j for j, r in enumerate(rows[c]) if r != 0]
                if members:
                    per_round.append(restate.get_aggregated_model([self.round_deltas[t][j] for j in members],
                                                                  [rows[c][j] for j in members]))
            sd = restate.model_agg_lazy(self.w0, per_round)
            _, _, det = restate.evaluation(sd, self.cfg, self.images, self.labels, return_details=True)
            correct.append(int(det["correct"]))
            loss.append(float(det["loss_sum"]))
        return correct, loss


def test_product_host_logic_matches_reference_fixture():
    """compute_utilities_lazy of the product (same signature as the reference) with the oracle as evaluator."""
    from shapleyserver.fed_client_contribution.utils_fed_shapley import compute_utilities_lazy

    meta, cfg, w0, round_deltas, selection, n_train, images, labels = _inputs()
    args = types.SimpleNamespace(num_clients=len(n_train))
    clients = [ClientBase(i, {}, None, synth.SizedStub(n)) for i, n in enumerate(n_train)]
    server = ServerBase({}, None, clients, None, None, None)
    subsets = powerset(range(len(n_train)))
    for case in meta["cases"]:
        rounds = [t for t in range(case["current_round"] + 1) if t >= case["include_from_round"]]
        ev = OracleRoundsEvaluator(cfg, w0, [round_deltas[t] for t in rounds], images, labels)
        util, util_dict = compute_utilities_lazy(args, meta["previous_utility"], round_deltas, selection, server, clients,
                                                 None, subsets, 2, case["current_round"], case["include_from_round"],
                                                 evaluator=ev)
        assert list(util[0]) == pytest.approx(case["acc"], abs=1e-12)
        assert list(util[1]) == pytest.approx(case["loss"], abs=2e-6)
        assert util_dict[0][(0, 2)] == pytest.approx(case["acc"][subsets[(0, 2)]], abs=1e-12)


@pytest.mark.gpu
def test_gpu_multi_round_models_bit_exact_and_utilities():
    """svit_aggregate_onto chain == the reference's per-round reconstruction, bit for bit in fp32; the
    utilities of the fp32 mode reproduce the reference fixture (exact counts), f16 stays within its gate."""
    from shapley_vit_b200.engine import CoalitionEngine
    from shapley_vit_b200.fed_shapley import compute_utilities_lazy, round_ratio_rows
    from shapley_vit_b200.layout import pack_state_dict, plan_layout

    meta, cfg, w0, round_deltas, selection, n_train, images, labels = _inputs()
    n = len(n_train)
    clients = [ClientBase(i, {}, None, synth.SizedStub(m)) for i, m in enumerate(n_train)]
    server = ServerBase({}, None, clients, None, None, None)
    subsets = powerset(range(n))
    lay = plan_layout(cfg)
    zero = [{k: v * 0 for k, v in w0.items()} for _ in range(n)]
    for prec, tol_acc, tol_loss in (("f32", 0.0, 1e-5), ("f16c8", 1 / meta["n_val"], 1e-4), ("f16", 4 / meta["n_val"], 2e-3)):
        eng = CoalitionEngine(cfg, w0, zero, images, labels, precision=prec, coalition_batch=4, image_chunk=64,
                              device="cuda:0")
        for case in meta["cases"]:
            rounds = [t for t in range(case["current_round"] + 1) if t >= case["include_from_round"]]
            if prec == "f32":   # the reconstructed weights themselves
                eng.set_round_deltas([round_deltas[t] for t in rounds])
                rows = [round_ratio_rows(list(subsets), selection[t], server, clients, n) for t in rounds]
                got = eng.aggregated_rows_rounds(rows).cpu()
                for i, S in enumerate(subsets):
                    sd = restate.lazy_state_dict(w0, round_deltas, selection, n_train, S, case["current_round"],
                                                 case["include_from_round"])
                    assert torch.equal(got[i], pack_state_dict(lay, sd)), (S, case)
            util, _ = compute_utilities_lazy({"num_clients": n}, meta["previous_utility"], round_deltas, selection, server,
                                             clients, None, subsets, 2, case["current_round"], case["include_from_round"],
                                             engine=eng)
            assert max(abs(a - b) for a, b in zip(util[0], case["acc"])) <= tol_acc + 1e-12
            assert max(abs(a - b) for a, b in zip(util[1], case["loss"])) <= tol_loss


def test_no_round_included_scores_w0_for_every_subset():
    """include_from_round > current_round: the reference's loop (utils_fed_shapley.py:166-176) adds no round, so
    every subset's model is W_0 and every utility is evaluation(W_0) - previous_utility = 0."""
    from shapley_vit_b200.fed_shapley import compute_utilities_lazy

    meta, cfg, w0, round_deltas, selection, n_train, images, labels = _inputs()
    clients = [ClientBase(i, {}, None, synth.SizedStub(n)) for i, n in enumerate(n_train)]
    server = ServerBase({}, None, clients, None, None, None)
    subsets = powerset(range(len(n_train)))

    class W0Only(OracleRoundsEvaluator):
        def evaluate_rounds(self, rows_per_round, n_coalitions=None):
            assert rows_per_round == []
            _, _, det = restate.evaluation(self.w0, self.cfg, self.images, self.labels, return_details=True)
            return [int(det["correct"])] * n_coalitions, [float(det["loss_sum"])] * n_coalitions

    ev = W0Only(cfg, w0, [], images, labels)
    util, _ = compute_utilities_lazy({"num_clients": len(n_train)}, meta["previous_utility"], round_deltas, selection, server,
                                     clients, None, subsets, 2, 0, 1, evaluator=ev)
    assert list(util[0]) == pytest.approx([0.0] * len(subsets), abs=1e-12)
    assert list(util[1]) == pytest.approx([0.0] * len(subsets), abs=2e-6)
