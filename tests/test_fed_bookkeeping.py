"""Host-side consumers of the multi-round utilities (reference utils_fed_shapley.py:29-91, :253-259) against what
the REFERENCE's functions produced on the same seeded tables (tests/golden/fed_bookkeeping.json, written by
``python -m oracle.make_golden fedbook``)."""
import json
import os
import types

import numpy as np
import pytest

from shapley_vit_b200 import fed_shapley
from shapley_vit_b200.estimators import powerset

GOLD = os.path.join(os.path.dirname(__file__), "golden", "fed_bookkeeping.json")


def cases():
    with open(GOLD) as f:
        return json.load(f)


@pytest.mark.parametrize("case", cases(), ids=lambda c: f"n{c['n']}")
def test_bookkeeping_matches_reference(case):
    n, T, part = case["n"], case["T"], case["participants"]
    args = types.SimpleNamespace(num_clients=n, num_users=n, epochs=T)
    table = {tuple(k): v for k, v in case["table"]}
    all_subsets = powerset(range(n))
    assert list(all_subsets) == [tuple(k) for k, _ in case["table"]]          # same subset order / columns
    assert fed_shapley.compute_shapley_value_baseline(args, table, part) == pytest.approx(case["baseline"], abs=1e-14)
    assert fed_shapley.compute_shapley_value_groundtruth(args, table) == pytest.approx(case["groundtruth"], abs=1e-14)
    assert fed_shapley.roundly_mask(part, all_subsets).tolist() == case["mask"]
    got = fed_shapley.compute_shapley_value_from_matrix(args, np.array(case["matrix"]), all_subsets)
    assert got == pytest.approx(case["from_matrix"], abs=1e-13)
    assert {str(k): v for k, v in fed_shapley.get_selection_dict(n, part).items()} == case["selection"]


def test_non_participants_get_zero_and_module_reexports():
    from shapleyserver.fed_client_contribution import utils_fed_shapley as ufs

    case = cases()[1]
    args = {"num_clients": case["n"]}
    table = {tuple(k): v for k, v in case["table"]}
    v = ufs.compute_shapley_value_baseline(args, table, case["participants"])
    assert all(v[i] == 0 for i in range(case["n"]) if i not in case["participants"])
    for name in ("compute_utilities_lazy", "roundly_mask", "get_selection_dict", "compute_shapley_value_from_matrix",
                 "compute_shapley_value_groundtruth", "powerset", "ncr"):
        assert hasattr(ufs, name)


@pytest.mark.parametrize("case", cases(), ids=lambda c: f"n{c['n']}")
def test_comfedsv_matches_reference(case):
    """compared_methods.py:17-73: per-round values from the utility matrix, and one matrix row through a game."""
    from oracle.toy_games import ToyGame
    from shapleyserver.fed_client_contribution import compared_methods as cm

    n, T, part = case["n"], case["T"], case["participants"]
    all_subsets = powerset(range(n))
    per_round, seconds = cm.comfedsv(types.SimpleNamespace(num_clients=n, rounds=T), np.array(case["matrix"]), all_subsets)
    assert len(per_round) == len(seconds) == T
    for got, want in zip(per_round, case["comfedsv"]):
        assert [got[c] for c in range(n)] == pytest.approx(want, abs=1e-14)
    game = ToyGame(n, seed=n, selection=[c in part for c in range(n)])
    util, mask = cm.call_comfedsv(game, all_subsets, None)
    assert [u.tolist() for u in util] == case["call_comfedsv"]["utilities"]
    assert mask.tolist() == case["call_comfedsv"]["mask"] == cm.roundly_mask(part, all_subsets).tolist()
