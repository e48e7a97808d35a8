"""The product's estimators (host logic) against outputs of the reference's own estimators
(tests/golden/estimators.json and the estimator block of cfg1_tiny.json), same seeds."""

import numpy as np
import pytest

from helpers import TableGame, load_golden, sv_lists
from oracle.toy_games import toy_table
from shapley_vit_b200 import compared, estimators

TOYS = ("toy5", "toy7")


def toy(meta):
    n = meta["n"]
    return TableGame(n, {tuple(k): v for k, v in toy_table(n, meta["table_seed"]).items()}, meta.get("selection"))


@pytest.fixture(scope="module")
def gold():
    return load_golden("estimators")[0]


@pytest.mark.parametrize("name", TOYS)
def test_exact_family(gold, name):
    g = gold[name]
    assert sv_lists(estimators.shapley_exact(toy(g))) == g["exact"]
    assert sv_lists(estimators.shapley_exact_own(toy(g))) == g["exact_own"]
    for ui in (0, 1):
        sv = compared.MR(ui).compute_shapley_value(toy(g), 0)
        assert [sv[c] for c in sorted(sv)] == g[f"MR_{ui}"]
        sv = compared.TMR(ui).compute_shapley_value(toy(g), 0)
        assert [sv[c] for c in sorted(sv)] == g[f"TMR_{ui}"]


@pytest.mark.parametrize("name", TOYS)
def test_sampling_estimators_reproduce_reference_streams(gold, name):
    g = gold[name]
    assert sv_lists(estimators.shapley_monte_carlo(toy(g), g["m_mc"], seed=g["seed"])) == g["monte_carlo"]
    assert sv_lists(estimators.shapley_comp_contrib(toy(g), g["m_cc"], seed=g["seed"])) == g["comp_contrib"]


@pytest.mark.parametrize("name", TOYS)
def test_gtg_truncated_permutations(gold, name):
    g = gold[name]
    for ui in (0, 1):
        np.random.seed(g["seed"])
        est = compared.GTG(ui)
        sv = est.compute_shapley_value(toy(g), 0)
        assert [sv[c] for c in sorted(sv)] == g[f"GTG_{ui}"]
        assert len(est.Contribution_records) == g[f"GTG_{ui}_records"]


@pytest.mark.parametrize("name", TOYS)
def test_group_testing_sampling_phase(gold, name):
    g = gold[name]
    for ui in (0, 1):
        np.random.seed(g["seed"])
        est = compared.Fed_SV(ui)
        game = toy(g)
        sv = est.compute_shapley_value(game, 0)
        assert np.array_equal(np.asarray(est.UD, dtype=np.float64), np.asarray(g[f"FedSV_{ui}_UD"]))
        # the LP solution is feasible: efficiency + pairwise differences within the final eps
        x = np.array([sv[c + 1] for c in range(g["n"])])
        assert x.sum() == pytest.approx(g[f"FedSV_{ui}_uN"], abs=1e-9)
        assert len(est.Ut[0]) == g[f"FedSV_{ui}_evals"]   # Ut only holds () and N, like the reference


def test_group_testing_rng_rewind_matches_sequential_draws(gold):
    """Replaying the reference's two draws per iteration for k iterations must leave the RNG in
    the state our look-ahead implementation leaves it."""
    g = gold["toy5"]
    np.random.seed(g["seed"])
    est = compared.Fed_SV(0)
    est.compute_shapley_value(toy(g), 0)
    mine_next = np.random.random()
    # sequential replay
    np.random.seed(g["seed"])
    N = g["n"]
    Z = 2 * sum(1 / s for s in range(1, N))
    q = np.array([N / (s * (N - s) * Z) for s in range(1, N)])
    for _ in range(est.iterations):
        size = np.random.choice(np.arange(1, N), p=q)
        np.random.choice(list(range(N)), size=size, replace=False)
    assert np.random.random() == mine_next


def test_partial_participation(gold):
    g = gold["toy6_partial"]
    assert sv_lists(estimators.shapley_exact(toy(g))) == g["exact"]
    assert sv_lists(estimators.shapley_exact_own(toy(g))) == g["exact_own"]
    assert sv_lists(estimators.shapley_comp_contrib(toy(g), 300, seed=g["seed"])) == g["comp_contrib"]


def test_cfg1_game_estimators_from_golden_utilities():
    """All estimators on the utility table the reference measured for BASELINE config 1."""
    meta, arr = load_golden("cfg1_tiny")
    table = {tuple(S): list(arr["utility"][i]) for i, S in enumerate(meta["coalitions"])}
    est = meta["estimators"]
    n = meta["n_clients"]
    mk = lambda: TableGame(n, table)
    assert sv_lists(estimators.shapley_exact(mk())) == est["exact"]
    assert sv_lists(estimators.shapley_exact_own(mk())) == est["exact_own"]
    assert sv_lists(estimators.shapley_monte_carlo(mk(), est["m_mc"], seed=est["seed"])) == est["monte_carlo"]
    assert sv_lists(estimators.shapley_comp_contrib(mk(), est["m_cc"], seed=est["seed"])) == est["comp_contrib"]
    for ui in (0, 1):
        np.random.seed(est["seed"])
        sv = compared.GTG(ui).compute_shapley_value(mk(), 0)
        assert [sv[c] for c in sorted(sv)] == est[f"GTG_{ui}"]
        sv = compared.TMR(ui).compute_shapley_value(mk(), 0)
        assert [sv[c] for c in sorted(sv)] == est[f"TMR_{ui}"]   # dim 0 hits the round truncation


def test_estimators_batch_their_queries():
    g = TableGame(5, {tuple(k): v for k, v in toy_table(5, 11).items()})
    estimators.shapley_exact(g)
    assert g.batched_calls == [31]
    g = TableGame(5, {tuple(k): v for k, v in toy_table(5, 11).items()})
    estimators.shapley_comp_contrib(g, 100, seed=1)
    assert g.batched_calls == [200]


def test_efficiency_axiom_and_null_player():
    g = TableGame(5, {tuple(k): v for k, v in toy_table(5, 11).items()})
    sv = estimators.shapley_exact(g)
    vN = g.eval_utility(range(5))
    assert sum(sv[0].values()) == pytest.approx(vN[0], abs=1e-12)
    assert sv[0][4] == pytest.approx(0.0, abs=1e-12)   # the toy game's last player is null


def test_dispatcher_default_is_comp_contrib(capsys):
    g = TableGame(5, {tuple(k): v for k, v in toy_table(5, 11).items()})
    args = {"seed": 3}
    sv = estimators.call_shapley_computation_method(args, g, None)
    assert args["approximation_method"] == "comp_contrib"
    assert sv_lists(sv) == sv_lists(estimators.shapley_comp_contrib(g, 250, seed=3))
    assert "Comp contrib" in capsys.readouterr().out
    with pytest.raises(ValueError):
        estimators.call_shapley_computation_method({"approximation_method": "nope"}, g, None)


def test_dispatcher_group_testing_values_every_client(capsys):
    """The additive 'group_testing' branch: 0-based membership and keys, so client 0 is valued and the vector sums
    to v(N) (the class used directly keeps the reference's 1-based conventions, compared_methods.py:165)."""
    g = TableGame(5, {tuple(k): v for k, v in toy_table(5, 11).items()})
    sv = estimators.call_shapley_computation_method({"approximation_method": "group_testing", "seed": 2}, g, None)
    vN = g.eval_utility(range(5))
    for dim in range(2):
        assert sorted(sv[dim]) == list(range(5))
        assert sum(sv[dim].values()) == pytest.approx(vN[dim], abs=1e-6)
        assert sv[dim][0] != 0
        # (a 200-sample group test is a noisy estimator; only its constraints are asserted)
    assert "group_testing" in capsys.readouterr().out


def test_fed_sv_solve_feasible_respects_the_constraints():
    rng = np.random.RandomState(0)
    x_true = rng.uniform(-0.2, 0.4, size=6)
    UD = np.subtract.outer(x_true, x_true).astype(np.float32)
    x = compared.Fed_SV(0).solveFeasible(6, float(x_true.sum()), UD)
    assert sum(x) == pytest.approx(float(x_true.sum()), abs=1e-6)
    assert np.allclose(x, x_true, atol=1e-5)
