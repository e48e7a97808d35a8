"""Host logic of the drop-in Game (memo, dedupe, empty coalition, selection filter, NaN) with a
CPU evaluator standing in for the CUDA engine, checked against the oracle's Game."""
import math

import pytest

from helpers import synthetic_game
from oracle import restate
from oracle_evaluator import OracleEvaluator
from shapley_vit_b200 import estimators
from shapley_vit_b200.fl import ClientBase, ServerBase
from shapley_vit_b200.game import Game
from shapley_vit_b200.synth import SizedStub


def build(n_clients=3, selection=None, n_val=48):
    cfg, w0, clients_sd, deltas, n_train, images, labels = synthetic_game(n_clients=n_clients, n_val=n_val, layers=1)
    og = restate.OracleGame(cfg, w0, deltas, n_train, images, labels, selection=selection)
    clients = [ClientBase(i, {}, None, SizedStub(n)) for i, n in enumerate(n_train)]
    server = ServerBase({}, w0, clients, None, None, None)
    game = Game(clients, server, w0, deltas, selection or [True] * n_clients, list(og.previous_utility), 2, {})
    game._evaluator = OracleEvaluator(cfg, w0, deltas, images, labels)
    return game, og


def test_attributes_match_reference_contract():
    game, og = build()
    assert (game.n, game._n_all, game.selected_clients, game.utility_dim) == (3, 3, [0, 1, 2], 2)
    assert game.default_shapley_value == [{0: 0, 1: 0, 2: 0}, {0: 0, 1: 0, 2: 0}]
    assert game.eval_utility([]) == [0, 0] and game.eval_utility(()) == [0, 0]


def test_utilities_equal_oracle_and_are_memoised():
    game, og = build()
    for S in ([0], [2, 0], (1, 2), [0, 1, 2]):
        assert game.eval_utility(S) == pytest.approx(og.eval_utility(S), abs=1e-9)
    n_calls = len(game._evaluator.calls)
    assert game.eval_utility([0, 2]) == game.eval_utility((2, 0))       # frozenset key
    assert len(game._evaluator.calls) == n_calls                          # memo hit, no evaluation
    assert frozenset({0, 2}) in game.utility[0] and frozenset({0, 2}) in game.utility[1]


def test_batched_queries_are_deduplicated():
    game, og = build()
    out = game.eval_utilities([[0], [1], [0], [], (1,), [0, 1]])
    assert game._evaluator.calls == [3]
    assert out[0] == out[2] and out[1] == out[4] and out[3] == [0, 0]
    assert game.n_evaluated == 3


def test_selection_vector_filters_members():
    sel = [True, False, True]
    game, og = build(selection=sel)
    assert game.selected_clients == [0, 2] and game.n == 2
    # client 1 is not selected: {0,1} aggregates like {0}
    assert game.eval_utility([0, 1]) == pytest.approx(og.eval_utility([0, 1]), abs=1e-9)
    game.eval_utility([0])
    assert game.counts[frozenset({0, 1})] == game.counts[frozenset({0})]
    sv = estimators.shapley_exact(game)
    ref = restate.shapley_exact(og)
    for d in range(2):
        for c in range(3):
            assert sv[d][c] == pytest.approx(ref[d][c], abs=1e-9)


def test_full_exact_shapley_equals_oracle():
    game, og = build()
    sv, ref = estimators.shapley_exact(game), restate.shapley_exact(og)
    for d in range(2):
        for c in range(3):
            assert sv[d][c] == pytest.approx(ref[d][c], abs=1e-9)
    assert game._evaluator.calls == [7]                                   # one batch of 2^3 - 1


def test_nan_loss_raises_value_error():
    game, _ = build()

    class Bad:
        n_val = 48

        def evaluate(self, rows):
            return [0] * len(rows), [math.nan] * len(rows)

    game._evaluator = Bad()
    with pytest.raises(ValueError, match="loss is nan"):
        game.eval_utility([0])


def test_server_ratio_hook_is_used():
    game, _ = build()
    seen = []
    orig = game.server.get_agg_ratio
    game.server.get_agg_ratio = lambda selected_clients=None: (seen.append(len(selected_clients)) or orig(selected_clients))
    game.eval_utility([0, 1])
    assert seen == [2]
    assert game.server.get_agg_ratio(selected_clients=game.clients[:2]) == [1000 / 3000, 2000 / 3000]


def test_eval_cache_holds_the_loader_and_is_bounded(monkeypatch):
    """the cache key is id(loader); the entry must keep the loader alive (no id reuse) and evict."""
    from shapley_vit_b200 import fl

    made = []

    class FakeVS:
        def __init__(self, tag):
            self.tag = tag

    monkeypatch.setattr(fl.ValidationSet, "from_loader", staticmethod(lambda cfg, loader, p, d: made.append(loader) or FakeVS(len(made))))
    fl._EvalCache.clear()
    loaders = [object() for _ in range(6)]
    first = fl._EvalCache.get(loaders[0], "cfg", 0, "cuda:0")
    assert fl._EvalCache.get(loaders[0], "cfg", 0, "cuda:0") is first and len(made) == 1
    for ld in loaders[1:]:
        fl._EvalCache.get(ld, "cfg", 0, "cuda:0")
    assert len(fl._EvalCache.sets) == fl._EvalCache.MAX_SETS
    assert all(entry[0] in loaders for entry in fl._EvalCache.sets.values())      # loaders are referenced
    assert fl._EvalCache.get(loaders[0], "cfg", 0, "cuda:0") is not first            # evicted, rebuilt
    fl._EvalCache.clear()
