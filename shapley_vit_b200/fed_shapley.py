"""Multi-round "lazy" utilities: drop-in for the reference's
``fed_client_contribution/utils_fed_shapley.py:146-196`` (``compute_utilities_lazy``).

For every non-empty subset S of the clients (powerset order) the reference rebuilds

    W_S = W_0 + sum_{t = include_from_round .. current_round} FedAvg_t(S restricted to the clients selected in round t)

(one ``get_aggregated_model`` per round, added in round order by ``ServerBase.model_agg_lazy``,
server2.py:121-127) and scores it with ``evaluation``.  Here all subsets of a batch are rebuilt by
``svit_aggregate_onto`` -- one launch per round folds that round's aggregate onto the per-coalition
partial models, in the reference's order of operations -- and scored by the batched forward; under
torch.distributed the subsets are sharded across ranks like ``Game.eval_utilities`` does.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib, dist
from .engine import CoalitionEngine, ValidationSet
from .estimators import powerset
from .fl import _clean_keys, _state_dict_of, config_of


def _num_clients(args, clients_all) -> int:
    n = getattr(args, "num_clients", None)
    if n is None and isinstance(args, dict):
        n = args.get("num_clients")
    return int(n) if n is not None else len(clients_all)


def round_ratio_rows(subsets: Sequence[Sequence[int]], selection_row: Sequence[bool], server, clients_all,
                     n_all: int) -> List[List[float]]:
    """Dense FedAvg ratio rows of every subset for ONE round: members = subset filtered by the round's
    selection vector (utils_fed_shapley.py:173-174), ratios from ``server.get_agg_ratio`` (:179)."""
    rows = []
    for S in subsets:
        members = [j for j in S if selection_row[j]]
        row = [0.0] * n_all
        if members:
            ratio = server.get_agg_ratio(selected_clients=[clients_all[j] for j in members])
            for j, r in zip(members, ratio):
                row[j] = float(r)
        rows.append(row)
    return rows


def compute_utilities_lazy(args, previous_utility, client_model_all_rounds, client_model_selection_matrix, fake_server,
                           clients_all, init_global_model, all_subsets: Dict[Tuple[int, ...], int], utility_dim,
                           current_round, include_from_round, engine: Optional[CoalitionEngine] = None,
                           evaluator=None):
    """Same signature and return value as the reference: ``(utilities, utilities_dict)`` with
    ``utilities[d]`` an array indexed by ``all_subsets[subset]`` and ``utilities_dict[d][subset]``,
    d = 0 accuracy gain, d = 1 loss gain over ``previous_utility``.

    ``client_model_all_rounds[t][j]`` is the delta (name -> tensor) of client j in round t,
    ``client_model_selection_matrix[t][j]`` whether it took part.  ``engine`` / ``evaluator`` are
    optional injections (an existing CoalitionEngine; a test double with ``evaluate_rounds``)."""
    assert utility_dim == 2
    n_all = _num_clients(args, clients_all)
    subsets = list(powerset(range(n_all)).keys())
    rounds = [t for t in range(current_round + 1) if t >= include_from_round]
    if evaluator is None:
        if engine is None:
            a = args if isinstance(args, dict) else getattr(args, "__dict__", {})
            precision = _lib.PRECISIONS[a.get("precision", _lib.DEFAULT_PRECISION)]
            device = a.get("device", dist.default_device())
            cfg = config_of(init_global_model, a.get("heads"))
            loader = fake_server.valid_loader
            val = loader if isinstance(loader, ValidationSet) else ValidationSet.from_loader(cfg, loader, precision, device)
            w0 = _clean_keys(_state_dict_of(init_global_model))
            engine = CoalitionEngine(cfg, w0, None, val, precision=precision, n_clients=n_all,
                                     coalition_batch=a.get("coalition_batch", 8), image_chunk=a.get("image_chunk", 128),
                                     device=device)
        engine.set_round_deltas([[(_clean_keys(d) if d is not None else None) for d in client_model_all_rounds[t]]
                                 for t in rounds])
        evaluator = engine
    rows_per_round = [round_ratio_rows(subsets, client_model_selection_matrix[t], fake_server, clients_all, n_all)
                      for t in rounds]
    n_val = evaluator.n_val

    def evaluate(flat_rows):
        # dist.sharded_evaluate shards a flat list of rows; a row here is the per-round stack of one subset
        if not rounds:      # no round included: every subset's model is W_0 (the reference then scores W_0 2^N - 1 times)
            return evaluator.evaluate_rounds([], n_coalitions=len(flat_rows))
        return evaluator.evaluate_rounds([[fr[t] for fr in flat_rows] for t in range(len(rounds))])

    per_subset = [[rows_per_round[t][i] for t in range(len(rounds))] for i in range(len(subsets))]
    correct, loss_sum = dist.sharded_evaluate(evaluate, per_subset)
    utilities = [np.zeros(len(all_subsets)) for _ in range(utility_dim)]
    utilities_dict: List[Dict[Tuple[int, ...], float]] = [{} for _ in range(utility_dim)]
    for S, c, l in zip(subsets, correct, loss_sum):
        if math.isnan(l):
            raise ValueError("loss is nan")
        u = (c / n_val - previous_utility[0], l / n_val - previous_utility[1])
        for d in range(utility_dim):
            utilities[d][all_subsets[S]] = u[d]
            utilities_dict[d][S] = u[d]
    return utilities, utilities_dict


# --------------------------------------------------------------------------------------------------- #
# Bookkeeping on the utility tables (host-side consumers of the utilities above; numpy only).
# Drop-ins for utils_fed_shapley.py:29-91 and :253-259.  ``all_subsets`` / the utility dictionaries are
# keyed by sorted tuples and, like the reference's ``powerset`` (utils_shapley.py:141-145), hold the
# NON-EMPTY subsets only: the marginal contribution over the empty coalition is not part of these sums
# (the reference's convention, kept so the numbers are the reference's).
# --------------------------------------------------------------------------------------------------- #
from .estimators import ncr  # noqa: E402


def _marginal_terms(players: Sequence[int]):
    """For every player i of ``players``: [(S, S + {i}, 1 / C(n - 1, |S|)) for the non-empty S without i]."""
    n = len(players)
    for pos, i in enumerate(players):
        others = list(players[:pos]) + list(players[pos + 1:])
        yield i, [(S, tuple(sorted(S + (i,))), 1.0 / ncr(n - 1, len(S))) for S in powerset(others)]


def compute_shapley_value_baseline(args, utilities_dict: Dict[Tuple[int, ...], float], idxs_users) -> np.ndarray:
    """Per-round Shapley values of the round's PARTICIPANTS (utils_fed_shapley.py:29-42); zeros elsewhere."""
    out = np.zeros(_num_clients(args, ()))
    players = list(idxs_users)
    for i, terms in _marginal_terms(players):
        out[i] = sum((utilities_dict[si] - utilities_dict[s]) * w for s, si, w in terms) / len(players)
    return out


def compute_shapley_value_groundtruth(args, utilities_dict: Dict[Tuple[int, ...], float]) -> np.ndarray:
    """The same over all ``args.num_users`` clients (utils_fed_shapley.py:45-58)."""
    n = int(args.num_users)
    out = np.zeros(n)
    for i, terms in _marginal_terms(list(range(n))):
        out[i] = sum((utilities_dict[si] - utilities_dict[s]) * w for s, si, w in terms) / n
    return out


def roundly_mask(idxs_users, all_subsets: Dict[Tuple[int, ...], int]) -> np.ndarray:
    """1 at the column of every subset of the round's participants (utils_fed_shapley.py:61-68)."""
    mask = np.zeros(len(all_subsets))
    mask[[all_subsets[s] for s in powerset(idxs_users)]] = 1
    return mask


def compute_shapley_value_from_matrix(args, utility_matrix: np.ndarray, all_subsets: Dict[Tuple[int, ...], int]) -> np.ndarray:
    """Shapley values summed over ``args.epochs`` rounds from the completed [rounds, subsets] utility matrix
    (ComFedSV bookkeeping, utils_fed_shapley.py:71-91)."""
    T, n = int(args.epochs), int(args.num_users)
    per_subset = np.asarray(utility_matrix)[:T].sum(axis=0)          # the rounds enter only through their sum
    out = np.zeros(n)
    for i, terms in _marginal_terms(list(range(n))):
        out[i] = sum((per_subset[all_subsets[si]] - per_subset[all_subsets[s]]) * w for s, si, w in terms) / n
    return out


def get_selection_dict(num_clients: int, idxs_participating_clients) -> Dict[int, bool]:
    """{client: took part} (utils_fed_shapley.py:253-259)."""
    chosen = set(int(i) for i in idxs_participating_clients)
    return {i: i in chosen for i in range(num_clients)}
