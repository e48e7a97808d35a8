"""Shapley estimators over a utility game, restructured as  plan -> prefetch -> accumulate.

Same estimators, RNG streams and floating-point accumulation order as the reference's
``fed_client_contribution/utils_shapley.py`` (powerset :141-144, ncr :148-152,
shapley_exact_own :156-182, shapley_exact :185-203, shapley_monte_carlo :248-269,
_cc_shap_task :273-304, shapley_comp_contrib :333-362, call_shapley_computation_method
:13-51), so that under identical seeds they query the same coalitions and return the same
vectors.  The difference is *when* utilities are computed: every estimator first lists all the
coalitions it is going to need (its sampling never depends on utility values), hands the
de-duplicated list to ``game.eval_utilities`` -- one batched, multi-GPU evaluation -- and only
then runs its scalar bookkeeping against the memo.  Games without ``eval_utilities`` (e.g. the
reference's own Game) work too, one query at a time.

Unlike the reference, estimators do not accumulate into the shared
``game.default_shapley_value``: each call starts from zeros, and
``shapley_monte_carlo`` shuffles a copy of ``game.selected_clients`` instead of the list itself.
"""
from __future__ import annotations

import math
import random
from itertools import chain, combinations
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np


# ---- combinatorics ------------------------------------------------------------------------
def powerset(iterable) -> Dict[Tuple[int, ...], int]:
    """Non-empty subsets -> running index, ordered by size, then lexicographically."""
    items = list(iterable)
    subsets = chain.from_iterable(combinations(items, r) for r in range(1, len(items) + 1))
    return {tuple(sorted(sub)): idx for idx, sub in enumerate(subsets)}


def ncr(n: int, r: int) -> int:
    return math.comb(n, r)


def _prefetch(game, coalitions: Iterable[Iterable[int]]) -> None:
    batch = getattr(game, "eval_utilities", None)
    if batch is not None:
        batch([tuple(int(j) for j in c) for c in coalitions])


def _shared_seed(seed: Optional[int]) -> Optional[int]:
    """Under torch.distributed every rank must draw the same coalitions (dist.shared_seed); a single process keeps
    the reference's behaviour (None = OS entropy, utils_shapley.py:253, 278)."""
    from . import dist

    return dist.shared_seed(seed)


def _zero_vector(game) -> List[Dict[int, float]]:
    return [{cid: 0 for cid in range(game._n_all)} for _ in range(game.utility_dim)]


# ---- exact --------------------------------------------------------------------------------
def shapley_exact(game):
    """phi_i = sum_{S contains i} c(|S|-1) v(S) - sum_{S without i} c(|S|) v(S),
    c(s) = s! (n-s-1)! / n!, subsets in powerset order."""
    players = list(game.selected_clients)
    n = game.n
    phi = _zero_vector(game)
    f = math.factorial
    weight = [f(s) * f(n - s - 1) / f(n) for s in range(n)]
    subsets = list(powerset(players))
    _prefetch(game, subsets)
    everyone = set(players)
    for S in subsets:
        u = game.eval_utility(S)
        size = len(S)
        for dim in range(game.utility_dim):
            for j in S:
                phi[dim][j] += weight[size - 1] * u[dim]
            for j in everyone - set(S):
                phi[dim][j] -= weight[size] * u[dim]
    return phi


def shapley_exact_own(game):
    """Marginal-contribution form: phi_c = (1/n) [ v({c}) + sum_{s subset of others, s != {}}
    (v(s + c) - v(s)) / C(n-1, |s|) ]."""
    players = list(game.selected_clients)
    n = game.n
    phi = _zero_vector(game)
    _prefetch(game, powerset(players))
    for c in players:
        others = [p for p in players if p != c]
        for s in powerset(others).keys():
            without, with_c = game.eval_utility(s), game.eval_utility(list(s) + [c])
            denom = ncr(n - 1, len(s))
            for dim in range(game.utility_dim):
                phi[dim][c] += (with_c[dim] - without[dim]) / denom
        alone = game.eval_utility([c])
        for dim in range(game.utility_dim):
            phi[dim][c] += alone[dim]
            phi[dim][c] /= n
    return phi


# ---- permutation sampling -------------------------------------------------------------------
def shapley_monte_carlo(game, m: int, seed: Optional[int] = None):
    """m sampled permutations (cumulative ``RandomState.shuffle`` of the player list);
    phi_i = mean marginal contribution of i over the permutations."""
    n = game.n
    rs = np.random.RandomState(_shared_seed(seed))
    order = list(game.selected_clients)
    perms: List[List[int]] = []
    for _ in range(m):
        rs.shuffle(order)
        perms.append(list(order))
    _prefetch(game, (p[:j] for p in perms for j in range(1, n + 1)))
    phi = _zero_vector(game)
    for p in perms:
        prev = [0, 0]
        for j in range(1, n + 1):
            cur = game.eval_utility(p[:j])
            for dim in range(game.utility_dim):
                phi[dim][p[j - 1]] += cur[dim] - prev[dim]
                prev[dim] = cur[dim]
    for dim in range(game.utility_dim):
        for j in order:
            phi[dim][j] /= m
    return phi


# ---- complementary contributions (the reference's default method) ---------------------------
def _cc_draws(n: int, m: int, seed: Optional[int]):
    """The reference's draw sequence: one private-RandomState shuffle of arange(n) (cumulative)
    then one ``random.randint(1, n)`` from the GLOBAL python RNG per sample."""
    seed = _shared_seed(seed)
    rs = np.random.RandomState(seed)
    if seed is not None:
        random.seed(seed)
    idxs = np.arange(n)
    draws = []
    for _ in range(m):
        rs.shuffle(idxs)
        j = random.randint(1, n)
        draws.append((idxs.copy(), j))
    return draws


def _cc_shap_task(game, local_m: int, seed: Optional[int] = None):
    n = game.n
    players = np.array(game.selected_clients)
    draws = _cc_draws(n, local_m, seed)
    _prefetch(game, chain.from_iterable((players[ix[:j]], players[ix[j:]]) for ix, j in draws))
    utility = [np.zeros((n + 1, n)) for _ in range(game.utility_dim)]
    count = np.zeros((n + 1, n))
    for ix, j in draws:
        u_in, u_out = game.eval_utility(players[ix[:j]]), game.eval_utility(players[ix[j:]])
        inside, outside = ix[:j], ix[j:]
        count[j, inside] += 1
        count[n - j, outside] += 1
        for dim in range(game.utility_dim):
            utility[dim][j, inside] += u_in[dim] - u_out[dim]
            utility[dim][n - j, outside] += u_out[dim] - u_in[dim]
    return utility, count


def shapley_comp_contrib(game, m: int, proc_num: int = 1, seed: Optional[int] = None):
    """phi_i = (1/n) sum over strata j of the mean complementary contribution of i at size j."""
    if proc_num < 0:
        raise ValueError("Invalid proc num.")
    n = game.n
    utility, count = _cc_shap_task(game, m, seed)
    per_player = [np.zeros(n) for _ in range(game.utility_dim)]
    for stratum in range(n + 1):
        for p in range(n):
            if count[stratum][p] == 0:
                continue
            for dim in range(game.utility_dim):
                per_player[dim][p] += utility[dim][stratum][p] / count[stratum][p]
    phi = _zero_vector(game)
    for dim in range(game.utility_dim):
        per_player[dim] /= n
        for pos, cid in enumerate(game.selected_clients):
            phi[dim][cid] = per_player[dim][pos]
    return phi


# ---- dispatcher -----------------------------------------------------------------------------
METHODS = ("comp_contrib", "exact", "exact_own", "monte_carlo", "gtg", "mr", "tmr", "group_testing")


def call_shapley_computation_method(args, game, logger=None):
    """The reference hard-wires ``comp_contrib`` with m = 50 n (utils_shapley.py:14-17); that
    stays the default.  Additive keys in ``args``: 'approximation_method' (one of METHODS),
    'm', 'seed', 'utility_index'."""
    args = args if isinstance(args, dict) else {}
    method = args.setdefault("approximation_method", "comp_contrib")
    seed = args.get("seed")
    if method == "comp_contrib":
        shapley_value = shapley_comp_contrib(game, args.get("m", 50 * game.n), seed=seed)
        print(f"Comp contrib: {shapley_value}")
    elif method == "exact":
        shapley_value = shapley_exact(game)
        print(f"Exact: {shapley_value}")
    elif method == "exact_own":
        shapley_value = shapley_exact_own(game)
        print(f"Exact own: {shapley_value}")
    elif method == "monte_carlo":
        shapley_value = shapley_monte_carlo(game, args.get("m", 100), seed=seed)
        print(f"Monte carlo: {shapley_value}")
    elif method in ("gtg", "mr", "tmr", "group_testing"):
        from . import compared

        cls = {"gtg": compared.GTG, "mr": compared.MR, "tmr": compared.TMR, "group_testing": compared.Fed_SV}[method]
        shapley_value = []
        seed = _shared_seed(seed)
        for dim in range(game.utility_dim):
            if seed is not None:
                np.random.seed(seed)
            if method == "group_testing":
                # The class keeps the reference's conventions when used directly (1-based result keys and the
                # ``S.count(i + 1)`` membership test, compared_methods.py:165, under which client 0 is never
                # valued).  The reference's dispatcher has no group-testing branch, so this additive one uses the
                # intended 0-based membership and re-keys the result: every client is valued, sum = v(N).
                sv = compared.Fed_SV(dim, one_based_membership=False).compute_shapley_value(game, 0)
                sv = {cid: sv[cid + 1] for cid in range(game._n_all)}
            else:
                sv = cls(dim).compute_shapley_value(game, 0)
            shapley_value.append({cid: sv.get(cid, 0) for cid in range(game._n_all)})
        print(f"{method}: {shapley_value}")
    else:
        raise ValueError("Unknown Shapley value approximation method")
    print(f"Shapley value sum for each utility: {[sum(list(shapley_value[i].values())) for i in range(2)]}")
    return shapley_value
