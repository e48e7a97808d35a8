"""Baseline estimators: exact MR / round-truncated TMR, truncated guided permutation sampling
(GTG) and group testing (Fed_SV), with batched utility evaluation.

Semantics, RNG streams (the GLOBAL ``np.random``) and accumulation order follow the reference's
``fed_client_contribution/compared_methods.py``: shapley_value :81-91, Fed_SV :106-243,
GTG :251-346, MR :354-388, TMR :396-432.  What changes is scheduling:

* MR / TMR list the power set up front and evaluate it in one batch;
* GTG is adaptive *within* a permutation (truncation once |v(N) - v_{j-1}| < eps) but its RNG
  draws are unconditional within a sweep, so each sweep draws its n guided permutations, evaluates
  all their prefixes speculatively in one batch, then replays the reference's sequential walk
  against the memo (a truncated prefix simply is not read);
* Fed_SV's sample draws never depend on utilities: the first 200 iterations are drawn and
  evaluated as one batch, later ones in small look-ahead chunks; the global RNG is rewound to
  the state the reference would have left.  The feasibility problem the reference hands to a
  Wolfram kernel (:200-243) is solved as an LP with SciPy.
"""
from __future__ import annotations

import copy
from typing import Dict, List, Tuple

import numpy as np
from scipy.special import comb

from .estimators import _prefetch, ncr, powerset


def shapley_value(utility: Dict[Tuple[int, ...], float], game) -> Dict[int, float]:
    """Exact Shapley from a full utility table: marginals weighted 1 / (C(N-1, |S|-1) N)."""
    N = len(game.selected_clients)
    sv = {cid: 0 for cid in range(game._n_all)}
    for S, u_S in utility.items():
        if S == ():
            continue
        for cid in S:
            rest = tuple(i for i in S if i != cid)
            sv[cid] += (u_S - utility[rest]) / (comb(N - 1, len(S) - 1) * N)
    return sv


class ShapleyValue:
    def __init__(self):
        self.FL_name = "Null"
        self.SV = {}


class MR(ShapleyValue):
    """Exact Shapley of one utility dimension over the full power set."""

    def __init__(self, utility_index):
        super().__init__()
        self.SV_t, self.Ut = {}, {}
        self.utility_index = utility_index
        self.full_set = ()

    def compute_shapley_value(self, game, t):
        subsets = list(powerset(game.selected_clients))
        _prefetch(game, subsets)
        util = {S: game.eval_utility(S)[self.utility_index] for S in subsets}
        util[()] = game.eval_utility(())[self.utility_index]
        self.full_set = subsets[-1]
        self.SV_t[t] = shapley_value(util, game)
        self.Ut[t] = copy.deepcopy(util)
        return self.SV_t[t]


class TMR(ShapleyValue):
    """MR with round truncation: all zeros when |v(N) - v({})| <= 0.01."""

    def __init__(self, utility_index):
        super().__init__()
        self.SV_t, self.Ut = {}, {}
        self.utility_index = utility_index
        self.round_trunc_threshold = 0.01

    def compute_shapley_value(self, game, t):
        subsets = list(powerset(game.selected_clients))
        util = {(): game.eval_utility(())[self.utility_index]}
        grand = subsets[-1]
        util[grand] = game.eval_utility(grand)[self.utility_index]
        if abs(util[grand] - util[()]) <= self.round_trunc_threshold:
            return {cid: 0 for cid in range(game._n_all)}
        _prefetch(game, subsets)
        for S in subsets:
            util[S] = game.eval_utility(S)[self.utility_index]
        self.SV_t[t] = shapley_value(util, game)
        self.Ut[t] = copy.deepcopy(util)
        return self.SV_t[t]


class GTG(ShapleyValue):
    """Guided truncated gradient Shapley: every client leads one permutation per sweep."""

    def __init__(self, utility_index):
        super().__init__()
        self.Ut, self.SV_t = {}, {}
        self.utility_index = utility_index
        self.Contribution_records: List[List[float]] = []
        self.eps = 0.001
        self.round_trunc_threshold = 0.01
        self.CONVERGE_MIN_K = 3 * 10
        self.last_k = 10
        self.CONVERGE_CRITERIA = 0.05

    def _running_means(self):
        rec = np.cumsum(self.Contribution_records, 0)
        return rec / np.reshape(np.arange(1, len(self.Contribution_records) + 1), (-1, 1))

    def isnotconverge(self, k):
        if k <= self.CONVERGE_MIN_K:
            return True
        tail = self._running_means()[-self.last_k:]
        errors = np.mean(np.abs(tail[-self.last_k:] - tail[-1:]) / (np.abs(tail[-1:]) + 1e-12), -1)
        return bool(np.max(errors) > self.CONVERGE_CRITERIA)

    def compute_shapley_value(self, game, t):
        players = game.selected_clients
        n_all, N = game._n_all, len(players)
        ui = self.utility_index
        self.Contribution_records = []
        util = {(): game.eval_utility(())[ui]}
        grand = tuple(players)
        util[grand] = game.eval_utility(grand)[ui]
        if abs(util[grand] - util[()]) <= self.round_trunc_threshold:
            return {cid: 0 for cid in range(n_all)}
        k = 0
        while self.isnotconverge(k):
            # one sweep: the n draws below do not depend on any utility
            sweep = [np.concatenate((np.array([lead]), np.random.permutation([p for p in players if p != lead])))
                     for lead in players]
            _prefetch(game, (perm[:j] for perm in sweep for j in range(1, N + 1)
                             if tuple(np.sort(perm[:j], kind="mergesort")) not in util))
            for perm in sweep:
                k += 1
                v = [0 for _ in range(N + 1)]
                v[0] = util[()]
                marginal = {cid: 0 for cid in range(n_all)}
                for j in range(1, N + 1):
                    key = tuple(np.sort(perm[:j], kind="mergesort"))
                    if abs(util[grand] - v[j - 1]) >= self.eps:      # not truncated
                        cached = util.get(key)
                        v[j] = cached if cached is not None else game.eval_utility(key)[ui]
                    else:
                        v[j] = v[j - 1]
                    util[key] = v[j]
                    marginal[perm[j - 1]] = v[j] - v[j - 1]
                self.Contribution_records.append([marginal[cid] for cid in range(n_all)])
        values = self._running_means()[-1:].tolist()[0]
        self.SV_t[t] = {cid: sv for cid, sv in enumerate(values)}
        self.Ut[t] = copy.deepcopy(util)
        return self.SV_t[t]


class Fed_SV(ShapleyValue):
    """Group-testing estimator: sample coalition sizes k ~ q(k) proportional to 1/k + 1/(N-k),
    update the pairwise-difference matrix UD, then find x with sum x = v(N) and
    |x_i - x_j - UD_ij| <= eps."""

    LOOKAHEAD = 16

    def __init__(self, utility_index, one_based_membership: bool = True):
        super().__init__()
        self.Ut, self.SV_t = {}, {}
        self.utility_index = utility_index
        self.Contribution_records = []
        self.CONVERGE_MIN_K = 200
        self.last_k = 10
        self.CONVERGE_CRITERIA = 0.05
        # The reference tests membership with ``S.count(i + 1)`` on 0-based ids
        # (compared_methods.py:165); True reproduces that, False uses the intended ``i in S``.
        self.one_based_membership = one_based_membership
        self.UD, self.iterations = None, 0

    def isnotconverge_Group(self, last_uds, UD):
        if len(last_uds) <= self.CONVERGE_MIN_K:
            return True
        for past in last_uds[-self.last_k:]:
            if np.sum(np.abs(UD - past), axis=(0, 1)) / len(UD[0]) > self.CONVERGE_CRITERIA:
                return True
        return False

    def _draw(self, ids, N, q):
        size = np.random.choice(np.arange(1, N), p=q)
        S = np.random.choice(ids, size=size, replace=False)
        return tuple(np.sort(S, kind="mergesort"))

    def compute_shapley_value(self, game, t):
        ui = self.utility_index
        ids = list(range(game._n_all))
        N = len(ids)
        util = {(): game.eval_utility(())[ui]}
        grand = tuple(ids)
        util[grand] = game.eval_utility(grand)[ui]
        Z = 0
        for s in range(1, N):
            Z += 1 / s
        Z *= 2
        UD = np.zeros([N, N], dtype=np.float32)
        q = np.array([N / (s * (N - s) * Z) for s in range(1, N)])
        shift = 1 if self.one_based_membership else 0
        last_uds: List[np.ndarray] = []
        k = 0
        queue: List[Tuple[Tuple[int, ...], tuple]] = []   # (sample, RNG state after drawing it)
        while self.isnotconverge_Group(last_uds, UD) or k < self.CONVERGE_MIN_K:
            if not queue:   # draw a chunk ahead and evaluate it as one batch
                ahead = max(self.CONVERGE_MIN_K - k, self.LOOKAHEAD)
                for _ in range(ahead):
                    S = self._draw(ids, N, q)
                    queue.append((S, np.random.get_state()))
                _prefetch(game, (S for S, _ in queue if S not in util))
            S, rng_state = queue.pop(0)
            k += 1
            # (the reference never stores sampled utilities in ``util``; the game's memo does that)
            u_S = util[S] if util.get(S) is not None else game.eval_utility(S)[ui]
            UD = (k - 1) / k * UD
            beta = [S.count(i + shift) for i in range(N)]
            for i in range(N):
                for j in range(N):
                    delta_beta = beta[i] - beta[j]
                    if delta_beta != 0:
                        UD[i, j] += delta_beta * u_S * Z / k      # python float into an fp32 cell
            last_uds.append(UD)
        if queue:                      # rewind the global RNG past the unused look-ahead draws
            np.random.set_state(rng_state)
        self.UD, self.iterations = UD, k
        x = self.solveFeasible(N, util[grand], UD)
        self.Ut[t] = copy.deepcopy(util)
        self.SV_t[t] = {cid + 1: sv for cid, sv in enumerate(x)}   # 1-based keys, as the reference
        return self.SV_t[t]

    def solveFeasible(self, agentNum, u_N, UD, lower=None):
        """Feasibility LP replacing the reference's Wolfram FindInstance: find x with
        sum x = u_N and |x_i - x_j - UD_ij| <= eps (i < j), eps starting at 1/(2 N sqrt N) and
        growing by 10 % until feasible; among feasible points take the one minimising the total
        slack.  The reference also imposes x_i > 0.05, which is infeasible for small or negative
        utilities; it is applied only when ``lower`` is given."""
        from scipy.optimize import linprog

        n = agentNum
        pairs = [(i, j) for i in range(n) for j in range(i + 1, n)]
        eps = 1 / np.sqrt(n) / n / 2.0
        for _ in range(400):
            # variables: x (n), slack s_ij (len(pairs)); minimise sum s with s_ij <= eps
            nv = n + len(pairs)
            A_ub, b_ub = [], []
            for p, (i, j) in enumerate(pairs):
                row = np.zeros(nv); row[i], row[j], row[n + p] = 1, -1, -1
                A_ub.append(row); b_ub.append(float(UD[i, j]))
                row = np.zeros(nv); row[i], row[j], row[n + p] = -1, 1, -1
                A_ub.append(row); b_ub.append(-float(UD[i, j]))
            A_eq = np.zeros((1, nv)); A_eq[0, :n] = 1
            bounds = [(lower, None)] * n + [(0, eps)] * len(pairs)
            c = np.concatenate([np.zeros(n), np.ones(len(pairs))])
            res = linprog(c, A_ub=np.array(A_ub), b_ub=np.array(b_ub), A_eq=A_eq, b_eq=[float(u_N)],
                          bounds=bounds, method="highs")
            if res.status == 0:
                return [float(v) for v in res.x[:n]]
            eps *= 1.1
        raise RuntimeError("group-testing feasibility problem has no solution")


# --------------------------------------------------------------------------------------------------- #
# ComFedSV bookkeeping (compared_methods.py:17-73)
# --------------------------------------------------------------------------------------------------- #
def roundly_mask(idxs_users, all_subsets):
    """1 at the column of every non-empty subset of the round's participants (compared_methods.py:66-73)."""
    mask = np.zeros(len(all_subsets))
    mask[[all_subsets[s] for s in powerset(idxs_users)]] = 1
    return mask


def comfedsv(args, utility_matrix, all_subsets):
    """Per-round Shapley values from the (completed) [rounds, subsets] utility matrix; unlike the per-table sums of
    utils_fed_shapley.py this one counts the singleton's own utility as its margin over the empty coalition
    (compared_methods.py:17-44).  Returns (list of {client: value} per round, seconds per round)."""
    import time

    T, N = int(args.rounds), int(args.num_clients)
    terms = []
    for cid in range(N):
        others = [j for j in range(N) if j != cid]
        terms.append([(all_subsets[S], all_subsets[tuple(sorted(S + (cid,)))], 1.0 / ncr(N - 1, len(S)))
                      for S in powerset(others)])
    per_round, seconds = [], []
    for t in range(T):
        t0 = time.time()
        row = utility_matrix[t]
        per_round.append({cid: (sum((row[b] - row[a]) * w for a, b, w in terms[cid]) + row[all_subsets[(cid,)]]) / N
                          for cid in range(N)})
        seconds.append(time.time() - t0)
    return per_round, seconds


def call_comfedsv(game, all_subsets, logger=None):
    """One row of the ComFedSV utility matrix: every subset of the round's selected clients through the game
    (one batched evaluation), plus the round's mask (compared_methods.py:47-62)."""
    sets = list(powerset(game.selected_clients))
    _prefetch(game, sets)
    utilities = [np.zeros(len(all_subsets)) for _ in range(game.utility_dim)]
    for S in sets:
        u = game.eval_utility(S)
        for i in range(game.utility_dim):
            utilities[i][all_subsets[S]] = u[i]
    return utilities, roundly_mask(game.selected_clients, all_subsets)

