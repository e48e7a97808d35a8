"""TEST INFRASTRUCTURE ONLY -- table-driven toy games for pinning the Shapley
estimators (host logic) against the reference's own estimator outputs.

``ToyGame`` exposes exactly the attributes the reference's estimators read from a
``Game`` (reference fed_client_contribution/game.py:19-36):
``n``, ``_n_all``, ``selected_clients``, ``client_selection_vector``,
``default_shapley_value``, ``utility_dim`` and ``eval_utility``.
"""
from __future__ import annotations

import math
from typing import Dict, FrozenSet, Iterable, List

import numpy as np


def toy_table(n: int, seed: int) -> Dict[FrozenSet[int], List[float]]:
    """Saturating two-dimensional utility with one null player (the last one),
    so that GTG's within-permutation truncation (compared_methods.py:306-312)
    actually fires, plus a small seeded interaction term."""
    rng = np.random.RandomState(seed)
    w = rng.uniform(0.2, 1.0, size=n)
    w[n - 1] = 0.0
    noise = rng.normal(0.0, 0.004, size=1 << n)
    table: Dict[FrozenSet[int], List[float]] = {}
    for mask in range(1, 1 << n):
        members = [j for j in range(n) if mask >> j & 1]
        s = float(sum(w[j] for j in members))
        core = mask & ~(1 << (n - 1))          # null player never changes the value
        acc = 0.3 * (1.0 - math.exp(-1.2 * s)) + (noise[core] if core else 0.0)
        loss = -0.8 * (1.0 - math.exp(-0.9 * s)) + 0.5 * (noise[core] if core else 0.0)
        table[frozenset(members)] = [acc, loss]
    return table


class ToyGame:
    def __init__(self, n: int, seed: int, selection=None):
        self._n_all = n
        self.client_selection_vector = list(selection) if selection is not None else [True] * n
        self.selected_clients = [i for i in range(n) if self.client_selection_vector[i]]
        self.n = len(self.selected_clients)
        self.utility_dim = 2
        self.table = toy_table(n, seed)
        self.calls = 0
        self.utility = [{}, {}]
        self.default_shapley_value = [{c: 0 for c in range(n)} for _ in range(2)]

    def eval_utility(self, coalition: Iterable[int]):
        fs = frozenset(int(j) for j in coalition)
        if not fs:
            return [0, 0]
        self.calls += 1
        u = self.table[fs]
        self.utility[0][fs], self.utility[1][fs] = u
        return list(u)
