"""Round selection for multi-round ("lazy") Shapley: which FL rounds get a Shapley computation.

Drop-in for the three selectors of the reference's ``fed_client_contribution/milp.py`` (SURVEY.md
section 8(f) N4) -- same class names, constructor arguments, ``solve() -> (success, fun, x)`` and
the same solver (``scipy.optimize.milp``, HiGHS), so the chosen rounds are the reference's:

* ``MILP_Shapley``                  (milp.py:8-93)    rounds weighted by who took part in them
* ``MILP_Shapley_Two_Sided``        (milp.py:96-208)  + pairwise participation balance as auxiliary variables
* ``MILP_Shapley_Two_Sided_Approx`` (milp.py:211-305) the balance term folded into the round weights

``selection_matrix[t][i]`` is 1 when client i took part in round t.  All three pick binary
``x_t`` with ``1 <= sum_t x_t <= max_shapley_computation``.  Host-side and tiny (T rounds, n clients);
the selected rounds feed ``fed_shapley.compute_utilities_lazy`` (the GPU path).

A client that never took part makes its column sum zero; like the reference this divides by it
(numpy warning, NaN weights) and the solver then reports failure: ``(False, None, None)``.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
from scipy import optimize


def participation_share(selection_matrix: np.ndarray) -> np.ndarray:
    """Column-normalised selection matrix: share[t][i] = fraction of client i's rounds that is round t."""
    sel = np.asarray(selection_matrix, dtype=float)
    return sel / sel.sum(axis=0)


class _RoundSelector:
    """Common part: argument handling, the cardinality constraint, the HiGHS call."""

    def __init__(self, selection_matrix, max_shapley_computation: Optional[int] = None, gamma: float = 0.5,
                 weight_epochs=None, verbose: bool = False):
        self.selection_matrix = selection_matrix
        self.num_epochs, self.num_clients = selection_matrix.shape[0], selection_matrix.shape[1]
        self.max_shapley_computation = self.num_epochs if max_shapley_computation is None else max_shapley_computation
        assert 0 <= gamma <= 1
        self.gamma = gamma
        self.weight_epochs = np.ones(self.num_epochs) / self.num_epochs if weight_epochs is None else weight_epochs
        self.verbose = verbose
        self.n_aux = 0  # continuous auxiliary variables after the T binary ones

    # --- the linear program: minimise objective . v, lower <= constraint_builder v <= upper ---
    def build_objective(self):
        self.objective = -1 * np.asarray(self.weight_epochs)

    def build_constraints(self):
        self.constraint_builder = np.ones((1, self.num_epochs))

    def build_upper_and_lower_bound(self):
        self.lower_bound = np.array([1])
        self.upper_bound = np.array([self.max_shapley_computation])

    def solve(self) -> Tuple[bool, Optional[float], Optional[np.ndarray]]:
        self.build_upper_and_lower_bound()
        self.build_objective()
        self.build_constraints()
        integrality = np.concatenate([np.ones(self.num_epochs), np.zeros(self.n_aux)])
        res = optimize.milp(c=self.objective, integrality=integrality, bounds=optimize.Bounds(0, 1),
                            constraints=optimize.LinearConstraint(A=self.constraint_builder, lb=self.lower_bound,
                                                                  ub=self.upper_bound))
        if not res.success:
            return res.success, None, None
        x = res.x[:self.num_epochs]
        if self.verbose:
            print(f"rounds <= {self.max_shapley_computation}: value {res.fun}, "
                  f"round weight {self.objective[:self.num_epochs] @ x}, x = {res.x} ({res.message})")
        return res.success, res.fun, x


class MILP_Shapley(_RoundSelector):
    """Round weight = gamma * prior + (1 - gamma) * (normalised sum of the participants' shares)."""

    def __init__(self, selection_matrix, max_shapley_computation=None, gamma=0.5, weight_epochs=None, verbose=False):
        super().__init__(selection_matrix, max_shapley_computation, gamma, weight_epochs, verbose)
        per_round = participation_share(selection_matrix).sum(axis=1)
        self.weight_epochs = self.weight_epochs * self.gamma + per_round / per_round.sum() * (1 - self.gamma)
        if verbose:
            print(f"weight epochs: {self.weight_epochs}")


class MILP_Shapley_Two_Sided(_RoundSelector):
    """Adds one variable z_ij >= |sum_t x_t (share[t][i] - share[t][j])| / n per client pair and charges
    their mean: selected rounds should cover the clients evenly."""

    def __init__(self, selection_matrix, max_shapley_computation=None, gamma=0.5, weight_epochs=None, verbose=False):
        super().__init__(selection_matrix, max_shapley_computation, gamma, weight_epochs, verbose)
        self.auxialiary_variable_dim = self.n_aux = int(self.num_clients * (self.num_clients - 1) / 2)

    def build_objective(self):
        self.objective = np.concatenate([self.gamma * -1 * np.asarray(self.weight_epochs),
                                         (1 - self.gamma) * np.ones(self.n_aux) / self.n_aux])

    def build_constraints(self):
        T, m = self.num_epochs, self.n_aux
        share = participation_share(self.selection_matrix)
        rows = [np.concatenate([np.ones(T), np.zeros(m)])]
        pair = 0
        for i in range(self.num_clients):
            for j in range(i + 1, self.num_clients):
                diff = (share[:, i] - share[:, j]) / self.num_clients
                z = np.zeros(m)
                z[pair] = 1
                rows.append(np.concatenate([-diff, z]))
                rows.append(np.concatenate([diff, z]))
                pair += 1
        self.constraint_builder = np.stack(rows)

    def build_upper_and_lower_bound(self):
        self.lower_bound = np.concatenate([[1], np.zeros(2 * self.n_aux)])
        self.upper_bound = np.concatenate([[self.max_shapley_computation], np.ones(2 * self.n_aux)])


class MILP_Shapley_Two_Sided_Approx(_RoundSelector):
    """The balance term as a per-round penalty: round weight = gamma * prior - (1 - gamma) * (normalised
    sum of pairwise |share[t][i] - share[t][j]|)."""

    def __init__(self, selection_matrix, max_shapley_computation=None, gamma=0.5, weight_epochs=None, verbose=False):
        super().__init__(selection_matrix, max_shapley_computation, gamma, weight_epochs, verbose)
        share = participation_share(selection_matrix)
        spread = np.array([np.abs(row[:, None] - row[None, :])[np.triu_indices(self.num_clients, 1)].sum() for row in share])
        self.weight_epochs = self.weight_epochs * self.gamma - spread / spread.sum() * (1 - self.gamma)
        if verbose:
            print(f"weight epochs: {self.weight_epochs}")
