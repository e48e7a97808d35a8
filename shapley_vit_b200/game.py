"""The utility game: drop-in for the reference's ``Game`` with a batched, GPU-resident back end.

Mirrors reference fed_client_contribution/game.py:6-36 (constructor and the attributes the
estimators read: ``n``, ``_n_all``, ``selected_clients``, ``client_selection_vector``,
``default_shapley_value``, ``utility_dim``, ``utility``) and :73-114 (``eval_utility``:
memoised v(S) = [acc(S) - acc_0, loss(S) - loss_0], v({}) = [0, 0], members filtered by the
selection vector, ``ValueError('loss is nan')`` on a NaN loss).

Additive: ``eval_utilities(list_of_coalitions)`` evaluates all not-yet-memoised coalitions in
batches (and, under torch.distributed, sharded across ranks with one all-gather of the
per-coalition (correct, loss_sum) pairs), filling the same memo.
"""
from __future__ import annotations

import math
from typing import Dict, FrozenSet, Iterable, List, Optional, Sequence

import torch

from . import _lib, dist, lora
from .engine import CoalitionEngine, ValidationSet
from .fl import _clean_keys, _state_dict_of, config_of


class Game:
    # a shapley game with n players
    def __init__(self, clients, server, init_server_model, client_models, client_selection_vector,
                 previous_utility, utility_dim, server_args):
        self.server = server
        self.clients = clients
        self.init_server_model = init_server_model
        self.client_models = client_models            # list of delta dicts (name -> tensor)
        self.client_selection_vector = client_selection_vector
        self._n_all = len(self.clients)
        self.selected_clients = [i for i in range(self._n_all) if self.client_selection_vector[i]]
        self.n = len(self.selected_clients)
        self.previous_utility = previous_utility
        self.utility_dim = utility_dim
        assert self.utility_dim == 2
        self.server_args = server_args if isinstance(server_args, dict) else {}
        self.utility: List[Dict[FrozenSet[int], float]] = [{} for _ in range(self.utility_dim)]
        self.counts: Dict[FrozenSet[int], tuple] = {}   # coalition -> (correct, loss_sum), for parity checks
        self.compute_default_shapley_value()
        self._engine: Optional[CoalitionEngine] = None
        self._evaluator = None                        # test hook: callable(ratio_rows) -> (correct, loss_sum)
        self.n_evaluated = 0

    def compute_default_shapley_value(self):
        self.default_shapley_value = [{client_id: 0 for client_id in range(self._n_all)}
                                      for _ in range(self.utility_dim)]

    def get_default_shapley_value(self):
        return self.default_shapley_value

    # ------------------------------------------------------------------ #
    @property
    def engine(self) -> CoalitionEngine:
        if self._engine is None:
            a = self.server_args
            precision = _lib.PRECISIONS[a.get("precision", _lib.DEFAULT_PRECISION)]
            device = a.get("device", dist.default_device())
            w0 = _clean_keys(_state_dict_of(self.init_server_model))
            is_lora = lora.is_lora_state_dict(w0)          # PEFT-wrapped model (reference start.py:274-283)
            if is_lora:
                from .models.vit import infer_config
                cfg = getattr(self.init_server_model, "cfg", None) or infer_config(lora.split_state_dict(w0)[0],
                                                                                  heads=a.get("heads"))
            else:
                cfg = config_of(self.init_server_model, a.get("heads"))
            deltas = []
            for j in range(self._n_all):
                d = self.client_models[j]
                deltas.append(_clean_keys(d) if d is not None else {k: torch.zeros_like(v) for k, v in w0.items()})
            loader = self.server.valid_loader
            val = loader if isinstance(loader, ValidationSet) else ValidationSet.from_loader(cfg, loader, precision, device)
            kw = dict(precision=precision, coalition_batch=a.get("coalition_batch", 8),
                      image_chunk=a.get("image_chunk", 128), device=device)
            if is_lora:
                self._engine = lora.LoraCoalitionEngine(cfg, w0, deltas, val, lora_alpha=a.get("lora_alpha", 8.0), **kw)
            else:
                self._engine = CoalitionEngine(cfg, w0, deltas, val, **kw)
        return self._engine

    def _n_val(self) -> int:
        if self._evaluator is not None:
            return self._evaluator.n_val
        return self.engine.n_val

    def _ratio_row(self, coalition: FrozenSet[int]) -> Optional[List[float]]:
        participating = [j for j in coalition if self.client_selection_vector[j]]
        if not participating:
            return None
        ratio = self.server.get_agg_ratio(selected_clients=[self.clients[j] for j in participating])
        row = [0.0] * self._n_all
        for j, r in zip(participating, ratio):
            row[j] = float(r)
        return row

    def eval_utilities(self, coalitions: Sequence[Iterable[int]]) -> List[List[float]]:
        keys = [frozenset(int(j) for j in c) for c in coalitions]
        missing: List[FrozenSet[int]] = []
        seen = set()
        for k in keys:
            if len(k) and k not in self.utility[0] and k not in seen:
                seen.add(k)
                missing.append(k)
        if missing:
            rows, row_keys, w0_only = [], [], []
            for k in missing:
                row = self._ratio_row(k)
                if row is None:
                    w0_only.append(k)      # no selected member: the model is W_0 itself
                    row = [0.0] * self._n_all
                rows.append(row)
                row_keys.append(k)
            ev = self._evaluator if self._evaluator is not None else self.engine
            n_val = self._n_val()
            # fewer pending coalitions than ranks: split the validation set instead (if the evaluator can)
            can_split = self._evaluator is None or getattr(self._evaluator, "supports_image_range", False)
            correct, loss_sum = dist.sharded_evaluate(ev.evaluate, rows, n_val if can_split else 0,
                                                      evaluate_into=getattr(ev, "evaluate_into", None))
            for k, c, l in zip(row_keys, correct, loss_sum):
                if math.isnan(l):
                    raise ValueError("loss is nan")
                self.counts[k] = (int(c), float(l))
                self.utility[0][k] = c / n_val - self.previous_utility[0]   # acc
                self.utility[1][k] = l / n_val - self.previous_utility[1]   # loss
            self.n_evaluated += len(row_keys)
        out = []
        for k in keys:
            if len(k) == 0:
                out.append([0 for _ in range(self.utility_dim)])
            else:
                out.append([self.utility[i][k] for i in range(self.utility_dim)])
        return out

    def eval_utility(self, coalition):
        return self.eval_utilities([coalition])[0]
