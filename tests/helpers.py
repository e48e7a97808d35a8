"""Shared test helpers (test infrastructure; may import oracle/)."""
import json
import os

import numpy as np
import torch

from oracle import restate
from shapley_vit_b200 import layout, synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    meta = json.load(open(os.path.join(GOLDEN, name + ".json")))
    arrays = np.load(os.path.join(GOLDEN, name + ".npz")) if os.path.exists(os.path.join(GOLDEN, name + ".npz")) else None
    return meta, arrays


def synthetic_game(vit="tiny", image=32, n_cls=10, n_clients=4, n_val=1000, seed=0, layers=None):
    """The synthetic inputs of SURVEY 8(d): (cfg, w0, client_sds, deltas, n_train, images, labels)."""
    cfg = layout.vit_preset(vit, image=image, n_cls=n_cls, layers=layers)
    w0 = synth.make_state_dict(cfg, seed)
    clients = [synth.make_client_state_dict(w0, j, seed) for j in range(n_clients)]
    deltas = [restate.get_difference_between_network_weights(sd, w0) for sd in clients]
    images, labels = synth.make_val_set(cfg, n_val, seed)
    return cfg, w0, clients, deltas, synth.client_sizes(n_clients), images, labels


class TableGame:
    """A game over a fixed utility table (coalition tuple -> [acc, loss])."""

    def __init__(self, n_all, table, selection=None):
        self._n_all = n_all
        self.client_selection_vector = list(selection) if selection is not None else [True] * n_all
        self.selected_clients = [i for i in range(n_all) if self.client_selection_vector[i]]
        self.n = len(self.selected_clients)
        self.utility_dim = 2
        self.table = {frozenset(k): list(v) for k, v in table.items()}
        self.default_shapley_value = [{c: 0 for c in range(n_all)} for _ in range(2)]
        self.batched_calls = []

    def eval_utility(self, coalition):
        fs = frozenset(int(j) for j in coalition)
        return [0, 0] if not fs else list(self.table[fs])

    def eval_utilities(self, coalitions):
        self.batched_calls.append(len(coalitions))
        return [self.eval_utility(c) for c in coalitions]


def sv_lists(sv):
    return [[float(d[c]) for c in sorted(d)] for d in sv]
