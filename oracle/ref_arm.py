"""TEST INFRASTRUCTURE ONLY -- the reference's OWN utility loop, timed on the host cores.

Drives the unmodified reference classes (through ``oracle/ref_shim.py``: from ``/root/reference`` in the
build container, from the staged copy ``oracle/_ref`` on the GPU box) exactly as reference start.py:163-187
wires them: HF ``ViTForImageClassification`` models, ``get_difference_between_network_weights``,
``ClientBase`` / ``ServerBase`` / ``Game``; one STEP = ``Game.eval_utility(S)`` of a not yet memoised
coalition (game.py:73-114: ``get_aggregated_model`` with its deepcopy, ``model_agg_lazy`` with its
deepcopy + ``load_state_dict``, ``evaluation`` over the loader, federated_learning/utils.py:864-926).

Only ``bench.py`` (``--impl reference`` and the ``cpu_baseline`` leg) and ``tests/`` import this file.
The process must not see a GPU (the reference would put its models on ``cuda:1`` / ``cuda:0``,
server2.py:17, utils.py:865): bench.py sets ``CUDA_VISIBLE_DEVICES=""`` before importing torch.
"""
from __future__ import annotations

import os
import time
from typing import List, Sequence


class ReferenceArm:
    """``n_sample`` validation images are scored per step (a bounded sample of the workload); the aggregation
    and ``load_state_dict`` parts run at full size.  ``step()`` returns the seconds ONE coalition evaluation
    over ``n_val`` images takes: everything outside ``evaluation`` as measured + ``evaluation`` scaled by
    ``n_val / n_sample`` (its cost is linear in the images: same batches of 128)."""

    def __init__(self, cfg, n_clients: int, n_val: int, n_sample: int, seed: int = 0, threads: int | None = None):
        import torch
        from torch.utils.data import DataLoader

        from oracle import ref_shim
        from oracle.hf_model import build_hf_vit
        from shapley_vit_b200 import synth

        assert not torch.cuda.is_available(), "the reference arm is the CPU path: hide the GPUs (CUDA_VISIBLE_DEVICES='')"
        torch.set_num_threads(threads or os.cpu_count() or 1)
        self.cores = torch.get_num_threads()
        self.ref = ref = ref_shim.load()
        self.ref_root = ref_shim.REFERENCE_ROOT
        self.quiet = ref_shim.quiet
        self.n_val, self.n_sample = n_val, min(n_sample, n_val)
        w0 = synth.make_state_dict(cfg, seed)
        init_model = build_hf_vit(cfg, w0)
        deltas = []
        for j in range(n_clients):
            m = build_hf_vit(cfg, synth.make_client_state_dict(w0, j, seed))
            deltas.append(ref.get_difference_between_network_weights(m, init_model))     # F1
            del m
        images, labels = synth.make_val_set(cfg, self.n_sample, seed)
        self.loader = DataLoader(synth.DictSampleDataset(images, labels), batch_size=128, shuffle=False)
        n_warm = min(16, self.n_sample)      # warm-up steps: the same call on a 16-image loader (thread pools, allocator)
        self.warm_loader = DataLoader(synth.DictSampleDataset(images[:n_warm], labels[:n_warm]), batch_size=128, shuffle=False)
        args = {}
        clients = [ref.ClientBase(i, args, init_model, synth.SizedStub(n)) for i, n in enumerate(synth.client_sizes(n_clients))]
        server = ref.ServerBase(args, init_model, clients, None, self.loader, None)
        self.game = ref.Game(clients, server, init_model, deltas, [True] * n_clients, [0.0, 0.0], 2, args)
        # time evaluation() separately: it is the part that scales with the number of images
        import importlib

        self._gmod = importlib.import_module(ref.Game.__module__)
        self._eval = self._gmod.evaluation
        self._t_eval = 0.0

        def timed_evaluation(*a, **k):
            t0 = time.perf_counter()
            try:
                return self._eval(*a, **k)
            finally:
                self._t_eval += time.perf_counter() - t0

        self._gmod.evaluation = timed_evaluation
        from itertools import combinations

        self.coalitions: List[Sequence[int]] = [c for r in range(1, n_clients + 1) for c in combinations(range(n_clients), r)]
        self.sample = (f"the reference's Game.eval_utility + evaluation (HF ViT fp32 eager, batch 128, {self.cores} threads), "
                       f"1 coalition per step: full-size get_aggregated_model + model_agg_lazy, evaluation on {self.n_sample} of "
                       f"{n_val} images scaled linearly in images")

    def step(self, i: int, warm: bool = False) -> float:
        self.game.server.valid_loader = self.warm_loader if warm else self.loader
        n_sample = len(self.game.server.valid_loader.dataset)
        S = self.coalitions[(37 * i + len(self.coalitions) // 2) % len(self.coalitions)]
        key = frozenset(S)
        for d in self.game.utility:          # never serve a step from the memo
            d.pop(key, None)
        self._t_eval = 0.0
        t0 = time.perf_counter()
        with self.quiet():
            self.game.eval_utility(S)
        total = time.perf_counter() - t0
        self.last = {"coalition": list(S), "seconds_total": total, "seconds_evaluation": self._t_eval}
        return (total - self._t_eval) + self._t_eval * (self.n_val / n_sample)
