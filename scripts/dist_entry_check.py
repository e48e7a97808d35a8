#!/usr/bin/env python3
"""Two NCCL ranks through the estimators with seed=None: the ranks must draw the same coalitions (dist.shared_seed)
and end with identical Shapley vectors; an explicit seed must reproduce the single-rank vector."""
import os
import sys

import torch
import torch.distributed as td

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shapley_vit_b200 import dist, estimators, layout, synth
from shapley_vit_b200.engine import CoalitionEngine
from shapley_vit_b200.fl import ClientBase, ServerBase
from shapley_vit_b200.game import Game

rank, ws, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
td.init_process_group("nccl", rank=rank, world_size=ws, device_id=dev)
cfg = layout.vit_preset("tiny", image=32, n_cls=10, layers=2)
w0 = synth.make_state_dict(cfg, 1)
n = 6
deltas = [{k: v - w0[k] for k, v in synth.make_client_state_dict(w0, j, 1).items()} for j in range(n)]
images, labels = synth.make_val_set(cfg, 256, 1)
eng = CoalitionEngine(cfg, w0, deltas, images, labels, precision="f32", coalition_batch=8, image_chunk=128, device=dev)


def game():
    clients = [ClientBase(i, {}, None, synth.SizedStub(s)) for i, s in enumerate(synth.client_sizes(n))]
    g = Game(clients, ServerBase({}, None, clients, None, eng.val, None), None, [None] * n, [True] * n, [0.0, 0.0], 2, {"precision": "f32"})
    g._engine = eng
    return g


def flat(phi):
    return [phi[d][c] for d in range(2) for c in range(n)]


cc = flat(estimators.shapley_comp_contrib(game(), 12, seed=None))
mc = flat(estimators.shapley_monte_carlo(game(), 5, seed=None))
cc7 = flat(estimators.shapley_comp_contrib(game(), 12, seed=7))
t = torch.tensor(cc + mc + cc7, dtype=torch.float64, device=dev)
lo, hi = t.clone(), t.clone()
td.all_reduce(lo, op=td.ReduceOp.MIN)
td.all_reduce(hi, op=td.ReduceOp.MAX)
same = bool((lo == hi).all())
# single-rank reference for the seeded run: evaluate without the sharded path
g1 = game()
g1_eval = g1.eval_utilities
import shapley_vit_b200.dist as D
real_world = D.world
D.world = lambda: (0, 1)
cc7_single = flat(estimators.shapley_comp_contrib(g1, 12, seed=7))
D.world = real_world
if rank == 0:
    print("ranks agree (seed None):", same)
    print("seeded 2-rank == single-rank:", cc7 == cc7_single)
td.barrier(device_ids=[local])
td.destroy_process_group()
