"""Worker for test_dist_gloo.py: one rank of a world_size-2 gloo job on CPU."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch.distributed as td  # noqa: E402

from helpers import synthetic_game  # noqa: E402
from oracle import restate  # noqa: E402
from oracle_evaluator import OracleEvaluator  # noqa: E402
from shapley_vit_b200 import dist, estimators  # noqa: E402
from shapley_vit_b200.fl import ClientBase, ServerBase  # noqa: E402
from shapley_vit_b200.game import Game  # noqa: E402
from shapley_vit_b200.synth import SizedStub  # noqa: E402


def main():
    out_path = sys.argv[1]
    distributed = "RANK" in os.environ
    if distributed:
        td.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
    cfg, w0, _, deltas, n_train, images, labels = synthetic_game(n_clients=3, n_val=32, layers=1)
    prev = list(restate.evaluation(w0, cfg, images, labels))
    clients = [ClientBase(i, {}, None, SizedStub(n)) for i, n in enumerate(n_train)]
    game = Game(clients, ServerBase({}, w0, clients, None, None, None), w0, deltas, [True] * 3, prev, 2, {})
    game._evaluator = OracleEvaluator(cfg, w0, deltas, images, labels)
    sv = estimators.shapley_exact(game)
    rank, ws = dist.world()
    # one pending coalition, two ranks: the validation-split axis (every rank scores its half, one all-reduce)
    game2 = Game(clients, ServerBase({}, w0, clients, None, None, None), w0, deltas, [True] * 3, prev, 2, {})
    game2._evaluator = OracleEvaluator(cfg, w0, deltas, images, labels)
    u_single = game2.eval_utility((0, 2))
    split = dict(utility=u_single, counts=list(game2.counts[frozenset((0, 2))]), ranges=game2._evaluator.ranges)
    # stochastic estimators under world_size > 1: seed None must still give every rank the same draws
    # (dist.shared_seed: one entropy draw of rank 0, broadcast), an explicit seed the single-process result
    def fresh():
        g = Game(clients, ServerBase({}, w0, clients, None, None, None), w0, deltas, [True] * 3, prev, 2, {})
        g._evaluator = OracleEvaluator(cfg, w0, deltas, images, labels)
        return g

    as_lists = lambda phi: [[phi[d][c] for c in range(3)] for d in range(2)]
    g_cc, g_mc, g_cc5 = fresh(), fresh(), fresh()
    stochastic = dict(cc_none=as_lists(estimators.shapley_comp_contrib(g_cc, 4, seed=None)),
                      mc_none=as_lists(estimators.shapley_monte_carlo(g_mc, 2, seed=None)),
                      cc_seed5=as_lists(estimators.shapley_comp_contrib(g_cc5, 4, seed=5)),
                      cc_none_keys=sorted(",".join(map(str, sorted(k))) for k in g_cc.counts))
    res = dict(rank=rank, world=ws, stochastic=stochastic, evaluated_here=sum(game._evaluator.calls), split=split,
               counts={",".join(map(str, sorted(k))): v for k, v in game.counts.items()},
               sv=[[sv[d][c] for c in range(3)] for d in range(2)],
               bounds=[dist.shard_bounds(7, r, ws) for r in range(ws)])
    with open(out_path, "w") as f:
        json.dump(res, f)
    if distributed:
        dist.barrier()
        td.destroy_process_group()


if __name__ == "__main__":
    main()
