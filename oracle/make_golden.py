"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/ from the REAL reference.

Run in the build container (needs /root/reference and HF transformers):

    python -m oracle.make_golden            # writes tests/golden/*.npz, *.json

The reference has no tests or golden vectors of its own,
so parity is pinned on outputs of the reference's own code driven here through
``oracle/ref_shim.py``:

* ``cfg1_tiny``   -- BASELINE config 1 (4 clients, ViT-Ti/16, 32x32, 1 000 images,
  10 classes): the reference's ``Game.eval_utility`` (game.py:73-114) for all 15
  coalitions -> utilities, correct counts, per-image predictions, logits for a
  few coalitions, fingerprints of the aggregated weights, and the Shapley
  vectors of every reference estimator on that game;
* ``base_probe``  -- ViT-B/16 at 224 px geometry (T = 197, 12 heads), 3 clients,
  8 images: logits for three coalitions (pins the restated forward at the
  BASELINE config 2 geometry);
* ``estimators``  -- every reference estimator on table-driven toy games;
* ``fed_bookkeeping`` -- the numpy bookkeeping on utility tables of utils_fed_shapley.py
  (:29-91, :253-259);
* ``round_select`` -- the three MILP round selectors (fed_client_contribution/milp.py) on seeded
  selection matrices;
* ``lazy_rounds`` -- the multi-round reconstruction ``compute_utilities_lazy``
  (utils_fed_shapley.py:146-196) on 3 clients x 3 FL rounds with a selection matrix.

RNG protocol for the stochastic estimators:
``np.random.RandomState(None)`` is replaced by ``RandomState(seed)``,
``random.seed(seed)`` and ``np.random.seed(seed)`` are called first.
"""
from __future__ import annotations

import contextlib
import hashlib
import json
import os
import random
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim, restate, toy_games  # noqa: E402
from oracle.hf_model import build_hf_vit  # noqa: E402
from shapley_vit_b200 import layout, synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SAMPLE_STRIDE = 997


@contextlib.contextmanager
def seeded(seed: int):
    real = np.random.RandomState
    np.random.RandomState = ref_shim.seeded_randomstate_factory(seed)
    random.seed(seed)
    np.random.seed(seed)
    try:
        yield
    finally:
        np.random.RandomState = real


def sv_to_lists(sv):
    """[{cid: v}, {cid: v}] -> [[v...], [v...]] ordered by client id."""
    return [[float(d[c]) for c in sorted(d)] for d in sv]


def flat_fp32(sd, keys):
    return torch.cat([sd[k].detach().reshape(-1).to(torch.float32) for k in keys])


def fingerprint(sd, keys):
    flat = flat_fp32(sd, keys).contiguous()
    return hashlib.sha256(flat.numpy().tobytes()).hexdigest(), flat[::SAMPLE_STRIDE].clone().numpy()


def run_estimators(ref, make_game, seed: int, m_mc: int, m_cc: int):
    """Every reference estimator, each on a fresh game (they alias
    ``game.default_shapley_value`` -- SURVEY.md section 8(c)(1))."""
    ush, cm = ref.utils_shapley, ref.compared_methods
    out = {"seed": seed, "m_mc": m_mc, "m_cc": m_cc}
    with ref_shim.quiet():
        out["exact"] = sv_to_lists(ush.shapley_exact(make_game()))
        out["exact_own"] = sv_to_lists(ush.shapley_exact_own(make_game()))
        with seeded(seed):
            out["monte_carlo"] = sv_to_lists(ush.shapley_monte_carlo(make_game(), m_mc))
        with seeded(seed):
            out["comp_contrib"] = sv_to_lists(ush.shapley_comp_contrib(make_game(), m_cc))
        for ui in (0, 1):
            g = make_game()
            sv = cm.MR(ui).compute_shapley_value(g, 0)
            out[f"MR_{ui}"] = [float(sv[c]) for c in sorted(sv)]
            g = make_game()
            sv = cm.TMR(ui).compute_shapley_value(g, 0)
            out[f"TMR_{ui}"] = [float(sv[c]) for c in sorted(sv)]
            g = make_game()
            with seeded(seed):
                gtg = cm.GTG(ui)
                sv = gtg.compute_shapley_value(g, 0)
            out[f"GTG_{ui}"] = [float(sv[c]) for c in sorted(sv)]
            out[f"GTG_{ui}_records"] = len(gtg.Contribution_records)
            # group testing: record the sampling phase; the feasibility solve needs a
            # Wolfram kernel (compared_methods.py:200-243) which does not exist here.
            rec = {}

            def fake_solve(self, agentNum, u_N, UD, rec=rec):
                rec["u_N"], rec["UD"] = float(u_N), np.array(UD, dtype=np.float64).tolist()
                return [0.0] * agentNum

            g = make_game()
            real_solve = cm.Fed_SV.solveFeasible
            cm.Fed_SV.solveFeasible = fake_solve
            try:
                with seeded(seed):
                    fed = cm.Fed_SV(ui)
                    fed.compute_shapley_value(g, 0)
            finally:
                cm.Fed_SV.solveFeasible = real_solve
            out[f"FedSV_{ui}_uN"] = rec["u_N"]
            out[f"FedSV_{ui}_UD"] = rec["UD"]
            out[f"FedSV_{ui}_evals"] = len(fed.Ut[0])
    return out


# --------------------------------------------------------------------------- #

def build_reference_game(ref, cfg, w0, client_sds, n_train, images, labels):
    """The object graph of reference start.py:163-187 with distinct models."""
    from torch.utils.data import DataLoader

    ds = synth.DictSampleDataset(images, labels)
    loader = DataLoader(ds, batch_size=128, shuffle=False)
    init_model = build_hf_vit(cfg, w0)
    deltas = []
    for sd in client_sds:
        m = build_hf_vit(cfg, sd)
        deltas.append(ref.get_difference_between_network_weights(m, init_model))
    args = {}
    with ref_shim.quiet():
        acc0, loss0 = ref.evaluation(args, init_model, loader)
    clients = [ref.ClientBase(i, args, init_model, synth.SizedStub(n)) for i, n in enumerate(n_train)]
    server = ref.ServerBase(args, init_model, clients, None, loader, None)

    def make_game(memo=None):
        g = ref.Game(clients, server, init_model, deltas, [True] * len(clients), [acc0, loss0], 2, args)
        if memo is not None:
            g.utility = memo
        return g

    return make_game, server, deltas, (acc0, loss0)


def golden_cfg1(ref):
    t0 = time.time()
    n_clients, n_val, seed = 4, 1000, 0
    cfg = layout.vit_preset("tiny", image=32, n_cls=10)
    keys = [k for k, _ in layout.state_dict_spec(cfg)]
    w0 = synth.make_state_dict(cfg, seed)
    client_sds = [synth.make_client_state_dict(w0, j, seed) for j in range(n_clients)]
    n_train = synth.client_sizes(n_clients)
    images, labels = synth.make_val_set(cfg, n_val, seed)
    make_game, server, deltas, (acc0, loss0) = build_reference_game(
        ref, cfg, w0, client_sds, n_train, images, labels)

    game = make_game()
    coalitions = list(ref.utils_shapley.powerset(game.selected_clients))
    util = np.zeros((len(coalitions), 2))
    pred = np.zeros((len(coalitions), n_val), dtype=np.int8)
    logits = np.zeros((len(coalitions), n_val, cfg.n_cls), dtype=np.float32)
    orders, shas, samples = [], [], []
    for ci, S in enumerate(coalitions):
        with ref_shim.quiet():
            u = game.eval_utility(S)
        util[ci] = u
        orders.append([int(j) for j in frozenset(S)])
        gm_sd = server.global_model.state_dict()      # = W_S (server2.py:121-127)
        sha, samp = fingerprint(gm_sd, keys)
        shas.append(sha)
        samples.append(samp)
        with torch.no_grad():
            lg = torch.cat([server.global_model(images[s:s + 128]).logits for s in range(0, n_val, 128)])
        logits[ci] = lg.numpy()
        pred[ci] = lg.argmax(dim=1).numpy().astype(np.int8)
        print(f"  cfg1 coalition {S}: u={u}", flush=True)

    # cross-check the restatement against the reference right here
    og = restate.OracleGame(cfg, w0, [restate.get_difference_between_network_weights(sd, w0)
                                      for sd in client_sds], n_train, images, labels)
    assert abs(og.previous_utility[0] - acc0) == 0 and abs(og.previous_utility[1] - loss0) < 1e-6
    for ci, S in enumerate(coalitions):
        u = og.eval_utility(S)
        assert abs(u[0] - util[ci, 0]) < 1e-12 and abs(u[1] - util[ci, 1]) < 1e-5, (S, u, util[ci])

    memo = game.utility
    est = run_estimators(ref, lambda: make_game(memo), seed=1234, m_mc=20, m_cc=50 * n_clients)

    np.savez_compressed(
        os.path.join(GOLD, "cfg1_tiny.npz"),
        coalition_mask=np.array([[j in S for j in range(n_clients)] for S in coalitions]),
        utility=util, pred=pred, logits=logits[[0, 7, 14]], logits_rows=np.array([0, 7, 14]),
        agg_sample=np.stack(samples), labels=labels.numpy().astype(np.int8))
    meta = dict(
        config="BASELINE config 1", vit="tiny", image=32, n_cls=10, n_clients=n_clients, n_val=n_val,
        seed=seed, n_train=n_train, acc0=acc0, loss0=loss0, coalitions=[list(S) for S in coalitions],
        frozenset_order=orders, agg_sha256=shas, sample_stride=SAMPLE_STRIDE, estimators=est,
        transformers=__import__("transformers").__version__, torch=torch.__version__,
        generated_by="oracle/make_golden.py via the reference's Game.eval_utility (game.py:73-114)")
    with open(os.path.join(GOLD, "cfg1_tiny.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(f"cfg1_tiny done in {time.time() - t0:.1f}s")


def golden_base_probe(ref):
    t0 = time.time()
    n_clients, n_val, seed = 3, 8, 3
    cfg = layout.vit_preset("base", image=224, n_cls=10)
    w0 = synth.make_state_dict(cfg, seed)
    client_sds = [synth.make_client_state_dict(w0, j, seed) for j in range(n_clients)]
    n_train = synth.client_sizes(n_clients)
    images, labels = synth.make_val_set(cfg, n_val, seed)
    make_game, server, deltas, (acc0, loss0) = build_reference_game(
        ref, cfg, w0, client_sds, n_train, images, labels)
    game = make_game()
    coalitions = [(0,), (0, 2), (0, 1, 2)]
    logits, util = [], []
    for S in coalitions:
        with ref_shim.quiet():
            util.append(game.eval_utility(S))
        with torch.no_grad():
            logits.append(server.global_model(images).logits.numpy())
    # restatement cross-check
    d = [restate.get_difference_between_network_weights(sd, w0) for sd in client_sds]
    for S, lg in zip(coalitions, logits):
        sd = restate.coalition_state_dict(w0, d, n_train, restate.reference_member_order(S))
        mine = restate.vit_forward(sd, cfg, images).numpy()
        err = np.abs(mine - lg).max()
        print(f"  base_probe {S}: restatement max|dlogit| = {err:.3e}")
        assert err < 2e-5
    np.savez_compressed(os.path.join(GOLD, "base_probe.npz"), logits=np.stack(logits),
                        utility=np.array(util), labels=labels.numpy())
    with open(os.path.join(GOLD, "base_probe.json"), "w") as f:
        json.dump(dict(vit="base", image=224, n_cls=10, n_clients=n_clients, n_val=n_val, seed=seed,
                       n_train=n_train, acc0=acc0, loss0=loss0,
                       coalitions=[list(S) for S in coalitions]), f, indent=1)
    print(f"base_probe done in {time.time() - t0:.1f}s")


def golden_estimators(ref):
    out = {}
    for name, n, tseed, seed in (("toy5", 5, 11, 7), ("toy7", 7, 5, 99)):
        out[name] = dict(n=n, table_seed=tseed,
                         **run_estimators(ref, lambda: toy_games.ToyGame(n, tseed), seed=seed,
                                          m_mc=40, m_cc=50 * n))
    # partial participation: client 2 not selected (Game filters members, game.py:90-91)
    sel = [True, True, False, True, True, True]
    ush = ref.utils_shapley
    with ref_shim.quiet():
        out["toy6_partial"] = dict(
            n=6, table_seed=21, selection=sel,
            exact=sv_to_lists(ush.shapley_exact(toy_games.ToyGame(6, 21, sel))),
            exact_own=sv_to_lists(ush.shapley_exact_own(toy_games.ToyGame(6, 21, sel))))
        with seeded(5):
            out["toy6_partial"]["comp_contrib"] = sv_to_lists(
                ush.shapley_comp_contrib(toy_games.ToyGame(6, 21, sel), 300))
        out["toy6_partial"]["seed"] = 5
    with open(os.path.join(GOLD, "estimators.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("estimators done")


def golden_lazy(ref):
    """compute_utilities_lazy (reference utils_fed_shapley.py:146-196) on 3 clients x 3 FL rounds with a
    per-round selection matrix, 2-layer ViT-Ti/16 @ 32 px, 200 images."""
    import importlib
    import types

    from torch.utils.data import DataLoader

    ufs = importlib.import_module("refshapleyserver.fed_client_contribution.utils_fed_shapley")
    cfg, w0, round_sds, selection, n_train, images, labels = synth.lazy_rounds_inputs()
    loader = DataLoader(synth.DictSampleDataset(images, labels), batch_size=128, shuffle=False)
    init_model = build_hf_vit(cfg, w0)
    round_deltas = [[ref.get_difference_between_network_weights(build_hf_vit(cfg, sd), init_model) for sd in sds]
                    for sds in round_sds]
    args = types.SimpleNamespace(num_clients=len(n_train))
    with ref_shim.quiet():
        acc0, loss0 = ref.evaluation(args, init_model, loader)
    clients = [ref.ClientBase(i, args, init_model, synth.SizedStub(n)) for i, n in enumerate(n_train)]
    server = ref.ServerBase(args, init_model, clients, None, loader, None)
    subsets = ref.utils_shapley.powerset(range(len(n_train)))
    out = {"seed": 11, "n_clients": len(n_train), "n_rounds": len(selection), "n_val": int(images.shape[0]),
           "selection": selection, "n_train": list(n_train), "previous_utility": [acc0, loss0],
           "subsets": [list(k) for k in subsets.keys()], "cases": []}
    for current_round, include_from in ((2, 0), (2, 1), (1, 0)):
        with ref_shim.quiet():
            util, _ = ufs.compute_utilities_lazy(args, [acc0, loss0], round_deltas, selection, server, clients, init_model,
                                                 subsets, 2, current_round, include_from)
        out["cases"].append({"current_round": current_round, "include_from_round": include_from,
                             "acc": [float(x) for x in util[0]], "loss": [float(x) for x in util[1]]})
    with open(os.path.join(GOLD, "lazy_rounds.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("lazy_rounds done")


def round_select_cases():
    """Seeded selection matrices (every client takes part at least once) and selector settings."""
    cases = []
    rng = np.random.RandomState(7)
    for T, n, kmax, gamma, weighted in ((6, 3, 2, 0.5, False), (10, 4, 3, 0.3, False), (8, 5, None, 0.5, True),
                                        (12, 4, 5, 0.9, True), (5, 3, 1, 0.0, False), (9, 6, 4, 1.0, False)):
        sel = (rng.rand(T, n) < 0.6).astype(float)
        for i in range(n):
            if sel[:, i].sum() == 0:
                sel[rng.randint(T), i] = 1.0
        w = None
        if weighted:
            w = rng.rand(T)
            w = w / w.sum()
        cases.append({"selection": sel.tolist(), "kmax": kmax, "gamma": gamma, "weights": None if w is None else w.tolist()})
    return cases


def golden_round_select(ref):
    """The three MILP round selectors of the reference (fed_client_contribution/milp.py) on seeded cases."""
    import importlib

    milp = importlib.import_module("refshapleyserver.fed_client_contribution.milp")
    out = []
    for case in round_select_cases():
        sel = np.array(case["selection"])
        rec = dict(case)
        for name in ("MILP_Shapley", "MILP_Shapley_Two_Sided", "MILP_Shapley_Two_Sided_Approx"):
            w = None if case["weights"] is None else np.array(case["weights"])
            ok, fun, x = getattr(milp, name)(sel, case["kmax"], case["gamma"], w).solve()
            rec[name] = {"success": bool(ok), "fun": None if fun is None else float(fun),
                         "x": None if x is None else [float(v) for v in x]}
        out.append(rec)
    with open(os.path.join(GOLD, "round_select.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("round_select done")


def golden_fed_bookkeeping(ref):
    """The numpy bookkeeping of utils_fed_shapley.py (:29-91, :253-259) on seeded utility tables."""
    import importlib
    import types

    ufs = importlib.import_module("refshapleyserver.fed_client_contribution.utils_fed_shapley")
    rng = np.random.RandomState(3)
    out = []
    for n, T, part in ((4, 3, [0, 2, 3]), (5, 4, [1, 2]), (3, 2, [0, 1, 2])):
        all_subsets = ref.utils_shapley.powerset(range(n))
        table = {k: float(rng.rand()) for k in all_subsets}
        matrix = rng.rand(T, len(all_subsets))
        args = types.SimpleNamespace(num_clients=n, num_users=n, epochs=T)
        out.append({"n": n, "T": T, "participants": part, "table": [[list(k), v] for k, v in table.items()],
                    "matrix": matrix.tolist(),
                    "baseline": ufs.compute_shapley_value_baseline(args, table, part).tolist(),
                    "groundtruth": ufs.compute_shapley_value_groundtruth(args, table).tolist(),
                    "mask": ufs.roundly_mask(part, all_subsets).tolist(),
                    "from_matrix": ufs.compute_shapley_value_from_matrix(args, matrix, all_subsets).tolist(),
                    "selection": {str(k): v for k, v in ufs.get_selection_dict(n, part).items()}})
    # ComFedSV (compared_methods.py:17-73): per-round values from the matrix, and one matrix row through a game
    from oracle import toy_games
    cm = ref.compared_methods
    for rec, (n, T, part) in zip(out, ((4, 3, [0, 2, 3]), (5, 4, [1, 2]), (3, 2, [0, 1, 2]))):
        all_subsets = ref.utils_shapley.powerset(range(n))
        args = types.SimpleNamespace(num_clients=n, rounds=T)
        with ref_shim.quiet():
            per_round, _ = cm.comfedsv(args, np.array(rec["matrix"]), all_subsets)
        rec["comfedsv"] = [[v[c] for c in range(n)] for v in per_round]
        game = toy_games.ToyGame(n, seed=n, selection=[c in part for c in range(n)])
        with ref_shim.quiet():
            util, mask = cm.call_comfedsv(game, all_subsets, None)
        rec["call_comfedsv"] = {"utilities": [u.tolist() for u in util], "mask": mask.tolist()}
    with open(os.path.join(GOLD, "fed_bookkeeping.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("fed_bookkeeping done")


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    ref = ref_shim.load()
    which = sys.argv[1:] or ["estimators", "cfg1", "base", "lazy", "rounds", "fedbook"]
    if "fedbook" in which:
        golden_fed_bookkeeping(ref)
    if "rounds" in which:
        golden_round_select(ref)
    if "estimators" in which:
        golden_estimators(ref)
    if "cfg1" in which:
        golden_cfg1(ref)
    if "base" in which:
        golden_base_probe(ref)
    if "lazy" in which:
        golden_lazy(ref)


if __name__ == "__main__":
    main()
