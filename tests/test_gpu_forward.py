"""End-to-end parity of the CUDA path against the oracle and the reference-generated fixtures:
batched forward logits, per-coalition utilities, top-1 agreement and Shapley vectors.

Tolerances are the north-star's: aggregated weights 1e-6 relative (here: bit-exact in fp32),
>= 99.9 % top-1 agreement per coalition, utility within one sample's accuracy, Shapley within
1e-3 absolute.  The PARITY modes -- f32 (fp32 operands, CUDA cores), f16x3 (fp16 hi/lo planes, three
tensor-core passes) and f16c8 (fp16 pass + e4m3 compensation passes; the library default) -- are
held to ALL of them.  The single-pass THROUGHPUT modes round GEMM operands to 11 (fp16, tf32) or 8
(bf16) significant bits; on these RANDOM-INIT weights the 1st-percentile top-1 margin is ~0.006
(SURVEY.md section 0), which is below what such an operand can resolve: they are held to the
logit / loss / Shapley tolerances only, and their top-1 agreement is REPORTED (printed), not gated."""
# per-precision gates: (min top-1 agreement, max |d correct| in samples, max |d mean loss|); None = reported only
PARITY_MODES = ("f32", "f16x3", "f16c8")
GATES = {"f32": (0.999, 0, 1e-5), "f16x3": (0.999, 1, 2e-5), "f16c8": (0.999, 1, 1e-4),
         "tf32": (None, None, 2e-3), "f16": (None, None, 2e-3), "bf16": (None, None, 1e-2)}
import pytest
import torch

from helpers import load_golden, sv_lists, synthetic_game
from oracle import restate

pytestmark = pytest.mark.gpu

PRECS = ["f32", "f16x3", "f16c8", "tf32", "f16", "bf16"]
# max |dlogit| vs the fp32 oracle that each operand precision is expected to stay under
LOGIT_TOL = {"f32": 2e-4, "f16x3": 2e-4, "f16c8": 5e-4, "tf32": 1.5e-2, "f16": 1.5e-2, "bf16": 8e-2}


def make_engine(cfg, w0, deltas, images, labels, prec, **kw):
    from shapley_vit_b200.engine import CoalitionEngine

    return CoalitionEngine(cfg, w0, deltas, images, labels, precision=prec, keep_logits=True, **kw)


def ratio_rows(coalitions, n_train):
    rows = []
    for S in coalitions:
        r = restate.get_agg_ratio([n_train[j] for j in S])
        row = [0.0] * len(n_train)
        for j, v in zip(S, r):
            row[j] = v
        rows.append(row)
    return rows


@pytest.mark.parametrize("prec", PRECS)
def test_forward_logits_small(prec):
    """2-layer ViT-Ti/16@32, 3 clients: logits of every coalition vs oracle/restate.py."""
    cfg, w0, _, deltas, n_train, images, labels = synthetic_game(n_clients=3, n_val=300, layers=2, seed=4)
    coalitions = [(0,), (1,), (2,), (0, 1), (0, 2), (1, 2), (0, 1, 2)]
    eng = make_engine(cfg, w0, deltas, images, labels, prec, coalition_batch=7, image_chunk=128)
    correct, loss = eng.evaluate(ratio_rows(coalitions, n_train))
    logits = eng.last_logits.cpu()
    worst = 0.0
    for ci, S in enumerate(coalitions):
        sd = restate.coalition_state_dict(w0, deltas, n_train, list(S))
        want = restate.vit_forward(sd, cfg, images)
        worst = max(worst, (logits[ci] - want).abs().max().item())
        if prec in PARITY_MODES:
            assert abs(int(correct[ci]) - int((want.argmax(1) == labels).sum())) <= GATES[prec][1]
        if prec == "f32":
            ce = torch.nn.functional.cross_entropy(want.double(), labels, reduction="sum").item()
            assert loss[ci] == pytest.approx(ce, rel=1e-5)
    print(f"[{prec}] max |dlogit| = {worst:.3e}")
    assert worst < LOGIT_TOL[prec]


@pytest.mark.parametrize("prec", PRECS)
def test_cfg1_against_reference_fixture(prec):
    """BASELINE config 1 end to end through the drop-in Game, against what the REFERENCE produced
    (tests/golden/cfg1_tiny): utilities, top-1 agreement, exact Shapley."""
    from shapley_vit_b200 import estimators
    from shapley_vit_b200.engine import ValidationSet
    from shapley_vit_b200.fl import ClientBase, ServerBase
    from shapley_vit_b200.game import Game
    from shapley_vit_b200.synth import SizedStub
    from shapley_vit_b200._lib import PRECISIONS

    meta, arr = load_golden("cfg1_tiny")
    cfg, w0, _, deltas, n_train, images, labels = synthetic_game(
        meta["vit"], meta["image"], meta["n_cls"], meta["n_clients"], meta["n_val"], meta["seed"])
    val = ValidationSet(cfg, images, labels, PRECISIONS[prec], "cuda:0")
    clients = [ClientBase(i, {}, None, SizedStub(n)) for i, n in enumerate(n_train)]
    server = ServerBase({}, w0, clients, None, val, None)
    game = Game(clients, server, w0, deltas, [True] * 4, [meta["acc0"], meta["loss0"]], 2,
                {"precision": prec, "coalition_batch": 15, "image_chunk": 500, "heads": cfg.heads})
    game.engine.keep_logits = True
    sv = estimators.shapley_exact(game)
    n = meta["n_val"]
    logits = game.engine.last_logits.cpu()
    agree = []
    for ci, S in enumerate(meta["coalitions"]):
        u = game.eval_utility(S)
        pred = logits[ci].argmax(1).numpy()
        agree.append(float((pred == arr["pred"][ci]).mean()))
        if GATES[prec][1] is not None:
            assert abs(u[0] - arr["utility"][ci, 0]) <= GATES[prec][1] / n + 1e-12
        assert abs(u[1] - arr["utility"][ci, 1]) < GATES[prec][2]
    print(f"[{prec}] min top-1 agreement over 15 coalitions = {min(agree):.4f}"
          + ("" if GATES[prec][0] else "  (throughput mode: reported, not gated)"))
    if GATES[prec][0] is not None:
        assert min(agree) >= GATES[prec][0]
    ref = meta["estimators"]["exact"]
    err = max(abs(a - b) for got, want in zip(sv_lists(sv), ref) for a, b in zip(got, want))
    print(f"[{prec}] max |dShapley| = {err:.3e}")
    assert err < 1e-3


@pytest.mark.parametrize("prec", ["f32", "f16x3", "f16c8", "f16"])
def test_vit_base_geometry_against_reference_fixture(prec):
    """ViT-B/16 @ 224 (T = 197, 12 heads): logits for three coalitions vs the reference's."""
    meta, arr = load_golden("base_probe")
    cfg, w0, _, deltas, n_train, images, labels = synthetic_game(
        meta["vit"], meta["image"], meta["n_cls"], meta["n_clients"], meta["n_val"], meta["seed"])
    eng = make_engine(cfg, w0, deltas, images, labels, prec, coalition_batch=3, image_chunk=8)
    eng.evaluate(ratio_rows([tuple(S) for S in meta["coalitions"]], n_train))
    err = (eng.last_logits.cpu().numpy() - arr["logits"]).__abs__().max()
    print(f"[{prec}] ViT-B max |dlogit| = {err:.3e}")
    assert err < LOGIT_TOL[prec]


@pytest.mark.parametrize("prec", ["f32", "f16x3", "f16c8", "f16", "bf16"])
def test_vit_large_geometry_against_oracle(prec):
    """ViT-L/16 @ 224 (h = 1024, 16 heads, ff = 4096; BASELINE config 4 runs it in bf16), 2 layers,
    3 clients: logits of three coalitions vs oracle/restate.py."""
    cfg, w0, _, deltas, n_train, images, labels = synthetic_game("large", 224, 10, 3, 6, 7, layers=2)
    coalitions = [(0,), (1, 2), (0, 1, 2)]
    eng = make_engine(cfg, w0, deltas, images, labels, prec, coalition_batch=3, image_chunk=4)
    eng.evaluate(ratio_rows(coalitions, n_train))
    logits = eng.last_logits.cpu()
    worst = 0.0
    for ci, S in enumerate(coalitions):
        want = restate.vit_forward(restate.coalition_state_dict(w0, deltas, n_train, list(S)), cfg, images)
        worst = max(worst, (logits[ci] - want).abs().max().item())
    print(f"[{prec}] ViT-L max |dlogit| = {worst:.3e}")
    assert worst < LOGIT_TOL[prec]


def test_aggregated_model_rows_bit_exact_vs_oracle():
    """K1 through the engine on the real parameter layout: W_S == oracle W_S, bit for bit."""
    from shapley_vit_b200 import layout

    cfg, w0, _, deltas, n_train, images, labels = synthetic_game(n_clients=4, n_val=8, layers=2, seed=9)
    eng = make_engine(cfg, w0, deltas, images, labels, "f32", coalition_batch=4, image_chunk=8)
    coalitions = [(0,), (1, 3), (0, 1, 2), (0, 1, 2, 3)]
    rows = eng.aggregated_rows(ratio_rows(coalitions, n_train)).cpu()
    lay = layout.plan_layout(cfg)
    for ci, S in enumerate(coalitions):
        sd = restate.coalition_state_dict(w0, deltas, n_train, list(S))
        got = layout.unpack_row(lay, rows[ci])
        for k in sd:
            assert torch.equal(got[k], sd[k]), k


def test_chunking_and_batching_do_not_change_results():
    """Per-coalition results are independent of how coalitions / images are grouped (the property
    that makes 1/2/4/8-GPU sharding bit-identical)."""
    cfg, w0, _, deltas, n_train, images, labels = synthetic_game(n_clients=3, n_val=200, layers=2, seed=2)
    coalitions = [(0,), (1,), (2,), (0, 1), (0, 2), (1, 2), (0, 1, 2)]
    rows = ratio_rows(coalitions, n_train)
    for prec in ("f16", "f16c8"):
        a = make_engine(cfg, w0, deltas, images, labels, prec, coalition_batch=7, image_chunk=200).evaluate(rows)
        b = make_engine(cfg, w0, deltas, images, labels, prec, coalition_batch=2, image_chunk=64).evaluate(rows)
        assert a == b


def test_module_forward_and_evaluation_api():
    """net(img).logits and evaluation(args, net, loader) -- the reference's call shapes."""
    from torch.utils.data import DataLoader

    from shapley_vit_b200 import synth
    from shapley_vit_b200.fl import evaluation
    from shapley_vit_b200.models.vit import ViTForImageClassification

    cfg, w0, _, _, _, images, labels = synthetic_game(n_clients=1, n_val=100, layers=2, seed=6)
    net = ViTForImageClassification(cfg, precision="f32")
    net.load_state_dict(w0)
    want = restate.vit_forward(w0, cfg, images[:16])
    got = net(images[:16].cuda()).logits.cpu()
    assert (got - want).abs().max() < 2e-4
    loader = DataLoader(synth.DictSampleDataset(images, labels), batch_size=128, shuffle=False)
    acc, loss = evaluation({"precision": "f32"}, net, loader)
    racc, rloss = restate.evaluation(w0, cfg, images, labels)
    assert acc == racc and loss == pytest.approx(rloss, rel=1e-5)


def test_validation_split_axis_sums_to_the_whole():
    """engine.evaluate(rows, image_range=...) -- the axis dist.sharded_evaluate takes when there are fewer pending
    coalitions than ranks: the per-slice (correct, loss_sum) pairs add up to the whole-set pair (counts exactly)."""
    cfg, w0, _, deltas, n_train, images, labels = synthetic_game(n_clients=3, n_val=200, layers=2, seed=6)
    eng = make_engine(cfg, w0, deltas, images, labels, "f32", coalition_batch=2, image_chunk=64)
    rows = ratio_rows([(0, 1), (2,), (0, 1, 2)], n_train)
    whole_c, whole_l = eng.evaluate(rows)
    parts = [eng.evaluate(rows, image_range=r) for r in ((0, 67), (67, 134), (134, 200), (200, 200))]
    for ci in range(3):
        assert sum(p[0][ci] for p in parts) == whole_c[ci]
        assert sum(p[1][ci] for p in parts) == pytest.approx(whole_l[ci], rel=1e-12)
    assert parts[3] == ([0, 0, 0], [0.0, 0.0, 0.0])
