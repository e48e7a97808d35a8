"""LoRA-aware coalition path.

The oracle for this row is a RESTATEMENT (PEFT is not installed in the build container, parity unpinned):
``oracle/restate.py`` averages every state_dict entry of the PEFT-wrapped models with the reference's own
aggregation (A and B separately, get_aggregated_model + model_agg_lazy) and runs the forward with the
UNMERGED low-rank branch, y = x W^T + b + (alpha / r) (x A^T) B^T, as peft's lora.Linear does in eval mode.
The product merges W + (alpha / r) B_S A_S on the GPU and runs the dense batched forward."""
import pytest
import torch

from oracle import restate
from shapley_vit_b200 import layout, lora, synth

R, ALPHA = 4, 8.0


def peft_game(n_clients=3, n_val=96, seed=3, frozen=True, layers=2):
    cfg = layout.vit_preset("tiny", image=32, n_cls=10, layers=layers)
    w0, clients = synth.make_peft_state_dicts(cfg, n_clients, seed=seed, r=R, frozen_base=frozen)
    deltas = [restate.get_difference_between_network_weights(sd, w0) for sd in clients]
    images, labels = synth.make_val_set(cfg, n_val, seed)
    return cfg, w0, deltas, synth.client_sizes(n_clients), images, labels


def oracle_logits(cfg, w0, deltas, n_train, members, images, r=R):
    sd = restate.coalition_state_dict(w0, deltas, n_train, list(members))     # every PEFT key, entry by entry
    hf, lo = restate.split_peft_state_dict(sd)
    return restate.vit_forward(hf, cfg, images, lora=lo, lora_scaling=ALPHA / r), hf, lo


def ratio_rows(coalitions, n_train):
    rows = []
    for S in coalitions:
        r = restate.get_agg_ratio([n_train[j] for j in S])
        row = [0.0] * len(n_train)
        for j, v in zip(S, r):
            row[j] = v
        rows.append(row)
    return rows


def test_split_and_pack_agree_with_the_oracle_parser():
    cfg, w0, deltas, *_ = peft_game()
    hf, lo = lora.split_state_dict(w0)
    hf_o, lo_o = restate.split_peft_state_dict(w0)
    assert list(hf) == list(hf_o) and all(torch.equal(hf[k], hf_o[k]) for k in hf)
    assert set(hf) == {k for k, _ in layout.state_dict_spec(cfg)}             # a plain HF ViT state_dict again
    assert torch.equal(hf["classifier.weight"], w0["module.base_model.model.classifier.modules_to_save.default.weight"])
    assert lora.lora_rank(lo) == R and len(lo) == cfg.layers * 2 * 2
    row = lora.pack_lora(cfg, lo, R, ALPHA / R).view(cfg.layers, 2, 2, cfg.hidden, R)
    for (layer, proj), (a, b) in lo_o.items():
        t = lora.TARGETS.index(proj)
        assert torch.equal(row[layer, t, 0], a.t()) and torch.equal(row[layer, t, 1], b * (ALPHA / R))
    with pytest.raises(ValueError):
        lora.split_state_dict({"vit.encoder.layer.0.attention.attention.key.lora_A.default.weight": torch.zeros(R, 8)})


def test_oracle_unmerged_branch_equals_merged_weights():
    """x W^T + s (x A^T) B^T == x (W + s B A)^T: the identity the GPU path relies on (fp64)."""
    cfg, w0, deltas, n_train, images, _ = peft_game(n_val=8)
    want, hf, lo = oracle_logits(cfg, w0, deltas, n_train, (0, 2), images)
    merged = {k: v.double() for k, v in hf.items()}
    for (layer, proj), (a, b) in lo.items():
        key = f"vit.encoder.layer.{layer}.attention.attention.{proj}.weight"
        merged[key] = merged[key] + (ALPHA / R) * b.double() @ a.double()
    got = restate.vit_forward(merged, cfg, images.double())
    assert (got - want.double()).abs().max() < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("frozen", [True, False])
def test_merged_projection_weights_match_oracle(frozen):
    cfg, w0, deltas, n_train, images, labels = peft_game(frozen=frozen)
    coalitions = [(0,), (1, 2), (0, 1, 2)]
    eng = lora.LoraCoalitionEngine(cfg, w0, deltas, images, labels, lora_alpha=ALPHA, precision="f32", coalition_batch=3,
                                   image_chunk=32, shared_base=False)      # the dense-merge path
    assert eng.base_frozen == frozen and not eng.shared
    got = eng.merged_rows(ratio_rows(coalitions, n_train)).cpu().double()
    for ci, S in enumerate(coalitions):
        _, hf, lo = oracle_logits(cfg, w0, deltas, n_train, S, images[:1])
        for layer in range(cfg.layers):
            for t, proj in enumerate(lora.TARGETS):
                a, b = lo[(layer, proj)]
                want = hf[f"vit.encoder.layer.{layer}.attention.attention.{proj}.weight"].double() + (ALPHA / R) * b.double() @ a.double()
                assert (got[ci, layer * 2 + t] - want).abs().max() < 2e-6


@pytest.mark.gpu
@pytest.mark.parametrize("prec,tol", [("f32", 2e-4), ("f16x3", 2e-4), ("f16c8", 5e-4), ("f16", 1.5e-2)])
@pytest.mark.parametrize("frozen", [True, False])
def test_lora_coalition_logits_and_counts(prec, tol, frozen):
    cfg, w0, deltas, n_train, images, labels = peft_game(frozen=frozen)
    coalitions = [(0,), (1,), (2,), (0, 1), (0, 2), (1, 2), (0, 1, 2)]
    eng = lora.LoraCoalitionEngine(cfg, w0, deltas, images, labels, lora_alpha=ALPHA, precision=prec, coalition_batch=4,
                                   image_chunk=32, keep_logits=True)
    assert eng.shared == frozen                        # frozen base: ONE shared weight region + the K-extension (N1)
    rows = ratio_rows(coalitions, n_train)
    worst = 0.0
    for s0 in range(0, len(rows), 4):                  # two batches: the frozen matrix region must survive a batch
        correct, loss = eng.evaluate(rows[s0:s0 + 4])
        logits = eng.last_logits.cpu()
        for ci, S in enumerate(coalitions[s0:s0 + 4]):
            want, _, _ = oracle_logits(cfg, w0, deltas, n_train, S, images)
            worst = max(worst, (logits[ci] - want).abs().max().item())
            if prec != "f16":
                assert int(correct[ci]) == int((want.argmax(1) == labels).sum())
    print(f"[lora {prec} frozen={frozen}] max |dlogit| = {worst:.3e}")
    assert worst < tol


@pytest.mark.gpu
@pytest.mark.parametrize("prec,tol", [("f32", 1e-5), ("f16x3", 3e-5), ("f16c8", 3e-4), ("f16", 8e-3)])
def test_shared_weight_forward_equals_the_dense_merge(prec, tol):
    """N1: with a frozen base the shared-weight forward (one mat region for every coalition, the factors as a
    K-extension of the QKV GEMM: x W^T + (x A_S^T)(s B_S)^T) and the dense per-coalition merge W + s B_S A_S are the
    same function; ViT-B geometry (pair GEMM kernels), rank 16 as in the reference's start.py:275."""
    cfg = layout.vit_preset("base", image=224, n_cls=10, layers=2)
    w0, clients = synth.make_peft_state_dicts(cfg, 3, seed=5, r=16, frozen_base=True)
    deltas = [restate.get_difference_between_network_weights(sd, w0) for sd in clients]
    n_train = synth.client_sizes(3)
    images, labels = synth.make_val_set(cfg, 6, 5)
    rows = ratio_rows([(0,), (1, 2), (0, 1, 2)], n_train)
    out = {}
    for shared in (None, False):
        eng = lora.LoraCoalitionEngine(cfg, w0, deltas, images, labels, lora_alpha=ALPHA, precision=prec, coalition_batch=3,
                                       image_chunk=6, keep_logits=True, shared_base=shared)
        assert eng.shared == (shared is None)
        out[shared] = (eng.evaluate(rows), eng.last_logits.cpu())
        del eng
    worst = (out[None][1] - out[False][1]).abs().max().item()
    print(f"[lora shared vs dense, {prec}] max |dlogit| = {worst:.3e}")
    assert worst < tol
    if prec in ("f32", "f16x3"):
        assert out[None][0][0] == out[False][0][0]
    want, _, _ = oracle_logits(cfg, w0, deltas, n_train, (1, 2), images, r=16)   # and against the oracle's unmerged branch
    assert (out[None][1][1] - want).abs().max().item() < max(tol, 2e-4)


@pytest.mark.gpu
def test_lora_engine_scores_one_explicit_model_and_keeps_its_frozen_base():
    cfg, w0, deltas, n_train, images, labels = peft_game()
    eng = lora.LoraCoalitionEngine(cfg, w0, deltas, images, labels, lora_alpha=ALPHA, precision="f32", coalition_batch=2,
                                   image_chunk=32)
    rows = ratio_rows([(0, 1), (2,)], n_train)
    before = eng.evaluate(rows)
    client0 = {k: w0[k] + deltas[0][k] for k in w0}
    c, l = eng.evaluate_state_dict(client0)
    hf, lo = restate.split_peft_state_dict(client0)
    want = restate.vit_forward(hf, cfg, images, lora=lo, lora_scaling=ALPHA / R)
    assert c == int((want.argmax(1) == labels).sum())
    assert l == pytest.approx(torch.nn.functional.cross_entropy(want.double(), labels, reduction="sum").item(), rel=1e-5)
    assert eng.evaluate(rows) == before          # the matrix region still holds W_0 where the frozen path expects it


@pytest.mark.gpu
def test_game_detects_peft_state_dicts():
    """The drop-in Game takes the PEFT-keyed init model and client deltas as they are (start.py:285-288)."""
    from shapley_vit_b200 import estimators
    from shapley_vit_b200.fl import ClientBase, ServerBase
    from shapley_vit_b200.game import Game

    cfg, w0, deltas, n_train, images, labels = peft_game()
    loader = torch.utils.data.DataLoader(synth.DictSampleDataset(images, labels), batch_size=32, shuffle=False)
    clients = [ClientBase(i, {}, None, synth.SizedStub(n)) for i, n in enumerate(n_train)]
    server = ServerBase({}, None, clients, None, loader, None)
    n = images.shape[0]
    hf0, _ = restate.split_peft_state_dict(w0)
    acc0, loss0 = restate.evaluation(hf0, cfg, images, labels)                # B_0 = 0: the initial model is the base
    game = Game(clients, server, w0, deltas, [True] * 3, [acc0, loss0], 2,
                {"precision": "f32", "coalition_batch": 4, "image_chunk": 32, "lora_alpha": ALPHA, "heads": cfg.heads})
    assert isinstance(game.engine, lora.LoraCoalitionEngine) and game.engine.r == R
    sv = estimators.shapley_exact(game)
    for S in [(0,), (1, 2), (0, 1, 2)]:
        want, _, _ = oracle_logits(cfg, w0, deltas, n_train, S, images)
        acc = (want.argmax(1) == labels).sum().item() / n
        assert game.eval_utility(S)[0] == pytest.approx(acc - acc0, abs=1e-12)
    assert len(sv) == 2 and set(sv[0]) == {0, 1, 2}
