"""The reference's outer seam: ``python mainShapley.py [flags]`` -> start() ->
getInitialShapleyValue -> Game -> the estimator, on synthetic inputs, checked against the CPU oracle driven
through the same estimators with the same seeds."""
import ast
import os
import re
import subprocess
import sys

import pytest
import torch

from helpers import sv_lists
from oracle import restate
from oracle_evaluator import OracleEvaluator
from shapley_vit_b200 import estimators, layout, lora, synth
from shapley_vit_b200.fl import ClientBase, ServerBase
from shapley_vit_b200.game import Game

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLAGS = ["--synthetic", "--vit_size", "tiny", "--image_size", "32", "--num_classes", "10", "--val_size", "96",
         "--num_clients", "3", "--dtype", "f32", "--seed", "3", "--approximation_method", "exact"]


def run_main(extra, tmp_path):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "mainShapley.py"), *FLAGS, *extra, "--exp_dir", str(tmp_path)],
                         capture_output=True, text=True, cwd=str(tmp_path), timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    m = re.search(r"^Exact: (.*)$", out.stdout, re.M)
    assert m, out.stdout[-2000:]
    return ast.literal_eval(m.group(1)), out.stdout


def oracle_exact(w0_hf, delta_hfs, cfg, images, labels, n_train):
    prev = list(restate.evaluation(w0_hf, cfg, images, labels))
    clients = [ClientBase(i, {}, None, synth.SizedStub(n)) for i, n in enumerate(n_train)]
    game = Game(clients, ServerBase({}, w0_hf, clients, None, None, None), w0_hf, delta_hfs, [True] * len(n_train), prev, 2, {})
    game._evaluator = OracleEvaluator(cfg, w0_hf, delta_hfs, images, labels)
    return sv_lists(estimators.shapley_exact(game))


@pytest.mark.gpu
def test_main_shapley_plain_vit(tmp_path):
    sv, stdout = run_main([], tmp_path)
    cfg = layout.vit_preset("tiny", image=32, n_cls=10)
    w0 = synth.make_state_dict(cfg, 3)
    deltas = [restate.get_difference_between_network_weights(synth.make_client_state_dict(w0, j, 3), w0) for j in range(3)]
    images, labels = synth.make_val_set(cfg, 96, 3)
    want = oracle_exact(w0, deltas, cfg, images, labels, synth.client_sizes(3))
    assert "Previous utility" in stdout and "Shapley value sum for each utility" in stdout
    for d in range(2):
        assert [sv[d][c] for c in range(3)] == pytest.approx(want[d], abs=1e-5 if d else 1e-12)


@pytest.mark.gpu
def test_main_shapley_lora_wrapped_vit(tmp_path):
    """--lora_rank: the PEFT-keyed model container, separate A / B aggregation and the GPU merge, end to end; the
    oracle side evaluates every coalition's MERGED dense model (identity checked in test_lora.py)."""
    sv, _ = run_main(["--lora_rank", "4", "--lora_alpha", "8"], tmp_path)
    cfg = layout.vit_preset("tiny", image=32, n_cls=10)
    w0, clients = synth.make_peft_state_dicts(cfg, 3, 3, r=4, prefix="base_model.model.")
    images, labels = synth.make_val_set(cfg, 96, 3)
    n_train = synth.client_sizes(3)

    class MergedOracle(OracleEvaluator):      # FedAvg over the PEFT entries, then W + (alpha / r) B A, then the forward
        def evaluate(self, rows, image_range=None):
            correct, loss = [], []
            for row in rows:
                members = [j for j, r in enumerate(row) if r != 0]
                agg = restate.get_aggregated_model([self.deltas[j] for j in members], [row[j] for j in members])
                sd = lora.merged_state_dict(restate.model_agg_lazy(self.w0, [agg] if agg is not None else []), 8.0)
                _, _, det = restate.evaluation(sd, self.cfg, self.images, self.labels, return_details=True)
                correct.append(int(det["correct"]))
                loss.append(float(det["loss_sum"]))
            return correct, loss

    deltas = [restate.get_difference_between_network_weights(sd, w0) for sd in clients]
    prev = list(restate.evaluation(lora.merged_state_dict(w0, 8.0), cfg, images, labels))
    cl = [ClientBase(i, {}, None, synth.SizedStub(n)) for i, n in enumerate(n_train)]
    game = Game(cl, ServerBase({}, None, cl, None, None, None), None, deltas, [True] * 3, prev, 2, {})
    game._evaluator = MergedOracle(cfg, w0, deltas, images, labels)
    want = sv_lists(estimators.shapley_exact(game))
    for d in range(2):
        assert [sv[d][c] for c in range(3)] == pytest.approx(want[d], abs=1e-5 if d else 1e-12)


def test_lora_container_takes_peft_checkpoints():
    """CPU: the parameter container has PEFT's keys, loads a DataParallel-prefixed checkpoint, and its merged
    state_dict is the plain HF layout."""
    from shapley_vit_b200.models.vit import LoraViTForImageClassification, ViTForImageClassification

    cfg = layout.vit_preset("tiny", image=32, n_cls=10, layers=2)
    m = LoraViTForImageClassification(cfg, r=4, lora_alpha=8.0)
    w0, clients = synth.make_peft_state_dicts(cfg, 1, 5, r=4)                     # 'module.base_model.model.' keys
    assert {f"module.{k}" for k in m.state_dict()} == set(w0)
    m.load_state_dict(clients[0])
    assert all(torch.equal(m.state_dict()[k[len("module."):]], v) for k, v in clients[0].items())
    merged = lora.merged_state_dict(m.state_dict(), 8.0)
    assert [k for k, _ in layout.state_dict_spec(cfg)] == sorted(merged, key=[k for k, _ in layout.state_dict_spec(cfg)].index)
    q = "vit.encoder.layer.1.attention.attention.query"
    a, b = clients[0][f"module.base_model.model.{q}.lora_A.default.weight"], clients[0][f"module.base_model.model.{q}.lora_B.default.weight"]
    assert torch.allclose(merged[q + ".weight"], clients[0][f"module.base_model.model.{q}.base_layer.weight"] + 2.0 * b @ a)
    plain = ViTForImageClassification(cfg)
    plain.load_state_dict({f"module.{k}": v for k, v in merged.items()})        # DataParallel prefix on the plain ViT too
    # PEFT's initial state: B = 0, so the wrapped model IS the base model
    fresh = LoraViTForImageClassification(cfg, r=4)
    fm = lora.merged_state_dict(fresh.state_dict(), 8.0)
    assert torch.equal(fm[q + ".weight"], fresh.state_dict()[f"base_model.model.{q}.base_layer.weight"])


@pytest.mark.gpu
def test_main_shapley_reads_client_checkpoints_and_waits_for_a_late_one(tmp_path):
    """N3 (reference start.py:134-151, 198-222): without --synthetic the client models are read from
    ``<cwd>/shapleyserver/local_training/client_<i>_model/ViT_epoch_9.pth.tar`` -- ``{'state_dict': ...}`` files whose
    keys carry DataParallel's ``module.`` prefix -- and a checkpoint that is not there yet (its writer has not released
    it) is waited for.  Only the validation data is synthetic."""
    import threading

    cfg = layout.vit_preset("tiny", image=32, n_cls=10)
    w0 = synth.make_state_dict(cfg, 3)
    client_sds = [synth.make_client_state_dict(w0, j, 3) for j in range(3)]
    init_path = tmp_path / "init_global.pth.tar"
    torch.save({"state_dict": w0, "epoch": 0}, init_path)
    finals = []
    for j, sd in enumerate(client_sds):
        d = tmp_path / "shapleyserver" / "local_training" / f"client_{j + 1}_model"
        d.mkdir(parents=True)
        final = d / "ViT_epoch_9.pth.tar"
        target = final if j < 2 else d / "ViT_epoch_9.pth.tar.writing"      # the third one is still being written
        torch.save({"state_dict": {f"module.{k}": v for k, v in sd.items()}, "epoch": 9}, target)
        finals.append((target, final))
    flags = ["--synthetic_data", "--vit_size", "tiny", "--image_size", "32", "--num_classes", "10", "--val_size", "96",
             "--num_clients", "3", "--dtype", "f32", "--seed", "3", "--approximation_method", "exact",
             "-loadModel", str(init_path), "--exp_dir", str(tmp_path)]
    # the writer releases the third checkpoint only once the reader has been seen waiting for it (event-driven, so the
    # test does not depend on how long the interpreter takes to start on a cold box)
    proc = subprocess.Popen([sys.executable, "-u", os.path.join(ROOT, "mainShapley.py"), *flags], stdout=subprocess.PIPE,
                            stderr=subprocess.PIPE, text=True, cwd=str(tmp_path))
    lines, released = [], []

    def pump():
        for ln in proc.stdout:
            lines.append(ln)
            if "Waiting for the file to be unlocked..." in ln and not released:
                released.append(True)
                os.replace(*finals[2])

    err = []
    t_out, t_err = threading.Thread(target=pump), threading.Thread(target=lambda: err.append(proc.stderr.read()))
    t_out.start(), t_err.start()
    try:
        rc = proc.wait(timeout=600)
    finally:
        if proc.poll() is None:
            proc.kill()
        t_out.join(), t_err.join()

    import types

    out = types.SimpleNamespace(returncode=rc, stdout="".join(lines), stderr="".join(err))
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.count("Waiting for the file to be unlocked...") >= 1
    assert out.stdout.count("Model loaded!") == 3
    sv = ast.literal_eval(re.search(r"^Exact: (.*)$", out.stdout, re.M).group(1))
    deltas = [restate.get_difference_between_network_weights(sd, w0) for sd in client_sds]
    images, labels = synth.make_val_set(cfg, 96, 3)
    # the reference hands every client the validation dataset as its "training set" (start.py:163-170): equal sizes
    want = oracle_exact(w0, deltas, cfg, images, labels, [96, 96, 96])
    for d in range(2):
        assert [sv[d][c] for c in range(3)] == pytest.approx(want[d], abs=1e-5 if d else 1e-12)


def test_check_local_training_model_exist_waits(tmp_path, capsys):
    """CPU: the wait-for-unlock loop (reference start.py:198-222) returns once the file has appeared."""
    import threading

    from shapleyserver.start import checkLocalTrainingModelExist

    path = tmp_path / "ViT_epoch_9.pth.tar"
    threading.Timer(0.3, lambda: path.write_bytes(b"x")).start()
    assert checkLocalTrainingModelExist(str(path), poll_seconds=0.05) is True
    assert "Waiting for the file to be unlocked..." in capsys.readouterr().out
