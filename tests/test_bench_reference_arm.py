"""The reference arm of bench.py (`--impl reference`, the `cpu_baseline` leg): the UNMODIFIED reference classes driven
through oracle/ref_shim.py -- from /root/reference in the build container, from the staged copy oracle/_ref on the GPU
box -- on a tiny configuration; and the staging recipe itself."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim, stage_ref  # noqa: E402

needs_reference = pytest.mark.skipif(not ref_shim.available(), reason="neither /root/reference nor oracle/_ref is present")


@needs_reference
def test_reference_arm_line_on_a_tiny_configuration():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--vit", "tiny", "--image", "32",
                          "--clients", "3", "--val", "200", "--cpu-sample-images", "64", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["gpu_launches"] == 0 and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["value"] > 0 and line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0,
                                                  "d2h_bytes_per_step": 0}
    assert "Game.eval_utility" in line["cpu_baseline"]["sample"] and "64 of 200" in line["cpu_baseline"]["sample"]


@needs_reference
def test_reference_arm_evaluates_what_the_oracle_restates():
    """One step of the arm = the reference's Game.eval_utility; its utility equals the restatement's on the same inputs."""
    import torch

    from oracle import restate
    from oracle.ref_arm import ReferenceArm
    from shapley_vit_b200 import layout, synth

    if torch.cuda.is_available():
        pytest.skip("the arm asserts that the process sees no GPU (bench.py hides them)")
    cfg = layout.vit_preset("tiny", image=32, n_cls=10, layers=2)
    arm = ReferenceArm(cfg, n_clients=3, n_val=48, n_sample=48, seed=4, threads=2)
    arm.step(0)
    S = tuple(arm.last["coalition"])
    got = [arm.game.utility[d][frozenset(S)] for d in range(2)]
    w0 = synth.make_state_dict(cfg, 4)
    deltas = [restate.get_difference_between_network_weights(synth.make_client_state_dict(w0, j, 4), w0) for j in range(3)]
    images, labels = synth.make_val_set(cfg, 48, 4)
    sd = restate.coalition_state_dict(w0, deltas, synth.client_sizes(3), restate.reference_member_order(S))
    acc, loss = restate.evaluation(sd, cfg, images, labels)
    assert got[0] == pytest.approx(acc, abs=1e-12) and got[1] == pytest.approx(loss, rel=1e-5)


@pytest.mark.skipif(not os.path.isdir("/root/reference/shapleyserver/fed_client_contribution"), reason="build container only")
def test_staged_copy_is_byte_identical(tmp_path):
    import hashlib

    manifest = stage_ref.stage("/root/reference", str(tmp_path / "_ref"))
    assert any(k.endswith("fed_client_contribution/game.py") for k in manifest)
    assert any(k.endswith("federated_learning/utils.py") for k in manifest)
    for rel, digest in manifest.items():
        for base in ("/root/reference", str(tmp_path / "_ref")):
            with open(os.path.join(base, rel), "rb") as f:
                assert hashlib.sha256(f.read()).hexdigest() == digest
