#!/usr/bin/env python
"""Benchmark of the client-contribution utility loop (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port + HF ViT)

Workload (BASELINE config 2): 8-client FedAvg, ViT-B/16 at 224 px, 10 classes, 10 000-image
synthetic validation set, exact-Shapley coalition enumeration (255 non-empty coalitions).
A STEP is one batch of `--coalition-batch` coalitions per GPU taken from that enumeration:
aggregate their weights (K1), run the batched ViT forward over the whole validation set (K2-K4)
and score (K5).  metric = coalition utility evaluations per second, whole job (all ranks).

Multi-GPU: one process per GPU (torchrun); rank 0 builds the stacked client deltas and W0 and
NCCL-broadcasts them once; every rank evaluates its own coalition slice (weak scaling: per-GPU
work per step is fixed); the timed region is bracketed by barrier + synchronize and the elapsed
time is the max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "coalition utility evals/sec (ViT-B/16, 8 clients)"
UNIT = "coalition-evals/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--vit", default="base")
    ap.add_argument("--image", type=int, default=224)
    ap.add_argument("--classes", type=int, default=10)
    ap.add_argument("--clients", type=int, default=8)
    ap.add_argument("--val", type=int, default=10000)
    ap.add_argument("--coalition-batch", type=int, default=8)
    ap.add_argument("--image-chunk", type=int, default=128)
    ap.add_argument("--precision", default="f16c8", choices=["f16c8", "f16x3", "f32", "f16", "bf16", "tf32"],
                    help="f16c8 (default), f16x3, f32: parity modes (>= 99.9 %% top-1 agreement with fp32); "
                         "f16, bf16, tf32: single-pass throughput modes outside that gate")
    ap.add_argument("--lora-rank", type=int, default=0,
                    help="> 0: the clients are PEFT-LoRA models (rank r on query / value, classifier saved) over a FROZEN base "
                         "(reference start.py:274-283): the shared-weight forward of SURVEY section 8(f) N1")
    ap.add_argument("--lora-path", default="shared", choices=["shared", "dense"],
                    help="with --lora-rank: shared-weight forward + K-extension (N1), or the dense per-coalition merge")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-images", type=int, default=256)
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-throughput-mode", action="store_true")
    ap.add_argument("--time-to-shapley", action="store_true",
                    help="also time the WHOLE exact-Shapley job (all 2^N - 1 coalitions, strong scaling over the ranks); "
                         "on by default when --gpus >= 4 (at 1-2 GPUs it adds minutes)")
    ap.add_argument("--no-time-to-shapley", action="store_true")
    ap.add_argument("--parity-clients", type=int, default=4)
    ap.add_argument("--parity-images", type=int, default=10000)
    return ap.parse_args()


def workload_name(a):
    cfg2 = (a.clients, a.vit, a.image, a.val, a.classes) == (8, "base", 224, 10000, 10)
    cfg4 = (a.clients, a.vit, a.image, a.coalition_batch) == (10, "large", 224, 32)
    tag = "BASELINE config 2" if cfg2 else "BASELINE config 4 geometry" if cfg4 else "not a BASELINE configuration"
    if a.lora_rank:
        tag += (f"; clients are PEFT-LoRA models, rank {a.lora_rank} on query / value over a frozen base, "
                + ("dense per-coalition merge" if a.lora_path == "dense" else "shared-weight forward + K-extension (N1)"))
    return (f"{a.clients}-client FedAvg ViT-{a.vit}/16 @{a.image}px, exact-Shapley enumeration "
            f"({2 ** a.clients - 1} coalitions), {a.val}-image synthetic val set, {a.classes} classes ({tag})")


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p, "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi, 200 ms) -- started before and stopped after the timed region
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.proc, self.lines = None, []
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
                power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "power_w_max": max(power) if power else None, "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's OWN Game.eval_utility + evaluation (oracle/ref_arm.py drives the
# unmodified reference classes, staged by build() into oracle/_ref for the GPU box) around the reference's actual
# third-party model (HF ViTForImageClassification, fp32 eager).  Falls back to the restatement (oracle/restate.py,
# kind "port") only when neither /root/reference nor oracle/_ref exists.
# ----------------------------------------------------------------------------------------------
class PortPath:
    """The restated path (oracle/restate.py) -- used only when the reference itself cannot be imported."""

    kind = "port"

    def __init__(self, a):
        import torch

        from oracle import restate
        from oracle.hf_model import build_hf_vit
        from shapley_vit_b200 import layout, synth

        self.torch, self.restate = torch, restate
        torch.set_num_threads(os.cpu_count() or 1)
        self.cores = torch.get_num_threads()
        self.a = a
        self.cfg = layout.vit_preset(a.vit, image=a.image, n_cls=a.classes)
        self.w0 = synth.make_state_dict(self.cfg, a.seed)
        clients = [synth.make_client_state_dict(self.w0, j, a.seed) for j in range(a.clients)]
        self.deltas = [restate.get_difference_between_network_weights(sd, self.w0) for sd in clients]
        del clients
        self.n_train = synth.client_sizes(a.clients)
        self.n_img = min(a.cpu_sample_images, a.val)
        self.images, self.labels = synth.make_val_set(self.cfg, self.n_img, a.seed)
        self.model = build_hf_vit(self.cfg, self.w0)
        self.coalitions = list(restate.powerset(range(a.clients)))
        self.sample = (f"restated path (oracle/restate.py), 1 coalition per step: full-size FedAvg aggregation + load_state_dict, "
                       f"forward/score on {self.n_img} of {a.val} images (HF ViT fp32 eager, batch 128), scaled linearly in images")

    def step(self, i: int, warm: bool = False) -> float:
        """Seconds one full coalition evaluation would take (measured on the sample, extrapolated)."""
        torch, restate = self.torch, self.restate
        n_img = min(16, self.n_img) if warm else self.n_img
        S = self.coalitions[(37 * i + len(self.coalitions) // 2) % len(self.coalitions)]
        t0 = time.perf_counter()
        members = restate.reference_member_order(S)
        sd = restate.coalition_state_dict(self.w0, self.deltas, self.n_train, members)
        self.model.load_state_dict(sd)
        t_agg = time.perf_counter() - t0
        t0 = time.perf_counter()
        correct, loss = 0, 0.0
        for s in range(0, n_img, 128):                  # evaluation(), utils.py:864-926 (autograd on)
            out = self.model(self.images[s:min(s + 128, n_img)]).logits
            pred = out.argmax(dim=1)
            correct += pred.eq(self.labels[s:min(s + 128, n_img)]).sum().item()
            loss += torch.nn.functional.cross_entropy(out, self.labels[s:min(s + 128, n_img)], reduction="sum").item()
        t_fwd = time.perf_counter() - t0
        return t_agg + t_fwd * (self.a.val / n_img)


def make_cpu_path(a):
    from oracle import ref_shim

    if ref_shim.available():
        from oracle.ref_arm import ReferenceArm
        from shapley_vit_b200 import layout

        arm = ReferenceArm(layout.vit_preset(a.vit, image=a.image, n_cls=a.classes), a.clients, a.val, a.cpu_sample_images, a.seed)
        arm.kind = "reference"
        return arm
    return PortPath(a)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cpu = make_cpu_path(a)
    for i in range(a.warmup):
        cpu.step(i, warm=True)
    t0 = time.perf_counter()
    per = [cpu.step(a.warmup + i) for i in range(a.steps)]
    wall = time.perf_counter() - t0
    sec_per_eval = sum(per) / len(per)
    value = 1.0 / sec_per_eval
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * wall / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "step": "one coalition evaluation on a bounded sample, scaled linearly in images",
                   "device": "host CPU", "reference_root": getattr(cpu, "ref_root", None)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.cores, "kind": cpu.kind, "sample": cpu.sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_subprocess(a):
    """The same arm in a child process that sees no GPU (the reference would otherwise move its models to
    cuda:1 / cuda:0, server2.py:17, utils.py:865): one warm-up step and one timed step."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "1",
           "--vit", a.vit, "--image", str(a.image), "--classes", str(a.classes), "--clients", str(a.clients),
           "--val", str(a.val), "--seed", str(a.seed), "--cpu-sample-images", str(a.cpu_sample_images)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    for ln in reversed(r.stdout.splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)["cpu_baseline"]
    return {"error": (r.stderr or r.stdout)[-400:]}


# ----------------------------------------------------------------------------------------------
def parity_leg(a, cfg, lay, deltas, w0, images_host, labels_host, dev, val_main=None):
    """The second half of the metric ("Shapley abs err") and north_star's gates (>= 99.9 % top-1 agreement per
    coalition, utilities within one sample, Shapley within 1e-3), on the bench's own device buffers.

    (1) The exact-Shapley game of the first `--parity-clients` clients (every coalition) over `--parity-images`
        validation images (default: all 10 000, i.e. >= 10 000 predictions per coalition) in the benchmarked
        precision and in an fp32-grade mode (`f16x3`: fp16 hi+lo operands, three tcgen05 passes, logits within
        ~1e-5 of fp32; `f16c8` when f16x3 itself is benchmarked): top-1 agreement, utility gap, Shapley gap.
    (2) anchor_f32: the same coalitions on the first 256 images in the library's fp32 mode (fp32 operands,
        CUDA-core FMA GEMMs) against every mode of (1).
    (3) anchor_oracle: two coalitions x 64 images through the CPU restatement of the reference path
        (oracle/restate.py: the reference's aggregation order + the HF ViT forward, pinned on reference outputs by
        tests/golden) against every mode: max |dlogit| and top-1 agreement.  The oracle is the checker here."""
    import torch

    from shapley_vit_b200 import layout, synth
    from shapley_vit_b200.engine import CoalitionEngine
    from shapley_vit_b200.estimators import powerset, shapley_exact
    from shapley_vit_b200.fl import ClientBase, ServerBase
    from shapley_vit_b200.game import Game

    n_c, n_img = min(a.parity_clients, a.clients), min(a.parity_images, a.val)
    n_head = min(256, n_img)
    n_train = synth.client_sizes(a.clients)[:n_c]
    coalitions = list(powerset(range(n_c)))
    d_sub = deltas[:n_c].contiguous()
    t0 = time.perf_counter()
    throughput_mode = a.precision in ("f16", "bf16", "tf32")
    ref_mode = "f16c8" if a.precision == "f16x3" else "f16x3"
    modes = [a.precision] + (["f16c8"] if throughput_mode else []) + ([ref_mode] if a.precision != "f32" else [])
    sv, preds, util, head = {}, {}, {}, {}

    def run(prec, images, labels, cb, chunk, val=None):
        eng = CoalitionEngine(cfg, w0, d_sub, val if val is not None else images, labels, precision=prec, coalition_batch=cb,
                              image_chunk=chunk, device=dev, keep_logits=True)
        clients = [ClientBase(i, {}, None, synth.SizedStub(n)) for i, n in enumerate(n_train)]
        server = ServerBase({}, None, clients, None, eng.val, None)
        rows = []
        for S in coalitions:
            r = server.get_agg_ratio(selected_clients=[clients[j] for j in S])
            row = [0.0] * n_c
            for j, v in zip(S, r):
                row[j] = v
            rows.append(row)
        game = Game(clients, server, None, [None] * n_c, [True] * n_c, [0.0, 0.0], 2, {"precision": prec})
        p, lg, counts = [], [], ([], [])
        for s0 in range(0, len(rows), cb):
            c, l = eng.evaluate(rows[s0:s0 + cb])
            counts[0].extend(c), counts[1].extend(l)
            p.append(eng.last_logits.argmax(dim=2).cpu())
            lg.append(eng.last_logits[:, :n_head].cpu())
        n = eng.n_val
        for S, c, l in zip(coalitions, *counts):       # fill the memo from the evaluation above: no second pass
            game.utility[0][frozenset(S)] = c / n
            game.utility[1][frozenset(S)] = l / n
        phi = shapley_exact(game)
        del eng
        torch.cuda.empty_cache()
        return ([[phi[d][c] for c in range(n_c)] for d in range(2)], torch.cat(p), [game.eval_utility(S) for S in coalitions],
                torch.cat(lg))

    for prec in modes:
        reuse = val_main if (val_main is not None and prec == a.precision and n_img == a.val) else None
        sv[prec], preds[prec], util[prec], head[prec] = run(prec, images_host[:n_img], labels_host[:n_img], 8, 128, reuse)

    def versus(m, r, sl=slice(None)):
        agree = (preds[m][:, sl] == preds[r][:, sl]).float().mean(dim=1)
        return {"shapley_abs_err": max(abs(x - y) for d in range(2) for x, y in zip(sv[m][d], sv[r][d])),
                "top1_agreement_min": float(agree.min()), "top1_agreement_mean": float(agree.mean()),
                "accuracy_utility_abs_err_max": max(abs(u[0] - v[0]) for u, v in zip(util[m], util[r]))}

    out = {}
    if a.precision != "f32":
        out.update(versus(a.precision, ref_mode))
        out.update({"predictions_per_coalition": n_img, "one_sample": 1.0 / n_img,
                    "reference": f"this library's {ref_mode} mode on the same {n_img} images (fp32-grade split-precision tensor-core mode), "
                                 "itself anchored to the fp32 CUDA-core mode (anchor_f32) and to the CPU oracle (anchor_oracle)",
                    "game": f"{n_c} clients, {len(coalitions)} coalitions, {n_img} images, {a.vit} @ {a.image}px, exact Shapley"})
        if throughput_mode:
            out["gate_mode"] = dict(versus("f16c8", ref_mode), precision="f16c8",
                                    note="the default precision (fp16 + e4m3-compensated tensor-core mode) on the same game")

    # ---- (2) the fp32 CUDA-core mode on the first n_head images ------------------------------------
    sv["f32"], preds["f32"], util["f32"], head["f32"] = run("f32", images_host[:n_head], labels_host[:n_head], 5, min(32, n_head))
    anchor = {"images": n_head, "coalitions": len(coalitions)}
    for m in modes:
        agree = (preds[m][:, :n_head] == preds["f32"]).float().mean(dim=1)
        anchor[m] = {"top1_agreement_min": float(agree.min()), "max_abs_dlogit": float((head[m] - head["f32"]).abs().max())}
    out["anchor_f32"] = anchor

    # ---- (3) the CPU oracle on two coalitions x 64 images ------------------------------------------
    try:
        from oracle import restate

        n_or = min(64, n_head)
        w0_sd = layout.unpack_row(lay, w0.cpu())
        d_sds = [layout.unpack_row(lay, d_sub[j].cpu()) for j in range(n_c)]
        picks = [0, len(coalitions) - 1] if len(coalitions) > 1 else [0]
        orc = {"images": n_or, "coalitions": [list(coalitions[i]) for i in picks],
               "oracle": "oracle/restate.py (reference aggregation order + HF ViT forward, fp32, CPU)"}
        want = []
        for i in picks:
            sd = restate.coalition_state_dict(w0_sd, d_sds, n_train, restate.reference_member_order(coalitions[i]))
            want.append(restate.vit_forward(sd, cfg, images_host[:n_or].float()))
        want = torch.stack(want)
        for m in modes + ["f32"]:
            got = head[m][picks][:, :n_or]
            orc[m] = {"max_abs_dlogit": float((got - want).abs().max()),
                      "top1_agreement": float((got.argmax(2) == want.argmax(2)).float().mean())}
        out["anchor_oracle"] = orc
    except Exception as e:
        out["anchor_oracle"] = {"error": repr(e)}
    out["seconds"] = time.perf_counter() - t0
    return out


# ----------------------------------------------------------------------------------------------
def time_to_shapley(a, eng, deltas, w0, clients, server, coalitions, dev, rank, ws, per_gpu_rate):
    """Strong scaling: wall-clock seconds from "rank 0 holds the client weights" to "every rank holds the exact
    Shapley vector" for the WHOLE game of the workload (BASELINE config 2: all 255 coalitions over the 10 000 images):
    NCCL broadcast of the stacked deltas + W0, the coalition slices, ONE all-gather of the packed per-coalition records
    (written by K5 straight into the send buffer), the fp64 accumulation on every rank.  Max over ranks.
    Then rank 0 recomputes a sample of coalitions that OTHER ranks owned and checks the gathered records bit for bit."""
    import torch

    from shapley_vit_b200 import dist
    from shapley_vit_b200.estimators import shapley_exact
    from shapley_vit_b200.game import Game

    N = a.clients
    torch.cuda.synchronize(dev)
    dist.barrier()
    t0 = time.perf_counter()
    dist.broadcast_(deltas)                       # eng.deltas / eng.w0 are these tensors: received in place
    dist.broadcast_(w0)
    torch.cuda.synchronize(dev)
    t_bcast = time.perf_counter() - t0
    game = Game(clients, server, None, [None] * N, [True] * N, [0.0, 0.0], 2, {"precision": a.precision})
    game._engine = eng
    phi = shapley_exact(game)                     # plan -> eval_utilities (sharded, one all-gather) -> accumulation
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - t0
    if ws > 1:
        import torch.distributed as td

        t = torch.tensor([wall, t_bcast], dtype=torch.float64, device=dev)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        wall, t_bcast = float(t[0]), float(t[1])
    out = None
    if rank == 0:
        n = len(coalitions)
        ideal = n / ws / per_gpu_rate
        out = {"value": wall, "unit": "s", "coalitions": n, "n_gpus": ws, "scaling": "strong",
               "broadcast_s": t_bcast, "broadcast_bytes": int(deltas.numel() * 4 + w0.numel() * 4),
               "ideal_s": ideal, "vs_ideal": wall / ideal,
               "ideal": "coalitions / n_gpus / (this line's per-GPU rate with everything resident)",
               "shapley_acc": [phi[0][c] for c in range(N)],
               "collectives": "2 broadcasts (client weights, once) + 1 all-gather of 16-byte records"}
        if ws > 1:                                 # cross-rank identity on hardware: recompute what others computed
            keys = list(coalitions)
            picks = []
            for r in range(1, ws):
                lo, hi = dist.shard_bounds(n, r, ws)
                if hi > lo:
                    picks.append(lo + (hi - lo) // 2)
            picks = picks[:4]
            rows = [game._ratio_row(frozenset(keys[i])) for i in picks]
            c, l = eng.evaluate(rows)
            same = all((int(ci), float(li)) == game.counts[frozenset(keys[i])] for i, ci, li in zip(picks, c, l))
            out["cross_rank_identity"] = {"coalitions_recomputed_on_rank0": [list(keys[i]) for i in picks],
                                          "bit_identical": bool(same)}
    dist.barrier()
    return out


# ----------------------------------------------------------------------------------------------
def run_ours(a):
    import torch

    from shapley_vit_b200 import _lib, dist, layout, synth
    from shapley_vit_b200.engine import CoalitionEngine, ValidationSet
    from shapley_vit_b200.fl import ClientBase, ServerBase
    from shapley_vit_b200.game import Game

    rank = int(os.environ.get("RANK", "0"))
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a B200 (there is no CPU path)")
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if ws > 1:
        import torch.distributed as td

        # keep stdout to the one JSON line: NCCL_DEBUG=VERSION printf()s "NCCL version ..." to stdout
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

        td.init_process_group("nccl", rank=rank, world_size=ws, device_id=dev)
    _lib.device_info()

    cfg = layout.vit_preset(a.vit, image=a.image, n_cls=a.classes)
    lay = layout.plan_layout(cfg)
    N, Cb = a.clients, a.coalition_batch
    prec = _lib.PRECISIONS[a.precision]

    # ---- client weights: built on rank 0, broadcast once over NVLink (NCCL) -------------------
    deltas = torch.empty((N if not a.lora_rank else 0, lay.total), dtype=torch.float32, device=dev)
    w0 = torch.empty(lay.total, dtype=torch.float32, device=dev)
    if rank == 0 and not a.lora_rank:
        w0_sd = synth.make_state_dict(cfg, a.seed)
        w0.copy_(layout.pack_state_dict(lay, w0_sd))
        row = torch.empty(lay.total, dtype=torch.float32).pin_memory()
        for j in range(N):
            cj = synth.make_client_state_dict(w0_sd, j, a.seed)
            layout.pack_state_dict(lay, {k: cj[k] - w0_sd[k] for k in cj}, out=row)   # F1: delta_j = W_j - W_0
            deltas[j].copy_(row)
        del w0_sd
    if not a.lora_rank:                      # (LoRA line: every rank builds the same seeded PEFT state_dicts below)
        dist.broadcast_(deltas)
        dist.broadcast_(w0)

    # ---- validation set: replicated (same seed on every rank), host copy kept for the e2e leg --
    g = torch.Generator(device=dev).manual_seed(a.seed + 424243)
    try:
        images_host = torch.empty((a.val, cfg.channels, cfg.image, cfg.image), dtype=torch.float32).pin_memory()
    except RuntimeError:
        images_host = torch.empty((a.val, cfg.channels, cfg.image, cfg.image), dtype=torch.float32)
    for s in range(0, a.val, 1000):
        n = min(1000, a.val - s)
        images_host[s:s + n].copy_(torch.randn((n, cfg.channels, cfg.image, cfg.image), generator=g, device=dev))
    labels_host = torch.randint(0, cfg.n_cls, (a.val,), generator=g, device=dev).cpu()
    val = ValidationSet(cfg, images_host, labels_host, prec, dev)
    if a.lora_rank:
        # N1: every rank builds the same seeded PEFT state_dicts (frozen base); the engine aggregates A and B separately
        from shapley_vit_b200 import lora as lora_mod

        w0_sd, client_sds = synth.make_peft_state_dicts(cfg, N, a.seed, r=a.lora_rank)
        delta_sds = [{k: c[k] - w0_sd[k] for k in c} for c in client_sds]
        del client_sds
        eng = lora_mod.LoraCoalitionEngine(cfg, w0_sd, delta_sds, val, lora_alpha=8.0, precision=prec, coalition_batch=Cb,
                                           image_chunk=a.image_chunk, device=dev, shared_base=a.lora_path == "shared")
        del delta_sds
        deltas, w0 = eng.deltas, eng.w0
    else:
        eng = CoalitionEngine(cfg, w0, deltas, val, precision=prec, coalition_batch=Cb, image_chunk=a.image_chunk,
                              device=dev)
    eng.profile = True

    # ---- coalition enumeration (exact Shapley order), FedAvg ratio rows -----------------------
    from shapley_vit_b200.estimators import powerset

    n_train = synth.client_sizes(N)
    coalitions = list(powerset(range(N)))
    clients = [ClientBase(i, {}, None, synth.SizedStub(n)) for i, n in enumerate(n_train)]
    server = ServerBase({}, None, clients, None, val, None)

    def rows_for(step: int, r: int):
        out = []
        for q in range(Cb):
            S = coalitions[((step * ws + r) * Cb + q) % len(coalitions)]
            ratio = server.get_agg_ratio(selected_clients=[clients[j] for j in S])
            row = [0.0] * N
            for j, v in zip(S, ratio):
                row[j] = v
            out.append(row)
        return out

    # ---- warm-up ------------------------------------------------------------------------------
    for i in range(a.warmup):
        eng.evaluate(rows_for(i, rank))
    torch.cuda.synchronize(dev)
    dist.barrier()

    # ---- timed region: K steps, inputs resident in HBM ----------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    eng.kernel_launches = 0
    eng.agg_spans.clear()
    eng.plan.timing_begin()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    dist.barrier()
    ev0.record()
    results = [eng._run_batch(rows_for(a.warmup + i, rank)) for i in range(a.steps)]
    ev1.record()
    torch.cuda.synchronize(dev)
    dist.barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler else None
    timing = eng.plan.timing_end()
    launches = eng.kernel_launches
    agg_ms = sum(s.elapsed_time(e) for s, e in eng.agg_spans)
    agg_launches = len(eng.agg_spans)
    for c, l in results:
        assert not torch.isnan(l).any(), "loss is nan"
    if ws > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    value = a.steps * Cb * ws / (elapsed_ms / 1e3)

    # ---- e2e: the public API (Game.eval_utilities) with HOST buffers every step ---------------
    e2e = None
    if not a.no_e2e:
        game = Game(clients, server, None, [None] * N, [True] * N, [0.0, 0.0], 2, {"precision": a.precision})
        game._engine = eng
        per_rank = Cb

        # Double-buffered input pipeline: the upload (pinned host -> HBM, + patchify) of step i+1 runs on a
        # copy stream while step i computes; the first upload of the timed region is not overlapped.
        val_b = ValidationSet(cfg, images_host, labels_host, prec, dev)
        vals = [val, val_b]
        copy_stream = torch.cuda.Stream(device=dev)
        ready = [[], []]          # per buffer: one event per 1 024 uploaded images (+ the labels, recorded first)
        step_imgs = ValidationSet.UPLOAD_STEP

        def start_upload(i: int) -> int:
            copy_stream.wait_stream(torch.cuda.current_stream(dev))      # the buffer's previous readers are done
            with torch.cuda.stream(copy_stream):
                ready[i & 1] = []
                h2d = vals[i & 1].upload(images_host, labels_host, events=ready[i & 1])   # images + labels, host -> device
            return h2d

        def e2e_step(i: int, last: bool):
            evs = ready[i & 1]
            cur = torch.cuda.current_stream(dev)
            # the forward of images [lo, hi) waits only for the upload events that cover them: compute starts on the
            # first 1 024 images while the rest of the set is still crossing PCIe
            eng.chunk_ready = lambda lo, hi: [cur.wait_event(evs[k]) for k in range(lo // step_imgs, (hi - 1) // step_imgs + 1)]
            eng.val = vals[i & 1]
            h2d = 0 if last else start_upload(i + 1)
            todo = [coalitions[((i * ws + r) * per_rank + q) % len(coalitions)] for r in range(ws) for q in range(per_rank)]
            game.utility = [{}, {}]
            game.eval_utilities(todo)                                    # ratios H2D, (correct, loss) D2H inside
            cur.wait_event(evs[-1])                                      # (the scoring reads the labels of the same buffer)
            eng.chunk_ready = None
            return h2d

        start_upload(0)
        e2e_step(0, True)                                                # warm-up
        torch.cuda.synchronize(dev)
        dist.barrier()
        t0 = time.perf_counter()
        h2d_b = start_upload(0)
        for i in range(a.steps):
            e2e_step(i, i == a.steps - 1)
        torch.cuda.synchronize(dev)
        wall = time.perf_counter() - t0
        h2d_b += eng.upload_bytes_per_batch()
        d2h_b = per_rank * 16
        eng.val = val
        if ws > 1:
            t = torch.tensor([wall], dtype=torch.float64, device=dev)
            td.all_reduce(t, op=td.ReduceOp.MAX)
            wall = float(t.item())
        e2e = {"value": a.steps * Cb * ws / wall, "unit": UNIT, "h2d_bytes_per_step": int(h2d_b),
               "d2h_bytes_per_step": int(d2h_b),
               "note": "per step: validation images+labels re-uploaded from pinned host memory and patchified (double-buffered: "
                       "the upload of step i+1 overlaps the compute of step i; the first step starts computing on the first 1 024 "
                       "images while the rest of its upload is in flight), ratio rows H2D, "
                       "per-coalition (correct, loss_sum) D2H through Game.eval_utilities; client deltas/W0 stay resident"}

    # ---- strong scaling: the whole job, "time to the Shapley vector" ---------------------------------
    tts = None
    if (a.time_to_shapley or ws >= 4) and not a.no_time_to_shapley:
        tts = time_to_shapley(a, eng, deltas, w0, clients, server, coalitions, dev, rank, ws, value / ws)
    elif rank == 0:
        tts = {"value": None, "note": f"not run at {ws} GPU(s) by default ({len(coalitions)} coalitions / {value:.2f} evals/s = "
                                      f"{len(coalitions) / value:.0f} s): pass --time-to-shapley; runs by default from 4 GPUs"}

    if rank != 0:
        if ws > 1:
            td.destroy_process_group()
        return

    # ---- the single-pass fp16 throughput mode on the same workload (outside the top-1 gate; reported, not the headline)
    throughput_mode = None
    if ws == 1 and not a.no_throughput_mode and a.precision in ("f16c8", "f16x3") and not a.lora_rank:
        try:
            p16 = _lib.PRECISIONS["f16"]
            val16 = ValidationSet(cfg, images_host, labels_host, p16, dev)
            eng16 = CoalitionEngine(cfg, w0, deltas, val16, precision=p16, coalition_batch=Cb, image_chunk=a.image_chunk, device=dev)
            eng16._run_batch(rows_for(0, 0))
            torch.cuda.synchronize(dev)
            t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0e.record()
            for i in range(2):
                eng16._run_batch(rows_for(a.warmup + i, 0))
            t1e.record()
            torch.cuda.synchronize(dev)
            throughput_mode = {"precision": "f16", "value": 2 * Cb / (t0e.elapsed_time(t1e) / 1e3), "unit": UNIT, "steps": 2, "warmup": 1,
                               "note": "single fp16 tcgen05 pass per product: ~99.2-99.7 % top-1 agreement with fp32 on random-init weights, "
                                       "i.e. OUTSIDE north_star's 99.9 % gate; `python bench.py --precision f16` is the full line"}
            del eng16, val16
            torch.cuda.empty_cache()
        except Exception as e:
            throughput_mode = {"error": repr(e)}

    parity = None
    if ws == 1 and not a.no_parity and not a.lora_rank:
        try:
            parity = parity_leg(a, cfg, lay, deltas, w0, images_host, labels_host, dev, val_main=val)
        except Exception as e:  # never lose the throughput line to the parity leg
            parity = {"error": repr(e)}

    # ---- roofline of the dominant kernel (the tcgen05 grouped GEMM) ---------------------------
    peaks, peak_src = load_peaks()
    g_ms, g_flops, g_n = timing["gemm"]
    tensor_peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    achieved = g_flops / (g_ms / 1e3) / 1e12 if g_ms > 0 else 0.0
    # DRAM bytes per launch: NOT measured in this run -- read from the committed `ncu --set full` capture of this
    # configuration and precision (profiles/r2_traffic.json, else round 1's f16 capture)
    traffic, traffic_agg, traffic_src = None, None, None
    if a.vit == "base" and a.coalition_batch == 8 and a.image_chunk == 128:
        for fn in ("r2_traffic.json", "r1_traffic.json"):
            try:
                with open(os.path.join(ROOT, "profiles", fn)) as f:
                    tj = json.load(f)
                tj = tj.get(a.precision, tj if (fn.startswith("r1") and a.precision in ("f16", "bf16")) else None)
                if tj:
                    traffic, traffic_agg = tj["gemm_avg_dram_bytes_per_launch"], tj["aggregate_dram_bytes_per_launch"]
                    traffic_src = f"profiles/{fn}: committed ncu --set full capture of this configuration (dram__bytes_read+write per launch), not this run"
                    break
            except Exception:
                pass
    roofline = {"bound": "tensor", "kernel": "gemm_tc2_kernel (tcgen05 cta_group::2 grouped GEMM)" if a.precision != "f32" else "gemm_simt_kernel",
                "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s", "frac": achieved / tensor_peak,
                "peak_source": f"{peak_src}; sustained figure (kernel timed inside a long step); burst = {peaks['bf16_tflops']}",
                "traffic": traffic, "traffic_source": traffic_src, "algorithmic_flops_per_launch": g_flops / max(g_n, 1),
                "launches": int(g_n), "avg_launch_ms": g_ms / max(g_n, 1),
                "share_of_step": g_ms / elapsed_ms}
    if a.precision in ("f16x3", "f16c8"):
        # every algorithmic product is 3 fp16 passes (f16x3) or 1 fp16 + 2 e4m3 passes at twice the rate (f16c8)
        passes = 3.0 if a.precision == "f16x3" else 2.0
        roofline.update({"mma_pass_equivalents": passes, "tensor_pipe_tflops": passes * achieved,
                         "tensor_pipe_frac": passes * achieved / tensor_peak,
                         "note": "achieved / frac count ALGORITHMIC flops against the bf16 dense peak; the tensor pipe is busy for "
                                 f"{passes:g} fp16-pass equivalents per product (tensor_pipe_frac)"})
    a_ms, a_flops, a_n = timing["attention"]
    l_ms, l_bytes, l_n = timing["layernorm"]
    es = 2 if a.precision in ("f16", "bf16") else 4   # bytes per weight element over all planes
    agg_bytes = a.steps * (4.0 * lay.total * (N + 1) + (4.0 * lay.vec_size + es * lay.mat_size) * Cb)
    if a.lora_rank and getattr(eng, "shared", False):   # vec region + the packed factor rows only
        lw = eng.ext_deltas.shape[1]
        agg_bytes = a.steps * (4.0 * (lay.vec_size + lw) * (N + 1) + (4.0 * lay.vec_size + es * lw) * Cb)
    breakdown = {
        "gemm_ms": g_ms, "attention_ms": a_ms, "attention_tflops": a_flops / (a_ms / 1e3) / 1e12 if a_ms else None,
        "layernorm_ms": l_ms, "layernorm_gbs": l_bytes / (l_ms / 1e3) / 1e9 if l_ms else None,
        "aggregate_ms": agg_ms, "forward_ms": timing["forward"][0], "step_ms_total": elapsed_ms,
    }
    roofline_agg = {"bound": "hbm", "kernel": "aggregate_kernel (K1)", "achieved": agg_bytes / (agg_ms / 1e3) / 1e9 if agg_ms else None,
                    "peak": peaks["hbm_gbs"], "unit": "GB/s", "launches": agg_launches,
                    "frac": (agg_bytes / (agg_ms / 1e3) / 1e9 / peaks["hbm_gbs"]) if agg_ms else None,
                    "algorithmic_bytes_per_step": agg_bytes / a.steps, "traffic": traffic_agg}
    model_flops_nominal = cfg.flops_per_image() * a.val * Cb * a.steps     # HF forward, every token of every layer
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ws, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": elapsed_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": a.precision, "data": "synthetic",
        "config": {"workload": workload_name(a), "step": f"{Cb} coalitions per GPU: aggregate + forward over {a.val} images + score",
                   "coalition_batch": Cb, "image_chunk": a.image_chunk, "parallelism": f"coalition-sharded x{ws}",
                   "l2": "inputs per step (2.7 GB delta stack, 3 GB patch matrix) exceed the 126 MB L2; no flush needed",
                   # executed GEMM + attention FLOPs (the last layer computes only the [CLS] rows past K and V)
                   "model_tflops_per_s_per_gpu": (g_flops + a_flops) / (elapsed_ms / 1e3) / 1e12,
                   "hf_equivalent_tflops_per_s_per_gpu": model_flops_nominal / (elapsed_ms / 1e3) / 1e12},
        "roofline": roofline, "roofline_aggregate": roofline_agg, "breakdown": breakdown,
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "parity": parity, "throughput_mode": throughput_mode,
        "time_to_shapley_s": tts["value"] if tts else None, "time_to_shapley": tts,
    }
    if ws == 1 and not a.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline_subprocess(a)
        except Exception as e:  # never lose the throughput line to the CPU leg
            line["cpu_baseline"] = {"error": repr(e)}
    print(json.dumps(line), flush=True)
    if ws > 1:
        td.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == "reference":
        os.environ["CUDA_VISIBLE_DEVICES"] = ""   # the CPU arm: the reference would otherwise use cuda:1 / cuda:0
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
