"""Deterministic synthetic inputs for the utility loop.

Everything is generated on the CPU from ``torch.Generator`` seeds so that the
oracle (CPU) and the CUDA path see bit-identical inputs:

* ``W0``      -- random-init ViT state_dict, N(0, 0.02) weights; biases and
  LayerNorm parameters are *also* perturbed (HF's default zeros/ones would hide
  bias / gain bugs in a parity test);
* client ``j`` -- ``W0 + sigma * N(0, 1)`` per tensor;
* ``n_j``     -- ``1000 * (j + 1)`` training samples, so FedAvg ratios differ;
* val set     -- N(0,1) images, uniform labels, served as the dict samples
  ``{'image', 'label', 'image_name'}`` the reference's ``evaluation`` expects
  (reference ``federated_learning/utils.py:880``).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Tuple

import torch

from .layout import VitConfig, state_dict_spec


def make_state_dict(cfg: VitConfig, seed: int = 0, std: float = 0.02) -> "OrderedDict[str, torch.Tensor]":
    g = torch.Generator().manual_seed(1_000_003 * seed + 17)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for key, shape in state_dict_spec(cfg):
        t = torch.randn(shape, generator=g, dtype=torch.float32) * std
        if key.endswith("layernorm_before.weight") or key.endswith("layernorm_after.weight") \
                or key == "vit.layernorm.weight":
            t += 1.0
        sd[key] = t
    return sd


def make_client_state_dict(w0: Dict[str, torch.Tensor], client: int, seed: int = 0,
                           sigma: float = 0.02) -> "OrderedDict[str, torch.Tensor]":
    g = torch.Generator().manual_seed(1_000_003 * seed + 7919 * (client + 1))
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for key, t in w0.items():
        out[key] = t + sigma * torch.randn(t.shape, generator=g, dtype=torch.float32)
    return out


def make_peft_state_dicts(cfg: VitConfig, n_clients: int, seed: int = 0, r: int = 16, frozen_base: bool = True,
                          prefix: str = "module.base_model.model.", sigma: float = 0.02):
    """Synthetic stand-in for the author's setup (reference start.py:274-283): a PEFT-LoRA wrapped ViT
    state_dict (query / value wrapped, rank r; classifier in modules_to_save) for the initial model and
    for ``n_clients`` locally trained models.  Initial model: A random, B = 0 (PEFT's init); clients: A, B
    and the classifier moved, the base frozen unless ``frozen_base`` is False.
    Returns (w0_sd, [client_sd, ...]) with PEFT key names."""
    base = make_state_dict(cfg, seed)
    h = cfg.hidden

    def wrap(sd, lora, cls_w, cls_b):
        out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
        for k, v in sd.items():
            m = k.endswith("attention.attention.query.weight") or k.endswith("attention.attention.value.weight") \
                or k.endswith("attention.attention.query.bias") or k.endswith("attention.attention.value.bias")
            if m:
                stem, leaf = k.rsplit(".", 1)
                out[f"{prefix}{stem}.base_layer.{leaf}"] = v
                if leaf == "weight":
                    a, b = lora[stem]
                    out[f"{prefix}{stem}.lora_A.default.weight"] = a
                    out[f"{prefix}{stem}.lora_B.default.weight"] = b
            elif k.startswith("classifier."):
                leaf = k.split(".", 1)[1]
                out[f"{prefix}classifier.original_module.{leaf}"] = v
                out[f"{prefix}classifier.modules_to_save.default.{leaf}"] = cls_w if leaf == "weight" else cls_b
            else:
                out[prefix + k] = v
        return out

    g = torch.Generator().manual_seed(1_000_003 * seed + 31)
    stems = [f"vit.encoder.layer.{i}.attention.attention.{t}" for i in range(cfg.layers) for t in ("query", "value")]
    lora0 = {st: (torch.randn(r, h, generator=g) * (1.0 / h) ** 0.5, torch.zeros(h, r)) for st in stems}
    w0 = wrap(base, lora0, base["classifier.weight"], base["classifier.bias"])
    clients = []
    for j in range(n_clients):
        gj = torch.Generator().manual_seed(1_000_003 * seed + 7919 * (j + 1) + 5)
        lj = {st: (a + sigma * torch.randn(a.shape, generator=gj), 4 * sigma * torch.randn(b.shape, generator=gj))
              for st, (a, b) in lora0.items()}
        bj = base if frozen_base else make_client_state_dict(base, j, seed, sigma)
        clients.append(wrap(bj, lj, base["classifier.weight"] + sigma * torch.randn(cfg.n_cls, h, generator=gj),
                            base["classifier.bias"] + sigma * torch.randn(cfg.n_cls, generator=gj)))
    return w0, clients


def client_sizes(n_clients: int) -> List[int]:
    return [1000 * (j + 1) for j in range(n_clients)]


def make_val_set(cfg: VitConfig, n: int, seed: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    g = torch.Generator().manual_seed(1_000_003 * seed + 424_243)
    images = torch.randn((n, cfg.channels, cfg.image, cfg.image), generator=g, dtype=torch.float32)
    labels = torch.randint(0, cfg.n_cls, (n,), generator=g, dtype=torch.int64)
    return images, labels


class DictSampleDataset(torch.utils.data.Dataset):
    """Dataset yielding the reference's sample dicts."""

    def __init__(self, images: torch.Tensor, labels: torch.Tensor):
        assert images.shape[0] == labels.shape[0]
        self.images, self.labels = images, labels

    def __len__(self) -> int:
        return self.images.shape[0]

    def __getitem__(self, i: int):
        return {"image": self.images[i], "label": self.labels[i], "image_name": f"synthetic_{i:06d}"}


class SizedStub:
    """Stands in for a client's training set: only ``len()`` is ever read
    (reference ``federated_learning/client2.py:14-15``)."""

    def __init__(self, n: int):
        self._n = int(n)

    def __len__(self) -> int:
        return self._n


def lazy_rounds_inputs(seed: int = 11, n_clients: int = 3, n_rounds: int = 3, n_val: int = 200):
    """Seeded inputs of the multi-round ("lazy") fixture tests/golden/lazy_rounds.json:
    (cfg, w0, round_sds[t][j], selection[t][j], n_train, images, labels)."""
    from . import layout

    cfg = layout.vit_preset("tiny", image=32, n_cls=10, layers=2)
    w0 = make_state_dict(cfg, seed)
    round_sds = [[make_client_state_dict(w0, j, seed + 100 * (t + 1)) for j in range(n_clients)]
                 for t in range(n_rounds)]
    selection = [[True, True, False], [True, False, True], [False, True, True]][:n_rounds]
    images, labels = make_val_set(cfg, n_val, seed)
    return cfg, w0, round_sds, selection, client_sizes(n_clients), images, labels
