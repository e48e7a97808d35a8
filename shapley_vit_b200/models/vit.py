"""ViT definition whose ``state_dict()`` is key-for-key the HuggingFace
``ViTForImageClassification`` the reference builds (reference start.py:258-267), so client
checkpoints (``{'state_dict': ...}``, start.py:146-151) load unchanged.

The module is a parameter container.  Its ``forward`` returns an object with ``.logits`` like
HF's, computed by libsvit's batched forward with a single coalition (C = 1) -- it exists so
that code written against ``net(img).logits`` (reference federated_learning/utils.py:886)
keeps working; the utility loop itself never goes through nn.Module.forward.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, Optional

import torch
from torch import nn

from .._lib import DEFAULT_PRECISION
from ..layout import VitConfig, pack_state_dict, plan_layout, state_dict_spec, vit_preset


class ViTForImageClassification(nn.Module):
    def __init__(self, cfg: VitConfig, precision: str = DEFAULT_PRECISION):
        super().__init__()
        self.cfg = cfg
        self.precision = precision
        self._names = []
        for key, shape in state_dict_spec(cfg):
            p = nn.Parameter(torch.empty(shape, dtype=torch.float32), requires_grad=False)
            self._register(key, p)
        self.reset_parameters()
        self._plan = None

    # parameters are registered on nested holder modules so state_dict() yields HF's dotted keys
    def _register(self, dotted: str, p: nn.Parameter) -> None:
        mod: nn.Module = self
        parts = dotted.split(".")
        for name in parts[:-1]:
            if not hasattr(mod, name):
                mod.add_module(name, nn.Module())
            mod = getattr(mod, name)
        mod.register_parameter(parts[-1], p)
        self._names.append(dotted)

    @torch.no_grad()
    def reset_parameters(self, std: float = 0.02, seed: Optional[int] = None) -> None:
        """HF default init: N(0, initializer_range) weights, zero biases, unit LayerNorm."""
        g = torch.Generator().manual_seed(seed) if seed is not None else None
        for name, p in self.named_parameters():
            if name.endswith("bias"):
                p.zero_()
            elif "layernorm" in name and name.endswith("weight"):
                p.fill_(1.0)
            else:
                p.copy_(torch.randn(p.shape, generator=g) * std)

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        """Accepts checkpoints saved from ``nn.DataParallel`` / PEFT wrappers: leading ``module.`` (and, for the
        plain ViT, ``base_model.model.``) prefixes are dropped when the bare key is one of this module's."""
        own = set(self.state_dict().keys())
        fixed = {}
        for k, v in state_dict.items():
            kk = k
            while kk not in own and (kk.startswith("module.") or kk.startswith("base_model.model.")):
                kk = kk[len("module."):] if kk.startswith("module.") else kk[len("base_model.model."):]
            fixed[kk] = v
        return super().load_state_dict(fixed, strict=strict, **kw)

    @property
    def config(self):  # a few HF-style attribute names
        c = self.cfg
        return SimpleNamespace(hidden_size=c.hidden, num_hidden_layers=c.layers, num_attention_heads=c.heads,
                               intermediate_size=c.ff, image_size=c.image, patch_size=c.patch,
                               num_channels=c.channels, num_labels=c.n_cls, layer_norm_eps=c.ln_eps)

    @torch.no_grad()
    def forward(self, pixel_values: torch.Tensor):
        from .. import ops
        from .._lib import PRECISIONS

        if not pixel_values.is_cuda:
            raise RuntimeError("ViTForImageClassification.forward runs on libsvit (CUDA, sm_100a) only")
        dev = pixel_values.device
        B = pixel_values.shape[0]
        lay = plan_layout(self.cfg)
        if self._plan is None or self._plan.max_images < B or self._plan.device != dev:
            self._plan = ops.Plan(self.cfg, PRECISIONS[self.precision], 1, max(B, 1), dev)
        plan = self._plan
        row = pack_state_dict(lay, {k: v.detach() for k, v in self.state_dict().items()}).to(dev).unsqueeze(0)
        one = torch.ones((1, 1), dtype=torch.float32)
        V = lay.vec_size
        wvec = ops.aggregate(row[:, :V], None, one, out_dtype=torch.float32, P=V)
        wmat = ops.aggregate(row[:, V:], None, one, out=plan.operand_array((1, lay.mat_size)), P=lay.mat_size)
        patches = plan.patchify(pixel_values.to(torch.float32))
        logits = torch.empty((1, B, self.cfg.n_cls), dtype=torch.float32, device=dev)
        plan.forward(wvec, wmat, patches, 0, B, logits)
        return SimpleNamespace(logits=logits[0])


def peft_state_dict_spec(cfg: VitConfig, r: int):
    """(key, shape) of a PEFT-LoRA wrapped HF ViT as the reference builds it (start.py:274-283): ``query`` and
    ``value`` wrapped (``base_layer`` + ``lora_A/B.default``), classifier in ``modules_to_save``."""
    out = []
    h = cfg.hidden
    for key, shape in state_dict_spec(cfg):
        stem, leaf = key.rsplit(".", 1)
        if stem.endswith("attention.attention.query") or stem.endswith("attention.attention.value"):
            out.append((f"base_model.model.{stem}.base_layer.{leaf}", shape))
            if leaf == "weight":
                out.append((f"base_model.model.{stem}.lora_A.default.weight", (r, h)))
                out.append((f"base_model.model.{stem}.lora_B.default.weight", (h, r)))
        elif key.startswith("classifier."):
            out.append((f"base_model.model.classifier.original_module.{leaf}", shape))
            out.append((f"base_model.model.classifier.modules_to_save.default.{leaf}", shape))
        else:
            out.append((f"base_model.model.{key}", shape))
    return out


class LoraViTForImageClassification(ViTForImageClassification):
    """Parameter container with the state_dict keys of ``get_peft_model(vit, LoraConfig(r, lora_alpha,
    target_modules=['query', 'value'], modules_to_save=['classifier']))`` (reference start.py:274-276), so the
    author's client checkpoints load unchanged.  PEFT's init: A Kaiming-uniform, B zero.  ``forward`` scores the
    merged model W + (alpha / r) B A."""

    def __init__(self, cfg: VitConfig, r: int = 16, lora_alpha: float = 8.0, precision: str = DEFAULT_PRECISION):
        nn.Module.__init__(self)
        self.cfg, self.precision, self.r, self.lora_alpha = cfg, precision, r, lora_alpha
        self._names = []
        for key, shape in peft_state_dict_spec(cfg, r):
            self._register(key, nn.Parameter(torch.empty(shape, dtype=torch.float32), requires_grad=False))
        self.reset_parameters()
        self._plan = None

    @torch.no_grad()
    def reset_parameters(self, std: float = 0.02, seed: Optional[int] = None) -> None:
        g = torch.Generator().manual_seed(seed) if seed is not None else None
        for name, p in self.named_parameters():
            if ".lora_B." in name or name.endswith("bias"):
                p.zero_()
            elif ".lora_A." in name:
                bound = (1.0 / p.shape[1]) ** 0.5          # kaiming_uniform_(a = sqrt(5)) on [r, h]
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * bound)
            elif "layernorm" in name and name.endswith("weight"):
                p.fill_(1.0)
            else:
                p.copy_(torch.randn(p.shape, generator=g) * std)

    @torch.no_grad()
    def forward(self, pixel_values: torch.Tensor):
        from ..lora import merged_state_dict

        plain = ViTForImageClassification(self.cfg, precision=self.precision)
        plain.load_state_dict(merged_state_dict(self.state_dict(), self.lora_alpha))
        return plain(pixel_values)


def infer_config(sd: Dict[str, torch.Tensor], heads: Optional[int] = None, ln_eps: float = 1e-12) -> VitConfig:
    """Recover the geometry from an HF ViT state_dict (head count is not recoverable from
    shapes; default = hidden / 64 as in every ViT-Ti/S/B/L)."""
    def strip(k):
        for pre in ("module.", "base_model.model."):
            while k.startswith(pre):
                k = k[len(pre):]
        return k

    sd = {strip(k): v for k, v in sd.items()}
    proj = sd["vit.embeddings.patch_embeddings.projection.weight"]
    hidden, channels, patch = proj.shape[0], proj.shape[1], proj.shape[2]
    T = sd["vit.embeddings.position_embeddings"].shape[1]
    side = int(round((T - 1) ** 0.5))
    layers = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("vit.encoder.layer."))
    ff = sd["vit.encoder.layer.0.intermediate.dense.weight"].shape[0]
    n_cls = sd["classifier.weight"].shape[0]
    return VitConfig(hidden=hidden, layers=layers, heads=heads or max(1, hidden // 64), ff=ff, image=side * patch,
                     n_cls=n_cls, patch=patch, channels=channels, ln_eps=ln_eps)


def vit(name: str = "base", image: int = 224, n_cls: int = 4, **kw) -> ViTForImageClassification:
    return ViTForImageClassification(vit_preset(name, image=image, n_cls=n_cls), **kw)
