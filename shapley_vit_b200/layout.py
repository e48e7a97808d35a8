"""ViT geometry, the HuggingFace state_dict key order, and the *plan layout*.

The reference never defines the ViT itself: it instantiates HuggingFace
``ViTForImageClassification`` (reference ``shapleyserver/start.py:258-267``) and
treats a model as its ``state_dict()`` -- an ordered ``name -> fp32 tensor`` map
(``federated_learning/utils.py:735-749, 781-792``; ``server2.py:121-127``).
FedAvg aggregation is element-wise over that map, so any fixed permutation of
the flattened parameters is an equally valid representation.  This module fixes
ONE such permutation, the *plan layout*, chosen for the sm_100a kernels:

* a **vector region** (always fp32): cls token, position embeddings, every bias,
  every LayerNorm gain/offset and the (tiny) classifier head;
* a **matrix region** (fp32 / fp16 / bf16, the GEMM operand type): the
  patch-projection and the four dense weights of every encoder layer, with the
  query/key/value weights stored back to back so they form one ``[3h, h]``
  K-major operand.

Every segment starts on a 64-element boundary, which makes every GEMM operand
16-byte aligned for TMA in both 2- and 4-byte element types.  The same table is
computed in C (``csrc/layout.cu``, exported as ``svit_layout_*``); a test asserts
that both agree.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Sequence, Tuple

ALIGN = 64  # elements; see module docstring

# region ids (shared with include/svit.h)
REGION_VEC = 0
REGION_MAT = 1


@dataclass(frozen=True)
class VitConfig:
    """Geometry of a HF ``ViTForImageClassification`` (qkv_bias=True, gelu-erf,
    pre-LN, no pooler, dropout 0).  Mirrors ``svit_vit_cfg`` in include/svit.h."""

    hidden: int
    layers: int
    heads: int
    ff: int
    image: int
    n_cls: int
    patch: int = 16
    channels: int = 3
    ln_eps: float = 1e-12

    def __post_init__(self):
        if self.hidden % self.heads:
            raise ValueError("hidden must be divisible by heads")
        if self.image % self.patch:
            raise ValueError("image must be divisible by patch")
        if self.hidden % ALIGN or self.ff % ALIGN:
            raise ValueError(f"hidden and ff must be multiples of {ALIGN}")

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads

    @property
    def n_patches(self) -> int:
        return (self.image // self.patch) ** 2

    @property
    def tokens(self) -> int:
        return self.n_patches + 1

    @property
    def patch_dim(self) -> int:
        return self.channels * self.patch * self.patch

    def flops_per_image(self) -> int:
        """2*MAC of one forward."""
        h, T, L = self.hidden, self.tokens, self.layers
        mac = L * T * (4 * h * h + 2 * h * self.ff + 2 * T * h)
        mac += self.n_patches * self.patch_dim * h + h * self.n_cls
        return 2 * mac


PRESETS: Dict[str, Tuple[int, int, int, int]] = {
    # name: (hidden, layers, heads, ff)
    "tiny": (192, 12, 3, 768),
    "small": (384, 12, 6, 1536),
    "base": (768, 12, 12, 3072),
    "large": (1024, 24, 16, 4096),
}


def vit_preset(name: str, image: int = 224, n_cls: int = 10, layers: int | None = None) -> VitConfig:
    hidden, nl, heads, ff = PRESETS[name]
    return VitConfig(hidden=hidden, layers=layers if layers is not None else nl,
                     heads=heads, ff=ff, image=image, n_cls=n_cls)


# --------------------------------------------------------------------------- #
# HuggingFace state_dict order (transformers ViTForImageClassification)
# --------------------------------------------------------------------------- #

def state_dict_spec(cfg: VitConfig) -> List[Tuple[str, Tuple[int, ...]]]:
    """``[(key, shape)]`` in the iteration order of the HF module's state_dict."""
    h, ff, T = cfg.hidden, cfg.ff, cfg.tokens
    spec: List[Tuple[str, Tuple[int, ...]]] = [
        ("vit.embeddings.cls_token", (1, 1, h)),
        ("vit.embeddings.position_embeddings", (1, T, h)),
        ("vit.embeddings.patch_embeddings.projection.weight", (h, cfg.channels, cfg.patch, cfg.patch)),
        ("vit.embeddings.patch_embeddings.projection.bias", (h,)),
    ]
    for i in range(cfg.layers):
        p = f"vit.encoder.layer.{i}."
        for nm in ("query", "key", "value"):
            spec.append((p + f"attention.attention.{nm}.weight", (h, h)))
            spec.append((p + f"attention.attention.{nm}.bias", (h,)))
        spec += [
            (p + "attention.output.dense.weight", (h, h)),
            (p + "attention.output.dense.bias", (h,)),
            (p + "intermediate.dense.weight", (ff, h)),
            (p + "intermediate.dense.bias", (ff,)),
            (p + "output.dense.weight", (h, ff)),
            (p + "output.dense.bias", (h,)),
            (p + "layernorm_before.weight", (h,)),
            (p + "layernorm_before.bias", (h,)),
            (p + "layernorm_after.weight", (h,)),
            (p + "layernorm_after.bias", (h,)),
        ]
    spec += [
        ("vit.layernorm.weight", (h,)),
        ("vit.layernorm.bias", (h,)),
        ("classifier.weight", (cfg.n_cls, h)),
        ("classifier.bias", (cfg.n_cls,)),
    ]
    return spec


def num_params(cfg: VitConfig) -> int:
    n = 0
    for _, shp in state_dict_spec(cfg):
        k = 1
        for s in shp:
            k *= s
        n += k
    return n


# --------------------------------------------------------------------------- #
# plan layout
# --------------------------------------------------------------------------- #

# segment kinds (shared with include/svit.h: enum svit_seg_kind)
(K_CLS, K_POS, K_PATCH_B, K_LN1_G, K_LN1_B, K_BQ, K_BK, K_BV, K_BO, K_LN2_G, K_LN2_B,
 K_B1, K_B2, K_LNF_G, K_LNF_B, K_HEAD_W, K_HEAD_B,
 K_PATCH_W, K_WQ, K_WK, K_WV, K_WO, K_W1, K_W2) = range(24)


@dataclass(frozen=True)
class Segment:
    kind: int
    layer: int          # -1 for non-layer segments
    region: int         # REGION_VEC / REGION_MAT
    offset: int         # element offset inside the region
    size: int           # elements
    key: str            # HF state_dict key
    shape: Tuple[int, ...]


def _round_up(x: int, a: int) -> int:
    return (x + a - 1) // a * a


@dataclass
class PlanLayout:
    cfg: VitConfig
    segments: List[Segment] = field(default_factory=list)
    vec_size: int = 0     # padded, multiple of ALIGN
    mat_size: int = 0     # multiple of ALIGN

    @property
    def total(self) -> int:
        """Row length of a model in plan layout: [vec_size | mat_size]."""
        return self.vec_size + self.mat_size

    def by_key(self) -> Dict[str, Segment]:
        return {s.key: s for s in self.segments}

    def find(self, kind: int, layer: int = -1) -> Segment:
        for s in self.segments:
            if s.kind == kind and s.layer == layer:
                return s
        raise KeyError((kind, layer))

    def flat_offset(self, seg: Segment) -> int:
        return seg.offset if seg.region == REGION_VEC else self.vec_size + seg.offset


def plan_layout(cfg: VitConfig) -> PlanLayout:
    h, ff, T = cfg.hidden, cfg.ff, cfg.tokens
    lay = PlanLayout(cfg)
    off = [0, 0]

    def add(kind, layer, region, key, shape):
        size = 1
        for s in shape:
            size *= s
        lay.segments.append(Segment(kind, layer, region, off[region], size, key, tuple(shape)))
        off[region] += _round_up(size, ALIGN)

    e = "vit.embeddings."
    add(K_CLS, -1, REGION_VEC, e + "cls_token", (1, 1, h))
    add(K_POS, -1, REGION_VEC, e + "position_embeddings", (1, T, h))
    add(K_PATCH_B, -1, REGION_VEC, e + "patch_embeddings.projection.bias", (h,))
    for i in range(cfg.layers):
        p = f"vit.encoder.layer.{i}."
        add(K_LN1_G, i, REGION_VEC, p + "layernorm_before.weight", (h,))
        add(K_LN1_B, i, REGION_VEC, p + "layernorm_before.bias", (h,))
        add(K_BQ, i, REGION_VEC, p + "attention.attention.query.bias", (h,))
        add(K_BK, i, REGION_VEC, p + "attention.attention.key.bias", (h,))
        add(K_BV, i, REGION_VEC, p + "attention.attention.value.bias", (h,))
        add(K_BO, i, REGION_VEC, p + "attention.output.dense.bias", (h,))
        add(K_LN2_G, i, REGION_VEC, p + "layernorm_after.weight", (h,))
        add(K_LN2_B, i, REGION_VEC, p + "layernorm_after.bias", (h,))
        add(K_B1, i, REGION_VEC, p + "intermediate.dense.bias", (ff,))
        add(K_B2, i, REGION_VEC, p + "output.dense.bias", (h,))
    add(K_LNF_G, -1, REGION_VEC, "vit.layernorm.weight", (h,))
    add(K_LNF_B, -1, REGION_VEC, "vit.layernorm.bias", (h,))
    add(K_HEAD_W, -1, REGION_VEC, "classifier.weight", (cfg.n_cls, h))
    add(K_HEAD_B, -1, REGION_VEC, "classifier.bias", (cfg.n_cls,))

    add(K_PATCH_W, -1, REGION_MAT, e + "patch_embeddings.projection.weight",
        (h, cfg.channels, cfg.patch, cfg.patch))
    for i in range(cfg.layers):
        p = f"vit.encoder.layer.{i}."
        add(K_WQ, i, REGION_MAT, p + "attention.attention.query.weight", (h, h))
        add(K_WK, i, REGION_MAT, p + "attention.attention.key.weight", (h, h))
        add(K_WV, i, REGION_MAT, p + "attention.attention.value.weight", (h, h))
        add(K_WO, i, REGION_MAT, p + "attention.output.dense.weight", (h, h))
        add(K_W1, i, REGION_MAT, p + "intermediate.dense.weight", (ff, h))
        add(K_W2, i, REGION_MAT, p + "output.dense.weight", (h, ff))
    lay.vec_size = off[REGION_VEC]
    lay.mat_size = off[REGION_MAT]
    assert {s.key for s in lay.segments} == {k for k, _ in state_dict_spec(cfg)}
    return lay


# --------------------------------------------------------------------------- #
# pack / unpack between a state_dict and one plan-layout row
# --------------------------------------------------------------------------- #

def pack_state_dict(lay: PlanLayout, sd, out=None):
    """Flatten ``sd`` (name -> tensor; extra prefixes such as ``module.`` are the
    caller's problem) into one fp32 row of length ``lay.total`` (padding = 0)."""
    import torch

    if out is None:
        out = torch.zeros(lay.total, dtype=torch.float32)
    else:
        out.zero_()
    for s in lay.segments:
        t = sd[s.key]
        if tuple(t.shape) != s.shape:
            raise ValueError(f"{s.key}: expected {s.shape}, got {tuple(t.shape)}")
        o = lay.flat_offset(s)
        out[o:o + s.size].copy_(t.detach().reshape(-1).to(torch.float32))
    return out


def unpack_row(lay: PlanLayout, row, order: Sequence[str] | None = None):
    """Inverse of :func:`pack_state_dict`; returns an ordered dict in HF key order."""
    segs = lay.by_key()
    keys: Iterable[str] = order if order is not None else [k for k, _ in state_dict_spec(lay.cfg)]
    out = {}
    for k in keys:
        s = segs[k]
        o = lay.flat_offset(s)
        out[k] = row[o:o + s.size].reshape(s.shape)
    return out
