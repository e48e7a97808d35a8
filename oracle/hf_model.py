"""TEST INFRASTRUCTURE ONLY -- build the reference's third-party model
(HuggingFace ``ViTForImageClassification``, reference start.py:258-267) from a
raw state_dict.  Used by ``oracle/make_golden.py`` (driving the real reference)
and by ``bench.py``'s CPU-baseline leg (the reference's own eager fp32 path)."""
from __future__ import annotations


def build_hf_vit(cfg, sd=None):
    from transformers import ViTConfig, ViTForImageClassification

    hf_cfg = ViTConfig(
        hidden_size=cfg.hidden, num_hidden_layers=cfg.layers, num_attention_heads=cfg.heads,
        intermediate_size=cfg.ff, image_size=cfg.image, patch_size=cfg.patch,
        num_channels=cfg.channels, num_labels=cfg.n_cls, layer_norm_eps=cfg.ln_eps,
        hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, qkv_bias=True, hidden_act="gelu")
    model = ViTForImageClassification(hf_cfg)
    if sd is not None:
        model.load_state_dict(sd, strict=True)
    return model.eval()
