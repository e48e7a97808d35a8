"""The ViT definition (HF ViTForImageClassification key-compatible); the reference builds it
from HuggingFace at start.py:258-267."""
from shapley_vit_b200.models.vit import ViTForImageClassification, infer_config, vit  # noqa: F401
