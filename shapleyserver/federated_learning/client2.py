"""Drop-in for reference shapleyserver/federated_learning/client2.py (ClientBase :7-42)."""
from shapley_vit_b200.fl import ClientBase  # noqa: F401
