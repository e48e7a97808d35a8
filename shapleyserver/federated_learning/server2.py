"""Drop-in for reference shapleyserver/federated_learning/server2.py (ServerBase :15-127)."""
from shapley_vit_b200.fl import ServerBase  # noqa: F401
