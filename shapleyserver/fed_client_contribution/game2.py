"""Drop-in for reference shapleyserver/fed_client_contribution/game2.py -- the copy start.py:15
imports.  It hard-codes three clients (game2.py:25); the general Game has the same semantics."""
from shapley_vit_b200.game import Game  # noqa: F401
