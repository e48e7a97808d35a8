"""Drop-in for reference shapleyserver/fed_client_contribution/game.py (Game :4-116)."""
from shapley_vit_b200.game import Game  # noqa: F401
