"""Drop-in for reference shapleyserver/fed_client_contribution/compared_methods.py."""
from shapley_vit_b200.compared import GTG, MR, TMR, Fed_SV, ShapleyValue, shapley_value  # noqa: F401
