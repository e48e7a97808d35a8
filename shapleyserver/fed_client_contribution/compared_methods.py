"""Drop-in for reference shapleyserver/fed_client_contribution/compared_methods.py."""
from shapley_vit_b200.compared import (  # noqa: F401
    GTG, MR, TMR, Fed_SV, ShapleyValue, call_comfedsv, comfedsv, roundly_mask, shapley_value)
