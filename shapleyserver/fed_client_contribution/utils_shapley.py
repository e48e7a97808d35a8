"""Drop-in for reference shapleyserver/fed_client_contribution/utils_shapley.py."""
from shapley_vit_b200.estimators import (  # noqa: F401
    METHODS, _cc_shap_task, call_shapley_computation_method, ncr, powerset, shapley_comp_contrib,
    shapley_exact, shapley_exact_own, shapley_monte_carlo)
