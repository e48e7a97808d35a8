"""Drop-in for the reference's fed_client_contribution/utils_fed_shapley.py (multi-round utilities)."""
from shapley_vit_b200.estimators import ncr, powerset  # noqa: F401
from shapley_vit_b200.fed_shapley import (  # noqa: F401
    compute_shapley_value_baseline, compute_shapley_value_from_matrix, compute_shapley_value_groundtruth,
    compute_utilities_lazy, get_selection_dict, roundly_mask)
