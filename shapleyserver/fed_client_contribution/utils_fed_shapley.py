"""Drop-in for the reference's fed_client_contribution/utils_fed_shapley.py (multi-round utilities)."""
from shapley_vit_b200.estimators import ncr, powerset  # noqa: F401
from shapley_vit_b200.fed_shapley import compute_utilities_lazy  # noqa: F401
