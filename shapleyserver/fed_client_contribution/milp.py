"""Drop-in for the reference's fed_client_contribution/milp.py (round selection for lazy Shapley)."""
from shapley_vit_b200.round_select import (  # noqa: F401
    MILP_Shapley, MILP_Shapley_Two_Sided, MILP_Shapley_Two_Sided_Approx)
