"""Drop-in for the hot-path part of reference shapleyserver/federated_learning/utils.py
(get_difference_between_network_weights :735, get_aggregated_model :781, evaluation :864)."""
from shapley_vit_b200.fl import (  # noqa: F401
    evaluation, get_aggregated_model, get_difference_between_network_weights)
