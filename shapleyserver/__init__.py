"""Drop-in mirror of the reference package path; implementation lives in shapley_vit_b200."""
