from efficientat_b200.models.ensemble import EnsemblerModel, get_ensemble_model  # noqa: F401
