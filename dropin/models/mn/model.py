from efficientat_b200.models.mn.model import *  # noqa: F401,F403
from efficientat_b200.models.mn.model import MN, get_model, mobilenet_v3, pretrained_models  # noqa: F401
# windowed_inference.py:8 imports `get_ensemble_model` from models.mn.model (the reference defines it in models/ensemble.py
# only, so that script cannot be imported against the reference's own package); exporting it here lets it run unchanged
from efficientat_b200.models.ensemble import get_ensemble_model  # noqa: E402,F401
