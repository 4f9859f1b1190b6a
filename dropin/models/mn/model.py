from efficientat_b200.models.mn.model import *  # noqa: F401,F403
from efficientat_b200.models.mn.model import MN, get_model, mobilenet_v3, pretrained_models  # noqa: F401
