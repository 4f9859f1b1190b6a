from efficientat_b200.models.dymn.model import *  # noqa: F401,F403
from efficientat_b200.models.dymn.model import DyMN, dymn, get_model, pretrained_models  # noqa: F401
