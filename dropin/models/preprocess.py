from efficientat_b200.models.preprocess import *  # noqa: F401,F403
from efficientat_b200.models.preprocess import AugmentMelSTFT  # noqa: F401
