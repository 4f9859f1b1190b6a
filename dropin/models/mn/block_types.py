from efficientat_b200.models.mn.block_types import *  # noqa: F401,F403
from efficientat_b200.models.mn.block_types import InvertedResidual, InvertedResidualConfig  # noqa: F401
