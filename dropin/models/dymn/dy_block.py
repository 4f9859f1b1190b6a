from efficientat_b200.models.dymn.dy_block import *  # noqa: F401,F403
from efficientat_b200.models.dymn.dy_block import DY_Block, DynamicConv, DynamicInvertedResidualConfig  # noqa: F401
