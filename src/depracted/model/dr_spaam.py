from planar_optical_flow_b200.model.dr_spaam import *  # noqa: F401,F403
from planar_optical_flow_b200.model.dr_spaam import DROW, SpatialDROW, _SpatialAttention  # noqa: F401
