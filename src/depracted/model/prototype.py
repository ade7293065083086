from planar_optical_flow_b200.model.prototype import *  # noqa: F401,F403
from planar_optical_flow_b200.model.prototype import Prototype, flow_loss  # noqa: F401
