from planar_optical_flow_b200.utils import *  # noqa: F401,F403
from planar_optical_flow_b200.utils import (scans_to_cutout, scans_to_cutout_torch, nms_predicted_center, get_laser_phi,  # noqa: F401
                                            rphi_to_xy, canonical_to_global, global_to_canonical, scans_to_cutout_batch,
                                            scans_to_cutout_original, scans_to_polar_grid)
