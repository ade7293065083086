from planar_optical_flow_b200.dataset_dr_spaam import *  # noqa: F401,F403
from planar_optical_flow_b200.dataset_dr_spaam import create_dataloader, create_test_dataloader, SyntheticDROWDataset  # noqa: F401
