from planar_optical_flow_b200 import utils, train_utils, eval_utils, dataset_dr_spaam  # noqa: F401
