from planar_optical_flow_b200.eval_utils import *  # noqa: F401,F403
from planar_optical_flow_b200.eval_utils import (make_model_fn_obj_det, model_fn_obj_det, eval_dr_spaam, batch_cutouts, model_fn,  # noqa: F401
                                                 model_fn_eval, loss_fn_eval)
