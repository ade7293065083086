from planar_optical_flow_b200.train_utils import *  # noqa: F401,F403
from planar_optical_flow_b200.train_utils import Trainer, load_checkpoint, save_checkpoint, checkpoint_state, lr_scheduler, create_tb_logger, LucasScheduler  # noqa: F401
