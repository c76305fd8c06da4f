"""Drop-in for the hot-path part of the reference's `model` package (model/__init__.py:1-10).
CheckpointIO (torch.save/load plumbing) is intentionally not re-implemented: state_dict keys and
shapes are identical, so the reference's model/checkpoints.py works on these modules unchanged."""
from .network import nope_nerf
from .training import Trainer
from .rendering import Renderer
from .config import get_model
from .official_nerf import OfficialStaticNerf
from .poses import LearnPose
from .intrinsics import LearnFocal
from .eval_pose_one_epoch import Trainer_pose
from .distortions import Learn_Distortion
from .losses import Loss, Loss_Eval
from .extracting_images import Extract_Images
from .eval_images import Eval_Images
from . import common
