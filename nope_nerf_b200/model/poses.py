"""LearnPose (reference: model/poses.py:6-33): per-view axis-angle r + translation t;
c2w = [Exp(r) t; 0 1] @ init_c2w[id], evaluated by nnb_pose_fwd/bwd."""
import torch
import torch.nn as nn
from .. import ops


class LearnPose(nn.Module):
    def __init__(self, num_cams, learn_R, learn_t, cfg, init_c2w=None):
        super().__init__()
        self.num_cams = num_cams
        self.init_c2w = None
        if init_c2w is not None:
            self.init_c2w = nn.Parameter(init_c2w, requires_grad=False)
        self.r = nn.Parameter(torch.zeros(size=(num_cams, 3), dtype=torch.float32), requires_grad=learn_R)
        self.t = nn.Parameter(torch.zeros(size=(num_cams, 3), dtype=torch.float32), requires_grad=learn_t)

    def forward(self, cam_id):
        cam_id = int(cam_id)
        init = None if self.init_c2w is None else self.init_c2w.detach().contiguous().float()
        return ops.pose_c2w(self.r, self.t, init, cam_id)

    def get_t(self):
        return self.t
