"""Learn_Distortion (reference: model/distortions.py:4-27).  Tiny; stays torch (feeds the
kernels through device scalars).  The reference's `if scale<0.01` host sync is replaced by a
sync-free clamp with identical value and gradient semantics."""
import torch
import torch.nn as nn


class Learn_Distortion(nn.Module):
    def __init__(self, num_cams, learn_scale, learn_shift, cfg):
        super().__init__()
        self.global_scales = nn.Parameter(torch.ones(size=(num_cams, 1), dtype=torch.float32), requires_grad=learn_scale)
        self.global_shifts = nn.Parameter(torch.zeros(size=(num_cams, 1), dtype=torch.float32), requires_grad=learn_shift)
        self.fix_scaleN = cfg['distortion']['fix_scaleN']
        self.num_cams = num_cams

    def forward(self, cam_id):
        cam_id = int(cam_id)
        scale = self.global_scales[cam_id]
        # distortions.py:21-22: scale<0.01 -> constant 0.01 (no gradient); clamp has the same forward
        # value and the same zero/one gradient pattern without a device->host sync
        scale = torch.clamp(scale, min=0.01)
        if self.fix_scaleN and cam_id == (self.num_cams - 1):
            scale = torch.ones(1, device=self.global_scales.device)
        shift = self.global_shifts[cam_id]
        return scale, shift
