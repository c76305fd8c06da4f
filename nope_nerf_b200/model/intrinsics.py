"""LearnFocal (reference: model/intrinsics.py:5-70): learnable fx, fy with order-1/2
parametrisation.  Six scalars; stays torch and feeds the kernels via camera_mat."""
import numpy as np
import torch
import torch.nn as nn


class LearnFocal(nn.Module):
    def __init__(self, req_grad, fx_only, order=2, init_focal=None):
        super().__init__()
        self.fx_only = fx_only
        self.order = order
        if order not in (1, 2):
            raise ValueError('Focal init order need to be 1 or 2.')

        def coef(v):
            v = float(v)
            return torch.tensor(np.sqrt(v) if order == 2 else v, requires_grad=False).float()
        if init_focal is None:
            fx0 = fy0 = torch.tensor(1.0, dtype=torch.float32)
        elif isinstance(init_focal, list):
            fx0, fy0 = coef(init_focal[0]), coef(init_focal[1])
        else:
            fx0 = fy0 = coef(init_focal)
        self.fx = nn.Parameter(fx0.clone(), requires_grad=req_grad)
        if not fx_only:
            self.fy = nn.Parameter(fy0.clone(), requires_grad=req_grad)

    def forward(self, i=None):
        fy = self.fx if self.fx_only else self.fy
        if self.order == 2:
            return torch.stack([self.fx ** 2, fy ** 2])
        return torch.stack([self.fx, fy])
