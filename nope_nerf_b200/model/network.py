"""nope_nerf wrapper module (reference: model/network.py:7-33): gathers the DPT prior at the
sampled pixels and calls the renderer.  The reference resizes the WHOLE depth map to (H,W) to
read N values; here only the N nearest-neighbour source indices are computed."""
import torch
import torch.nn as nn
from .common import nearest_prior_index


class nope_nerf(nn.Module):
    def __init__(self, cfg, renderer, depth_estimator=None, device=None, **kwargs):
        super().__init__()
        self.renderer = renderer.to(device)
        self.depth_estimator = depth_estimator.to(device) if depth_estimator is not None else None
        self.device = device

    def forward(self, p, ray_idx, camera_mat, world_mat, scale_mat, rendering_technique, it=0, eval_mode=False,
                depth_img=None, add_noise=True, img_size=None):
        if rendering_technique == 'nope_nerf':
            H, W = img_size
            h_d, w_d = depth_img.shape[-2:]
            src = nearest_prior_index(ray_idx.reshape(-1).to(depth_img.device), H, W, h_d, w_d)
            depth = depth_img.reshape(-1)[src].reshape(1, -1, 1)          # network.py:22-24
        else:
            depth = None
        return self.renderer(p, depth, camera_mat, world_mat, scale_mat, rendering_technique, eval_=eval_mode, it=it,
                             add_noise=add_noise)
