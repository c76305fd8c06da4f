"""Geometry helpers of the hot path's callers (reference: model/common.py:13-39,623-630).
Only host-side conveniences live here; ray generation itself is in the CUDA library."""
import numpy as np
import torch


def arange_pixels(resolution=(128, 128), batch_size=1, image_range=(-1., 1.), device=torch.device("cpu")):
    """common.py:13-39: integer pixel locations (x,y) and their [-1,1]-scaled floats."""
    h, w = resolution
    ys, xs = torch.meshgrid(torch.arange(0, h, device=device), torch.arange(0, w, device=device), indexing="ij")
    pixel_locations = torch.stack([xs, ys], dim=-1).long().view(1, -1, 2).repeat(batch_size, 1, 1)
    pixel_scaled = pixel_locations.clone().float()
    scale = (image_range[1] - image_range[0]); loc = scale / 2
    pixel_scaled[:, :, 0] = scale * pixel_scaled[:, :, 0] / (w - 1) - loc
    pixel_scaled[:, :, 1] = scale * pixel_scaled[:, :, 1] / (h - 1) - loc
    return pixel_locations, pixel_scaled


def mse2psnr(mse):
    """common.py:623-630"""
    mse = np.maximum(mse, 1e-10)
    return (-10.0 * np.log10(mse)).astype(np.float32)


def nearest_prior_index(ray_idx, H, W, h_d, w_d):
    """index into the (h_d,w_d) DPT map that F.interpolate(...,'nearest')[ray_idx] reads
    (model/network.py:22-24; ATen nearest: floor(dst * float(in/out)))."""
    row = torch.div(ray_idx, W, rounding_mode="floor"); col = ray_idx - row * W
    sr = torch.clamp((row.float() * float(np.float32(h_d) / np.float32(H))).floor().long(), max=h_d - 1)
    sc = torch.clamp((col.float() * float(np.float32(w_d) / np.float32(W))).floor().long(), max=w_d - 1)
    return sr * w_d + sc
