"""Loss / Loss_Eval (reference: model/losses.py).  Same forward() signature and return keys.
The dense chamfer term runs in nnb_chamfer (no (3,P,Q) tensor is materialised); the
photometric / depth terms of the training step are fused into nnb_loss_rgb_depth by the
Trainer — the torch expressions below serve direct callers of this class."""
import torch
from torch import nn
from torch.nn import functional as F
from .. import ops


class Loss_Eval(nn.Module):
    def forward(self, rgb_pred, rgb_gt):
        return {'loss': F.mse_loss(rgb_pred, rgb_gt)}


class Loss(nn.Module):
    def __init__(self, cfg=None):
        super().__init__()
        self.depth_loss_type = cfg['depth_loss_type']
        self.cfg = cfg          # 'with_auto_mask' only feeds get_reprojection_loss / get_DPT_reprojection_loss in the reference
                                # (losses.py:65-102), which nothing calls: like there, it has no effect on forward()

    def get_rgb_full_loss(self, rgb_values, rgb_gt, rgb_loss_type='l2'):     # losses.py:27-32
        d = rgb_values - rgb_gt
        s = d.abs().sum() if rgb_loss_type == 'l1' else (d * d).sum()
        return s / float(rgb_values.shape[1])

    def depth_loss_dpt(self, pred_depth, gt_depth, weight=None):              # losses.py:34-57
        t_pred = torch.median(pred_depth); s_pred = torch.mean(torch.abs(pred_depth - t_pred))
        t_gt = torch.median(gt_depth); s_gt = torch.mean(torch.abs(gt_depth - t_gt))
        return F.mse_loss((pred_depth - t_pred) / s_pred, (gt_depth - t_gt) / s_gt)

    def get_depth_loss(self, depth_pred, depth_gt):                           # losses.py:59-64
        if self.depth_loss_type == 'l1':
            return (depth_pred - depth_gt).abs().sum() / float(depth_pred.shape[0])
        if self.depth_loss_type == 'invariant':
            return self.depth_loss_dpt(depth_pred, depth_gt)
        raise ValueError(self.depth_loss_type)

    def mean_on_mask(self, diff, valid_mask):                                 # losses.py:77-85 (sync-free)
        mask = valid_mask.expand_as(diff)
        cnt = mask.sum()
        return torch.where(cnt > 0, (diff * mask).sum() / cnt.clamp(min=1), torch.zeros((), device=diff.device))

    def get_weight_dist_loss(self, t_list):                                   # losses.py:103-112
        dist = (t_list - t_list.roll(shifts=1, dims=0))[1:].norm(dim=1)
        dist_diff = (dist - dist.roll(shifts=1))[1:]
        return dist.mean(), dist_diff.pow(2.0).mean()

    def get_pc_loss(self, Xt, Yt):                                            # losses.py:114-148
        if self.cfg['match_method'] != 'dense':
            raise NotImplementedError(self.cfg['match_method'])
        return ops.chamfer(Xt[0], Yt[0])

    @staticmethod
    def ssim_loss_map(x, y):                                                  # losses.py:222-252 (class SSIM)
        """(1 - SSIM) / 2 per pixel and channel of two (B,C,H,W) images: 3x3 mean windows over the reflection-padded images,
        C1 = 0.01^2, C2 = 0.03^2, clamped to [0, 1]"""
        c1, c2 = 0.01 ** 2, 0.03 ** 2
        box = lambda t: F.avg_pool2d(F.pad(t, (1, 1, 1, 1), mode='reflect'), 3, 1)
        mx, my = box(x), box(y)
        vx, vy, cxy = box(x * x) - mx * mx, box(y * y) - my * my, box(x * y) - mx * my
        s = ((2 * mx * my + c1) * (2 * cxy + c2)) / ((mx * mx + my * my + c1) * (vx + vy + c2))
        return ((1 - s) / 2).clamp(0, 1)

    def get_rgb_s_loss(self, rgb1, rgb2, valid_points):                       # losses.py:150-157
        diff_img = (rgb1 - rgb2).abs().clamp(0, 1)
        if self.cfg.get('with_ssim', False):
            diff_img = 0.15 * diff_img + 0.85 * self.ssim_loss_map(rgb1, rgb2)
        return self.mean_on_mask(diff_img, valid_points)

    def get_depth_consistency_loss(self, d1_proj, d2, d2_proj=None, d1=None):
        loss = (d1_proj - d2).abs().sum() / float(d1_proj.shape[1])
        if d2_proj is not None:
            loss = 0.5 * loss + 0.5 * (d2_proj - d1).abs().sum() / float(d2_proj.shape[1])
        return loss

    def forward(self, rgb_pred, rgb_gt, depth_pred=None, depth_gt=None, t_list=None, X=None, Y=None, rgb_pc1=None,
                rgb_pc1_proj=None, valid_points=None, d1_proj=None, d2=None, d2_proj=None, d1=None, weights={},
                rgb_loss_type='l2', **kwargs):
        dev = rgb_gt.device
        zero = lambda: torch.zeros((), device=dev)
        w = weights
        rgb_full_loss = self.get_rgb_full_loss(rgb_pred, rgb_gt, rgb_loss_type) if w['rgb_weight'] != 0.0 else zero()
        depth_loss = self.get_depth_loss(depth_pred, depth_gt) if w['depth_weight'] != 0.0 else zero()
        if w['weight_dist_2nd_loss'] != 0.0 or w['weight_dist_1st_loss'] != 0.0:
            loss_dist_1st, loss_dist_2nd = self.get_weight_dist_loss(t_list)
        else:
            loss_dist_1st, loss_dist_2nd = zero(), zero()
        pc_loss = self.get_pc_loss(X, Y) if w['pc_weight'] != 0.0 else zero()
        rgb_s_loss = self.get_rgb_s_loss(rgb_pc1, rgb_pc1_proj, valid_points) if w['rgb_s_weight'] != 0.0 else zero()
        dc_loss = self.get_depth_consistency_loss(d1_proj, d2, d2_proj, d1) if w['depth_consistency_weight'] != 0.0 else zero()
        l2_mean = F.mse_loss(rgb_pred, rgb_gt) if (w['rgb_weight'] != 0.0 or w['depth_weight'] != 0.0) else zero()
        loss = (w['rgb_weight'] * rgb_full_loss + w['depth_weight'] * depth_loss + w['weight_dist_1st_loss'] * loss_dist_1st +
                w['weight_dist_2nd_loss'] * loss_dist_2nd + w['pc_weight'] * pc_loss + w['rgb_s_weight'] * rgb_s_loss +
                w['depth_consistency_weight'] * dc_loss)
        return {'loss': loss, 'loss_rgb': rgb_full_loss, 'loss_depth': depth_loss, 'l2_mean': l2_mean,
                'loss_dist_1st': loss_dist_1st, 'loss_dist_2nd': loss_dist_2nd, 'loss_pc': pc_loss,
                'loss_rgb_s': rgb_s_loss, 'loss_depth_consistency': dc_loss}
