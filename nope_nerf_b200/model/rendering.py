"""Renderer (reference: model/rendering.py:10-167).  forward() keeps the reference's
signature and output dict; the whole nope_nerf render (ray generation, sampling, encoding,
MLP, compositing) is one call into the CUDA library with autograd glue."""
import torch
import torch.nn as nn
from .. import ops


class Renderer(nn.Module):
    def __init__(self, model, cfg, device=None, **kwargs):
        super().__init__()
        self._device = device
        self.depth_range = cfg['depth_range']
        self.n_max_network_queries = cfg['n_max_network_queries']   # accepted, unused: nothing is chunked
        self.white_background = cfg['white_background']
        self.cfg = cfg
        self.model = model.to(device)
        self.engine = kwargs.get("engine", None)

    def to(self, device):
        model = super().to(device)
        model._device = device
        return model

    def forward(self, pixels, depth, camera_mat, world_mat, scale_mat, rendering_technique, add_noise=True, eval_=False,
                it=1000000):
        if rendering_technique == 'nope_nerf':
            return self.nope_nerf(pixels, depth, camera_mat, world_mat, scale_mat, it=it, add_noise=add_noise, eval_=eval_)
        if rendering_technique == 'phong_renderer':
            return self.phong_renderer(pixels, camera_mat, world_mat, scale_mat, it)
        raise ValueError(rendering_technique)

    def render_meta(self, N, eval_, extra=None):
        cfg = self.cfg
        ndc = cfg['sample_option'] == 'ndc'
        meta = dict(N=N, S=int(cfg['num_points']), near=0.0 if ndc else float(self.depth_range[0]),
                    far=1.0 if ndc else float(self.depth_range[1]),
                    flags=ops.flags_from_cfg(cfg, self.model.occ_activation, eval_=eval_),
                    engine=self.engine if self.engine is not None else ops.default_engine())
        if extra: meta.update(extra)
        return meta

    def nope_nerf(self, pixels, depth, camera_mat, world_mat, scale_mat, add_noise=False, it=100000, eval_=False,
                  c2w=None):
        """rendering.py:36-167.  camera_mat must be diag(kx,ky,-1,1) and scale_mat identity (the only
        forms the reference constructs: dataset.py:101-104,199; training.py:247-252)."""
        batch_size, n_points, _ = pixels.shape
        if batch_size != 1:
            raise NotImplementedError("batch_size must be 1 (configs/default.yaml:14)")
        if scale_mat is not None and not scale_mat.is_cuda:
            if not torch.equal(scale_mat.reshape(-1, 4, 4)[0], torch.eye(4)):
                raise NotImplementedError("scale_mat must be the identity")
        if c2w is None:
            c2w = torch.linalg.inv(world_mat.reshape(4, 4))       # common.py:139-141 (invert=True)
        cam = camera_mat.reshape(4, 4)
        S = int(self.cfg['num_points'])
        noise = None
        if add_noise and self.cfg['sample_option'] == 'uniform':
            noise = torch.rand(batch_size, n_points, S, device=pixels.device)[0]     # rendering.py:189
        meta = self.render_meta(n_points, eval_, dict(pixels=pixels.reshape(n_points, 2).contiguous().float(), noise=noise,
                                                      want_z_alpha=True))
        flat = self.model.flat_weights()
        d = depth.reshape(n_points)
        rgb, dp, dg, mask, z, alpha = ops.render_autograd(meta, flat, c2w, cam, d, None, None, list(self.model.parameters()))
        m = mask.bool()
        return {'rgb': rgb.reshape(1, n_points, 3), 'z_vals': z, 'normal': None,
                'depth_pred': dp[m], 'depth_gt': dg[m], 'alpha': alpha}

    # ---- geometry visualiser (reference rendering.py:198-418): evaluation-time only (render_visdata every
    #      `visualize_every` iterations).  Host logic mirrors the reference; every field evaluation (occupancy along the rays, secant
    #      refinement, normals = gradient(), surface colour) is a call into the CUDA library (nnb_field_fwd / nnb_field_bwd). ----
    @staticmethod
    def _rays_world(pixels, camera_mat, world_mat, scale_mat):
        """image_points_to_world / origin_to_world (common.py:186-237, invert=True): ray origins and unit directions"""
        b, n, _ = pixels.shape
        dev = pixels.device
        Ki = torch.inverse(camera_mat.to(dev)); Wi = torch.inverse(world_mat.to(dev)); Si = torch.inverse(scale_mat.to(dev))
        T = Si @ Wi @ Ki
        ph = torch.cat([pixels.permute(0, 2, 1), torch.ones(b, 2, n, device=dev)], dim=1)          # depth 1: (x, y, 1, 1)
        pw = (T @ ph)[:, :3].permute(0, 2, 1)
        o = torch.zeros(b, 4, n, device=dev); o[:, -1] = 1.
        cw = (T @ o)[:, :3].permute(0, 2, 1)
        d = pw - cw
        return cw, d / d.norm(2, 2).unsqueeze(-1)

    @staticmethod
    def _sphere_far(cam_loc, dirs, r):
        """far intersection depth of the rays with the sphere |x| = r (rendering.py:439-460), 0 where the ray misses"""
        dot = torch.bmm(dirs, cam_loc.unsqueeze(-1)).reshape(-1)
        under = dot ** 2 - (cam_loc.norm(2, 1) ** 2 - r ** 2).reshape(-1, 1).expand(-1, dirs.shape[1]).reshape(-1)
        far = torch.zeros_like(dot)
        m = under > 0
        far[m] = torch.sqrt(under[m]) - dot[m]
        return far.clamp_min(0.0).reshape(dirs.shape[0], dirs.shape[1])

    def phong_renderer(self, pixels, camera_mat, world_mat, scale_mat, it):
        batch_size, num_pixels, _ = pixels.shape
        dev = pixels.device
        rad = self.cfg['radius']
        camera_world, ray_vector = self._rays_world(pixels, camera_mat, world_mat, scale_mat)
        light_source = camera_world[0, 0]
        light = (light_source / light_source.norm(2)).unsqueeze(1)
        diffuse_per = torch.tensor([0.7, 0.7, 0.7], device=dev); ambiant = torch.tensor([0.3, 0.3, 0.3], device=dev)
        self.model.eval()
        with torch.no_grad():
            d_i = self.ray_marching(camera_world, ray_vector, self.model, n_secant_steps=8, n_steps=[512, 513], rad=rad)
            mask_zero_occupied = d_i == 0
            mask_pred = torch.isfinite(d_i)                                # common.get_mask
            dists = torch.ones_like(d_i)
            dists[mask_pred] = d_i[mask_pred]
            dists[mask_zero_occupied] = 0.
            network_object_mask = (mask_pred & ~mask_zero_occupied)[0]
            dists = dists[0]
            cw = camera_world.reshape(-1, 3); rv = ray_vector.reshape(-1, 3)
            points = cw + rv * dists.unsqueeze(-1)
            view_vol = -rv
            rgb_values = torch.ones_like(points)
            surface_points = points[network_object_mask]
            surface_view_vol = view_vol[network_object_mask]
        rgb_val = torch.zeros(batch_size * num_pixels, 3, device=dev)
        if surface_points.shape[0] > 0:
            grad = self.model.gradient(surface_points, it)[:, 0, :].detach()
            surface_normals = grad / grad.norm(2, 1, keepdim=True)
            diffuse = torch.mm(surface_normals, light).clamp_min(0).repeat(1, 3) * diffuse_per.unsqueeze(0)
            rgb_values[network_object_mask] = (ambiant.unsqueeze(0) + diffuse).clamp_max(1.0)
            with torch.no_grad():
                rgb_val[network_object_mask] = self.model(surface_points, surface_view_vol)
        return {'rgb': rgb_values.reshape(batch_size, -1, 3), 'normal': None, 'rgb_surf': rgb_val.reshape(batch_size, -1, 3)}

    def ray_marching(self, ray0, ray_direction, model, c=None, tau=0.5, n_steps=[128, 129], n_secant_steps=8, depth_range=[0., 2.4],
                     max_points=3500000, rad=1.0):
        """rendering.py:274-384: first sign change of (occupancy - tau) along each ray, refined by the secant method"""
        batch_size, n_pts, _ = ray0.shape
        dev = ray0.device
        tau = 0.5
        n_steps = int(torch.randint(n_steps[0], n_steps[1], (1,)).item())
        d_intersect = self._sphere_far(ray0[:, 0], ray_direction, rad)
        d_proposal = torch.linspace(0, 1, steps=n_steps, device=dev).view(1, 1, n_steps, 1)
        d_proposal = depth_range[0] * (1. - d_proposal) + d_intersect.view(1, -1, 1, 1) * d_proposal
        p_proposal = ray0.unsqueeze(2) + ray_direction.unsqueeze(2) * d_proposal
        with torch.no_grad():
            val = torch.cat([model(p_split, only_occupancy=True) - tau
                             for p_split in torch.split(p_proposal.reshape(batch_size, -1, 3), int(max_points / batch_size), dim=1)],
                            dim=1).view(batch_size, -1, n_steps)
        mask_0_not_occupied = val[:, :, 0] < 0
        sign_matrix = torch.cat([torch.sign(val[:, :, :-1] * val[:, :, 1:]), torch.ones(batch_size, n_pts, 1, device=dev)], dim=-1)
        cost_matrix = sign_matrix * torch.arange(n_steps, 0, -1, device=dev).float()
        values, indices = torch.min(cost_matrix, -1)
        mask_sign_change = values < 0
        bi = torch.arange(batch_size, device=dev).unsqueeze(-1); pi = torch.arange(n_pts, device=dev).unsqueeze(0)
        mask_neg_to_pos = val[bi, pi, indices] < 0
        mask = mask_sign_change & mask_neg_to_pos & mask_0_not_occupied
        n = batch_size * n_pts
        ar = torch.arange(n, device=dev)
        dp = d_proposal.expand(batch_size, n_pts, n_steps, 1).reshape(n, n_steps)
        vv = val.reshape(n, n_steps)
        d_low = dp[ar, indices.view(n)].view(batch_size, n_pts)[mask]
        f_low = vv[ar, indices.view(n)].view(batch_size, n_pts)[mask]
        indices = torch.clamp(indices + 1, max=n_steps - 1)
        d_high = dp[ar, indices.view(n)].view(batch_size, n_pts)[mask]
        f_high = vv[ar, indices.view(n)].view(batch_size, n_pts)[mask]
        ray0_masked = ray0[mask]; ray_direction_masked = ray_direction[mask]
        d_pred_out = torch.ones(batch_size, n_pts, device=dev)
        if ray0_masked.shape[0] != 0:
            d_pred_out[mask] = self.secant(f_low, f_high, d_low, d_high, n_secant_steps, ray0_masked, ray_direction_masked, tau)
        d_pred_out[mask == 0] = float('inf')
        d_pred_out[mask_0_not_occupied == 0] = 0
        return d_pred_out

    def secant(self, f_low, f_high, d_low, d_high, n_secant_steps, ray0_masked, ray_direction_masked, tau, it=0):
        """rendering.py:386-418"""
        d_pred = -f_low * (d_high - d_low) / (f_high - f_low) + d_low
        for _ in range(n_secant_steps):
            p_mid = ray0_masked + d_pred.unsqueeze(-1) * ray_direction_masked
            with torch.no_grad():
                f_mid = self.model(p_mid, batchwise=False, only_occupancy=True, it=it)[..., 0] - tau
            ind_low = f_mid < 0
            d_low = torch.where(ind_low, d_pred, d_low); f_low = torch.where(ind_low, f_mid, f_low)
            d_high = torch.where(ind_low, d_high, d_pred); f_high = torch.where(ind_low, f_high, f_mid)
            d_pred = -f_low * (d_high - d_low) / (f_high - f_low) + d_low
        return d_pred
