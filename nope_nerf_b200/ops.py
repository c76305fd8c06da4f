"""Torch-facing wrappers over the C ABI: tensors in, tensors out, everything enqueued on the
current CUDA stream.  torch is plumbing here (device memory, streams, autograd glue);
all arithmetic of the hot path happens in libnope_nerf_b200.so."""
import ctypes as C
import torch
from . import _lib as L

_DEFAULT_ENGINE = [L.ENGINE_TC]
_TC_BACKWARD = [True]     # tcgen05 backward (operand-plane stash); False -> fp32 SIMT backward after a TC forward
# weight-gradient operand planes of the tcgen05 backward: 'fp16' = one fp16 plane per operand, per-layer power-of-two dY scales
# from the previous step (NNB_WG16: half the stash traffic, a third of the MMAs, MLP weight gradients to ~3e-4);
# 'exact' = bf16 hi|lo planes, three MMAs per product (MLP weight gradients to fp32 round-off).  NNB_WGRAD overrides.
import os as _os
_WGRAD = [_os.environ.get("NNB_WGRAD", "fp16")]
# forward-precision EXPERIMENT of the tcgen05 engine (never the default): 0 = three-term split, 1 = without
# a_hi*b_lo (NNB_FWD_DROP_WLO), 2 = without a_lo*b_hi (NNB_FWD_DROP_ALO), 3 = a_hi*b_hi only
_FWD_DROP = [int(_os.environ.get("NNB_FWD_DROP", "0"))]


def set_forward_split_experiment(mode):
    if mode not in (0, 1, 2, 3):
        raise ValueError("forward split experiment mode must be 0..3")
    _FWD_DROP[0] = mode


def set_wgrad_precision(mode):
    if mode not in ("fp16", "exact"):
        raise ValueError("wgrad precision must be 'fp16' or 'exact'")
    _WGRAD[0] = mode


def wgrad_precision():
    return _WGRAD[0]



def set_tc_backward(on):
    _TC_BACKWARD[0] = bool(on)


def set_default_engine(name):
    _DEFAULT_ENGINE[0] = {"simt": L.ENGINE_SIMT, "tc": L.ENGINE_TC}[name]


def default_engine():
    return _DEFAULT_ENGINE[0]


def pick_engine(engine, S):
    """the tcgen05 engine tiles 128 samples per CTA and needs S in {32,64,128,256}; other sample counts
    run on the exact-fp32 SIMT engine (still CUDA; there is no CPU path)."""
    if engine is None:
        engine = _DEFAULT_ENGINE[0]
    if engine == L.ENGINE_TC and S not in (32, 64, 128, 256):
        return L.ENGINE_SIMT
    return engine


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(t, name, allow_pinned=False):
    """device tensors only; `allow_pinned` additionally admits page-locked HOST tensors, which the kernels read in
    place over PCIe through the unified address space (used for the frame / DPT map, of which a step touches N pixels)"""
    if t is None or t.is_cuda:
        return
    if allow_pinned and t.is_pinned():
        return
    raise RuntimeError("nope_nerf_b200: %s must live on a CUDA device (no CPU fallback exists)" % name)


def _f32c(t):
    if t is None:
        return None
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.contiguous().float()
    return t


def flags_from_cfg(cfg, occ_activation="softplus", eval_=False, shift_first=False):
    """cfg = the reference's cfg['rendering'] dict (configs/default.yaml:32-44)."""
    f = 0
    if cfg["dist_alpha"]: f |= L.DIST_ALPHA
    if cfg["sample_option"] == "ndc": f |= L.NDC
    elif cfg["sample_option"] != "uniform":
        raise ValueError("sample_option must be 'uniform' or 'ndc' (model/rendering.py:98-101)")
    if cfg["normalise_ray"]: f |= L.NORMALISE
    if cfg["use_ray_dir"]: f |= L.USE_DIR
    if cfg["white_background"]: f |= L.WHITE_BG
    if occ_activation == "softplus": f |= L.SOFTPLUS
    if eval_: f |= L.EVAL
    if shift_first: f |= L.SHIFT_FIRST
    if cfg.get("normal_loss", False):
        raise NotImplementedError("rendering.normal_loss=True is outside the hot path (SURVEY.md 8(f) rank 4)")
    if cfg.get("outside_steps", 0) != 0:
        raise NotImplementedError("rendering.outside_steps != 0 is not supported")
    return f


class _WorkspacePool:
    def __init__(self):
        self.free = {}

    def take(self, nbytes, device):
        key = (nbytes, device.index)
        lst = self.free.get(key)
        if lst:
            return lst.pop()
        return torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)

    def give(self, ws):
        self.free.setdefault((ws.numel(), ws.device.index), []).append(ws)


_pool = _WorkspacePool()


class RenderCall:
    """One forward launch + the state its backward needs (workspace with the stash)."""

    def __init__(self, weights, c2w, cam, *, N, S, flags, engine, near, far, ray_idx=None, pixels=None, depth=None,
                 depth_map=None, scale=None, shift=None, noise=None, H=0, W=0, want_z_alpha=False, stash=False, wgrad=None):
        for n_, t_ in (("weights", weights), ("c2w", c2w), ("camera_mat", cam), ("ray_idx", ray_idx), ("pixels", pixels),
                       ("depth", depth), ("noise", noise)):
            _need_cuda(t_, n_)
        _need_cuda(depth_map, "depth_map")
        dev = weights.device
        self.keep = [weights, c2w, cam, ray_idx, pixels, depth, depth_map, scale, shift, noise]
        a = L.RenderArgs()
        a.weights = L.ptr(weights); a.c2w = L.ptr(c2w); a.cam = L.ptr(cam)
        a.ray_idx = L.ptr(ray_idx); a.pixels = L.ptr(pixels); a.depth = L.ptr(depth); a.depth_map = L.ptr(depth_map)
        a.scale = L.ptr(scale); a.shift = L.ptr(shift); a.noise = L.ptr(noise)
        a.N, a.S, a.H, a.W = N, S, H, W
        if depth_map is not None:
            a.h_d, a.w_d = depth_map.shape[-2], depth_map.shape[-1]
        a.near_, a.far_ = float(near), float(far)
        engine = pick_engine(engine, S)
        if stash:
            flags |= L.STASH
            if engine == L.ENGINE_TC and _TC_BACKWARD[0]:
                flags |= L.TCBWD
                if (wgrad or _WGRAD[0]) == "fp16": flags |= L.WG16
        if engine == L.ENGINE_TC and _FWD_DROP[0]:
            flags |= (_FWD_DROP[0] & 1) * L.FWD_DROP_WLO | ((_FWD_DROP[0] >> 1) & 1) * L.FWD_DROP_ALO
        a.flags = flags; a.engine = engine
        self.rgb = torch.empty(N, 3, device=dev); self.depth_pred = torch.empty(N, device=dev)
        self.depth_gt = torch.empty(N, device=dev); self.mask = torch.empty(N, dtype=torch.uint8, device=dev)
        self.z_vals = torch.empty(N, S, device=dev) if want_z_alpha else None
        self.alpha = torch.empty(N, S, device=dev) if want_z_alpha else None
        a.rgb = L.ptr(self.rgb); a.depth_pred = L.ptr(self.depth_pred); a.depth_gt = L.ptr(self.depth_gt)
        a.mask = L.ptr(self.mask); a.z_vals = L.ptr(self.z_vals); a.alpha = L.ptr(self.alpha)
        nbytes = L.lib.nnb_workspace_bytes(N, S, flags, engine)
        self.ws = _pool.take(nbytes, dev)
        a.workspace = L.ptr(self.ws); a.workspace_bytes = self.ws.numel()
        self.args = a; self.N = N; self.S = S; self.stash = stash; self.pooled = True
        L.check(L.lib.nnb_render_fwd(C.byref(a), _stream()), "nnb_render_fwd")
        if not stash:
            self.release()

    def release(self):
        if self.ws is not None:
            if self.pooled: _pool.give(self.ws)
            self.ws = None

    def backward(self, g_rgb, g_depth_pred, g_depth_gt, g_weights, g_c2w, g_cam=None, g_depth=None, g_scale_shift=None, phase=0,
                 wg_state=None, wg_seed=False):
        """all outputs are accumulated into (caller-zeroed) buffers; g_weights may be None (pose only).  phase 1 / 2: the two halves
        of the tcgen05 backward (data gradients | weight gradients + ray adjoint), see include/nope_nerf_b200.h.  wg_state: the
        caller's persistent 32-float dY-scale state of the fp16 weight-gradient planes (None: stateless, seeded per call)."""
        if self.ws is None:
            raise RuntimeError("backward called without a stashed forward")
        b = L.RenderBwdArgs()
        b.phase = int(phase)
        b.fwd = self.args
        keep = [_f32c(g_rgb), _f32c(g_depth_pred), _f32c(g_depth_gt)]
        b.g_rgb, b.g_depth_pred, b.g_depth_gt = L.ptr(keep[0]), L.ptr(keep[1]), L.ptr(keep[2])
        b.g_weights = L.ptr(g_weights); b.g_c2w = L.ptr(g_c2w); b.g_cam = L.ptr(g_cam); b.g_depth = L.ptr(g_depth)
        b.g_scale_shift = L.ptr(g_scale_shift)
        b.wg_state = L.ptr(wg_state); b.wg_seed = int(bool(wg_seed))
        L.check(L.lib.nnb_render_bwd(C.byref(b), _stream()), "nnb_render_bwd")
        if phase != 1:
            self.release()


class _RenderFn(torch.autograd.Function):
    """autograd glue for the drop-in Renderer / nope_nerf modules.
    differentiable inputs: c2w (4,4), cam (4,4), depth (N,) | (scale, shift) and the 24 MLP parameters."""

    @staticmethod
    def forward(ctx, meta, flat, c2w, cam, depth, scale, shift, *params):
        # (grad mode is always off inside Function.forward: decide from the inputs' requires_grad)
        need_grad = any(ctx.needs_input_grad)
        call = RenderCall(flat, _f32c(c2w.detach()), _f32c(cam.detach()), stash=need_grad,
                          depth=None if depth is None else _f32c(depth.detach()),
                          scale=None if scale is None else _f32c(scale.detach()),
                          shift=None if shift is None else _f32c(shift.detach()), **meta)
        ctx.call = call; ctx.nparams = len(params)
        ctx.has = (depth is not None, scale is not None, shift is not None)
        ctx.params_need = any(p.requires_grad for p in params)
        ctx.mark_non_differentiable(call.mask)
        outs = (call.rgb, call.depth_pred, call.depth_gt, call.mask)
        if call.z_vals is not None:
            ctx.mark_non_differentiable(call.z_vals, call.alpha)
            outs += (call.z_vals, call.alpha)
        return outs

    @staticmethod
    def backward(ctx, g_rgb, g_dp, g_dg, *unused):
        call = ctx.call
        dev = call.rgb.device
        if g_rgb is None: g_rgb = torch.zeros(call.N, 3, device=dev)
        g_c2w = torch.zeros(4, 4, device=dev); g_cam = torch.zeros(4, 4, device=dev)
        g_w = torch.zeros(L.NUM_PARAMS, device=dev) if ctx.params_need else None
        has_depth, has_scale, has_shift = ctx.has
        g_depth = torch.empty(call.N, device=dev) if has_depth else None
        g_ss = torch.zeros(2, device=dev) if (has_scale or has_shift) else None
        call.backward(g_rgb, g_dp, g_dg, g_w, g_c2w, g_cam, g_depth, g_ss)
        gp = [None] * ctx.nparams
        if g_w is not None:
            from .model.official_nerf import PARAM_SLICES
            gp = [g_w[o:o + n].view(shape) for (o, n, shape) in PARAM_SLICES]
        return (None, None, g_c2w, g_cam, g_depth, g_ss[0] if has_scale else None, g_ss[1] if has_shift else None) + tuple(gp)


def render_autograd(meta, flat, c2w, cam, depth, scale, shift, params):
    return _RenderFn.apply(meta, flat, c2w, cam, depth, scale, shift, *params)


class _PoseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, r, t, init, cam_id):
        c2w = torch.empty(4, 4, device=r.device)
        L.check(L.lib.nnb_pose_fwd(L.ptr(r), L.ptr(t), L.ptr(init), cam_id, L.ptr(c2w), _stream()), "nnb_pose_fwd")
        ctx.save_for_backward(r, t, init if init is not None else r.new_empty(0))
        ctx.cam_id = cam_id; ctx.has_init = init is not None
        return c2w

    @staticmethod
    def backward(ctx, g):
        r, t, init = ctx.saved_tensors
        g_r = torch.zeros_like(r) if ctx.needs_input_grad[0] else None
        g_t = torch.zeros_like(t) if ctx.needs_input_grad[1] else None
        L.check(L.lib.nnb_pose_bwd(L.ptr(r), L.ptr(t), L.ptr(init) if ctx.has_init else None, ctx.cam_id, L.ptr(_f32c(g)),
                                   L.ptr(g_r), L.ptr(g_t), _stream()), "nnb_pose_bwd")
        return g_r, g_t, None, None


def pose_c2w(r, t, init, cam_id):
    _need_cuda(r, "LearnPose.r")
    return _PoseFn.apply(r, t, init, int(cam_id))


def pose_fwd_raw(r, t, init, cam_id, out):
    L.check(L.lib.nnb_pose_fwd(L.ptr(r), L.ptr(t), L.ptr(init), int(cam_id), L.ptr(out), _stream()), "nnb_pose_fwd")


def pose_bwd_raw(r, t, init, cam_id, g_c2w, g_r, g_t):
    L.check(L.lib.nnb_pose_bwd(L.ptr(r), L.ptr(t), L.ptr(init), int(cam_id), L.ptr(g_c2w), L.ptr(g_r), L.ptr(g_t), _stream()),
            "nnb_pose_bwd")


def loss_rgb_depth(rgb, depth_pred, depth_gt, mask, w_rgb, w_depth, rgb_l2, *, rgb_gt=None, img=None, ray_idx=None,
                   grad_scale=1.0):
    """returns (losses[4] = loss, loss_rgb, loss_depth, l2_mean ; g_rgb, g_depth_pred, g_depth_gt)"""
    N = rgb.shape[0]; dev = rgb.device
    _need_cuda(img, "img", allow_pinned=True); _need_cuda(rgb_gt, "rgb_gt")
    out = torch.empty(4, device=dev)
    g_rgb = torch.empty(N, 3, device=dev); g_dp = torch.empty(N, device=dev); g_dg = torch.empty(N, device=dev)
    HW = 0 if img is None else (img.shape[-1] if img.dim() == 2 else img.shape[-1] * img.shape[-2])
    L.check(L.lib.nnb_loss_rgb_depth(L.ptr(rgb), L.ptr(rgb_gt), L.ptr(img), L.ptr(ray_idx), HW, L.ptr(depth_pred), L.ptr(depth_gt),
                                     L.ptr(mask), N, float(w_rgb), float(w_depth), int(bool(rgb_l2)), float(grad_scale), L.ptr(out),
                                     L.ptr(g_rgb), L.ptr(g_dp), L.ptr(g_dg), _stream()), "nnb_loss_rgb_depth")
    return out, g_rgb, g_dp, g_dg


class _ChamferFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, Y):
        X = _f32c(X); Y = _f32c(Y)
        P, Q = X.shape[0], Y.shape[0]
        keys = torch.empty(P + Q, dtype=torch.int64, device=X.device)
        loss = torch.zeros(1, device=X.device)
        need = X.requires_grad or Y.requires_grad
        gX = torch.zeros_like(X) if need else None; gY = torch.zeros_like(Y) if need else None
        L.check(L.lib.nnb_chamfer(L.ptr(X), P, L.ptr(Y), Q, L.ptr(keys), None, None, L.ptr(loss), 1.0, L.ptr(gX), L.ptr(gY),
                                  _stream()), "nnb_chamfer")
        ctx.save_for_backward(gX if need else X.new_empty(0), gY if need else X.new_empty(0))
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        gX, gY = ctx.saved_tensors
        return (gX * g if ctx.needs_input_grad[0] else None), (gY * g if ctx.needs_input_grad[1] else None)


def chamfer(X, Y):
    """Loss.get_pc_loss (dense) for X (P,3), Y (Q,3) on the GPU."""
    _need_cuda(X, "X")
    return _ChamferFn.apply(X, Y)


_rs_ws = {}


def refstage_raw(c2w_cur, dist_cur, c2w_ref, dist_ref, dpt_cur, dpt_ref, *, H, W, img_cur=None, img_ref=None, img_pp=None, is_last=False,
                 kx=0.0, ky=0.0, cam=None, cam_idx_dev=None, num_cams=0, weights_dev=None, nearest_limit=0.01, pc_ratio=4, scale_pcs=True,
                 detach_rgbs_scale=False, shift_first=False, w_pc=1.0, w_rgb_s=1.0, losses=None, g_c2w=None, g_dist=None, g_kxy=None,
                 loss_total=None, grad_scale=1.0, workspace=None):
    """nnb_refstage: the reference-image stage of Trainer.compute_loss (model/training.py:280-365), forward + adjoint in one call.
    All tensor arguments are device tensors (frames may be page-locked host tensors); g_* / loss_total ACCUMULATE."""
    _need_cuda(c2w_cur, "c2w_cur")
    for n_, t_ in (("img_cur", img_cur), ("img_ref", img_ref)):
        _need_cuda(t_, n_, allow_pinned=True)
    dev = c2w_cur.device
    hd, wd = int(dpt_cur.shape[-2]), int(dpt_cur.shape[-1])
    nbytes = L.lib.nnb_refstage_workspace_bytes(hd, wd, int(pc_ratio))
    if nbytes == 0:
        raise ValueError("reference-image stage: DPT map %dx%d too small for pc_ratio %d" % (hd, wd, pc_ratio))
    if workspace is None:
        key = (nbytes, dev.index)
        workspace = _rs_ws.get(key)
        if workspace is None:
            workspace = _rs_ws[key] = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    if losses is None:
        losses = torch.zeros(2, device=dev)
    a = L.RefStageArgs()
    a.img_cur, a.img_ref, a.dpt_cur, a.dpt_ref = L.ptr(img_cur), L.ptr(img_ref), L.ptr(dpt_cur), L.ptr(dpt_ref)
    a.c2w_cur, a.c2w_ref, a.dist_cur, a.dist_ref = L.ptr(c2w_cur), L.ptr(c2w_ref), L.ptr(dist_cur), L.ptr(dist_ref)
    a.H, a.W, a.h_d, a.w_d, a.pc_ratio, a.is_last = int(H), int(W), hd, wd, int(pc_ratio), int(bool(is_last))
    a.flags = (1 if scale_pcs else 0) | (2 if detach_rgbs_scale else 0) | (4 if shift_first else 0)
    a.kx, a.ky, a.nearest_limit, a.w_pc, a.w_rgb_s = float(kx), float(ky), float(nearest_limit), float(w_pc), float(w_rgb_s)
    a.losses, a.g_c2w, a.g_dist = L.ptr(losses), L.ptr(g_c2w), L.ptr(g_dist)
    a.workspace, a.workspace_bytes = L.ptr(workspace), workspace.numel()
    a.img_pp, a.cam, a.cam_idx_dev, a.num_cams, a.weights_dev = L.ptr(img_pp), L.ptr(cam), L.ptr(cam_idx_dev), int(num_cams), L.ptr(weights_dev)
    a.g_kxy, a.loss_total, a.grad_scale = L.ptr(g_kxy), L.ptr(loss_total), float(grad_scale)
    L.check(L.lib.nnb_refstage(C.byref(a), _stream()), "nnb_refstage")
    return losses


class _RefStageFn(torch.autograd.Function):
    """w_pc * loss_pc + w_rgb_s * loss_rgb_s of the reference-image stage (model/training.py:280-365) as a function of the current
    view's pose matrix, effective distortion {scale, shift} and (kx, ky); forward and adjoint come out of ONE library call."""

    @staticmethod
    def forward(ctx, c2w_cur, dist_cur, kxy, c2w_ref, dist_ref, img_cur, img_ref, dpt_cur, dpt_ref, is_last, nearest_limit, pc_ratio,
                scale_pcs, detach_rgbs_scale, shift_first, w_pc, w_rgb_s):
        dev = c2w_cur.device
        f = lambda t: _f32c(t.detach())
        H, W = img_cur.shape[-2:]
        g_c2w = torch.zeros(4, 4, device=dev); g_dist = torch.zeros(2, device=dev); g_kxy = torch.zeros(2, device=dev)
        kx, ky = (float(v) for v in kxy.detach().cpu())
        losses = refstage_raw(f(c2w_cur), f(dist_cur), f(c2w_ref), f(dist_ref), _f32c(dpt_cur), _f32c(dpt_ref), H=H, W=W, img_cur=_f32c(img_cur),
                              img_ref=_f32c(img_ref), is_last=is_last, kx=kx, ky=ky, nearest_limit=nearest_limit, pc_ratio=pc_ratio,
                              scale_pcs=scale_pcs, detach_rgbs_scale=detach_rgbs_scale, shift_first=shift_first, w_pc=w_pc, w_rgb_s=w_rgb_s,
                              g_c2w=g_c2w, g_dist=g_dist, g_kxy=g_kxy, workspace=torch.empty(
                                  L.lib.nnb_refstage_workspace_bytes(int(dpt_cur.shape[-2]), int(dpt_cur.shape[-1]), int(pc_ratio)) or 1,
                                  dtype=torch.uint8, device=dev))
        ctx.save_for_backward(g_c2w, g_dist, g_kxy)
        ctx.mark_non_differentiable(losses)
        return float(w_pc) * losses[0] + float(w_rgb_s) * losses[1], losses

    @staticmethod
    def backward(ctx, g, _g_losses):
        g_c2w, g_dist, g_kxy = ctx.saved_tensors
        return (g_c2w * g if ctx.needs_input_grad[0] else None, g_dist * g if ctx.needs_input_grad[1] else None,
                g_kxy * g if ctx.needs_input_grad[2] else None) + (None,) * 14


def refstage(c2w_cur, dist_cur, c2w_ref, dist_ref, img_cur, img_ref, dpt_cur, dpt_ref, is_last, kx, ky, nearest_limit=0.01, pc_ratio=4,
             scale_pcs=True, detach_rgbs_scale=False, shift_first=False, w_pc=1.0, w_rgb_s=1.0):
    """(total, losses[2] = {loss_pc, loss_rgb_s}); total is differentiable w.r.t. c2w_cur (4,4), dist_cur = [scale_eff, shift] and, when
    kx / ky are tensors, the intrinsics (autograd wrapper over refstage_raw; the Trainer calls refstage_raw directly)."""
    _need_cuda(c2w_cur, "c2w_cur")
    if isinstance(kx, torch.Tensor) or isinstance(ky, torch.Tensor):
        kxy = torch.stack([torch.as_tensor(kx, device=c2w_cur.device).reshape(()), torch.as_tensor(ky, device=c2w_cur.device).reshape(())])
    else:
        kxy = torch.tensor([float(kx), float(ky)], device=c2w_cur.device)
    return _RefStageFn.apply(c2w_cur, dist_cur, kxy, c2w_ref, dist_ref, img_cur, img_ref, dpt_cur, dpt_ref, is_last, nearest_limit, pc_ratio,
                             scale_pcs, detach_rgbs_scale, shift_first, w_pc, w_rgb_s)


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    L.check(L.lib.nnb_adam_step(L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), p.numel(), int(step), float(lr), beta1, beta2, eps,
                                _stream()), "nnb_adam_step")


# ---- CUDA-graph friendly variants (device-resident per-step scalars) -----------------------------------------
def pose_fwd_dev(r, t, init, cam_idx_dev, out):
    L.check(L.lib.nnb_pose_fwd_dev(L.ptr(r), L.ptr(t), L.ptr(init), L.ptr(cam_idx_dev), L.ptr(out), _stream()), "nnb_pose_fwd_dev")


def pose_bwd_dev(r, t, init, cam_idx_dev, g_c2w, g_r, g_t):
    L.check(L.lib.nnb_pose_bwd_dev(L.ptr(r), L.ptr(t), L.ptr(init), L.ptr(cam_idx_dev), L.ptr(g_c2w), L.ptr(g_r), L.ptr(g_t), _stream()),
            "nnb_pose_bwd_dev")


def distortion_fwd_dev(scales, shifts, cam_idx_dev, fix_last, out2):
    L.check(L.lib.nnb_distortion_fwd_dev(L.ptr(scales), L.ptr(shifts), scales.shape[0], L.ptr(cam_idx_dev), int(bool(fix_last)), L.ptr(out2),
                                         _stream()), "nnb_distortion_fwd_dev")


def distortion_bwd_dev(scales, cam_idx_dev, fix_last, g_ss, g_scales, g_shifts):
    L.check(L.lib.nnb_distortion_bwd_dev(L.ptr(scales), scales.shape[0], L.ptr(cam_idx_dev), int(bool(fix_last)), L.ptr(g_ss), L.ptr(g_scales),
                                         L.ptr(g_shifts), _stream()), "nnb_distortion_bwd_dev")


def adam_step_dev(p, g, m, v, step_dev, lr_dev, beta1=0.9, beta2=0.999, eps=1e-8):
    L.check(L.lib.nnb_adam_step_dev(L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), p.numel(), L.ptr(step_dev), L.ptr(lr_dev), beta1, beta2, eps,
                                    _stream()), "nnb_adam_step_dev")


def counter_incr(counters):
    L.check(L.lib.nnb_counter_incr(L.ptr(counters), counters.numel(), _stream()), "nnb_counter_incr")


def loss_rgb_depth_indirect(rgb, depth_pred, depth_gt, mask, w_rgb, w_depth, rgb_l2, img_pp, ray_idx, HW, out, g_rgb, g_dp, g_dg, grad_scale=1.0,
                            w_dev=None):
    L.check(L.lib.nnb_loss_rgb_depth_indirect(L.ptr(rgb), L.ptr(img_pp), L.ptr(ray_idx), int(HW), L.ptr(depth_pred), L.ptr(depth_gt), L.ptr(mask),
                                              rgb.shape[0], float(w_rgb), float(w_depth), int(bool(rgb_l2)), float(grad_scale), L.ptr(out),
                                              L.ptr(g_rgb), L.ptr(g_dp), L.ptr(g_dg), L.ptr(w_dev), _stream()), "nnb_loss_rgb_depth_indirect")


def sample_pixels(hw, n, device):
    """n distinct pixel ids uniform in [0, hw) (distribution of torch.randperm(hw)[:n])"""
    u = torch.rand(2 * n, device=device)
    out = torch.empty(n, dtype=torch.int64, device=device)
    L.check(L.lib.nnb_sample_pixels(L.ptr(u), int(hw), int(n), L.ptr(out), _stream()), "nnb_sample_pixels")
    return out
