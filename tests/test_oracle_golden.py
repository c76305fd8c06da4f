"""CPU: pin the numpy oracle against fixtures produced by the unmodified reference
(tools/make_golden.py).  Tolerances: both sides are fp32 evaluations of the same
formulae in different operation orders -> 2e-5 relative-to-max on outputs, 2e-4 on
gradients (fp32 autograd vs fp32 hand adjoint; the fp64 oracle is checked tighter)."""
import numpy as np
import pytest
from oracle import nerf_oracle as O
from _util import load_golden, relmax, cfg_from_golden, check_param_digest

RENDER_CASES = ["render_tanks_noise", "render_tanks_r0", "render_eval_ones", "render_ndc_distalpha", "render_oddflags"]


@pytest.mark.parametrize("name", ["mlp_alpha_softplus", "mlp_sigma_relu"])
def test_mlp(name):
    g = load_golden(name)
    P = O.init_params(seed=int(g["seed"]))
    assert relmax(O.encode_position(g["pts"], 10), g["enc"]) < 1e-6
    rgb, a, cache = O.mlp_forward(P, g["pts"], g["dirs"], dist_alpha=bool(g["dist_alpha"]), occ_activation=str(g["occ"]))
    assert relmax(rgb, g["rgb"]) < 2e-6
    assert relmax(a, g["a"]) < 2e-5
    G, gp, gd = O.mlp_backward(P, cache, g["g_rgb"], g["g_a"])
    assert relmax(gp, g["g_pts"]) < 1e-4
    assert relmax(gd, g["g_dirs"]) < 1e-4
    assert check_param_digest(g, G) < 1e-4


def test_pose_expmap():
    g = load_golden("pose_expmap")
    V = g["r"].shape[0]
    for use_init in (0, 1):
        for v in range(V):
            init = g["init"][v] if use_init else None
            c = O.make_c2w(g["r"][v], g["t"][v], init)
            assert relmax(c, g["c2w_%d" % use_init][v]) < 1e-6
            gr, gt = O.make_c2w_bwd(g["r"][v], g["t"][v], init, g["G"][v])
            assert relmax(gr, g["gr_%d" % use_init][v]) < 1e-5, (use_init, v)
            assert relmax(gt, g["gt_%d" % use_init][v]) < 1e-5


def run_render(g, dtype, gates_from=None):
    """gates_from: mlp cache of another pass whose ReLU gates (sign of the pre-activations) replace this pass's in the backward"""
    cfg = cfg_from_golden(g)
    P = {k: v.astype(dtype) for k, v in O.init_params(seed=int(g["seed"]), white_bkgd=cfg["white_background"],
                                                      hf_damp=bool(g.get("hf_damp", False))).items()}
    H, W = int(g["H"]), int(g["W"])
    cam = int(g["cam_id"])
    init = g["init_c2w"][cam] if "init_c2w" in g else None
    r, t = g["r"][cam].astype(dtype), g["t"][cam].astype(dtype)
    c2w = O.make_c2w(r, t, init)
    raw, _ = O.gather_prior_depth(g["dpt"], g["ray_idx"], H, W)
    depth = (raw.astype(dtype) * dtype(g["scale"]) + dtype(g["shift"])).astype(dtype)
    pix = O.pixels_from_idx(g["ray_idx"], H, W, dtype)
    noise = g["noise"].astype(dtype) if "noise" in g else None
    out, cache = O.render_forward(P, pix, depth, c2w, dtype(g["kx"]), dtype(g["ky"]), cfg, noise=noise,
                                  eval_=bool(g["eval_mode"]))
    if dtype == np.float32:
        run_render.last_cache32 = cache["mcache"]
    if gates_from is not None:
        mc = cache["mcache"]
        tiny = np.finfo(np.float64).tiny
        for li in range(8):
            mc["Y"][li] = np.where(gates_from["Y"][li] > 0, np.abs(mc["Y"][li]) + tiny, -np.abs(mc["Y"][li]) - tiny)
        mc["yr"] = np.where(gates_from["yr"] > 0, np.abs(mc["yr"]) + tiny, -np.abs(mc["yr"]) - tiny)
    N = int(g["N"])
    m = out["mask"]
    gdp = np.zeros(N, dtype); gdg = np.zeros(N, dtype)
    gdp[m] = g["g_dp"]; gdg[m] = g["g_dg"]
    gr = O.render_backward(P, cache, g["g_rgb"].astype(dtype), gdp, gdg)
    g_r, g_t = O.make_c2w_bwd(r, t, init, gr["c2w"])
    return out, gr, g_r, g_t, raw, pix


@pytest.mark.parametrize("name", RENDER_CASES)
def test_render(name):
    """fp32 oracle vs the reference's fp32 autograd.  Gradients through sin(2^9 p) are
    sensitive to 1-ulp differences in p, so the reference's own fp32 result sits up to a few
    1e-3 from the fp64 truth on adversarial cotangents; the gate is therefore
    max(2e-4, 3 x |reference - fp64 oracle|) (the envelope of the reference's own rounding)."""
    g = load_golden(name)
    cam = int(g["cam_id"])
    out, gr, g_r, g_t, raw, pix = run_render(g, np.float32)
    out64, gr64, g_r64, g_t64, _, _ = run_render(g, np.float64)
    assert relmax(pix, g["pixels"]) < 1e-6
    assert relmax(out["z_vals"], g["z_vals"]) < 1e-6
    for o in (out, out64):
        assert relmax(o["rgb"], g["rgb"]) < 2e-5
        assert relmax(o["depth_pred"], g["depth_pred"]) < 2e-5
        assert relmax(o["depth_gt"], g["depth_gt"]) < 2e-5
        assert relmax(o["alpha"], g["alpha"]) < 5e-5

    # occ_activation='relu' makes the density gradient discontinuous at s = 0: a sample whose
    # logit sits within rounding of 0 flips its mask between two fp32 evaluation orders.
    floor = 2e-2 if cfg_from_golden(g)["occ_activation"] == "relu" else 2e-4

    def gate(a32, a64, ref):
        env = relmax(a64, ref)
        assert env < 2e-2, env
        assert relmax(a32, ref) < max(floor, 3 * env), (relmax(a32, ref), env)
    gate(gr["c2w"], gr64["c2w"], g["grad_c2w"])
    gate(g_r, g_r64, g["grad_r"][cam])
    gate(g_t, g_t64, g["grad_t"][cam])
    gate(np.array([(gr["depth"] * raw).sum(), gr["depth"].sum()]),
         np.array([(gr64["depth"] * raw).sum(), gr64["depth"].sum()]),
         np.array([g["grad_scale"], g["grad_shift"]]))
    env = check_param_digest(g, gr64["params"])
    assert check_param_digest(g, gr["params"]) < max(5e-4, floor, 3 * env)


@pytest.mark.parametrize("name", RENDER_CASES)
def test_render_gate_matched_and_kxy(name):
    """(1) d/d(kx,ky): the fp64 adjoint vs the reference's autograd through camera_mat (`grad_kxy`, what LearnFocal receives).
    (2) Where fp32 and fp64 gradients differ by more than rounding, the cause is hidden-unit ReLU gates: first-layer
    pre-activations carry ~1e-4 of noise from sin(2^9 p) of fp32-rounded points, so a handful of units with y ~ 0 switch between
    any two evaluation orders (ours, the reference's, fp64).  Evaluated in fp64 WITH the fp32 pass's gates, the adjoint agrees
    with the fp32 pass to rounding -- i.e. the adjoint arithmetic is right and the spread is the gates.  (render_oddflags:
    6.5e-3 on d c2w and 13 % on d kx between fp32 and fp64, 1e-4 / 3e-4 gate-matched.)"""
    g = load_golden(name)
    out, gr, g_r, g_t, raw, pix = run_render(g, np.float32)
    out64, gr64, _, _, _, _ = run_render(g, np.float64)
    assert relmax(gr64["kxy"], g["grad_kxy"]) < 2e-3, (gr64["kxy"], g["grad_kxy"])
    _, grm, _, _, _, _ = run_render(g, np.float64, gates_from=run_render.last_cache32)
    assert relmax(gr["c2w"], grm["c2w"]) < 3e-4, relmax(gr["c2w"], grm["c2w"])
    assert relmax(gr["kxy"], grm["kxy"]) < 1e-3, (gr["kxy"], grm["kxy"])


def _oracle_train(g, dtype):
    cfg = dict(O.DEFAULT_CFG); cfg["num_points"] = int(g["S"])
    state = dict(P={k: v.astype(dtype) for k, v in O.init_params(seed=int(g["seed"])).items()},
                 r=g["r0"].astype(dtype).copy(), t=g["t0"].astype(dtype).copy(),
                 scales=g["scales0"].astype(dtype).copy(), shifts=g["shifts0"].astype(dtype).copy())
    hist = []
    for it in range(int(g["steps"])):
        ld, grads, _ = O.train_step(state, g["img"].astype(dtype), g["dpt"].astype(dtype), g["ray_idx_%d" % it], g["noise_%d" % it].astype(dtype),
                                    int(g["idx"]), dtype(g["kx"]), dtype(g["ky"]), cfg, w_rgb=1.0, w_depth=0.04, rgb_loss_type="l1")
        hist.append((ld, grads))
    return state, hist


def test_train_step_render_only():
    """oracle.train_step (the function bench.py times as the CPU baseline and smoke() checks against) vs two consecutive
    reference Trainer.train_step calls (model/training.py:67-97; render + rgb L1 + depth L1, epoch 0 weights 1.0 / 0.04,
    torch.optim.Adam x3): losses, pose / distortion / MLP gradients of both steps, and every parameter after the two updates."""
    g = load_golden("train_render_only")
    idx = int(g["idx"])
    st32, h32 = _oracle_train(g, np.float32)
    st64, h64 = _oracle_train(g, np.float64)
    for it in range(int(g["steps"])):
        ld, gr = h32[it]; _, gr64 = h64[it]
        for k in ("loss", "loss_rgb", "loss_depth", "l2_mean"):
            ref = float(g["loss_%d.%s" % (it, k)])
            assert abs(float(ld[k]) - ref) / abs(ref) < 2e-5, (it, k, float(ld[k]), ref)
        for name, key in (("r", "grad_r_%d"), ("t", "grad_t_%d")):
            ref = g[key % it][idx]
            env = relmax(gr64[name], ref)
            assert relmax(gr[name], ref) < max(2e-4, 3 * env), (it, name, relmax(gr[name], ref), env)
            assert np.abs(np.delete(g[key % it], idx, 0)).max() == 0.0          # only the current view receives a pose gradient
        ref_ss = np.array([g["grad_scales_%d" % it][idx, 0], g["grad_shifts_%d" % it][idx, 0]])
        env = relmax(np.array([gr64["scale"], gr64["shift"]]), ref_ss)
        assert relmax(np.array([gr["scale"], gr["shift"]]), ref_ss) < max(2e-4, 3 * env)
        env = check_param_digest(g, gr64["P"], prefix="pg_%d." % it)
        assert check_param_digest(g, gr["P"], prefix="pg_%d." % it) < max(5e-4, 3 * env), it
    # Adam: parameters after the two updates (differences are measured on the UPDATE, which is ~lr in size)
    for key in ("r", "t", "scales", "shifts"):
        upd_ref = g[key + "_end"] - g[key + "0"]
        assert relmax(st32[key] - g[key + "0"], upd_ref) < 2e-3, key
    env = check_param_digest(g, st64["P"], prefix="pend.")
    assert check_param_digest(g, st32["P"], prefix="pend.") < max(1e-5, 3 * env)


def test_chamfer_dense():
    """oracle.chamfer vs the reference's Loss.get_pc_loss + autograd (model/losses.py:114-148) on two point clouds with one
    coincident pair (zero distance: sub-gradient 0) and one duplicated target (argmin tie -> first index)."""
    g = load_golden("chamfer_dense")
    loss, gX, gY, ixy, iyx = O.chamfer(g["X"], g["Y"])
    assert abs(float(loss) - float(g["loss"])) / float(g["loss"]) < 1e-6
    assert relmax(gX, g["gX"]) < 1e-5 and relmax(gY, g["gY"]) < 1e-5
    assert ixy[7] == 5 and not np.any(ixy == 11)          # the coincident pair; the duplicate at 11 never wins against index 3
    assert np.isfinite(gX).all() and np.isfinite(gY).all()


@pytest.mark.parametrize("name", ["train_full_losses", "train_full_lastview"])
def test_train_step_full_losses(name):
    """oracle.train_step with the reference-image stage (point-cloud chamfer + warped-RGB terms, training.py:280-365) vs the
    reference Trainer.train_step with the default loss set; `train_full_lastview` has the current view as the LAST camera
    (the two views swap roles, the scale of the last view is the fixed constant 1)."""
    g = load_golden(name)
    idx = int(g["idx"])

    def run(dtype):
        cfg = dict(O.DEFAULT_CFG); cfg["num_points"] = int(g["S"])
        state = dict(P={k: v.astype(dtype) for k, v in O.init_params(seed=int(g["seed"])).items()},
                     r=g["r0"].astype(dtype).copy(), t=g["t0"].astype(dtype).copy(),
                     scales=g["scales0"].astype(dtype).copy(), shifts=g["shifts0"].astype(dtype).copy())
        hist = []
        for it in range(int(g["steps"])):
            ref = dict(img=g["ref"].astype(dtype), dpt=g["rdpt"].astype(dtype), idx=int(g["ref_idx"]), w_pc=1.0, w_rgb_s=1.0)
            hist.append(O.train_step(state, g["img"].astype(dtype), g["dpt"].astype(dtype), g["ray_idx_%d" % it], g["noise_%d" % it].astype(dtype),
                                     idx, dtype(g["kx"]), dtype(g["ky"]), cfg, w_rgb=1.0, w_depth=0.04, rgb_loss_type="l1", ref=ref)[:2])
        return state, hist
    st32, h32 = run(np.float32)
    st64, h64 = run(np.float64)
    for it in range(int(g["steps"])):
        ld, gr = h32[it]; _, gr64 = h64[it]
        for k in ("loss", "loss_rgb", "loss_depth", "loss_pc", "loss_rgb_s"):
            ref = float(np.ravel(g["loss_%d.%s" % (it, k)])[0])
            assert abs(float(ld[k]) - ref) / abs(ref) < 5e-5, (it, k, float(ld[k]), ref)
        for nm, key in (("r", "grad_r_%d"), ("t", "grad_t_%d")):
            ref = g[key % it][idx]
            env = relmax(gr64[nm], ref)
            assert relmax(gr[nm], ref) < max(2e-4, 3 * env), (it, nm, relmax(gr[nm], ref), env)
        ref_ss = np.array([g["grad_scales_%d" % it][idx, 0], g["grad_shifts_%d" % it][idx, 0]])
        env = relmax(np.array([gr64["scale"], gr64["shift"]]), ref_ss)
        assert relmax(np.array([gr["scale"], gr["shift"]]), ref_ss) < max(2e-4, 3 * env), (it, gr["scale"], gr["shift"], ref_ss)
    for key in ("r", "t", "scales", "shifts"):
        upd_ref = g[key + "_end"] - g[key + "0"]
        assert relmax(st32[key] - g[key + "0"], upd_ref) < 2e-3, key


def test_train_step_learn_focal():
    """oracle.train_step with a learnable focal (LearnFocal order 2, training.py:247-252, intrinsics.py:59-70) and the full
    loss set vs two reference Trainer.train_step calls: d loss / d (fx, fy) collects the ray-generation, back-projection and
    projection terms of BOTH the render path and the reference-image stage."""
    g = load_golden("train_learn_focal")
    idx = int(g["idx"])

    def run(dtype):
        cfg = dict(O.DEFAULT_CFG); cfg["num_points"] = int(g["S"])
        state = dict(P={k: v.astype(dtype) for k, v in O.init_params(seed=int(g["seed"])).items()},
                     r=g["r0"].astype(dtype).copy(), t=g["t0"].astype(dtype).copy(), focal=g["focal0"].astype(dtype).copy(),
                     scales=g["scales0"].astype(dtype).copy(), shifts=g["shifts0"].astype(dtype).copy())
        hist = []
        for it in range(int(g["steps"])):
            ref = dict(img=g["ref"].astype(dtype), dpt=g["rdpt"].astype(dtype), idx=int(g["ref_idx"]), w_pc=1.0, w_rgb_s=1.0)
            hist.append(O.train_step(state, g["img"].astype(dtype), g["dpt"].astype(dtype), g["ray_idx_%d" % it], g["noise_%d" % it].astype(dtype),
                                     idx, None, None, cfg, w_rgb=1.0, w_depth=0.04, rgb_loss_type="l1", ref=ref,
                                     focal=dict(kx_gt=float(g["kx"]), ky_gt=float(g["ky"]), lr=1e-3))[:2])
        return state, hist
    st32, h32 = run(np.float32)
    st64, h64 = run(np.float64)
    for it in range(int(g["steps"])):
        ld, gr = h32[it]; _, gr64 = h64[it]
        for k in ("loss", "loss_rgb", "loss_depth", "loss_pc", "loss_rgb_s", "focalx", "focaly"):
            ref = float(np.ravel(g["loss_%d.%s" % (it, k)])[0])
            assert abs(float(ld[k]) - ref) / abs(ref) < 5e-5, (it, k, float(ld[k]), ref)
        env = relmax(gr64["focal"], g["grad_focal_%d" % it])
        assert relmax(gr["focal"], g["grad_focal_%d" % it]) < max(2e-4, 3 * env), (it, gr["focal"], g["grad_focal_%d" % it], env)
        assert env < 2e-3, env                              # the fp64 adjoint itself agrees with the reference's autograd
        for nm, key in (("r", "grad_r_%d"), ("t", "grad_t_%d")):
            ref = g[key % it][idx]
            assert relmax(gr[nm], ref) < max(2e-4, 3 * relmax(gr64[nm], ref)), (it, nm)
    assert relmax(st32["focal"] - g["focal0"], g["focal_end"] - g["focal0"]) < 2e-3
