"""Shared helpers for the parity tests (test infrastructure)."""
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def relmax(a, b):
    """max-norm error relative to the tensor max (SURVEY.md 8(d) parity gate)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    den = max(np.abs(b).max(), 1e-30)
    return np.abs(a - b).max() / den


def cfg_from_golden(g):
    from oracle import nerf_oracle as O
    cfg = dict(O.DEFAULT_CFG)
    cfg["num_points"] = int(g["S"])
    m = {"cfg.rendering.sample_option": "sample_option", "cfg.rendering.dist_alpha": "dist_alpha",
         "cfg.rendering.depth_range": "depth_range", "cfg.rendering.white_background": "white_background",
         "cfg.rendering.use_ray_dir": "use_ray_dir", "cfg.rendering.normalise_ray": "normalise_ray",
         "cfg.model.occ_activation": "occ_activation"}
    for k, v in m.items():
        if k in g:
            val = g[k]
            if val.dtype.kind in "US":
                val = str(val)
            elif val.ndim == 0:
                val = val.item()
            else:
                val = tuple(float(x) for x in val)
            cfg[v] = val
    return cfg


def check_param_digest(g, grads, prefix="pg.", tol=1e-4):
    """Compare a dict name->gradient against the stored digests. Returns worst rel error."""
    from oracle import nerf_oracle as O
    worst = 0.0
    for n in O.PARAM_NAMES:
        gg = np.asarray(grads[n], dtype=np.float64).reshape(-1)
        amax = float(g[prefix + n + ".amax"])
        idx = g[prefix + n + ".idx"]; val = g[prefix + n + ".val"]
        den = max(amax, 1e-30)
        e1 = np.abs(gg[idx] - val).max() / den
        e2 = abs(np.sqrt((gg * gg).sum()) - float(g[prefix + n + ".l2"])) / max(float(g[prefix + n + ".l2"]), 1e-30)
        worst = max(worst, e1, e2)
    return worst
