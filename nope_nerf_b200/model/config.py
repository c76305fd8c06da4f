"""get_model (reference: model/config.py:4-17).  The DPT monocular-depth network is an offline
preprocessing dependency and out of scope: depth.type must be None."""
from .network import nope_nerf


def get_model(renderer, cfg, device=None, **kwargs):
    if cfg['depth']['type'] == 'DPT':
        raise NotImplementedError("depth.type == 'DPT' (online DPT inference) is out of scope; precompute dpt/*.npz "
                                  "with the reference's preprocess/dpt_depth.py")
    return nope_nerf(cfg, renderer, None, device)
