"""OfficialStaticNerf.forward on explicit points (official_nerf.py:69-96) through the library:
every point is a 'ray' with one sample (pixels/depth unused: explicit points mode)."""
import torch


def field_query(net, p, ray_d):
    raise NotImplementedError("explicit-point field queries are served by the renderer path; "
                              "OfficialStaticNerf.forward(p, ray_d) lands with nnb_field_fwd (DESIGN.md, next)")
