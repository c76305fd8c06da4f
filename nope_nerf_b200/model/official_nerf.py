"""OfficialStaticNerf — parameter container + field evaluation (reference:
model/official_nerf.py:8-119).  Same nn.Linear submodules / state_dict keys as the
reference so checkpoints and train.py's reset_parameters() loop work unchanged; the
arithmetic runs in the CUDA library on a FLAT fp32 view of all 24 tensors."""
import math
import torch
import torch.nn as nn
from .. import _lib as L

_D, _PIN, _DIN = 256, 63, 27
_SHAPES = []
for _blk, _ins in (("layers0", [_PIN, _D, _D, _D]), ("layers1", [_D + _PIN, _D, _D, _D])):
    for _j, _i in enumerate((0, 2, 4, 6)):
        _SHAPES += [("%s.%d.weight" % (_blk, _i), (_D, _ins[_j])), ("%s.%d.bias" % (_blk, _i), (_D,))]
_SHAPES += [("fc_density.weight", (1, _D)), ("fc_density.bias", (1,)), ("fc_feature.weight", (_D, _D)),
            ("fc_feature.bias", (_D,)), ("rgb_layers.0.weight", (_D // 2, _D + _DIN)), ("rgb_layers.0.bias", (_D // 2,)),
            ("fc_rgb.weight", (3, _D // 2)), ("fc_rgb.bias", (3,))]
PARAM_NAMES = [n for n, _ in _SHAPES]
PARAM_SLICES = []  # (offset, numel, shape) in flat order == parameters() order
_o = 0
for _n, _s in _SHAPES:
    _k = int(math.prod(_s)); PARAM_SLICES.append((_o, _k, _s)); _o += _k
assert _o == L.NUM_PARAMS


class OfficialStaticNerf(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        D = cfg['model']['hidden_dim']
        if D != 256 or cfg['model']['pos_enc_levels'] != 10 or cfg['model']['dir_enc_levels'] != 4:
            raise NotImplementedError("the CUDA field is specialised for hidden_dim=256, pos/dir levels 10/4 "
                                      "(the reference hard-codes the levels, official_nerf.py:61,87)")
        pos_in, dir_in = 63, 27
        self.white_bkgd = cfg['rendering']['white_background']
        self.dist_alpha = cfg['rendering']['dist_alpha']
        self.occ_activation = cfg['model']['occ_activation']
        self.layers0 = nn.Sequential(nn.Linear(pos_in, D), nn.ReLU(), nn.Linear(D, D), nn.ReLU(),
                                     nn.Linear(D, D), nn.ReLU(), nn.Linear(D, D), nn.ReLU())
        self.layers1 = nn.Sequential(nn.Linear(D + pos_in, D), nn.ReLU(), nn.Linear(D, D), nn.ReLU(),
                                     nn.Linear(D, D), nn.ReLU(), nn.Linear(D, D), nn.ReLU())
        self.fc_density = nn.Linear(D, 1)
        self.fc_feature = nn.Linear(D, D)
        self.rgb_layers = nn.Sequential(nn.Linear(D + dir_in, D // 2), nn.ReLU())
        self.fc_rgb = nn.Linear(D // 2, 3)
        self.fc_density.bias.data = torch.tensor([0.1]).float()          # official_nerf.py:39
        self.sigmoid = nn.Sigmoid()
        self.fc_rgb.bias.data = torch.tensor([0.8 if self.white_bkgd else 0.02] * 3).float()   # :41-44
        self._flat = None
        self._flat_grad = None

    # ---- flat parameter view -------------------------------------------------------------
    def _plist(self):
        ps = list(self.parameters())
        assert len(ps) == len(PARAM_SLICES)
        return ps

    def flat_weights(self):
        """Flat fp32 tensor aliasing all parameters (re-packed if a .to()/load broke the aliasing)."""
        ps = self._plist()
        f = self._flat
        ok = f is not None and f.device == ps[0].device
        if ok:
            base = f.data_ptr()
            for p, (o, n, s) in zip(ps, PARAM_SLICES):
                if p.data_ptr() != base + 4 * o:
                    ok = False; break
        if not ok:
            f = torch.empty(L.NUM_PARAMS, dtype=torch.float32, device=ps[0].device)
            with torch.no_grad():
                for p, (o, n, s) in zip(ps, PARAM_SLICES):
                    f[o:o + n].copy_(p.detach().reshape(-1).float())
                    p.data = f[o:o + n].view(s)
            self._flat = f
        return f

    def flat_grad(self, zero=True, alias=None):
        """Flat gradient buffer whose slices are installed as every parameter's .grad.
        `alias` lets a caller (data-parallel trainer) provide the storage."""
        ps = self._plist()
        g = alias if alias is not None else self._flat_grad
        if g is None or g.device != ps[0].device:
            g = torch.zeros(L.NUM_PARAMS, dtype=torch.float32, device=ps[0].device)
            zero = False
        self._flat_grad = g
        if zero:
            g.zero_()
        for p, (o, n, s) in zip(ps, PARAM_SLICES):
            if p.requires_grad:
                p.grad = g[o:o + n].view(s)
        return g

    # ---- reference API -------------------------------------------------------------------
    def infer_occ(self, p):
        """official_nerf.py:60-67: (trunk features, density logit).  The trunk features never leave the kernel here; callers in the
        reference (gradient(), forward()) only use the logit, so the first element is None."""
        from ..field import field_query
        _, s = field_query(self, p, None, raw_density=True)
        return None, s

    def gradient(self, p, it):
        """official_nerf.py:46-58: -d(density logit)/d p, shape (N, 1, 3) (surface normals of the geometry view).  The reference
        differentiates infer_occ with autograd; here it is the data-gradient chain of the field kernel with cotangent 1 on the logit."""
        from ..field import field_query
        with torch.enable_grad():
            q = p.detach().clone().requires_grad_(True)
            _, s = field_query(self, q, None, raw_density=True)
            g, = torch.autograd.grad(s, q, torch.ones_like(s), create_graph=False, retain_graph=False, allow_unused=True)
        return -g.unsqueeze(1)

    def forward(self, p, ray_d=None, only_occupancy=False, return_logits=False, return_addocc=False,
                noise=False, it=100000, **kwargs):
        """Field query on explicit points (official_nerf.py:69-96); `noise`, `it`, `return_logits`
        are accepted and ignored exactly like the reference."""
        from ..field import field_query
        rgb, a = field_query(self, p, ray_d)
        if only_occupancy:
            return a
        if ray_d is not None:
            return (rgb, a) if return_addocc else rgb
        return None
