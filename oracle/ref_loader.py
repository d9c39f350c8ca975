"""TEST INFRASTRUCTURE — loads the UNMODIFIED reference (songrise/HyperNeRF-torch) in place, with the module stubs
of SURVEY.md Appendix C.  Only tests/, oracle/make_golden.py, __graft_entry__.smoke() and bench.py's reference /
cpu_baseline legs may import this; the product package never does.

The reference tree is resolved as $HN_REFERENCE_DIR -> /root/reference -> baseline/_ref.  It does not exist on the GPU
box unless baseline/_ref was populated; callers must handle `reference_available() == False`.
"""
import os
import sys
import types

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_dir():
    for cand in (os.environ.get("HN_REFERENCE_DIR"), "/root/reference", os.path.join(_ROOT, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "hypernerf", "models.py")):
            return cand
    return None


def reference_available():
    return reference_dir() is not None


_loaded = None


def load_reference():
    """Returns (models, model_utils) modules of the reference.  Stubs: immutabledict (models.py:21, dead code),
    torchsummary (modules.py:21, dead), torchsearchsorted (models/rendering.py:2), and on CPU-only hosts a no-op
    Tensor.cuda because channel counting calls .cuda() (model_utils.py:250,278)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    ref = reference_dir()
    if ref is None:
        raise RuntimeError("reference tree not found (set HN_REFERENCE_DIR)")
    sys.modules.setdefault('immutabledict', types.SimpleNamespace(immutabledict=dict))
    sys.modules.setdefault('torchsummary', types.ModuleType('torchsummary'))
    ts = types.ModuleType('torchsearchsorted')
    ts.searchsorted = lambda a, v, side='left': torch.searchsorted(a, v, right=(side == 'right'))
    sys.modules.setdefault('torchsearchsorted', ts)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    if ref not in sys.path:
        sys.path.insert(0, ref)
    from hypernerf import model_utils as ref_model_utils
    from hypernerf import models as ref_models
    _loaded = (ref_models, ref_model_utils)
    return _loaded


EMBEDDINGS = {'warp': list(range(100)), 'camera': [0], 'appearance': list(range(100)), 'time': list(range(100))}
EXTRA = dict(nerf_alpha=None, warp_alpha=None, hyper_alpha=None, hyper_sheet_alpha=None)


def cfg1_kwargs(n_fine=64, noise_std=1.0):
    """Constructor arguments of BASELINE.json configs 1-3 (train.py:48-67 with hyper_dim 2)."""
    return dict(near=0., far=1., n_samples_coarse=64, n_samples_fine=n_fine, noise_std=noise_std,
                hyper_slice_method='bendy_sheet', hyper_slice_out_dim=2, use_warp=True, use_nerf_embed=False,
                use_alpha_cond=False, use_rgb_cond=False, GLO_dim=8, share_GLO=True, xyz_fourier_dim=10,
                hyper_fourier_dim=6, view_fourier_dim=6)


def build_reference_model(seed=0, **overrides):
    ref_models, _ = load_reference()
    kw = cfg1_kwargs()
    kw.update(overrides)
    torch.manual_seed(seed)
    return ref_models.NerfModel(EMBEDDINGS, **kw)


class _DrawTape:
    """Records (or replays) the torch.rand / torch.randn calls the reference makes inside one forward, in order
    (SURVEY.md App. A.5): rand[B,Nc], randn(B,Nc,1), rand(B,Nf), randn(B,Nc+Nf,1)."""

    def __init__(self, replay=None):
        self.replay = list(replay) if replay is not None else None
        self.tape = []

    def __enter__(self):
        self._rand, self._randn = torch.rand, torch.randn

        def make(orig):
            def fn(*a, **k):
                if self.replay is not None:
                    t = self.replay.pop(0)
                    dev = k.get('device', None)
                    t = t.to(dev) if dev is not None else t
                else:
                    t = orig(*a, **k)
                self.tape.append(t.detach().clone())
                return t
            return fn
        torch.rand, torch.randn = make(self._rand), make(self._randn)
        return self

    def __exit__(self, *exc):
        torch.rand, torch.randn = self._rand, self._randn


def run_reference(model, rays, draws=None, taps=None, **forward_kwargs):
    """Forward of the unmodified reference NerfModel on ray rows (B,9).  draws: list of tensors to replay in call
    order (None = draw fresh and record).  taps: optional dict that receives the stage taps z_coarse / z_fine
    (the reference does not return its depth samples).  Returns (outputs, list_of_draws)."""
    _, ref_mu = load_reference()
    orig_sar, orig_pdf = ref_mu.sample_along_rays, ref_mu.sample_pdf

    def sar(*a, **k):
        z, p = orig_sar(*a, **k)
        if taps is not None:
            taps['z_coarse'] = z.detach().clone()
        return z, p

    def pdf(*a, **k):
        z, p = orig_pdf(*a, **k)
        if taps is not None:
            taps['z_fine'] = z.detach().clone()
        return z, p

    ref_mu.sample_along_rays, ref_mu.sample_pdf = sar, pdf
    try:
        with _DrawTape(draws) as tape:
            out = model(ref_mu.prepare_ray_dict(rays), dict(EXTRA), **forward_kwargs)
    finally:
        ref_mu.sample_along_rays, ref_mu.sample_pdf = orig_sar, orig_pdf
    return out, tape.tape


def draws_to_dict(tape, noise=True):
    if noise:
        return dict(u_coarse=tape[0], noise_coarse=tape[1], u_fine=tape[2], noise_fine=tape[3])
    return dict(u_coarse=tape[0], noise_coarse=None, u_fine=tape[1], noise_fine=None)
