"""Per-kernel CUDA-event times of one 8 192-ray chunk (fine level: 1 M samples): forward in inference and training
mode, data gradient, weight gradient.  python profiles/kernel_times.py [samples per ray]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypernerf_torch_b200 import _lib, synthetic  # noqa: E402
from hypernerf_torch_b200.models import NerfModel, _FusedMlp  # noqa: E402

dev = torch.device("cuda", 0)
emb = {'warp': list(range(100)), 'camera': [0], 'appearance': list(range(100)), 'time': list(range(100))}
model = NerfModel(emb, near=0., far=1., n_samples_coarse=64, n_samples_fine=64, noise_std=1.0,
                  hyper_slice_method='bendy_sheet', hyper_slice_out_dim=2, use_warp=True, use_nerf_embed=False,
                  use_alpha_cond=False, use_rgb_cond=False, GLO_dim=8, share_GLO=True, xyz_fourier_dim=10,
                  hyper_fourier_dim=6, view_fourier_dim=6)
model.load_state_dict(synthetic.make_state_dict(model, seed=0))
model = model.to(dev)
B, S = 8192, int(sys.argv[1]) if len(sys.argv) > 1 else 128     # samples per ray: 128 = 1 M samples per launch
rays, _ = synthetic.train_rays(B, seed=0, device=dev)
o, d, ids = rays[:, :3].contiguous(), rays[:, 3:6].contiguous(), rays[:, 8].long()
z, _ = torch.sort(torch.rand(B, S, device=dev), -1)
pts = (o[:, None] + z[..., None] * d[:, None]).contiguous()
params = model._canonical_params()


def show(tag):
    torch.cuda.synchronize()
    for name, n, a, b in _lib.profile:
        ms = a.elapsed_time(b)
        print(f"{tag:10s} {name:10s} {ms:7.3f} ms  {2 * 801536 * n / ms / 1e9:7.1f} TFLOP/s")
    _lib.profile = []


for it in range(3):
    _lib.profile = []
    # detached parameters: autograd reports needs_input_grad even under no_grad, and the shim allocates the stash
    # (training kernel) whenever a parameter requires grad
    _FusedMlp.apply(model, 1, pts, d, ids, None, 0.0, *[q.detach() for q in params])
    show("inference")
    sigma, rgb, warped = _FusedMlp.apply(model, 1, pts, d, ids, None, 0.0, *params)
    show("training")
    gs, gr = torch.randn_like(sigma), torch.randn_like(rgb)
    ((sigma * gs).sum() + (rgb * gr).sum()).backward()
    show("training")
_lib.profile = None
