"""Profiling driver: forward + backward of one chunk of the cfg2 workload (or, with a third argument `se3`, of cfg5), repeated,
for ncu:  python profiles/prof_chunk.py <rays> <iterations> [se3]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypernerf_torch_b200 import synthetic  # noqa: E402
from hypernerf_torch_b200 import train as hn_train  # noqa: E402
from hypernerf_torch_b200.models import NerfModel  # noqa: E402

n_rays = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
emb = {'warp': list(range(100)), 'camera': [0], 'appearance': list(range(100)), 'time': list(range(100))}
if len(sys.argv) > 3 and sys.argv[3] == "se3":   # BASELINE.json configs[4]: SE3 warp + axis-aligned slicing, 128+128
    model = NerfModel(emb, near=0., far=1., n_samples_coarse=128, n_samples_fine=128, noise_std=1.0, use_warp=True,
                      use_nerf_embed=False, hyper_slice_method='axis_aligned_plane', hyper_slice_out_dim=8,
                      view_fourier_dim=6, warp_field_type='se3')
else:
    model = NerfModel(emb, near=0., far=1., n_samples_coarse=64, n_samples_fine=64, noise_std=1.0,
                      hyper_slice_method='bendy_sheet', hyper_slice_out_dim=2, use_warp=True, use_nerf_embed=False,
                      use_alpha_cond=False, use_rgb_cond=False, GLO_dim=8, share_GLO=True, xyz_fourier_dim=10,
                      hyper_fourier_dim=6, view_fourier_dim=6)
model.load_state_dict(synthetic.make_state_dict(model, seed=0))
model = model.to(dev)
fg = hn_train.FlatGrads(model.parameters())
model.attach_flat_grads(fg)
rays, rgbs = synthetic.train_rays(n_rays, seed=0, device=dev)
for _ in range(iters):
    hn_train.train_step(model, rays, rgbs, fg, global_rays=n_rays, chunk=n_rays)
torch.cuda.synchronize()
print("done")
