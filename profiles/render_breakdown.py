"""Where a full-frame render (cfg3: 1008 x 756, 64 + 128 samples, no_grad) spends its time: CUDA-event time of the whole
frame vs the sum of the fused MLP kernels, plus a torch.profiler table of everything else."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypernerf_torch_b200 import _lib, synthetic  # noqa: E402
from hypernerf_torch_b200 import train as hn_train  # noqa: E402
from hypernerf_torch_b200.models import NerfModel  # noqa: E402

dev = torch.device("cuda", 0)
emb = {'warp': list(range(100)), 'camera': [0], 'appearance': list(range(100)), 'time': list(range(100))}
model = NerfModel(emb, near=0., far=1., n_samples_coarse=64, n_samples_fine=128, noise_std=None,
                  hyper_slice_method='bendy_sheet', hyper_slice_out_dim=2, use_warp=True, use_nerf_embed=False,
                  use_alpha_cond=False, use_rgb_cond=False, GLO_dim=8, share_GLO=True, xyz_fourier_dim=10,
                  hyper_fourier_dim=6, view_fourier_dim=6)
model.load_state_dict(synthetic.make_state_dict(model, seed=0))
model = model.to(dev).eval()
frame = synthetic.frame_rays(image_id=0, seed=0, device=dev)
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
hn_train.render_rays(model, frame[:chunk], chunk=chunk)
for it in range(2):
    _lib.profile = []
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    hn_train.render_rays(model, frame, chunk=chunk)
    b.record()
    torch.cuda.synchronize()
    mlp = sum(x.elapsed_time(y) for _, _, x, y in _lib.profile)
    print(f"frame {a.elapsed_time(b):.1f} ms, fused MLP kernels {mlp:.1f} ms in {len(_lib.profile)} launches, chunk {chunk}")
    _lib.profile = None
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    hn_train.render_rays(model, frame, chunk=chunk)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
