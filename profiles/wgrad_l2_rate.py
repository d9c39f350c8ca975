"""Per-CTA operand rate of the weight-gradient kernel when its stashes are L2-resident against streaming them from HBM:
8 CTAs (hn_set_sm_partition) over 4 096 samples (69 MB of X + dY stash), timed right after a run that left the stashes in
L2 and after an L2 flush.  Tells whether a CTA's ~45 GB/s is set by memory latency / outstanding requests or by the SM."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypernerf_torch_b200 import synthetic  # noqa: E402
from hypernerf_torch_b200._lib import lib, ptr, check  # noqa: E402
from hypernerf_torch_b200.models import NerfModel  # noqa: E402

dev = torch.device("cuda", 0)
emb = {'warp': list(range(100)), 'camera': [0], 'appearance': list(range(100)), 'time': list(range(100))}
model = NerfModel(emb, near=0., far=1., n_samples_coarse=64, n_samples_fine=64, noise_std=1.0,
                  hyper_slice_method='bendy_sheet', hyper_slice_out_dim=2, use_warp=True, use_nerf_embed=False,
                  use_alpha_cond=False, use_rgb_cond=False, GLO_dim=8, share_GLO=True, xyz_fourier_dim=10,
                  hyper_fourier_dim=6, view_fourier_dim=6)
model.load_state_dict(synthetic.make_state_dict(model, seed=0))
model = model.to(dev)
level = 1
desc = model._desc
d = C.byref(desc)
offs, total = model._grad_offsets()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(512 << 20, device=dev, dtype=torch.uint8)

for B, S, ctas in ((64, 64, 8), (128, 64, 8), (128, 64, 16), (32, 64, 4), (8192, 128, 0)):
    sizes = model._sizes(B * S)
    saved = torch.randint(0, 60, (sizes.saved_bytes,), device=dev, dtype=torch.uint8)   # small finite bf16 bit patterns
    work = torch.randint(0, 60, (sizes.workspace_bytes,), device=dev, dtype=torch.uint8)
    flat = torch.zeros(total, device=dev)
    check(lib().hn_set_sm_partition(0, ctas), "partition")
    mb = (sizes.saved_bytes + sizes.workspace_bytes) / 1e6
    n_cta = ctas or 148

    def once():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib().hn_mlp_bwd_weights(d, ptr(saved), B, S, level, offs, ptr(flat), ptr(work), st), "wgrad")
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    once()
    warm = min(once() for _ in range(3))
    cold = []
    for _ in range(3):
        flush.fill_(1)
        cold.append(once())
    cold = min(cold)
    print(f"{B * S:8d} samples, {mb:8.1f} MB of stash, {n_cta:3d} CTAs: L2-warm {warm:7.3f} ms = {mb / warm / n_cta:6.1f} GB/s per CTA,"
          f" after an L2 flush {cold:7.3f} ms = {mb / cold / n_cta:6.1f} GB/s per CTA", flush=True)
check(lib().hn_set_sm_partition(0, 0), "partition")
