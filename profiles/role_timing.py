"""Per-role cycle breakdown of the fused MLP kernels (hn_debug_set_timing_buffer): where each warp role waits."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypernerf_torch_b200 import _lib, synthetic  # noqa: E402
from hypernerf_torch_b200 import train as hn_train  # noqa: E402
from hypernerf_torch_b200.models import NerfModel, _FusedMlp  # noqa: E402

dev = torch.device("cuda", 0)
emb = {'warp': list(range(100)), 'camera': [0], 'appearance': list(range(100)), 'time': list(range(100))}
model = NerfModel(emb, near=0., far=1., n_samples_coarse=64, n_samples_fine=64, noise_std=1.0,
                  hyper_slice_method='bendy_sheet', hyper_slice_out_dim=2, use_warp=True, use_nerf_embed=False,
                  use_alpha_cond=False, use_rgb_cond=False, GLO_dim=8, share_GLO=True, xyz_fourier_dim=10,
                  hyper_fourier_dim=6, view_fourier_dim=6)
model.load_state_dict(synthetic.make_state_dict(model, seed=0))
model = model.to(dev)
B, S = 8192, 128
rays, _ = synthetic.train_rays(B, seed=0, device=dev)
o, d, ids = rays[:, :3].contiguous(), rays[:, 3:6].contiguous(), rays[:, 8].long()
z, _ = torch.sort(torch.rand(B, S, device=dev), -1)
pts = (o[:, None] + z[..., None] * d[:, None]).contiguous()
params = model._canonical_params()
dbg = torch.zeros(148 * 8 + 64, dtype=torch.int64, device=dev)
names = ["prod_wait_empty", "mma_wait_act_ready", "mma_wait_full", "mma_total", "epi_wait_acc", "epi_work", "epi_prologue",
         "epi_total"]


def report(tag):
    torch.cuda.synchronize()
    v = dbg[:148 * 8].view(148, 8).double()
    tot = v[:, 3].mean().item()
    if tot == 0:
        print(tag, '(library built without -DHN_ROLE_TIMING=1: no role counters)')
        return
    print(tag, " ".join(f"{n}={v[:, i].mean().item() / tot:.3f}" for i, n in enumerate(names)), f"cycles={tot:.3e}")
    lay = dbg[148 * 8:].view(32, 2).double()
    if lay.sum() > 0 and tag.startswith("fwd"):
        tiles = 8192 * 128 / 256 / 148
        print("   per layer of CTA 0, cycles per tile [wait acc | drain]:",
              " ".join(f"{int(a / tiles)}|{int(b / tiles)}" for a, b in lay.tolist() if a + b > 0))
    dbg.zero_()


_lib.lib().hn_debug_set_timing_buffer(_lib.ptr(dbg))
with torch.no_grad():
    for it in range(2):
        _lib.profile = []
        _FusedMlp.apply(model, 1, pts, d, ids, None, 0.0, *[q.detach() for q in params])
        report("fwd (inference, no stash)")
        for name, n, a, b in _lib.profile:
            print(f"   {name}: {a.elapsed_time(b):.3f} ms  {2 * 801536 * n / a.elapsed_time(b) / 1e9:.1f} TFLOP/s")
_lib.profile = None
for it in range(2):
    _lib.lib().hn_debug_set_timing_buffer(_lib.ptr(dbg))
    sigma, rgb, warped = _FusedMlp.apply(model, 1, pts, d, ids, None, 0.0, *params)
    report("fwd  ")
    gs, gr = torch.randn_like(sigma), torch.randn_like(rgb)
    _lib.profile = []
    ((sigma * gs).sum() + (rgb * gr).sum()).backward()
    report("dgrad")
    for name, n, a, b in _lib.profile:
        print(f"   {name}: {a.elapsed_time(b):.3f} ms  {2 * 801536 * n / a.elapsed_time(b) / 1e9:.1f} TFLOP/s")
    _lib.profile = None
    _lib.lib().hn_debug_set_timing_buffer(None)
