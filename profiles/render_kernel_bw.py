"""Achieved HBM bandwidth of the per-ray kernels (sample_along_rays, sample_pdf, compositing forward / backward) against
their algorithmic bytes: composite forward 24 S + 40 B/ray, backward 40 S + 40 B/ray (same reads +
g_weights 4 S in, 16 S out), sample_pdf 4 (Nc - 2) + 4 Nc + 4 Nf + 24 in, 16 (Nc + Nf) out (depths + points).
Inputs are larger than L2 (126 MB) at the default 262 144 rays; CUDA events, 20 launches after 3 warm-ups."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypernerf_torch_b200 import model_utils as mu  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
ITERS = int(sys.argv[2]) if len(sys.argv) > 2 else 20     # 1 for an ncu capture (one warm-up, one launch per kernel)
dev = torch.device("cuda", 0)
peak = 6540.5
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timed(fn, iters=None):
    iters = ITERS if iters is None else iters
    for _ in range(3 if iters > 1 else 1):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


def report(name, secs, nbytes):
    gbs = nbytes / secs / 1e9
    print(f"{name:34s} {secs * 1e6:8.1f} us  {nbytes / 1e6:8.1f} MB  {gbs:7.0f} GB/s  {gbs / peak:5.2f} of measured HBM peak")


o = torch.rand(B, 3, device=dev) - 0.5
d = torch.rand(B, 3, device=dev) - 0.5
for S in (64, 128):
    sigma = torch.rand(B, S, device=dev) * 5
    rgb = torch.rand(B, S, 3, device=dev)
    z, _ = torch.sort(torch.rand(B, S, device=dev), -1)
    sigma_g, rgb_g = sigma.clone().requires_grad_(True), rgb.clone().requires_grad_(True)
    with torch.no_grad():
        t = timed(lambda: mu.volumetric_rendering(rgb, sigma, z, d, False))
    report(f"composite fwd S={S}", t, B * (24 * S + 40))
    out = mu.volumetric_rendering(rgb_g, sigma_g, z, d, False)
    g_rgb = torch.rand_like(out['rgb'])

    def bwd():
        torch.autograd.grad(out['rgb'], (rgb_g, sigma_g), g_rgb, retain_graph=True)
    t = timed(bwd)
    report(f"composite bwd S={S} (rgb grad only)", t, B * (20 * S + 12 + 12 + 16 * S))
Nc = 64
for Nf in (64, 128):
    z, _ = torch.sort(torch.rand(B, Nc, device=dev), -1)
    w = torch.rand(B, Nc, device=dev)
    u = torch.rand(B, Nf, device=dev)
    t = timed(lambda: mu.sample_pdf_fused(z, w, o, d, Nf, u=u))
    report(f"sample_pdf {Nc}+{Nf}", t, B * (4 * (Nc - 2) + 4 * Nc + 4 * Nf + 24 + 16 * (Nc + Nf)))
t = timed(lambda: mu.sample_along_rays(o, d, Nc, 0., 1., True, False))
report("sample_along_rays 64 (incl. torch.rand)", t, B * (24 + 4 * Nc + 16 * Nc))
