"""UMMA issue-rate microbenchmark: cycles per K=16 UMMA (M=128), slope and per-commit constant separated by varying
the number of UMMAs between commits."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypernerf_torch_b200 import _lib  # noqa: E402

L = _lib.probe_lib()
out = torch.zeros(148, dtype=torch.int64, device="cuda")


def run(N, nsub, sw, inner, grid=148, reps=100, ks=16):
    _lib.check(L.hn_umma_rate(N, ks, reps, sw, nsub, inner, grid, _lib.ptr(out), _lib.stream()), "rate")
    torch.cuda.synchronize()
    return out[:grid].double().mean().item() / reps


for N in (256, 128, 64, 16):
    for nsub in (1, 2):
        for sw in (0, 1):
            a, b = run(N, nsub, sw, 1), run(N, nsub, sw, 8)
            n1, n8 = 16 * nsub, 8 * 16 * nsub
            slope = (b - a) / (n8 - n1)
            const = a - slope * n1
            print(f"N={N:3d} nsub={nsub} swizzle={'128B' if sw else 'none'}: {slope:6.1f} cycles/UMMA (floor {128 * N / 256:.0f}), "
                  f"commit+wait constant {const:6.0f} cycles")

print("commit frequency (un-waited tcgen05.commit inside the issue loop), N=256, 2 accumulators:")
for every in (0, 1, 2, 4):
    sw = 0 if every == 0 else every + 1
    a, b = run(256, 2, sw, 1), run(256, 2, sw, 8)
    slope = (b - a) / (8 * 32 - 32)
    print(f"  commit every {every or 'never':>5} k-steps: {slope:6.1f} cycles/UMMA")
