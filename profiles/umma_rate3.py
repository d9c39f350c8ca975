"""UMMA cost with a fully unrolled issue sequence (hn_umma_rate3)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypernerf_torch_b200 import _lib  # noqa: E402

L = _lib.probe_lib()
out = torch.zeros(148, dtype=torch.int64, device="cuda")


def run(N, nacc, inner, grid=148, reps=50):
    _lib.check(L.hn_umma_rate3(N, nacc, reps, inner, grid, _lib.ptr(out), _lib.stream()), "rate3")
    torch.cuda.synchronize()
    return out[:grid].double().mean().item() / reps


for grid in (1, 148):
    for N in (256, 128, 64, 16):
        for nacc in (1, 2):
            a, b = run(N, nacc, 1, grid), run(N, nacc, 9, grid)
            print(f"grid={grid:3d} N={N:3d} nacc={nacc}: {(b - a) / (8 * 16 * nacc):7.1f} cycles/UMMA (floor {N / 2:.0f}); 16-UMMA rep incl. commit+wait {a:7.0f}")
