"""UMMA cost by accumulator rotation, M, operand source and CTA pairing (hn_umma_rate2)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypernerf_torch_b200 import _lib  # noqa: E402

L = _lib.probe_lib()
out = torch.zeros(148, dtype=torch.int64, device="cuda")


def run(cg, M, N, nacc, order, a_src, inner, grid=148, reps=50):
    _lib.check(L.hn_umma_rate2(cg, M, N, nacc, order, a_src, reps, inner, grid, _lib.ptr(out), _lib.stream()), "rate2")
    torch.cuda.synchronize()
    v = out[:grid:cg].double()
    return v.mean().item() / reps


def slope(cg, M, N, nacc, order, a_src):
    a, b = run(cg, M, N, nacc, order, a_src, 1), run(cg, M, N, nacc, order, a_src, 5)
    return (b - a) / (4 * 16 * nacc)


print("cta_group M N nacc order a_src -> cycles/UMMA (floor = max(M,128)*N/(256*cg))")
for cg, M in ((1, 128), (1, 64), (2, 256), (2, 128)):
    for N in (256, 128, 64, 16):
        for nacc, order in ((1, 0), (2, 0), (2, 1), (3, 0)):
            if N * nacc > 384:
                continue
            for a_src in (0, 1, 2):
                if a_src == 2 and cg == 2:
                    continue
                s = slope(cg, M, N, nacc, order, a_src)
                print(f"cg={cg} M={M:3d} N={N:3d} nacc={nacc} order={order} a_src={a_src}: {s:7.1f}  (floor {max(M // cg, 128) * N / 256 / 1:.0f})")
