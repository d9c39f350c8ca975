"""TMEM -> register read throughput (tcgen05.ld.32x32b.x32) per SM."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypernerf_torch_b200 import _lib  # noqa: E402

L = _lib.probe_lib()
out = torch.zeros(4, dtype=torch.int64, device="cuda")
for nw in (1, 4, 8):
    for mode in (0, 1):
        reps, cols = 2000, 256
        _lib.check(L.hn_tmem_rate(nw, cols, reps, mode, _lib.ptr(out), _lib.stream()), "tmem_rate")
        torch.cuda.synchronize()
        cyc = out[0].item() / reps
        byts = nw * 32 * cols * 4
        print(f"warps={nw} mode={mode}: {cyc:8.1f} cycles per {cols}-column sweep  -> {byts / cyc:7.1f} B/clk/SM")
