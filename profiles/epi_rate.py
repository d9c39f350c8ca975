"""Which pipe bounds the forward epilogue: cycles per 256-column layer drain (8 warps, thread = row) by work subset."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypernerf_torch_b200 import _lib  # noqa: E402

L = _lib.probe_lib()
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 148
out = torch.zeros(grid * 2, dtype=torch.int64, device="cuda")
gout = torch.empty(grid * 256 * 512, dtype=torch.uint8, device="cuda")
NAMES = {1: "cvt", 2: "sts", 4: "stg", 8: "bias(lds+fadd)", 16: "2 tmem loads in flight", 32: "stg 64B/thread", 64: "bulk s2g 512B x32/warp", 128: "bulk s2g 64KB/sub"}
for mode in (16, 19, 23, 27, 31, 83, 147, 91, 155):
    reps = 200
    _lib.check(L.hn_epi_rate(mode, reps, grid, _lib.ptr(gout), _lib.ptr(out), _lib.stream()), "epi_rate")
    torch.cuda.synchronize()
    cyc = out[::2].double().mean().item() / reps
    print(f"mode {mode:2d} [{' + '.join(v for k, v in NAMES.items() if mode & k) or 'tmem load only'}]: {cyc:8.0f} cycles per 256-column layer")
