"""What an epilogue drain costs next to a busy tensor pipe and a busy weight ring (hn_overlap_rate): cycles for four warps to
drain a 128 x 256 accumulator, alone / with back-to-back UMMAs on the other sub-tile / with bulk copies into shared memory."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypernerf_torch_b200 import _lib  # noqa: E402

L = _lib.probe_lib()
grid, reps = 148, 200
gout = torch.empty(grid * 65536, dtype=torch.uint8, device="cuda")
gsrc = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
out = torch.zeros(grid * 4, dtype=torch.int64, device="cuda")
MODES = [(0, "tmem loads only"), (1, "+ pack"), (3, "+ st.shared"), (11, "+ bias"), (7, "pack + st.shared + st.global"),
         (15, "all (training drain)")]
print(f"{'drain work':32s} {'alone':>8s} {'+umma256':>9s} {'umma/128cyc':>11s} {'+umma+tma':>10s} {'umma/128cyc':>11s} {'+umma128':>9s}")
for mode, name in MODES:
    row = []
    for umma, n, tma in ((0, 256, 0), (1, 256, 0), (1, 256, 1), (1, 128, 0)):
        out.zero_()
        _lib.check(L.hn_overlap_rate(mode, umma, n, tma, reps, grid, _lib.ptr(gout), _lib.ptr(gsrc), _lib.ptr(out), _lib.stream()),
                   "overlap", L)
        torch.cuda.synchronize()
        v = out.view(grid, 4).double()
        drain = v[:, 0].mean().item() / reps
        rate = (v[:, 1] * 128.0 / v[:, 2].clamp(min=1)).mean().item() if umma else 0.0   # fraction of the N=256 floor (128 cycles / UMMA)
        row.append((drain, rate))
    print(f"{name:32s} {row[0][0]:8.0f} {row[1][0]:9.0f} {row[1][1]:11.2f} {row[2][0]:10.0f} {row[2][1]:11.2f} {row[3][0]:9.0f}")
