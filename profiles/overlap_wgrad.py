"""Does the HBM-bound weight-gradient kernel overlap with the shared-memory-bound forward / data-gradient kernels when the
SMs are partitioned between them (hn_set_sm_partition) and the weight gradient runs on a second stream?
Sequential: [fwd(A), dgrad(A), wgrad(B)] x n on one stream, full grids.  Partitioned: [fwd(A), dgrad(A)] x n on the main
stream at m CTAs beside [wgrad(B)] x n on a side stream at w CTAs.  1 M samples per launch.
python profiles/overlap_wgrad.py [n]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypernerf_torch_b200 import _lib, synthetic  # noqa: E402
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
n_it = int(sys.argv[1]) if len(sys.argv) > 1 else 6
B, S, level = 8192, 128, 1
desc = model._desc
d = C.byref(desc)
packed = model._packed_weights(level)
offs, total = model._grad_offsets()
sizes = model._sizes(B * S)


class Set:
    def __init__(self, seed):
        rays, _ = synthetic.train_rays(B, seed=seed, device=dev)
        o, self.d, self.ids = rays[:, :3].contiguous(), rays[:, 3:6].contiguous(), rays[:, 8].long().contiguous()
        z, _ = torch.sort(torch.rand(B, S, device=dev), -1)
        self.pts = (o[:, None] + z[..., None] * self.d[:, None]).contiguous()
        self.sigma = torch.empty(B, S, device=dev)
        self.rgb = torch.empty(B, S, 3, device=dev)
        self.warped = torch.empty(B, S, 3 + desc.hyper_dim, device=dev)
        self.saved = torch.empty(sizes.saved_bytes, device=dev, dtype=torch.uint8)
        self.work = torch.empty(sizes.workspace_bytes, device=dev, dtype=torch.uint8)
        self.gs, self.gr = torch.randn(B, S, device=dev), torch.randn(B, S, 3, device=dev)
        self.flat = torch.zeros(total, device=dev)

    def fwd(self, st):
        check(lib().hn_mlp_fwd(d, ptr(packed), ptr(self.pts), ptr(self.d), ptr(self.ids), None, 0.0, B, S, None, 0,
                               ptr(self.sigma), ptr(self.rgb), ptr(self.warped), ptr(self.saved), None, st), "fwd")

    def dgrad(self, st):
        check(lib().hn_mlp_bwd_data(d, ptr(packed), ptr(self.ids), ptr(self.sigma), ptr(self.rgb), ptr(self.warped),
                                    ptr(self.saved), ptr(self.gs), ptr(self.gr), None, B, S, None, 0, level, offs,
                                    ptr(self.flat), ptr(self.work), None, None, st), "dgrad")

    def wgrad(self, st):
        check(lib().hn_mlp_bwd_weights(d, ptr(self.saved), B, S, level, offs, ptr(self.flat), ptr(self.work), st), "wgrad")


A, Bs = Set(0), Set(1)
main = torch.cuda.current_stream()
side = torch.cuda.Stream()
hm, hs = C.c_void_p(main.cuda_stream), C.c_void_p(side.cuda_stream)
Bs.fwd(hm); Bs.dgrad(hm); Bs.wgrad(hm)
ref = Bs.flat.clone()
torch.cuda.synchronize()


def run(m, w, overlap):
    check(lib().hn_set_sm_partition(m, w), "partition")
    for rep in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        es = torch.cuda.Event(enable_timing=True)
        e0.record(main)
        if overlap:
            side.wait_event(e0)
        for _ in range(n_it):
            A.fwd(hm); A.dgrad(hm)
            Bs.wgrad(hs if overlap else hm)
        if overlap:
            es.record(side)
            em = torch.cuda.Event(enable_timing=True); em.record(main)
            main.wait_event(es)
        e1.record(main)
        torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / n_it
    extra = f" main {e0.elapsed_time(em) / n_it:6.3f} side {e0.elapsed_time(es) / n_it:6.3f}" if overlap else ""
    print(f"mlp_ctas {m:3d} wgrad_ctas {w:3d} {'two streams' if overlap else 'one stream '}: {t:6.3f} ms per 1 M samples"
          f" (fwd + dgrad + wgrad){extra}", flush=True)
    return t


t0 = run(0, 0, False)
for w in (28, 36, 44, 52, 60, 68):
    run(148 - w, w, True)
run(0, 0, True)    # both at full grids: the block scheduler serialises them
check(lib().hn_set_sm_partition(0, 0), "partition")
# results do not depend on the partition
Bs.flat.zero_(); check(lib().hn_set_sm_partition(100, 44), "partition"); Bs.fwd(hm); Bs.dgrad(hm); Bs.wgrad(hm)
torch.cuda.synchronize()
print("flat gradient, partitioned vs full grids: rel", float((Bs.flat - ref).norm() / ref.norm()))
