"""Pure-write / pure-read / copy HBM bandwidth with torch kernels (context for the stash-writing MLP kernels, whose
traffic is almost all writes): fill_ (write only), sum (read only), copy_ (read + write)."""
import torch

dev = torch.device("cuda", 0)
n = 4 << 30
x = torch.empty(n, dtype=torch.uint8, device=dev)
y = torch.empty(n, dtype=torch.uint8, device=dev)
xf = x.view(torch.float32)


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 1e3)
    return best


t = timed(lambda: xf.fill_(1.0)); print(f"fill_  (write only)   {n / t / 1e12:.2f} TB/s")
t = timed(lambda: xf.sum()); print(f"sum    (read only)    {n / t / 1e12:.2f} TB/s")
t = timed(lambda: y.copy_(x)); print(f"copy_  (read + write) {2 * n / t / 1e12:.2f} TB/s")
