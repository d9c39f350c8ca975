"""TEST INFRASTRUCTURE.  Training-trajectory parity: the fused train step + FusedAdam against the same optimisation carried out
with autograd of the fp32 oracle + torch.optim.Adam (the reference's arithmetic and optimizer, utils/__init__.py:22-41), and
against the oracle under fp16 autocast with a 2^16 loss scale (the precision the reference trains in, train.py:217-218).

One fixed batch of rays with a learnable target (a smooth function of the ray), the same initial weights, and — step by
step — the SAME random draws in every arm (the product's draws are recorded and replayed into the oracle arms).  Per-tensor
gradient distances (profiles/grad_parity.md) say how far one bf16-operand gradient is from the fp32 one; this says what that
does to training: whether the loss curves stay together.

`python tests/train_parity.py [rays] [steps]` writes gpurun_out/train_parity.md (copied to profiles/)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import hypernerf_oracle as orc  # noqa: E402
from oracle import ref_loader  # noqa: E402


def target_colours(rays):
    """A target the model can fit: smooth in the ray origin / direction (NDC rays: origins on z = -1, x, y in [-1, 1])."""
    o, d = rays[:, :3], rays[:, 3:6]
    return torch.stack([0.5 + 0.4 * torch.sin(3.0 * o[:, 0]), 0.5 + 0.4 * torch.cos(2.0 * o[:, 1]),
                        0.5 + 0.3 * torch.sin(2.0 * (o[:, 0] + o[:, 1]) + d[:, 0])], -1).contiguous()


def _draw_dict(tape):
    noise = len(tape) == 4
    return dict(u_coarse=tape[0], noise_coarse=tape[1] if noise else None, u_fine=tape[2 if noise else 1],
                noise_fine=tape[3] if noise else None)


def product_run(model, rays, rgbs, steps, lr):
    """`steps` train steps of the product (one chunk per step); returns the losses and every step's recorded draws."""
    from hypernerf_torch_b200 import train as hn_train
    fg = hn_train.FlatGrads(model.parameters())
    model.attach_flat_grads(fg)
    opt = hn_train.FusedAdam(fg, lr=lr)
    losses, tapes = [], []
    try:
        for _ in range(steps):
            with ref_loader._DrawTape() as tape:
                loss = hn_train.train_step(model, rays, rgbs, fg, chunk=rays.shape[0], optimizer=opt)
            losses.append(float(loss))
            tapes.append(tape.tape)
    finally:
        model.attach_flat_grads(None)
    return losses, tapes


def oracle_run(sd0, rays, rgbs, tapes, cfg, lr, mode, chunk=1024):
    """The same optimisation through the oracle: mode 'f32' or 'amp16' (autocast + loss scale 2^16, unscaled before Adam)."""
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in sd0.items()}
    opt = torch.optim.Adam(list(sd.values()), lr=lr, eps=1e-8)
    B = rays.shape[0]
    scale = 65536.0 if mode == 'amp16' else 1.0
    losses = []
    for tape in tapes:
        draws = _draw_dict(tape)
        opt.zero_grad(set_to_none=True)
        total = 0.0
        for i in range(0, B, chunk):
            sl = slice(i, i + chunk)
            dr = {k: (None if v is None else v[sl]) for k, v in draws.items()}
            with torch.autocast('cuda', dtype=torch.float16, enabled=(mode == 'amp16')):
                out = orc.forward(sd, rays[sl, :3], rays[sl, 3:6], rays[sl, 8].long(), dr, cfg)
                loss = ((out['coarse']['rgb'].float() - rgbs[sl]) ** 2).sum() / (3.0 * B) + \
                       ((out['fine']['rgb'].float() - rgbs[sl]) ** 2).sum() / (3.0 * B)
            (loss * scale).backward()
            total += float(loss.detach())
        if scale != 1.0:
            for p in sd.values():
                if p.grad is not None:
                    p.grad.div_(scale)
        opt.step()
        losses.append(total)
    return losses


def run(n_rays=2048, steps=40, lr=5e-4, seed=0, device="cuda", arms=('f32', 'amp16')):
    from hypernerf_torch_b200 import synthetic
    from hypernerf_torch_b200.models import NerfModel
    model = NerfModel(ref_loader.EMBEDDINGS, **ref_loader.cfg1_kwargs(n_fine=64, noise_std=1.0))
    model.load_state_dict(synthetic.make_state_dict(model, seed=seed, boosted=False))
    model = model.to(device)
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    rays, _ = synthetic.train_rays(n_rays, seed=seed + 5, device=device)
    rgbs = target_colours(rays)
    torch.manual_seed(seed + 11)
    out = {}
    out['kernel'], tapes = product_run(model, rays, rgbs, steps, lr)
    cfg = orc.default_cfg(n_fine=64, noise_std=1.0)
    for mode in arms:
        out[mode] = oracle_run(sd0, rays, rgbs, tapes, cfg, lr, mode)
    return out


def worst_rel(a, b):
    return max(abs(x - y) / abs(y) for x, y in zip(a, b))


def mean_abs_log_ratio(a, b):
    """Mean over the steps of |ln(a_i / b_i)|: how far two loss curves are apart, in relative terms."""
    import math
    return sum(abs(math.log(x / y)) for x, y in zip(a, b)) / len(a)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    tr = run(n, steps)
    k, f, a = tr['kernel'], tr['f32'], tr['amp16']
    w = max(1, steps // 10)
    md = ["# Training-trajectory parity", "",
          f"`python tests/train_parity.py {n} {steps}` on a B200: one fixed batch of {n} rays (cfg 1 / 2 model, 64+64 samples, noise_std 1),",
          "a smooth learnable target, Adam lr 5e-4.  `kernel` = `train.train_step` + `FusedAdam` (bf16 operands, fp32 accumulate);",
          "`fp32 oracle` = autograd of the oracle + `torch.optim.Adam`; `fp16 autocast oracle` = the same under",
          "`torch.autocast(float16)` with a 2^16 loss scale (the reference's training precision, train.py:217-218).  Same initial",
          "weights, and in every step the same random draws in all three arms.  Loss = MSE(coarse) + MSE(fine).", "",
          "Fitting one batch with Adam is a noisy optimisation (the curves spike), and it amplifies any perturbation: after ~15",
          "steps the three arms are different realisations of the same descent.  What can be asked is (a) that they coincide",
          "while the perturbation is still small, and (b) that the bf16-operand path stays as close to the fp32 curve as the",
          "reference's own fp16-autocast training does.", "",
          f"* first 10 steps, largest relative loss difference: kernel vs fp32 oracle **{worst_rel(k[:10], f[:10]):.2e}**, "
          f"fp16 autocast oracle vs fp32 oracle {worst_rel(a[:10], f[:10]):.2e}",
          f"* all {steps} steps, mean |ln(loss / loss_fp32)|: kernel **{mean_abs_log_ratio(k, f):.3f}**, fp16 autocast oracle "
          f"{mean_abs_log_ratio(a, f):.3f}",
          f"* loss {f[0]:.5f} -> mean of the last {w} steps {sum(f[-w:]) / w:.6f} (fp32 oracle), {sum(k[-w:]) / w:.6f} (kernel), "
          f"{sum(a[-w:]) / w:.6f} (fp16 autocast oracle)", "",
          f"| steps | kernel (mean loss) | fp32 oracle | fp16 autocast oracle | kernel / fp32 − 1 | fp16 / fp32 − 1 |", "|---|---|---|---|---|---|"]
    for i in range(0, steps, w):
        mk, mf, ma = (sum(x[i:i + w]) / len(x[i:i + w]) for x in (k, f, a))
        md.append(f"| {i}–{min(i + w, steps) - 1} | {mk:.6f} | {mf:.6f} | {ma:.6f} | {mk / mf - 1:+.2e} | {ma / mf - 1:+.2e} |")
    md += ["", "| step | kernel | fp32 oracle | fp16 autocast oracle |", "|---|---|---|---|"]
    for i in list(range(0, min(steps, 16))):
        md.append(f"| {i} | {k[i]:.6f} | {f[i]:.6f} | {a[i]:.6f} |")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", "train_parity.md"), "w").write("\n".join(md) + "\n")
    print("\n".join(md))
