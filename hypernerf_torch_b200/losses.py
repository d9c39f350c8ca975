"""losses.py of the reference (MSELoss, loss_dict) on the B200 path: one fused kernel (hn_mse_loss) computes
MSE(coarse.rgb) + MSE(fine.rgb), the gradient seed of both levels and the fine-level MSE that metrics.psnr needs."""
import torch
from torch import nn

from . import _lib
from ._lib import check, lib, ptr, stream


class _FusedMSE(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, coarse, fine, targets, inv_count):
        c = coarse.contiguous()
        f = None if fine is None else fine.contiguous()
        t = targets.to(torch.float32).contiguous()
        B = c.shape[0]
        sums = torch.zeros(2, device=c.device, dtype=torch.float32)
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        gc = torch.empty_like(c) if need else None
        gf = torch.empty_like(f) if (need and f is not None) else None
        check(lib().hn_mse_loss(ptr(c), ptr(f), ptr(t), B, float(inv_count), ptr(sums), ptr(gc), ptr(gf), stream()),
              "hn_mse_loss")
        _lib.count(1)
        ctx.save_for_backward(gc, gf)
        ctx.mark_non_differentiable(sums)
        return (sums[0] + sums[1]) * inv_count, sums

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, g_loss, _g_sums):
        gc, gf = ctx.saved_tensors
        return (None if gc is None else gc * g_loss, None if gf is None else gf * g_loss, None, None)


def mse_coarse_fine(inputs, targets, global_count=None):
    """Returns (loss, sums): loss = (sum_sq(coarse) + sum_sq(fine)) / global_count with global_count = 3 * B by default
    (losses.py:9-14), sums = per-level sums of squared errors (device tensor, no gradient)."""
    coarse = inputs['coarse']['rgb']
    fine = inputs['fine']['rgb'] if 'fine' in inputs else None
    n = targets.numel() if global_count is None else global_count
    return _FusedMSE.apply(coarse, fine, targets, 1.0 / float(n))


class MSELoss(nn.Module):
    """losses.py:4-14: MSE(inputs['coarse']['rgb'], targets) [+ MSE(inputs['fine']['rgb'], targets)], mean reduction."""

    def __init__(self):
        super().__init__()

    def forward(self, inputs, targets):
        return mse_coarse_fine(inputs, targets)[0]


loss_dict = {'mse': MSELoss}
