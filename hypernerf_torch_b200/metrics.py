"""Image metrics with the reference's names and arguments (metrics.py:4-13): `mse`, `psnr`; plus `psnr_from_sum`, which
turns the squared-error sum that hn_mse_loss already produced into a PSNR without a second pass over the image.  `ssim`
is kornia's in the reference (metrics.py:15-20), outside the per-ray hot path, and not built."""
import torch


def mse(image_pred, image_gt, valid_mask=None, reduction='mean'):
    sq = torch.square(image_pred - image_gt)
    sq = sq if valid_mask is None else sq[valid_mask]
    return sq.mean() if reduction == 'mean' else sq


def psnr(image_pred, image_gt, valid_mask=None, reduction='mean'):
    return torch.log10(mse(image_pred, image_gt, valid_mask, reduction)) * -10.0


def psnr_from_sum(sum_sq, count):
    return torch.log10(sum_sq / count) * -10.0


def ssim(image_pred, image_gt, reduction='mean'):
    raise NotImplementedError("ssim is not part of the B200 hot path (the reference delegates it to kornia)")
