"""metrics.py of the reference: mse / psnr (metrics.py:4-13).  ssim needs kornia in the reference (absent here) and is
not part of the per-ray hot path; it raises."""
import torch


def mse(image_pred, image_gt, valid_mask=None, reduction='mean'):
    value = (image_pred - image_gt) ** 2
    if valid_mask is not None:
        value = value[valid_mask]
    if reduction == 'mean':
        return torch.mean(value)
    return value


def psnr(image_pred, image_gt, valid_mask=None, reduction='mean'):
    return -10 * torch.log10(mse(image_pred, image_gt, valid_mask, reduction))


def psnr_from_sum(sum_sq, count):
    """PSNR from the sum of squared errors that hn_mse_loss already produced (no second pass over the image)."""
    return -10 * torch.log10(sum_sq / count)


def ssim(image_pred, image_gt, reduction='mean'):
    raise NotImplementedError("ssim is kornia's in the reference (metrics.py:15-20) and outside the per-ray hot path")
