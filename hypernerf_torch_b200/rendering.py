"""Static NeRF rendering (reference: models/rendering.py:58-244), B200 path: same signature and result keys as the
reference's `render_rays`; sampling, MLP (fused tcgen05 kernels) and compositing (warp-scan kernels) run on the GPU
through libhypernerf_b200.so.  No CPU fallback."""
import torch

from . import _lib
from . import model_utils as mu
from .nerf import NeRF

__all__ = ['render_rays', 'sample_pdf']


def sample_pdf(bins, weights, N_importance, det=False, eps=1e-5):
    """models/rendering.py:14-55 as a stand-alone utility (torch ops; the reference calls the third-party
    `torchsearchsorted.searchsorted(cdf, u, side='right')` here, i.e. torch.searchsorted(right=True)).  render_rays does
    not use it: it runs the whole resampling step in the hn_sample_pdf kernel."""
    N_rays, N_samples_ = weights.shape
    weights = weights + eps
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[:, :1]), cdf], -1)
    if det:
        u = torch.linspace(0, 1, N_importance, device=bins.device).expand(N_rays, N_importance)
    else:
        u = torch.rand(N_rays, N_importance, device=bins.device)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp_min(inds - 1, 0)
    above = torch.clamp_max(inds, N_samples_)
    inds_sampled = torch.stack([below, above], -1).view(N_rays, 2 * N_importance)
    cdf_g = torch.gather(cdf, 1, inds_sampled).view(N_rays, N_importance, 2)
    bins_g = torch.gather(bins, 1, inds_sampled).view(N_rays, N_importance, 2)
    denom = cdf_g[..., 1] - cdf_g[..., 0]
    denom[denom < eps] = 1
    return bins_g[..., 0] + (u - cdf_g[..., 0]) / denom * (bins_g[..., 1] - bins_g[..., 0])


def render_rays(models, embeddings, rays, N_samples=64, use_disp=False, perturb=0, noise_std=1, N_importance=0,
                chunk=1024 * 32, white_back=False, test_time=False, _taps=None):
    """Same contract as models/rendering.py:58-244.  `embeddings` only fixes the frequency counts (10 / 4, checked);
    `chunk` is accepted for signature compatibility (the fused kernels tile the samples themselves).  `_taps`: optional
    dict that receives the sample depths z_coarse / z_fine (test hook; the reference does not return them)."""
    if not rays.is_cuda:
        raise _lib.NativeLibraryError("render_rays needs CUDA tensors (no CPU path)")
    if embeddings[0].N_freqs != 10 or embeddings[1].N_freqs != 4:
        raise NotImplementedError("the kernels are instantiated for Embedding(3, 10) / Embedding(3, 4)")
    for m in models:
        if not isinstance(m, NeRF):
            raise TypeError("models must be hypernerf_torch_b200.nerf.NeRF instances")
    N_rays = rays.shape[0]
    rays_o, rays_d = rays[:, 0:3].contiguous(), rays[:, 3:6].contiguous()
    near, far = rays[:, 6:7], rays[:, 7:8]
    flags = _lib.HN_COMP_ACC_ALL | (_lib.HN_COMP_WHITE_BKGD if white_back else 0)

    def inference(model, xyz, z_vals, weights_only=False):
        # rendering.py:145: noise = randn(sigmas.shape) * noise_std is drawn even when noise_std == 0
        noise = torch.randn(z_vals.shape, device=rays.device)
        sigma, rgb = model.query(xyz, rays_d, noise, float(noise_std))
        out_rgb, depth, _, acc, weights, _ = mu._Composite.apply(rgb, sigma, z_vals, rays_d, flags, 1e-10, 1e10)
        if weights_only:
            return weights
        return out_rgb, depth, weights

    z_steps = torch.linspace(0, 1, N_samples, device=rays.device)
    if not use_disp:
        z_vals = near * (1 - z_steps) + far * z_steps
    else:
        z_vals = 1 / (1 / near * (1 - z_steps) + 1 / far * z_steps)
    z_vals = z_vals.expand(N_rays, N_samples)
    if perturb > 0:
        z_vals_mid = 0.5 * (z_vals[:, :-1] + z_vals[:, 1:])
        upper = torch.cat([z_vals_mid, z_vals[:, -1:]], -1)
        lower = torch.cat([z_vals[:, :1], z_vals_mid], -1)
        perturb_rand = perturb * torch.rand(z_vals.shape, device=rays.device)
        z_vals = lower + (upper - lower) * perturb_rand
    z_vals = z_vals.contiguous()
    xyz_coarse = rays_o.unsqueeze(1) + rays_d.unsqueeze(1) * z_vals.unsqueeze(2)
    if _taps is not None:
        _taps['z_coarse'] = z_vals.detach().clone()

    if test_time:
        weights_coarse = inference(models[0], xyz_coarse, z_vals, weights_only=True)
        result = {'opacity_coarse': weights_coarse.sum(1)}
    else:
        rgb_coarse, depth_coarse, weights_coarse = inference(models[0], xyz_coarse, z_vals)
        result = {'rgb_coarse': rgb_coarse, 'depth_coarse': depth_coarse, 'opacity_coarse': weights_coarse.sum(1)}

    if N_importance > 0:
        # rendering.py:223-233 in one launch (hn_sample_pdf): z_vals_mid, weights_coarse[:, 1:-1], inverse-CDF samples
        # (searchsorted right=True), sort(cat([z_vals, samples])) and the fine sample points; detached like the reference
        u = None if perturb > 0 else torch.linspace(0, 1, N_importance, device=rays.device).expand(N_rays, N_importance)
        z_vals, xyz_fine = mu.sample_pdf_fused(z_vals, weights_coarse.detach(), rays_o, rays_d, N_importance, u=u)
        if _taps is not None:
            _taps['z_fine'] = z_vals.detach().clone()
        rgb_fine, depth_fine, weights_fine = inference(models[1], xyz_fine, z_vals.contiguous())
        result['rgb_fine'] = rgb_fine
        result['depth_fine'] = depth_fine
        result['opacity_fine'] = weights_fine.sum(1)
    return result
