"""TEST INFRASTRUCTURE — not product code.  Plain fp32 torch restatement of the reference's STATIC NeRF baseline
(BASELINE.json configs[3]): models/nerf.py (Embedding :4-38, NeRF :41-123) and models/rendering.py (sample_pdf
:14-55, render_rays :58-244), with every random draw passed in explicitly.  Only tests/, oracle/make_golden.py and
bench legs may import it.

Pinning: tests/test_static_oracle.py checks it against the UNMODIFIED reference imported from /root/reference
(oracle/ref_loader.py; `torchsearchsorted`, a third-party extension the reference imports but does not ship or pin —
SURVEY.md §8(c) — is stubbed with torch.searchsorted(right=True), the semantics of its call site rendering.py:42) and
against tests/golden/static_*.pt generated from that reference by oracle/make_golden.py.
"""
import torch
import torch.nn.functional as F


def embed(x, n_freqs):
    """nerf.py:21-38, logscale bands 2**k: [x, sin(2^0 x), cos(2^0 x), ...]."""
    out = [x]
    for k in range(n_freqs):
        f = float(2 ** k)
        out += [torch.sin(f * x), torch.cos(f * x)]
    return torch.cat(out, -1)


def nerf_forward(sd, xyz_emb, dir_emb, skips=(4,), D=8):
    """nerf.py:84-123 on embedded inputs; sd = state_dict of one NeRF.  Returns rgb (..,3), raw sigma (..,)."""
    h = xyz_emb
    for i in range(D):
        if i in skips:
            h = torch.cat([xyz_emb, h], -1)
        h = F.relu(F.linear(h, sd[f'xyz_encoding_{i + 1}.0.weight'], sd[f'xyz_encoding_{i + 1}.0.bias']))
    sigma = F.linear(h, sd['sigma.weight'], sd['sigma.bias'])[..., 0]
    final = F.linear(h, sd['xyz_encoding_final.weight'], sd['xyz_encoding_final.bias'])
    d = F.relu(F.linear(torch.cat([final, dir_emb], -1), sd['dir_encoding.0.weight'], sd['dir_encoding.0.bias']))
    rgb = torch.sigmoid(F.linear(d, sd['rgb.0.weight'], sd['rgb.0.bias']))
    return rgb, sigma


def composite(rgbs, sigmas, z_vals, dirs, noise, white_back=False):
    """rendering.py:137-172: last delta 1e10, relu(sigma + noise), cumprod of (1 - alpha + 1e-10)."""
    deltas = z_vals[:, 1:] - z_vals[:, :-1]
    deltas = torch.cat([deltas, 1e10 * torch.ones_like(deltas[:, :1])], -1)
    deltas = deltas * torch.norm(dirs.unsqueeze(1), dim=-1)
    alphas = 1 - torch.exp(-deltas * torch.relu(sigmas + noise))
    shifted = torch.cat([torch.ones_like(alphas[:, :1]), 1 - alphas + 1e-10], -1)
    weights = alphas * torch.cumprod(shifted, -1)[:, :-1]
    rgb = torch.sum(weights.unsqueeze(-1) * rgbs, -2)
    depth = torch.sum(weights * z_vals, -1)
    if white_back:
        rgb = rgb + 1 - weights.sum(1).unsqueeze(-1)
    return rgb, depth, weights


def sample_pdf(bins, weights, u, eps=1e-5):
    """rendering.py:14-55 with the draws u (N_rays, N_importance) given."""
    n = weights.shape[1]
    weights = weights + eps
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[:, :1]), cdf], -1)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp_min(inds - 1, 0)
    above = torch.clamp_max(inds, n)
    g = torch.stack([below, above], -1).view(u.shape[0], -1)
    cdf_g = torch.gather(cdf, 1, g).view(*u.shape, 2)
    bins_g = torch.gather(bins, 1, g).view(*u.shape, 2)
    denom = cdf_g[..., 1] - cdf_g[..., 0]
    denom = torch.where(denom < eps, torch.ones_like(denom), denom)
    return bins_g[..., 0] + (u - cdf_g[..., 0]) / denom * (bins_g[..., 1] - bins_g[..., 0])


def render_rays(sds, rays, draws, n_samples=64, n_importance=64, perturb=1.0, noise_std=1.0, white_back=False,
                xyz_freqs=10, dir_freqs=4, z_fine_override=None):
    """rendering.py:174-244.  sds = [coarse_sd, fine_sd]; draws = dict(u_perturb (N,Nc), noise_coarse (N,Nc),
    u_pdf (N,Nf), noise_fine (N,Nc+Nf)) (u_pdf may be None for the deterministic linspace of perturb == 0).
    z_fine_override: evaluate the fine level at these sorted depths instead of the resampled ones (used by the GPU tests
    to separate the fine MLP / compositing error from the error inherited through resampling)."""
    o, d, near, far = rays[:, 0:3], rays[:, 3:6], rays[:, 6:7], rays[:, 7:8]
    N = rays.shape[0]
    dir_emb = embed(d, dir_freqs)
    steps = torch.linspace(0, 1, n_samples, device=rays.device)
    z = (near * (1 - steps) + far * steps).expand(N, n_samples)
    if perturb > 0:
        mid = 0.5 * (z[:, :-1] + z[:, 1:])
        upper = torch.cat([mid, z[:, -1:]], -1)
        lower = torch.cat([z[:, :1], mid], -1)
        z = lower + (upper - lower) * (perturb * draws['u_perturb'])
    out = {}

    def level(sd, zv, noise, tag):
        pts = o.unsqueeze(1) + d.unsqueeze(1) * zv.unsqueeze(2)
        S = zv.shape[1]
        rgb, sigma = nerf_forward(sd, embed(pts, xyz_freqs), dir_emb[:, None].expand(N, S, dir_emb.shape[-1]))
        c, dep, w = composite(rgb, sigma, zv, d, noise * noise_std, white_back)
        out['rgb_' + tag], out['depth_' + tag], out['opacity_' + tag] = c, dep, w.sum(1)
        return w

    w_c = level(sds[0], z, draws['noise_coarse'], 'coarse')
    out['z_coarse'] = z
    if n_importance > 0:
        mid = 0.5 * (z[:, :-1] + z[:, 1:])
        u = draws.get('u_pdf')
        if u is None:
            u = torch.linspace(0, 1, n_importance, device=rays.device).expand(N, n_importance)
        z_new = sample_pdf(mid, w_c[:, 1:-1], u).detach()
        z_f, _ = torch.sort(torch.cat([z, z_new], -1), -1)
        if z_fine_override is not None:   # stage-isolated parity: evaluate the fine level at given depths
            z_f = z_fine_override
        level(sds[1], z_f, draws['noise_fine'], 'fine')
        out['z_fine'] = z_f
    return out
