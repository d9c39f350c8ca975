"""TEST INFRASTRUCTURE — not product code.  A plain fp32 torch restatement (CPU or GPU tensors) of the reference's
per-ray HyperNeRF hot path, function by function, with every random draw passed in explicitly.  Only tests/,
oracle/make_golden.py, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may import it; the
product path (hypernerf_torch_b200/) never does and fails loudly without its CUDA library.

Pinning: tests/test_oracle.py checks this restatement against the UNMODIFIED reference imported from
/root/reference (oracle/ref_loader.py) on identical weights / rays / draws, and against the committed golden vectors
in tests/golden/ that oracle/make_golden.py generated from that reference.  The reference ships no tests or golden
vectors of its own, so those two checks are the pin.

All citations are file:line into songrise/HyperNeRF-torch.
"""
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------------------
# primitives
# ------------------------------------------------------------------------------------------------------------
def posenc_orig(x, n_freqs):
    """hypernerf/model_utils.py:234-246: [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(F-1) x), cos(2^(F-1) x)]."""
    out = [x]
    for k in range(n_freqs):
        f = float(2 ** k)
        out += [torch.sin(f * x), torch.cos(f * x)]
    return torch.cat(out, -1)


def posenc_scaled(x, min_deg, max_deg):
    """hypernerf/model_utils.py:255-273 with use_identity=False: scales 2**linspace(min_deg, max_deg, max_deg - min_deg) (NOT
    integer octaves: 1, 2.208, ..., 256 for (0, 8)), cos as sin(x + 0.5 * 3.1415926), the alpha window commented out in the
    reference (:262-265).  Channel order (*, F, 2, C) flattened: per scale [sin(s x_0..2), sin(s x_0..2 + pi/2)]."""
    scales = 2. ** torch.linspace(min_deg, max_deg, steps=max_deg - min_deg, device=x.device)
    xb = x[..., None, :] * scales[:, None]
    four = torch.sin(torch.stack((xb, xb + 0.5 * 3.1415926), dim=-2))
    return four.reshape(*x.shape[:-1], -1)


def _skew(w):
    """rigid_body.py:24-38 (Modern Robotics Eq. 3.30), batched."""
    z = torch.zeros_like(w[..., 0])
    return torch.stack([torch.stack([z, -w[..., 2], w[..., 1]], -1),
                        torch.stack([w[..., 2], z, -w[..., 0]], -1),
                        torch.stack([-w[..., 1], w[..., 0], z], -1)], -2)


def se3_transform(points, w, v):
    """SE3Field.warp's tail (warping.py:229-238) with the closed forms of rigid_body.py:55-83, batched over the leading
    dimensions: theta = |w|; the screw axis is (w, v) / theta; R = I + sin(theta) W + (1 - cos(theta)) W^2 (Rodrigues);
    p = (theta I + (1 - cos(theta)) W + (theta - sin(theta)) W^2) v (Modern Robotics Eq. 3.88); warped = R x + p (the
    homogeneous divide of rigid_body.py:91-93 is by 1).  RESTATEMENT, parity unpinned by the reference: rigid_body.skew only
    accepts one point, returns constants and severs autograd, so this is what the reference code
    says, batched, not what it runs.  Like the reference there is no epsilon: theta == 0 gives NaN here (the CUDA path
    evaluates the same map through its removable-singularity form and returns x + v there)."""
    theta = torch.norm(w, dim=-1)
    wh = w / theta[..., None]
    vh = v / theta[..., None]
    W = _skew(wh)
    th = theta[..., None, None]
    eye = torch.eye(3, dtype=points.dtype, device=points.device)
    WW = W @ W
    R = eye + torch.sin(th) * W + (1.0 - torch.cos(th)) * WW
    G = th * eye + (1.0 - torch.cos(th)) * W + (th - torch.sin(th)) * WW
    return (R @ points[..., None])[..., 0] + (G @ vh[..., None])[..., 0]


def se3_field(sd, points, q=None, gates=None):
    """SE3Field.warp, warping.py:212-240: trunk MLP (depth 6, width 128, skip 4, out 128 without activation) on
    posenc(points, 0, 8) — the metadata embedding is NOT an input (:223-224) —, w_net / v_net (depth 0 => ONE hidden 128 ReLU
    layer + 3 outputs each, modules.py:95-98), exp map.  Returns (warped xyz, w, v)."""
    g = gates or {}
    q = q or _ident
    feat = q(posenc_scaled(points, 0, 8))
    trunk = q(mlp(sd, "warp_field.trunk", feat, 6, q=q, gates=g.get('warp')))
    w = mlp(sd, "warp_field.w_net", trunk, 1, skips=(), q=q, gates=g.get('w_net'))
    v = mlp(sd, "warp_field.v_net", trunk, 1, skips=(), q=q, gates=g.get('v_net'))
    return se3_transform(points, w, v), w, v


def _ident(x):
    return x


def bf16_ste(x):
    """Round to bf16 with a straight-through gradient.  Used by the `emulate_bf16` mode, which places a rounding at
    exactly the points where the sm_100a kernels hold bf16 tensor-core operands (weights, layer inputs, post-ReLU
    activations, bottleneck), everything else fp32 — the north-star numerics (bf16 operands, fp32 accumulate)."""
    return x + (x.to(torch.bfloat16).to(x.dtype) - x).detach()


def _relu(x, gate):
    """ReLU, or — for tests that must share the kernels' ReLU gates exactly — multiplication by a given 0/1 gate."""
    return F.relu(x) if gate is None else x * gate.to(x.dtype)


def mlp(sd, prefix, x, depth, skips=(4,), out_act=None, q=_ident, gates=None):
    """hypernerf/modules.py:114-127: x = relu(linear_i(x)); concat [x, inputs] after layer i in skips; logit layer.
    x must already be q()-rounded by the caller.  gates: optional list of depth (+1 if out_act is relu) masks."""
    inputs = x
    for i in range(depth):
        pre = F.linear(x, q(sd[f"{prefix}.linears.{i}.weight"]), sd[f"{prefix}.linears.{i}.bias"])
        x = q(_relu(pre, None if gates is None else gates[i]))
        if i in skips:
            x = torch.cat([x, inputs], -1)
    x = F.linear(x, q(sd[f"{prefix}.logit_layer.weight"]), sd[f"{prefix}.logit_layer.bias"])
    if out_act is F.relu:
        return _relu(x, None if gates is None else gates[depth])
    return out_act(x) if out_act is not None else x


def nerf_mlp(sd, prefix, feat, rgb_cond, q=_ident, gates=None, alpha_cond=None):
    """hypernerf/modules.py:266-298: trunk (ReLU on the logit layer, :230), bottleneck (no activation), alpha Linear on
    [bottleneck | alpha condition] (:280-286), rgb MLP depth 4 (skip never fires) on [bottleneck | rgb condition] (:290-296)
    + Sigmoid (models.py:164)."""
    g = gates or {}
    x = q(mlp(sd, f"{prefix}.trunk_mlp", q(feat), 8, out_act=F.relu, q=q, gates=g.get('trunk')))
    bott = q(F.linear(x, q(sd[f"{prefix}.bottleneck_mlp.weight"]), sd[f"{prefix}.bottleneck_mlp.bias"]))
    alpha_in = bott
    if alpha_cond is not None:
        alpha_in = torch.cat([bott, q(alpha_cond)[:, None, :].expand(-1, bott.shape[1], -1)], -1)
    alpha = F.linear(alpha_in, q(sd[f"{prefix}.alpha_mlp.weight"]), sd[f"{prefix}.alpha_mlp.bias"])
    cond = q(rgb_cond)[:, None, :].expand(-1, bott.shape[1], -1)  # broadcast_condition, modules.py:254-264
    rgb = mlp(sd, f"{prefix}.rgb_mlp", torch.cat([bott, cond], -1), 4, out_act=torch.sigmoid, q=q, gates=g.get('rgb'))
    return rgb, alpha


def sample_along_rays(origins, directions, n, near, far, u):
    """hypernerf/model_utils.py:6-41 (stratified when u is given)."""
    t = torch.linspace(0., 1., n, device=origins.device)
    z = near * (1. - t) + far * t
    if u is not None:
        mids = .5 * (z[..., 1:] + z[..., :-1])
        upper = torch.cat([mids, z[..., -1:]], -1)
        lower = torch.cat([z[..., :1], mids], -1)
        z = lower + (upper - lower) * u
    else:
        z = z[None].expand(origins.shape[0], n)
    return z, origins[:, None, :] + z[..., None] * directions[:, None, :]


def volumetric_rendering(rgb, sigma, z, dirs, white_bkgd=False, sample_at_infinity=True, eps=1e-5):
    """hypernerf/model_utils.py:43-107 + compute_opaqueness_mask / depth index / depth map (:319-362)."""
    last = torch.full_like(z[..., :1], 1e7 if sample_at_infinity else 1e-7)
    dists = torch.cat([z[..., 1:] - z[..., :-1], last], -1) * torch.norm(dirs[:, None, :], dim=-1)
    alpha = 1.0 - torch.exp(-sigma * dists)
    T = torch.cat([torch.ones_like(alpha[..., :1]), torch.cumprod(1.0 - alpha[..., :-1] + eps, -1)], -1)
    w = alpha * T
    out_rgb = (w[..., None] * rgb).sum(-2)
    depth = (w * z).sum(-1)
    acc = w.sum(-1)
    if white_bkgd:
        out_rgb = out_rgb + (1. - acc[..., None])
    if sample_at_infinity:
        acc = w[..., :-1].sum(-1)
    opaque = torch.cumsum(w, -1) >= 0.5
    mask = torch.logical_xor(opaque, torch.cat([torch.zeros_like(opaque[..., :1]), opaque[..., :-1]], -1)).to(w.dtype)
    return {'rgb': out_rgb, 'depth': depth, 'med_depth': (mask * z).sum(-1), 'acc': acc, 'weights': w,
            'med_idx': torch.argmax(mask, -1)}


def piecewise_constant_pdf(bins, weights, u):
    """hypernerf/model_utils.py:160-204 with the arithmetic contract of DESIGN.md made explicit:
    S = fl32(sum w') and cdf_j = fl32(sum_{k<=j} pdf_k) with the sums carried in fp64 (what torch.cumsum does on
    CPU); every other op is one fp32 rounding.  Returns (samples, inds)."""
    eps = 1e-5
    nb = weights.shape[-1]
    w = weights + eps
    S = w.double().sum(-1, keepdim=True).float()
    pdf = w / S
    cdf = torch.cumsum(pdf.double(), -1).float()
    cdf = torch.cat([torch.zeros_like(cdf[:, :1]), cdf], -1)
    inds = torch.searchsorted(cdf, u.contiguous(), right=True)
    below = torch.clamp_min(inds - 1, 0)
    above = torch.clamp_max(inds, nb)
    c0, c1 = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    b0, b1 = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = c1 - c0
    denom = torch.where(denom < eps, torch.ones_like(denom), denom)
    samples = b0 + (u - c0) / denom * (b1 - b0)
    return samples.detach(), inds


def sample_pdf(bins, weights, origins, directions, z, u):
    """hypernerf/model_utils.py:206-232."""
    samples, inds = piecewise_constant_pdf(bins, weights, u)
    z_all, _ = torch.sort(torch.cat([z, samples], -1), -1)
    return z_all, origins[:, None, :] + z_all[..., None] * directions[:, None, :], inds


# ------------------------------------------------------------------------------------------------------------
# model
# ------------------------------------------------------------------------------------------------------------
def query_fields(sd, level, points, viewdirs, ids, cfg, noise=None, gates=None):
    """map_points + query_template (hypernerf/models.py:545-581, 447-493).  cfg['slice'] in {'bendy_sheet' (default),
    'axis_aligned_plane', 'none'}, cfg['use_warp'] (default True), cfg['cond_alpha'] / cfg['cond_rgb'] (template GLO
    conditioning, models.py:404-445), cfg['warp_field'] in {'translation' (what models.py:234 hard-codes), 'se3' (config 5: the
    SE3Field the reference defines but never instantiates — restated, see se3_transform)}.  ids: dict of metadata id rows {'time', 'warp'} or a single row used for both.
    Returns (rgb (B,S,3), sigma (B,S), warped_points (B,S,3[+H]), raw alpha)."""
    B, S, _ = points.shape
    q = bf16_ste if cfg.get('emulate_bf16') else _ident
    use_warp = cfg.get('use_warp', True)
    slice_method = cfg.get('slice', 'bendy_sheet')
    if not isinstance(ids, dict):
        ids = {'time': ids, 'warp': ids}
    g = gates or {}
    if use_warp:
        embed_b = sd["warp_embed.embed.weight"][ids['time'].reshape(-1)]   # GLOEmbed, modules.py:155-167; key 'time'
        embed = embed_b[:, None, :].expand(B, S, embed_b.shape[-1])        # models.py:627-632
        if cfg.get('warp_field', 'translation') == 'se3':
            warped, _, _ = se3_field(sd, points, q=q, gates=g)
        else:
            # TranslationField.warp, warping.py:90-96 (n_freq hard-coded 10)
            warp_in = q(torch.cat([posenc_orig(points, 10), embed], -1))
            warped = points + mlp(sd, "warp_field.mlp", warp_in, 6, q=q, gates=g.get('warp'))
        if slice_method == 'bendy_sheet':
            # HyperSheetMLP on the UNWARPED points, modules.py:331-337 (n_freq 7), models.py:571-572
            sheet_in = q(torch.cat([posenc_orig(points, 7), embed], -1))
            hyper = mlp(sd, "hyper_sheet_mlp.mlp", sheet_in, 6, q=q, gates=g.get('sheet'))
        else:
            hyper = embed            # axis_aligned_plane: hyper_points = hyper_embed = warp_embed (models.py:533-534, 618-619)
        warped_points = torch.cat([warped, hyper], -1)
        feat = torch.cat([posenc_orig(warped_points[..., :3], cfg['xyz_freq']),
                          posenc_orig(warped_points[..., 3:], cfg['hyper_freq'])], -1)   # models.py:458-478
    else:
        warped_points = points       # map_points, models.py:568-569
        feat = posenc_orig(points, cfg['xyz_freq'])
    rgb_cond = posenc_orig(viewdirs, cfg['view_freq'])                                # models.py:410-419
    alpha_cond = None
    if cfg.get('cond_alpha') or cfg.get('cond_rgb'):
        # get_condition_inputs, models.py:421-434: warp_embed[metadata['time']] with a warp, nerf_embed[metadata['warp']] without
        nerf_embed = embed_b if use_warp else sd["nerf_embed.embed.weight"][ids['warp'].reshape(-1)]
        if cfg.get('cond_alpha'):
            alpha_cond = nerf_embed
        if cfg.get('cond_rgb'):
            rgb_cond = torch.cat([rgb_cond, nerf_embed], -1)
    prefix = "nerf_mlps_fine" if level == 'fine' else "nerf_mlps_coarse"
    rgb, alpha = nerf_mlp(sd, prefix, feat, rgb_cond, q=q, gates=gates, alpha_cond=alpha_cond)
    if noise is not None:
        alpha = alpha + noise * cfg['noise_std']                                      # model_utils.py:312-316
    sigma = F.softplus(alpha.squeeze(-1))                                             # models.py:491
    return rgb, sigma, warped_points, alpha.squeeze(-1)


def filter_sigma(points, sigma, render_opts):
    """hypernerf/models.py:35-63."""
    if render_opts is None:
        return sigma
    if 'dust_threshold' in render_opts:
        sigma = (sigma >= render_opts.get('dust_threshold', 0.0)) * sigma
    if 'bounding_box' in render_opts:
        xmin, xmax, ymin, ymax, zmin, zmax = render_opts['bounding_box']
        mask = ((points[..., 0] >= xmin) & (points[..., 0] <= xmax) & (points[..., 1] >= ymin) & (points[..., 1] <= ymax)
                & (points[..., 2] >= zmin) & (points[..., 2] <= zmax))
        sigma = mask * sigma
    return sigma


def render_samples(sd, level, points, z, directions, viewdirs, ids, cfg, noise=None, render_opts=None):
    """hypernerf/models.py:587-671."""
    rgb, sigma, warped_points, alpha = query_fields(sd, level, points, viewdirs, ids, cfg, noise)
    sigma = filter_sigma(points, sigma, render_opts)                                  # models.py:650
    out = {'points': points, 'warped_points': warped_points, 'sigma': sigma, 'rgb_samples': rgb}
    out.update(volumetric_rendering(rgb, sigma, z, directions))
    out['med_points'] = torch.gather(warped_points, -2, out['med_idx'][..., None, None])   # models.py:664-669
    return out


def forward(sd, origins, directions, ids, draws, cfg, fine_z=None, render_opts=None):
    """hypernerf/models.py:673-780.  draws: dict(u_coarse (B,Nc), noise_coarse (B,Nc,1)|None, u_fine (B,Nf),
    noise_fine (B,Nc+Nf,1)|None) in the reference's RNG order.
    cfg: dict(near, far, n_coarse, n_fine, noise_std, xyz_freq, hyper_freq, view_freq).
    fine_z: stage-isolation hook for tests — evaluate the fine level at these depths instead of the resampled ones."""
    z, points = sample_along_rays(origins, directions, cfg['n_coarse'], cfg['near'], cfg['far'], draws['u_coarse'])
    coarse = render_samples(sd, 'coarse', points, z, directions, directions, ids, cfg, draws.get('noise_coarse'))
    coarse['z_vals'] = z
    z_mid = .5 * (z[..., 1:] + z[..., :-1])
    z_f, points_f, inds = sample_pdf(z_mid, coarse['weights'][..., 1:-1].detach(), origins, directions, z,
                                     draws['u_fine'])
    if fine_z is not None:
        z_f = fine_z
        points_f = origins[:, None, :] + z_f[..., None] * directions[:, None, :]
    fine = render_samples(sd, 'fine', points_f, z_f, directions, directions, ids, cfg, draws.get('noise_fine'),
                          render_opts=render_opts)                                    # fine level only, models.py:768
    fine['z_vals'] = z_f
    fine['pdf_inds'] = inds
    return {'coarse': coarse, 'fine': fine}


def default_cfg(n_fine=64, noise_std=1.0, emulate_bf16=False, **over):
    cfg = dict(near=0., far=1., n_coarse=64, n_fine=n_fine, noise_std=noise_std, xyz_freq=10, hyper_freq=6,
               view_freq=6, emulate_bf16=emulate_bf16, use_warp=True, slice='bendy_sheet', cond_alpha=False, cond_rgb=False,
               warp_field='translation')
    cfg.update(over)
    return cfg


def cfg_from_kwargs(kw, emulate_bf16=False):
    """Oracle configuration for a NerfModel constructor-argument dict (reference signature, models.py:111-127)."""
    nerf_embed = kw.get('use_nerf_embed', True)
    return default_cfg(n_fine=kw.get('n_samples_fine', 128), noise_std=kw.get('noise_std') or 0.0, emulate_bf16=emulate_bf16,
                       near=kw.get('near', 0.), far=kw.get('far', 1.), n_coarse=kw.get('n_samples_coarse', 64),
                       xyz_freq=kw.get('xyz_fourier_dim', 10), hyper_freq=kw.get('hyper_fourier_dim', 6),
                       view_freq=kw.get('view_fourier_dim', 4), use_warp=kw.get('use_warp', True),
                       slice=kw.get('hyper_slice_method') or 'none',
                       cond_alpha=nerf_embed and kw.get('use_alpha_cond', True),
                       cond_rgb=nerf_embed and kw.get('use_rgb_cond', False),
                       warp_field=kw.get('warp_field_type', 'translation'))


def mse_loss(out, target):
    """losses.py:9-14."""
    return F.mse_loss(out['coarse']['rgb'], target) + F.mse_loss(out['fine']['rgb'], target)
