"""Host-side mirror of hypernerf/model_utils.py: same function names, argument meaning and return values, with
the arithmetic done by the sm_100a kernels of libhypernerf_b200.so (no CPU or torch fallback).

    sample_along_rays      model_utils.py:6-41     -> hn_sample_coarse
    volumetric_rendering   model_utils.py:43-107   -> hn_composite_fwd / hn_composite_bwd (autograd)
    sample_pdf             model_utils.py:206-232  -> hn_sample_pdf
    compute_depth_index    model_utils.py:342-345  -> `_med_idx` output of volumetric_rendering(_return_index=True)
    prepare_ray_dict / extract_rays_batch / append_batch / concat_ray_batch: boundary glue (model_utils.py:365-461),
    same semantics, pure tensor slicing.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import check, lib, ptr, stream


def _f32c(t):
    return t.detach().to(torch.float32).contiguous()


# ------------------------------------------------------------------------------------------------------------
# sampling
# ------------------------------------------------------------------------------------------------------------
def sample_along_rays(origins, directions, num_coarse_samples, near, far, use_stratified_sampling,
                      use_linear_disparity):
    """Stratified sampling along the rays (model_utils.py:6-41).  Returns (z_vals (B,Nc), points (B,Nc,3))."""
    B = origins.shape[0]
    dev = origins.device
    o, d = _f32c(origins), _f32c(directions)
    z = torch.empty(B, num_coarse_samples, device=dev, dtype=torch.float32)
    pts = torch.empty(B, num_coarse_samples, 3, device=dev, dtype=torch.float32)
    lower, upper, depths = _stratum_bounds(num_coarse_samples, float(near), float(far), bool(use_linear_disparity), dev)
    if use_stratified_sampling:
        t_rand = torch.rand([B, num_coarse_samples], device=dev)
        check(lib().hn_sample_coarse(ptr(o), ptr(d), ptr(t_rand), ptr(lower), ptr(upper), B, num_coarse_samples,
                                     ptr(z), ptr(pts), stream()), "hn_sample_coarse")
        _lib.count(1)
    else:
        check(lib().hn_sample_coarse(ptr(o), ptr(d), None, ptr(depths), ptr(depths), B, num_coarse_samples, ptr(z), ptr(pts),
                                     stream()), "hn_sample_coarse")
        _lib.count(1)
    return z, pts


_bounds_cache = {}


def _stratum_bounds(n, near, far, disparity, dev):
    """(lower, upper) stratum bounds of model_utils.py:25-33 (non-stratified: lower = the depths themselves), computed once
    per (n, near, far, device) with the reference's own torch expressions (bit-identical linspace) and cached: they do not
    depend on the rays."""
    key = (n, near, far, disparity, str(dev))
    hit = _bounds_cache.get(key)
    if hit is None:
        t_vals = torch.linspace(0., 1., n, device=dev)
        if not disparity:
            z_vals = near * (1. - t_vals) + far * t_vals
        else:
            z_vals = 1. / (1. / near * (1. - t_vals) + 1. / far * t_vals)
        mids = .5 * (z_vals[..., 1:] + z_vals[..., :-1])
        upper = torch.cat([mids, z_vals[..., -1:]], dim=-1).contiguous()
        lower = torch.cat([z_vals[..., :1], mids], dim=-1).contiguous()
        hit = (lower, upper, z_vals.contiguous())
        _bounds_cache[key] = hit
    return hit


def _sample_pdf_impl(bins, weights, origins, directions, z_vals, n_new, use_stratified_sampling, want_points=True,
                     want_inds=False, u=None):
    B, nb = weights.shape
    dev = weights.device
    if u is None:
        if use_stratified_sampling:
            u = torch.rand(B, n_new, device=dev)
        else:
            u = torch.linspace(0, 1, n_new, device=dev).expand(B, n_new)
    u = u.contiguous()
    w = weights.detach()
    if w.dtype != torch.float32 or w.stride(-1) != 1:
        w = w.to(torch.float32).contiguous()
    zc = _f32c(z_vals)
    Nc = zc.shape[1]
    b = _f32c(bins)
    assert b.shape[1] == nb + 1, "bins must have one more entry than weights"
    S = Nc + n_new
    z_fine = torch.empty(B, S, device=dev, dtype=torch.float32)
    pts = torch.empty(B, S, 3, device=dev, dtype=torch.float32) if want_points else None
    inds = torch.empty(B, n_new, device=dev, dtype=torch.int32) if want_inds else None
    o = _f32c(origins) if want_points else None
    d = _f32c(directions) if want_points else None
    wptr = C.c_void_p(w.data_ptr())
    check(lib().hn_sample_pdf(ptr(zc), ptr(b), wptr, w.stride(0), ptr(u), ptr(o), ptr(d), B, Nc, nb, n_new, ptr(z_fine),
                              ptr(pts), ptr(inds), stream()), "hn_sample_pdf")
    _lib.count(1)
    return z_fine, pts, inds


def sample_pdf(bins, weights, origins, directions, z_vals, num_coarse_samples, use_stratified_sampling):
    """Hierarchical sampling (model_utils.py:206-232): returns (sorted z_vals (B,Nc+Nf), points (B,Nc+Nf,3))."""
    z_fine, pts, _ = _sample_pdf_impl(bins, weights, origins, directions, z_vals, num_coarse_samples,
                                      use_stratified_sampling)
    return z_fine, pts


def sample_pdf_fused(z_vals, coarse_weights, origins, directions, n_new, u=None, want_inds=False, want_ranks=False):
    """models.py:752-755 in one launch: bins = .5*(z[1:]+z[:-1]) and weights[...,1:-1] are formed in-kernel.
    want_ranks: also return (pos_coarse (B,Nc), pos_new (B,n_new)) int32 — where the merge put every coarse depth and
    the i-th smallest new sample in the sorted row (a permutation of 0..Nc+n_new-1 per ray)."""
    B, Nc = z_vals.shape
    dev = z_vals.device
    if u is None:
        u = torch.rand(B, n_new, device=dev)
    u = u.contiguous()
    zc, w = _f32c(z_vals), _f32c(coarse_weights)
    o, d = _f32c(origins), _f32c(directions)
    S = Nc + n_new
    z_fine = torch.empty(B, S, device=dev, dtype=torch.float32)
    pts = torch.empty(B, S, 3, device=dev, dtype=torch.float32)
    inds = torch.empty(B, n_new, device=dev, dtype=torch.int32) if want_inds else None
    wptr = C.c_void_p(w.data_ptr() + 4)  # weights[..., 1:-1]
    if want_ranks:
        pos_c = torch.empty(B, Nc, device=dev, dtype=torch.int32)
        pos_n = torch.empty(B, n_new, device=dev, dtype=torch.int32)
        check(lib().hn_sample_pdf_ranks(ptr(zc), None, wptr, Nc, ptr(u), ptr(o), ptr(d), B, Nc, Nc - 2, n_new, ptr(z_fine),
                                        ptr(pts), ptr(inds), ptr(pos_c), ptr(pos_n), stream()), "hn_sample_pdf_ranks")
        _lib.count(1)
        ranks = (pos_c, pos_n)     # int32 position tables: consumed as they are by hn_mlp_fwd / hn_mlp_fwd_trunk
        return (z_fine, pts, inds, ranks) if want_inds else (z_fine, pts, ranks)
    check(lib().hn_sample_pdf(ptr(zc), None, wptr, Nc, ptr(u), ptr(o), ptr(d), B, Nc, Nc - 2, n_new, ptr(z_fine),
                              ptr(pts), ptr(inds), stream()), "hn_sample_pdf")
    _lib.count(1)
    return (z_fine, pts, inds) if want_inds else (z_fine, pts)


# ------------------------------------------------------------------------------------------------------------
# compositing
# ------------------------------------------------------------------------------------------------------------
class _Composite(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, rgb, sigma, z_vals, dirs, flags, eps, last_delta):
        B, S = sigma.shape
        dev = sigma.device
        rgb_c, sigma_c, z_c, d_c = _f32c(rgb), _f32c(sigma), _f32c(z_vals), _f32c(dirs)
        out_rgb = torch.empty(B, 3, device=dev, dtype=torch.float32)
        depth = torch.empty(B, device=dev, dtype=torch.float32)
        med_depth = torch.empty(B, device=dev, dtype=torch.float32)
        acc = torch.empty(B, device=dev, dtype=torch.float32)
        weights = torch.empty(B, S, device=dev, dtype=torch.float32)
        med_idx = torch.empty(B, device=dev, dtype=torch.int64)
        check(lib().hn_composite_fwd(ptr(sigma_c), ptr(rgb_c), ptr(z_c), ptr(d_c), B, S, flags, eps, last_delta,
                                     ptr(out_rgb), ptr(depth), ptr(med_depth), ptr(acc), ptr(weights), ptr(med_idx),
                                     stream()), "hn_composite_fwd")
        _lib.count(1)
        ctx.save_for_backward(rgb_c, sigma_c, z_c, d_c)
        ctx.cfg = (flags, eps, last_delta)
        ctx.mark_non_differentiable(med_depth, med_idx)
        ctx.set_materialize_grads(False)
        return out_rgb, depth, med_depth, acc, weights, med_idx

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, g_rgb, g_depth, g_med, g_acc, g_weights, g_idx):
        rgb_c, sigma_c, z_c, d_c = ctx.saved_tensors
        flags, eps, last_delta = ctx.cfg
        B, S = sigma_c.shape
        g_sigma = torch.empty_like(sigma_c)
        g_rgb_s = torch.empty_like(rgb_c)

        keep = [None if g is None else _f32c(g) for g in (g_rgb, g_depth, g_acc, g_weights)]
        check(lib().hn_composite_bwd(ptr(sigma_c), ptr(rgb_c), ptr(z_c), ptr(d_c), B, S, flags, eps, last_delta,
                                     ptr(keep[0]), ptr(keep[1]), ptr(keep[2]), ptr(keep[3]),
                                     ptr(g_sigma), ptr(g_rgb_s), stream()), "hn_composite_bwd")
        _lib.count(1)
        return g_rgb_s, g_sigma, None, None, None, None, None


def volumetric_rendering(rgb, sigma, z_vals, dirs, use_white_background, sample_at_infinity=True, eps=1e-5,
                         _return_index=False):
    """Volumetric rendering (model_utils.py:43-107).  Same dictionary as the reference."""
    flags = (_lib.HN_COMP_WHITE_BKGD if use_white_background else 0) | (0 if sample_at_infinity else _lib.HN_COMP_ACC_ALL)
    last_delta = 1e7 if sample_at_infinity else 1e-7
    out_rgb, depth, med_depth, acc, weights, med_idx = _Composite.apply(rgb, sigma, z_vals, dirs, flags, float(eps),
                                                                       float(last_delta))
    out = {'rgb': out_rgb, 'depth': depth, 'med_depth': med_depth, 'acc': acc, 'weights': weights}
    if _return_index:
        out['_med_idx'] = med_idx
    return out


# ------------------------------------------------------------------------------------------------------------
# boundary glue (pure slicing; identical semantics to model_utils.py:365-461)
# ------------------------------------------------------------------------------------------------------------
def prepare_ray_dict(rays: torch.Tensor) -> dict:
    use_meta = rays.shape[-1] == 9
    if len(rays.shape) > 2:
        rays = rays.view(-1, 8)
    B = rays.shape[0]
    if B == 0:   # the reference reads rays[0, 6] / rays[0, 7] here (model_utils.py:389-390): same exception for an empty batch
        raise IndexError("index 0 is out of bounds for dimension 0 with size 0")
    idx = torch.ones((B, 1), dtype=torch.long, device=rays.device)
    if use_meta:
        idx = rays[:, 8].type(torch.long)
    # the reference hands out four clones of the id column (model_utils.py:392-401); nothing on the path writes to them, so
    # the four keys share one tensor here (4 fewer launches per chunk)
    metadata = {'warp': idx, 'camera': idx, 'appearance': idx, 'time': idx}
    return {"origins": rays[:, :3], "directions": rays[:, 3:6], "viewdirs": None, "metadata": metadata}


def extract_rays_batch(rays: dict, start: int, end: int, drop_last=True) -> dict:
    out = {k: None for k in rays.keys()}
    for key, val in rays.items():
        if key == 'metadata':
            out[key] = {k: (v[start:end] if v is not None else None) for k, v in val.items()}
        elif val is not None:
            out[key] = val[start:end]
    return out


def append_batch(all_ret, batch) -> dict:
    for k, v in all_ret.items():
        if v is None:
            all_ret[k] = batch[k]
        else:
            for kk, vv in batch[k].items():
                if vv is not None:
                    all_ret[k][kk] = torch.cat([all_ret[k][kk], vv], dim=0)
    return all_ret


def concat_ray_batch(rays: list) -> dict:
    result = {k: None for k in rays[0].keys()}
    for c in rays:
        for k, v in c.items():
            result[k] = v if result[k] is None else torch.cat([result[k], v], dim=0)
    return result
