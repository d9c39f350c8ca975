"""GPU parity of BASELINE.json config 5 (SE3Field warp + axis-aligned slicing, H = G = 8) against the restated oracle.

PARITY UNPINNED BY THE REFERENCE: songrise/HyperNeRF-torch defines warping.SE3Field but never instantiates it
(models.py:195 vs :234) and rigid_body.exp_se3 only accepts one point, returns constants and severs autograd
(SURVEY.md §8(c), App. B.1).  What is compared here is the CUDA path against oracle.hypernerf_oracle (se3_field /
se3_transform), whose network half is pinned to the reference's own SE3Field modules and whose exp map is pinned to
torch.linalg.matrix_exp (tests/test_oracle.py)."""
import pytest
import torch

import helpers as H
from oracle import hypernerf_oracle as orc
from oracle import ref_loader
from hypernerf_torch_b200 import model_utils as mu
from hypernerf_torch_b200 import synthetic
from hypernerf_torch_b200.models import NerfModel

pytestmark = pytest.mark.gpu
DEV = "cuda"
KW = dict(n_samples_coarse=64, n_samples_fine=64, noise_std=1.0, use_warp=True, use_nerf_embed=False,
          hyper_slice_method='axis_aligned_plane', hyper_slice_out_dim=8, view_fourier_dim=6, warp_field_type='se3')
TOL = 2e-3            # north star: rgb / depth / weights within 2e-3 max-abs
# boosted stress weights: rotations up to ~0.8 rad about the origin of points at |x| ~ 1.5; the bf16 operands of the w head
# (relative 2^-9) move the warped point by up to ~|x| |w| 2^-9 ~ 5e-3, which posenc then amplifies: 3x the bound, as for the
# translation field's boosted fixtures (tests/test_model.py BOOSTED_TOL)
BOOSTED_TOL = 3 * TOL
WARPED_TOL = {False: 2e-3, True: 1e-2}
FINE_POINT_TOL = 2e-2   # tests/test_model.py: one coarse bin (1/64)


def _model(seed, boosted, **over):
    kw = dict(KW)
    kw.update(over)
    model = NerfModel(H.EMB, **kw)
    sd = synthetic.make_state_dict(model, seed=seed, boosted=boosted)
    model.load_state_dict(sd)
    return model.to(DEV), {k: v.to(DEV) for k, v in sd.items()}, kw


def _draws(B, n_fine, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.rand(B, 64, generator=g), torch.randn(B, 64, 1, generator=g), torch.rand(B, n_fine, generator=g),
            torch.randn(B, 64 + n_fine, 1, generator=g)]


@pytest.mark.parametrize("boosted", [False, True])
def test_se3_forward_matches_restated_oracle(boosted):
    B = 96
    model, sd, kw = _model(3, boosted)
    rays, _ = synthetic.train_rays(B, seed=6, device=DEV)
    draws = [t.to(DEV) for t in _draws(B, 64, 1)]
    with torch.no_grad(), ref_loader._DrawTape(draws):
        out = model(mu.prepare_ray_dict(rays), dict(H.EXTRA))
    o, d = rays[:, :3], rays[:, 3:6]
    free = orc.forward(sd, o, d, rays[:, 8].long(), ref_loader.draws_to_dict(draws), orc.cfg_from_kwargs(kw))
    # the resampled depths inherit the coarse level's bf16 error through the inverse CDF: bounded by about one coarse bin
    perr = (out['fine']['points'] - free['fine']['points']).abs().max().item()
    print(f"se3 boosted={boosted} fine points max_abs_err {perr:.3e}")
    assert perr < FINE_POINT_TOL
    # stage isolation (as tests/test_oracle.py): the fine level of the restatement evaluated at the product's own depths
    z_f = ((out['fine']['points'] - o[:, None, :]) * d[:, None, :]).sum(-1) / (d * d).sum(-1, keepdim=True)
    ref = orc.forward(sd, o, d, rays[:, 8].long(), ref_loader.draws_to_dict(draws), orc.cfg_from_kwargs(kw), fine_z=z_f)
    tol = BOOSTED_TOL if boosted else TOL
    for lvl in ("coarse", "fine"):
        assert set(out[lvl]) == {'points', 'warped_points', 'rgb', 'depth', 'med_depth', 'acc', 'weights', 'med_points'}
        for k in ("rgb", "depth", "acc", "weights"):
            assert out[lvl][k].shape == ref[lvl][k].shape
            err = (out[lvl][k] - ref[lvl][k]).abs().max().item()
            print(f"se3 boosted={boosted} {lvl} {k} max_abs_err {err:.3e}")
            assert err < tol, (lvl, k, err)
    assert torch.equal(out['coarse']['points'], ref['coarse']['points'])
    werr = (out['coarse']['warped_points'] - ref['coarse']['warped_points']).abs().max().item()
    print(f"se3 boosted={boosted} coarse warped_points max_abs_err {werr:.3e}")
    assert werr < WARPED_TOL[boosted]
    # the hyper coordinates are the GLO vector itself (models.py:533-534): exact
    assert torch.equal(out['coarse']['warped_points'][..., 3:], ref['coarse']['warped_points'][..., 3:])


def test_se3_inference_and_training_forward_agree():
    """eval (no stash, folded biases) and training (stash, aux) kernels of the SE3 shape produce the same outputs."""
    B = 40
    model, _, _ = _model(4, True)
    rays, _ = synthetic.train_rays(B, seed=7, device=DEV)
    draws = [t.to(DEV) for t in _draws(B, 64, 2)]
    with torch.no_grad(), ref_loader._DrawTape(draws):
        a = model(mu.prepare_ray_dict(rays), dict(H.EXTRA))
    with ref_loader._DrawTape(draws):
        b = model(mu.prepare_ray_dict(rays), dict(H.EXTRA))
    # (fine per-sample tensors are compared through their composites: the two forwards carry their biases differently
    # — bf16 pairs inside the UMMAs vs fp32 adds —, and a 1e-4 change of a coarse weight already moves resampled depths)
    for lvl, keys in (("coarse", ("rgb", "depth", "acc", "weights", "warped_points")), ("fine", ("rgb", "depth", "acc"))):
        for k in keys:
            err = (a[lvl][k] - b[lvl][k].detach()).abs().max().item()
            assert err < 2e-3, (lvl, k, err)


# saved-activation slab map of the SE3 shape (hn_mlp_program.h make_slabs), in 8-column chunks
SE3_X_HWS, SE3_X_WV, SE3_X_T, SE3_X_R, SE3_X_TOTAL = 6, 118, 172, 482, 546


def _se3_gates(saved, B, S):
    """ReLU gates of every hidden layer decoded from the kernels' own activation stash (cf. helpers.gates_from_stash)."""
    def gate(chunk, ncols, lo=0, hi=None):
        x = H.decode_slab(saved, B * S, chunk, ncols, total_chunks=SE3_X_TOTAL)[:, lo:hi]
        return (x > 0).reshape(B, S, -1)
    return {'warp': [gate(SE3_X_HWS + 16 * l, 128) for l in range(6)],
            'w_net': [gate(SE3_X_WV, 256, 0, 128)], 'v_net': [gate(SE3_X_WV, 256, 128, 256)],
            'trunk': [gate(SE3_X_T + 32 * l, 256) for l in range(9)], 'rgb': [gate(SE3_X_R + 16 * l, 128) for l in range(4)]}


@pytest.mark.parametrize("B,S,level,boosted", [(6, 64, 0, True), (3, 128, 1, True), (5, 64, 1, False), (2, 100, 0, True)])
def test_se3_backward_parameter_gradients(B, S, level, boosted):
    """hn_mlp_bwd of the SE3 shape (exp-map chain rule, w / v heads, logit layer, trunk) against autograd of the restatement in
    bf16-emulation mode through the kernels' own ReLU gates — the same construction and bounds as
    tests/test_mlp.py::test_backward_parameter_gradients, which explains why the gates are shared."""
    from hypernerf_torch_b200.models import _FusedMlp
    model, sd, kw = _model(5, boosted)
    rays, _ = synthetic.train_rays(B, seed=6, device=DEV)
    o, d, ids = rays[:, :3].contiguous(), rays[:, 3:6].contiguous(), rays[:, 8].long()
    g = torch.Generator(device=DEV).manual_seed(9)
    z, _ = torch.sort(torch.rand(B, S, device=DEV, generator=g), -1)
    pts = (o[:, None, :] + z[..., None] * d[:, None, :]).contiguous()
    gs = torch.randn(B, S, device=DEV, generator=g)
    gr = torch.randn(B, S, 3, device=DEV, generator=g)
    gw = torch.randn(B, S, 11, device=DEV, generator=g) * 0.1
    params = model._canonical_params()
    with model.packed_frozen():
        sigma, rgb, warped = _FusedMlp.apply(model, level, pts, d, ids, None, 0.0, *params)
        gates = _se3_gates(sigma.grad_fn.saved_tensors[4].clone(), B, S)
        ((sigma * gs).sum() + (rgb * gr).sum() + (warped * gw).sum()).backward()
    cfg = orc.cfg_from_kwargs(kw, emulate_bf16=True)
    sd_r = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    rgb_r, sigma_r, wp_r, _ = orc.query_fields(sd_r, 'fine' if level else 'coarse', pts, d, ids, cfg, gates=gates)
    ((sigma_r * gs).sum() + (rgb_r * gr).sum() + (wp_r * gw).sum()).backward()
    print("fwd vs oracle[gates]: sigma", H.rel_err(sigma, sigma_r), "rgb", (rgb - rgb_r).abs().max().item(),
          "warped", (warped - wp_r).abs().max().item())
    other = "nerf_mlps_coarse" if level else "nerf_mlps_fine"
    bad = []
    for name, p in model.named_parameters():
        r = sd_r[name].grad
        if name.startswith(other) or r is None:
            assert p.grad is None or p.grad.abs().max() == 0, name
            continue
        e = H.rel_err(p.grad, r)
        print(f"grad {name:50s} same_gates {e:.3e} |ref| {r.norm().item():.3e}")
        # tests/test_mlp.py: 1e-2, the deepest tensors 2e-2.  With the boosted stress weights (rotations up to ~0.8 rad of
        # points at |x| ~ 1.5) a hidden activation that rounds to the neighbouring bf16 value (tensor-core vs torch summation
        # order) moves the warped point by ~1e-3, i.e. by ~0.8 rad of phase in posenc's top octave (2^9), whose cos / sin
        # the chain rule multiplies by 2^9: those samples' d(w, v) decorrelate, and on 200-400 samples that is 2-3 %
        lim = (4e-2 if boosted else 2e-2) if name.startswith(("warp_", "hyper_")) else 1e-2
        if e > lim:
            bad.append((name, e))
    assert not bad, bad


@pytest.mark.parametrize("B,S", [(1, 7), (3, 33), (2, 257), (5, 128)])
def test_se3_forward_ragged_shapes(B, S):
    """One level through hn_mlp_fwd at sample counts that do not fill a 256-row CTA tile (and one that spills into a second
    one), inference and training kernels, against the restatement in bf16-emulation mode."""
    from hypernerf_torch_b200.models import _FusedMlp
    model, sd, kw = _model(7, True)
    rays, _ = synthetic.train_rays(B, seed=9, device=DEV)
    o, d, ids = rays[:, :3].contiguous(), rays[:, 3:6].contiguous(), rays[:, 8].long()
    g = torch.Generator(device=DEV).manual_seed(1)
    z, _ = torch.sort(torch.rand(B, S, device=DEV, generator=g), -1)
    pts = (o[:, None, :] + z[..., None] * d[:, None, :]).contiguous()
    rgb_r, sigma_r, wp_r, _ = orc.query_fields(sd, 'coarse', pts, d, ids, orc.cfg_from_kwargs(kw, emulate_bf16=True))
    params = model._canonical_params()
    for train in (False, True):
        ps = params if train else [q.detach() for q in params]
        sigma, rgb, warped = _FusedMlp.apply(model, 0, pts, d, ids, None, 0.0, *ps)
        assert sigma.shape == (B, S) and rgb.shape == (B, S, 3) and warped.shape == (B, S, 11)
        assert (rgb.detach() - rgb_r).abs().max().item() < 2e-3
        assert (warped.detach() - wp_r).abs().max().item() < 2e-3
        assert H.rel_err(sigma.detach(), sigma_r) < 1e-2


# Gradients against fp32 autograd of the restatement at reference-init weights, 512 rays.  What separates a bf16 from an
# fp32 forward is ReLU gates that flip (profiles/grad_parity.md): at 8 192 rays the translation field's tensors sit at
# 5-12 % and the template's at 0.5-3 %; at 512 rays that noise is ~4x larger, and the bias of the v head is a sum with
# heavy cancellation.  So this test only holds size and direction per tensor and the whole-gradient distance; the exact
# check of the backward kernels is test_se3_backward_parameter_gradients above.  (The boosted weights are not compared with
# fp32 at all: there the warped point itself differs by ~4e-3 = more than a radian of posenc's top octave.)
GRAD_REL = {"warp_field": 0.4, "nerf_mlps": 0.2, "warp_embed": 0.2}
GRAD_COS = 0.95
WHOLE_GRAD_REL = 8e-2


def test_se3_gradients_match_restated_oracle():
    B = 512
    model, sd, kw = _model(5, False)
    rays, rgbs = synthetic.train_rays(B, seed=8, device=DEV)
    draws = [t.to(DEV) for t in _draws(B, 64, 3)]
    with ref_loader._DrawTape(draws):
        out = model(mu.prepare_ray_dict(rays), dict(H.EXTRA))
    loss = torch.nn.functional.mse_loss(out['coarse']['rgb'], rgbs) + torch.nn.functional.mse_loss(out['fine']['rgb'], rgbs)
    loss.backward()
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = orc.forward(sdg, rays[:, :3], rays[:, 3:6], rays[:, 8].long(), ref_loader.draws_to_dict(draws), orc.cfg_from_kwargs(kw))
    ref_loss = orc.mse_loss(ref, rgbs)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) < 2e-3
    bad, num, den = [], 0.0, 0.0
    for k, p in model.named_parameters():
        r = sdg[k].grad
        if r is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        assert p.grad is not None, k
        num += float((p.grad - r).double().pow(2).sum())
        den += float(r.double().pow(2).sum())
        rel = ((p.grad - r).norm() / (r.norm() + 1e-20)).item()
        cos = torch.nn.functional.cosine_similarity(p.grad.flatten(), r.flatten(), dim=0).item()
        print(f"se3 {k:52s} |g| {r.norm().item():.3e} rel {rel:.3e} cos {cos:.5f}")
        lim = next(v for pre, v in GRAD_REL.items() if k.startswith(pre))
        if r.norm().item() > 1e-7 and (rel > lim or (cos < GRAD_COS and r.numel() > 1)):
            bad.append((k, rel, cos))
    whole = (num / den) ** 0.5
    print(f"se3 whole flat gradient, relative L2 vs the fp32 restatement: {whole:.3e}")
    assert not bad, bad
    assert whole < WHOLE_GRAD_REL
