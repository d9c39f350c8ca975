"""Pins the oracle (oracle/hypernerf_oracle.py) against (a) the committed golden vectors generated from the
unmodified reference and (b) the live reference when /root/reference is present.  CPU only."""
import pytest
import torch

from conftest import cfg1_shapes, golden_state_dict, load_golden
from oracle import hypernerf_oracle as orc
from oracle import ref_loader

FIXTURES = ["cfg1_refinit_b32", "cfg1_boosted_b32", "cfg3_boosted_b16"]
KEYS = ["points", "warped_points", "rgb", "depth", "med_depth", "acc", "weights", "med_points"]


def _run_oracle(fix, requires_grad=False, isolate_fine=True):
    sd = golden_state_dict(fix, cfg1_shapes())
    if requires_grad:
        sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    rays = fix['rays']
    cfg = orc.default_cfg(n_fine=fix['n_fine'], noise_std=fix['noise_std'] or 0.0)
    draws = ref_loader.draws_to_dict(fix['draws'], noise=bool(fix['noise_std']))
    # stage isolation: the fine level is evaluated at the reference's own resampled depths, because the
    # reference's torch.sum makes its resampled z differ from the contract arithmetic by an ulp now and then
    out = orc.forward(sd, rays[:, :3], rays[:, 3:6], rays[:, 8].long(), draws, cfg,
                      fine_z=fix['taps']['z_fine'] if isolate_fine else None)
    return sd, out


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_matches_golden_outputs(name):
    fix = load_golden(name)
    _, out = _run_oracle(fix)
    for lvl in ("coarse", "fine"):
        for k in KEYS:
            ref = fix['out'][lvl][k]
            got = out[lvl][k]
            assert got.shape == ref.shape, (lvl, k)
            # same fp32 arithmetic, same op order: differences are summation-order noise only
            torch.testing.assert_close(got, ref, rtol=2e-5, atol=2e-6, msg=f"{name} {lvl} {k}")
    # resampling stage on its own: golden coarse z / weights + recorded draws -> golden fine z.  The explicit
    # arithmetic contract (fp64-carried sums) equals the reference's except for 1-ulp pdf differences caused by
    # torch.sum's SIMD cascade.
    o, d = fix['rays'][:, :3], fix['rays'][:, 3:6]
    z_c = fix['taps']['z_coarse']
    assert torch.equal(out['coarse']['z_vals'], z_c)
    u_fine = fix['draws'][2] if fix['noise_std'] else fix['draws'][1]
    bins = .5 * (z_c[..., 1:] + z_c[..., :-1])
    z_f, _, _ = orc.sample_pdf(bins, fix['out']['coarse']['weights'][..., 1:-1], o, d, z_c, u_fine)
    assert (z_f - fix['taps']['z_fine']).abs().max() < 5e-6
    assert (z_f == fix['taps']['z_fine']).float().mean() > 0.5


@pytest.mark.parametrize("name", FIXTURES[:2])
def test_oracle_matches_golden_grads(name):
    fix = load_golden(name)
    sd, out = _run_oracle(fix, requires_grad=True)
    loss = orc.mse_loss(out, fix['rgbs'])
    assert abs(float(loss) - fix['loss']) < 1e-6
    loss.backward()
    for k, n in fix['grad_norms'].items():
        g = sd[k].grad
        assert g is not None, k
        assert abs(float(g.double().norm()) - n) <= 1e-4 * n + 1e-12, k
    for k, ref in fix['grad_small'].items():
        torch.testing.assert_close(sd[k].grad, ref, rtol=1e-3, atol=1e-9, msg=k)


CONFIG_FIXTURES = ["optdefault_b16", "cond_h4_vf4_b16", "alphacond_h8_b16", "axis_h8_b16", "nowarp_b16", "nowarp_cond_b16"]


@pytest.mark.parametrize("name", CONFIG_FIXTURES)
def test_oracle_matches_config_goldens(name):
    """The configurations beyond cfg 1-3 (opt.py defaults with hyper_slice_out_dim 4, template GLO conditioning,
    axis-aligned slicing, no warp): oracle outputs and gradients against the unmodified reference's (make_golden.py configs)."""
    from hypernerf_torch_b200 import synthetic
    fix = load_golden(name)
    sd = synthetic.make_state_dict(fix['shapes'], seed=fix['weight_seed'], boosted=fix['boosted'])
    chk = float(sum(v.double().abs().sum() for v in sd.values()))
    assert abs(chk - fix['weight_checksum']) <= 1e-6 * abs(chk)
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    rays = fix['rays']
    cfg = orc.cfg_from_kwargs(fix['kw'])
    draws = ref_loader.draws_to_dict(fix['draws'], noise=bool(fix['noise_std']))
    out = orc.forward(sd, rays[:, :3], rays[:, 3:6], rays[:, 8].long(), draws, cfg, fine_z=fix['taps']['z_fine'])
    for lvl in ("coarse", "fine"):
        for k in KEYS:
            ref = fix['out'][lvl][k]
            assert out[lvl][k].shape == ref.shape, (lvl, k)
            torch.testing.assert_close(out[lvl][k], ref, rtol=2e-5, atol=2e-6, msg=f"{name} {lvl} {k}")
    loss = orc.mse_loss(out, fix['rgbs'])
    assert abs(float(loss.detach()) - fix['loss']) < 1e-6
    loss.backward()
    for k, n in fix['grad_norms'].items():
        if n is None:     # parameters the reference itself leaves without a gradient (unused embedding tables)
            assert sd[k].grad is None or float(sd[k].grad.abs().max()) == 0.0, k
            continue
        assert abs(float(sd[k].grad.double().norm()) - n) <= 1e-4 * n + 1e-12, k
    for k, ref in fix['grad_small'].items():
        torch.testing.assert_close(sd[k].grad, ref, rtol=1e-3, atol=1e-9, msg=k)


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not mounted")
def test_oracle_matches_live_reference():
    from hypernerf_torch_b200 import synthetic
    model = ref_loader.build_reference_model(seed=0)
    sd = synthetic.make_state_dict(model, seed=7, boosted=True)
    model.load_state_dict(sd)
    rays, _ = synthetic.train_rays(24, seed=3)
    torch.manual_seed(99)
    taps = {}
    ref_out, tape = ref_loader.run_reference(model, rays, taps=taps)
    out = orc.forward(sd, rays[:, :3], rays[:, 3:6], rays[:, 8].long(), ref_loader.draws_to_dict(tape), orc.default_cfg(),
                      fine_z=taps['z_fine'])
    for lvl in ("coarse", "fine"):
        for k in KEYS:
            torch.testing.assert_close(out[lvl][k], ref_out[lvl][k].detach(), rtol=2e-5, atol=2e-6, msg=f"{lvl} {k}")
    # resampling stage against the reference's own function on the same bins / weights / draws
    _, ref_mu = ref_loader.load_reference()
    z = taps['z_coarse']
    bins = .5 * (z[..., 1:] + z[..., :-1])
    w = ref_out['coarse']['weights'][..., 1:-1].detach()
    with ref_loader._DrawTape([tape[2]]):
        ref_samples = ref_mu.piecewise_constant_pdf(bins, w, tape[2].shape[1], True)
    samples, _ = orc.piecewise_constant_pdf(bins, w, tape[2])
    assert (samples - ref_samples).abs().max() < 5e-6


RENDER_OPTS = {'dust_threshold': 0.6, 'bounding_box': (-0.8, 0.9, -0.7, 0.8, -1.0, -0.2)}


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not mounted")
def test_oracle_filter_sigma_matches_live_reference():
    """render_opts (dust threshold + bounding box, models.py:35-63) reach the fine level only (models.py:768)."""
    from hypernerf_torch_b200 import synthetic
    model = ref_loader.build_reference_model(seed=0)
    sd = synthetic.make_state_dict(model, seed=9, boosted=True)
    model.load_state_dict(sd)
    rays, _ = synthetic.train_rays(16, seed=5)
    torch.manual_seed(7)
    taps = {}
    ref_out, tape = ref_loader.run_reference(model, rays, taps=taps, render_opts=RENDER_OPTS)
    plain, _ = ref_loader.run_reference(model, rays, draws=tape)
    assert (ref_out['fine']['weights'] - plain['fine']['weights']).abs().max() > 1e-3      # the options do something
    assert torch.equal(ref_out['coarse']['weights'], plain['coarse']['weights'])            # ... to the fine level only
    out = orc.forward(sd, rays[:, :3], rays[:, 3:6], rays[:, 8].long(), ref_loader.draws_to_dict(tape), orc.default_cfg(),
                      fine_z=taps['z_fine'], render_opts=RENDER_OPTS)
    for lvl in ("coarse", "fine"):
        for k in KEYS:
            torch.testing.assert_close(out[lvl][k], ref_out[lvl][k].detach(), rtol=2e-5, atol=2e-6, msg=f"{lvl} {k}")


# ------------------------------------------------------------------------------------------------------------------------
# config 5: SE3Field + axis-aligned slicing.  The reference never instantiates SE3Field and its exp map cannot run batched
# (SURVEY.md §8(c)): the restatement is pinned in two halves — the networks against the reference's own modules, the exp
# map against the matrix exponential of the twist — and "parity unpinned by the reference" stays true for the composition.
# ------------------------------------------------------------------------------------------------------------------------
def test_se3_networks_match_reference_golden():
    """posenc(points, 0, 8), trunk, w_net, v_net of the UNMODIFIED warping.SE3Field (oracle/make_golden.py se3net)."""
    from hypernerf_torch_b200 import synthetic
    fix = load_golden("se3_net_ref")
    sd = synthetic.make_state_dict(fix['shapes'], seed=fix['weight_seed'], boosted=True)
    chk = float(sum(v.double().abs().sum() for v in sd.values()))
    assert abs(chk - fix['weight_checksum']) <= 1e-6 * abs(chk)
    assert torch.equal(orc.posenc_scaled(fix['points'], 0, 8), fix['feat'])
    _, w, v = orc.se3_field(sd, fix['points'])
    torch.testing.assert_close(w, fix['w'], rtol=2e-5, atol=2e-6)
    torch.testing.assert_close(v, fix['v'], rtol=2e-5, atol=2e-6)


def test_se3_transform_is_the_matrix_exponential_of_the_twist():
    """se3_transform (Rodrigues + Modern Robotics Eq. 3.88, rigid_body.py:55-83 batched) == expm([[skew(w), v], [0, 0]])
    applied to the homogeneous point, in fp64, for rotation angles from 1e-3 to beyond pi."""
    g = torch.Generator().manual_seed(3)
    n = 512
    axis = torch.nn.functional.normalize(torch.randn(n, 3, generator=g, dtype=torch.float64), dim=-1)
    theta = torch.logspace(-3, 0.6, n, dtype=torch.float64)           # 1e-3 .. ~4 rad
    w = axis * theta[:, None]
    v = torch.randn(n, 3, generator=g, dtype=torch.float64)
    x = torch.randn(n, 3, generator=g, dtype=torch.float64)
    twist = torch.zeros(n, 4, 4, dtype=torch.float64)
    twist[:, :3, :3] = orc._skew(w)
    twist[:, :3, 3] = v
    T = torch.linalg.matrix_exp(twist)
    want = (T[:, :3, :3] @ x[..., None])[..., 0] + T[:, :3, 3]
    got = orc.se3_transform(x, w, v)
    torch.testing.assert_close(got, want, rtol=1e-9, atol=1e-9)
    # a rigid map: distances between points are preserved
    x2 = torch.randn(n, 3, generator=g, dtype=torch.float64)
    d0 = (x - x2).norm(dim=-1)
    d1 = (got - orc.se3_transform(x2, w, v)).norm(dim=-1)
    torch.testing.assert_close(d0, d1, rtol=1e-9, atol=1e-9)


def test_se3_config_forward_runs_and_differentiates():
    """The restated config-5 forward end to end on CPU: shapes of the reference's output dictionary, finite gradients for
    every SE3 tensor, none for the unused hyper_embed table."""
    from hypernerf_torch_b200 import synthetic
    from hypernerf_torch_b200.models import NerfModel
    kw = dict(n_samples_coarse=64, n_samples_fine=64, noise_std=1.0, use_warp=True, use_nerf_embed=False,
              hyper_slice_method='axis_aligned_plane', hyper_slice_out_dim=8, view_fourier_dim=6, warp_field_type='se3')
    model = NerfModel(ref_loader.EMBEDDINGS, **kw)     # parameter containers only: constructs without a GPU
    sd = {k: v.clone().requires_grad_(True) for k, v in synthetic.make_state_dict(model, seed=1, boosted=True).items()}
    rays, rgbs = synthetic.train_rays(8, seed=4)
    g = torch.Generator().manual_seed(0)
    draws = [torch.rand(8, 64, generator=g), torch.randn(8, 64, 1, generator=g), torch.rand(8, 64, generator=g),
             torch.randn(8, 128, 1, generator=g)]
    out = orc.forward(sd, rays[:, :3], rays[:, 3:6], rays[:, 8].long(), ref_loader.draws_to_dict(draws), orc.cfg_from_kwargs(kw))
    assert out['fine']['warped_points'].shape == (8, 128, 11) and out['fine']['rgb'].shape == (8, 3)
    orc.mse_loss(out, rgbs).backward()
    for k, v in sd.items():
        if k.startswith("hyper_embed"):
            assert v.grad is None
        else:
            assert v.grad is not None and torch.isfinite(v.grad).all() and float(v.grad.abs().max()) > 0, k


def test_synthetic_rays_are_llff_shaped():
    from hypernerf_torch_b200 import synthetic
    rays, rgbs = synthetic.train_rays(4096, seed=0)
    assert rays.shape == (4096, 9) and rgbs.shape == (4096, 3)
    assert torch.allclose(rays[:, 2], torch.full((4096,), -1.0), atol=1e-5)      # origins on the NDC near plane
    assert torch.allclose(rays[:, 5], torch.full((4096,), 2.0), atol=1e-5)       # d_z = 2
    assert rays[:, :2].abs().max() < 1.9 and rays[:, 3:5].abs().max() < 0.9
    ids = rays[:, 8]
    assert ids.min() >= 0 and ids.max() <= 99 and torch.equal(ids, ids.round())
