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


def test_synthetic_rays_are_llff_shaped():
    from hypernerf_torch_b200 import synthetic
    rays, rgbs = synthetic.train_rays(4096, seed=0)
    assert rays.shape == (4096, 9) and rgbs.shape == (4096, 3)
    assert torch.allclose(rays[:, 2], torch.full((4096,), -1.0), atol=1e-5)      # origins on the NDC near plane
    assert torch.allclose(rays[:, 5], torch.full((4096,), 2.0), atol=1e-5)       # d_z = 2
    assert rays[:, :2].abs().max() < 1.9 and rays[:, 3:5].abs().max() < 0.9
    ids = rays[:, 8]
    assert ids.min() >= 0 and ids.max() <= 99 and torch.equal(ids, ids.round())
