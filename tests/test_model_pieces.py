"""GPU: the pieces of NerfModel.render_samples as the reference exposes them (hypernerf/models.py:447-585: apply_warp,
map_points, map_spatial_points, map_hyper_points, query_template) and metadata_encoded=True (models.py:605-625), each
against the model's own ids path — same kernels, same values, so equality is exact — and that path is pinned to the
reference by tests/test_model.py."""
import pytest
import torch

import helpers as H
from oracle import ref_loader
from hypernerf_torch_b200 import model_utils as mu
from hypernerf_torch_b200 import synthetic
from hypernerf_torch_b200.models import NerfModel

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model(boosted=True, **over):
    kw = ref_loader.cfg1_kwargs(n_fine=64, noise_std=None)
    kw.update(over)
    m = NerfModel(H.EMB, **kw)
    m.load_state_dict(synthetic.make_state_dict(m, seed=2, boosted=boosted))
    return m.to(DEV)


def _points(B, S, seed=0):
    rays, _ = synthetic.train_rays(B, seed=seed, device=DEV)
    o, d, ids = rays[:, :3].contiguous(), rays[:, 3:6].contiguous(), rays[:, 8].long()
    g = torch.Generator(device=DEV).manual_seed(seed)
    z, _ = torch.sort(torch.rand(B, S, device=DEV, generator=g), -1)
    return rays, (o[:, None, :] + z[..., None] * d[:, None, :]).contiguous(), z, d, ids


@pytest.mark.parametrize("kw", [dict(), dict(hyper_slice_method='axis_aligned_plane', hyper_slice_out_dim=8),
                                dict(use_nerf_embed=True, use_alpha_cond=True, use_rgb_cond=True, hyper_slice_out_dim=4)])
def test_pieces_compose_to_render_samples(kw):
    B, S = 24, 64
    model = _model(**kw)
    _, pts, z, d, ids = _points(B, S, seed=3)
    meta = {'time': ids, 'warp': ids}
    with torch.no_grad():
        whole = model.render_samples('coarse', pts, z, d, d, meta, dict(H.EXTRA), use_sample_at_infinity=True)
        table = model.warp_embed.embed.weight
        embed = table[ids][:, None, :].expand(B, S, 8)                      # models.py:627-632
        warped, jac = model.map_points(pts, embed, embed, dict(H.EXTRA))
        assert jac is None and torch.equal(warped, whole['warped_points'])
        sp, _ = model.map_spatial_points(pts, embed, dict(H.EXTRA))
        assert torch.equal(sp, warped[..., :3])
        assert torch.equal(model.map_hyper_points(pts, embed, dict(H.EXTRA)), warped[..., 3:])
        assert torch.equal(model.apply_warp(pts, ids, dict(H.EXTRA))['warped_points'], warped[..., :3])
        # the sub-modules called on their own, as the reference's map_* methods call them (warping.py:98-125, modules.py:331-337)
        assert torch.equal(model.warp_field(pts, embed, dict(H.EXTRA))['warped_points'], warped[..., :3])
        if model.hyper_slice_method == 'bendy_sheet':
            assert torch.equal(model.hyper_sheet_mlp(pts, embed, alpha=None), warped[..., 3:])
        rgb, sigma = model.query_template('coarse', warped, d, meta, dict(H.EXTRA))
        assert rgb.shape == (B, S, 3) and sigma.shape == (B, S)
        comp = mu.volumetric_rendering(rgb, sigma, z, d, False, sample_at_infinity=True)
        for k in ("rgb", "depth", "acc", "weights"):
            assert torch.equal(comp[k], whole[k]), k
        # pieces that the configuration does not have
        assert model.map_points(pts, embed, embed, dict(H.EXTRA), use_warp=False)[0] is pts
        with pytest.raises(NotImplementedError):
            model.map_hyper_points(pts, embed, dict(H.EXTRA), hyper_point_override=embed)


def test_query_template_without_warp():
    B, S = 8, 64
    model = _model(use_warp=False, hyper_slice_method=None)
    _, pts, z, d, ids = _points(B, S, seed=4)
    with torch.no_grad():
        whole = model.render_samples('fine', pts, z, d, d, {}, dict(H.EXTRA), use_sample_at_infinity=True)
        rgb, sigma = model.query_template('fine', pts, d, {}, dict(H.EXTRA))
        comp = mu.volumetric_rendering(rgb, sigma, z, d, False, sample_at_infinity=True)
    assert torch.equal(comp['rgb'], whole['rgb'])
    assert model.map_points(pts, None, None, dict(H.EXTRA))[0] is pts and model.map_hyper_points(pts, None, dict(H.EXTRA)) is None


@pytest.mark.parametrize("kw", [dict(), dict(use_nerf_embed=True, use_alpha_cond=True, hyper_slice_out_dim=4)])
def test_metadata_encoded_equals_ids_path(kw):
    """forward(metadata_encoded=True) with 'encoded_*' = table[ids]: same outputs bit for bit, the same parameter gradients,
    and the gradient with respect to the encoded vectors scatter-adds to the table gradient of the ids path."""
    B = 48
    model = _model(noise_std=1.0, **kw)
    rays, rgbs = synthetic.train_rays(B, seed=5, device=DEV)
    g = torch.Generator().manual_seed(0)
    draws = [t.to(DEV) for t in (torch.rand(B, 64, generator=g), torch.randn(B, 64, 1, generator=g),
                                 torch.rand(B, 64, generator=g), torch.randn(B, 128, 1, generator=g))]
    ray_dict = mu.prepare_ray_dict(rays)
    with ref_loader._DrawTape(draws):
        a = model(ray_dict, dict(H.EXTRA))
    loss_a = torch.nn.functional.mse_loss(a['coarse']['rgb'], rgbs) + torch.nn.functional.mse_loss(a['fine']['rgb'], rgbs)
    loss_a.backward()
    grads_a = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    model.zero_grad(set_to_none=True)
    ids = rays[:, 8].long()
    enc = model.warp_embed.embed.weight.detach()[ids].clone().requires_grad_(True)
    enc_dict = dict(ray_dict)
    enc_dict['metadata'] = {'encoded_warp': enc, 'encoded_hyper': enc, 'encoded_nerf': enc}
    with ref_loader._DrawTape(draws):
        b = model(enc_dict, dict(H.EXTRA), metadata_encoded=True)
    for lvl in ("coarse", "fine"):
        for k in a[lvl]:
            assert torch.equal(a[lvl][k], b[lvl][k]), (lvl, k)
    loss_b = torch.nn.functional.mse_loss(b['coarse']['rgb'], rgbs) + torch.nn.functional.mse_loss(b['fine']['rgb'], rgbs)
    loss_b.backward()
    table_grad = torch.zeros_like(model.warp_embed.embed.weight).index_add_(0, ids, enc.grad)
    torch.testing.assert_close(table_grad, grads_a['warp_embed.embed.weight'], rtol=1e-4, atol=1e-9)
    for k, p in model.named_parameters():
        if k == 'warp_embed.embed.weight' or k not in grads_a:
            continue
        # (weight gradients are accumulated with floating-point atomics: equal up to summation order)
        torch.testing.assert_close(p.grad, grads_a[k], rtol=2e-3, atol=1e-7, msg=k)
    with pytest.raises(NotImplementedError):
        other = dict(enc_dict)
        other['metadata'] = {'encoded_warp': enc, 'encoded_hyper': enc.detach().clone(), 'encoded_nerf': enc}
        model(other, dict(H.EXTRA), metadata_encoded=True)


def test_sub_modules_outside_a_model_fail_loudly():
    from hypernerf_torch_b200 import modules
    pts = torch.zeros(2, 4, 3, device=DEV)
    emb = torch.zeros(2, 4, 8, device=DEV)
    for mod, args in ((modules.TranslationField(), (pts, emb, {})), (modules.HyperSheetMLP(out_ch=2), (pts, emb)),
                      (modules.SE3Field(), (pts, emb, {})), (modules.NerfMLP(89), (pts,))):
        with pytest.raises(NotImplementedError):
            mod.to(DEV)(*args)
    # a model survives deep copies / device moves with its back references in place (no cyclic module tree)
    import copy
    m = copy.deepcopy(_model())
    assert m.warp_field._owner() is not None and 'warp_field._owner_ref' not in m.state_dict()
