"""GPU: properties of the fused path at BASELINE.json sizes, where the fp32 oracle would take minutes.

* chunk invariance: a ray's outputs and the accumulated gradient do not depend on how the batch is cut into chunks
  or which tile / CTA a ray lands in (8 192 rays in one call vs 5 uneven calls; same draws);
* the inference and training instantiations of the forward kernel agree (to bf16 rounding of the activations);
* compositing invariants on the full 65 536-ray batch: weights in [0, 1], sum(weights) <= 1, acc / depth bounds,
  sorted fine samples, finite everything;
* gradient linearity: grad of (a * loss) == a * grad of loss (the backward kernels accumulate into the flat buffer).
"""
import pytest
import torch

import helpers as H
from hypernerf_torch_b200 import model_utils as mu
from hypernerf_torch_b200 import synthetic
from hypernerf_torch_b200 import train as hn_train
from oracle import ref_loader

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model(noise_std=1.0, n_fine=64):
    sd = synthetic.make_state_dict(synthetic.cfg1_state_dict_shapes(), seed=0, boosted=True)
    return H.make_model(n_fine=n_fine, noise_std=noise_std, sd=sd)


def _draws(B, Nc, Nf, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return [torch.rand(B, Nc, device=DEV, generator=g), torch.randn(B, Nc, 1, device=DEV, generator=g),
            torch.rand(B, Nf, device=DEV, generator=g), torch.randn(B, Nc + Nf, 1, device=DEV, generator=g)]


def _forward(model, rays, draws):
    with ref_loader._DrawTape(draws):
        return model(mu.prepare_ray_dict(rays), dict(H.EXTRA))


def test_chunk_invariance_of_outputs_and_gradients():
    B, Nc, Nf = 8192, 64, 64
    model = _model()
    rays, rgbs = synthetic.train_rays(B, seed=3, device=DEV)
    draws = _draws(B, Nc, Nf, seed=11)
    fg = hn_train.FlatGrads(model.parameters())
    model.attach_flat_grads(fg)

    def run(cuts):
        fg.zero()
        outs = []
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            out = _forward(model, rays[lo:hi], [d[lo:hi] for d in draws])
            loss = (torch.nn.functional.mse_loss(out['coarse']['rgb'], rgbs[lo:hi], reduction='sum') +
                    torch.nn.functional.mse_loss(out['fine']['rgb'], rgbs[lo:hi], reduction='sum')) / (3.0 * B)
            loss.backward()
            outs.append({k: out['fine'][k].detach() for k in ('rgb', 'depth', 'acc', 'weights')})
        return {k: torch.cat([o[k] for o in outs]) for k in outs[0]}, fg.flat.clone()

    whole, g_whole = run([0, B])
    parts, g_parts = run([0, 1000, 1001, 4097, 7777, B])
    for k in whole:                      # per-ray results are position independent: bit-identical
        assert torch.equal(whole[k], parts[k]), k
    # gradients are sums over samples in a different order (atomics / tile partition / wgrad job groups): equal up to
    # fp32 rounding of ~28 k-term running sums in the TMEM accumulators (measured 1.9e-5)
    rel = (g_whole - g_parts).norm() / g_whole.norm()
    assert rel < 1e-4, rel
    assert torch.isfinite(g_whole).all() and g_whole.abs().sum() > 0


def test_inference_and_training_forward_agree():
    B = 4096
    model = _model(noise_std=None)
    rays, _ = synthetic.train_rays(B, seed=5, device=DEV)
    d4 = _draws(B, 64, 64, seed=2)
    draws = [d4[0], d4[2]]                          # no sigma noise: only the two uniform draws are consumed
    with torch.no_grad():
        a = _forward(model, rays, draws)            # inference instantiation (no stash)
    b = _forward(model, rays, draws)                # training instantiation (stash + gate words)
    # the inference instantiation adds the biases inside the UMMAs (extra K step against a ones column, bias as two
    # bf16 terms), the training one in its epilogue: same math, different fp32 summation order, so activations can
    # differ by one bf16 ulp here and there
    for k in ('rgb', 'depth', 'acc', 'weights', 'warped_points'):
        assert torch.equal(a['coarse'][k], b['coarse'][k].detach()) or \
            (a['coarse'][k] - b['coarse'][k].detach()).abs().max() < 2e-3, ('coarse', k)
    # the fine level resamples from the coarse weights: its depths move with them (continuously), so do its outputs
    for k in ('rgb', 'depth', 'acc'):
        d = (a['fine'][k] - b['fine'][k].detach()).abs().max()
        assert d < 1e-2, ('fine', k, d)


def test_full_batch_invariants():
    B = 65536
    model = _model()
    rays, _ = synthetic.train_rays(B, seed=0, device=DEV)
    outs = []
    with torch.no_grad():
        for i in range(0, B, 16384):
            outs.append(model(mu.prepare_ray_dict(rays[i:i + 16384]), dict(H.EXTRA)))
    for lvl, S in (('coarse', 64), ('fine', 128)):
        w = torch.cat([o[lvl]['weights'] for o in outs])
        rgb = torch.cat([o[lvl]['rgb'] for o in outs])
        depth = torch.cat([o[lvl]['depth'] for o in outs])
        acc = torch.cat([o[lvl]['acc'] for o in outs])
        assert w.shape == (B, S) and torch.isfinite(w).all() and torch.isfinite(rgb).all()
        assert (w >= 0).all() and (w <= 1 + 1e-5).all()
        assert (w.sum(-1) <= 1 + 1e-3).all()            # sum alpha_i T_i <= 1 (up to the +1e-5 inside the cumprod)
        assert (acc >= 0).all() and (acc <= w.sum(-1) + 1e-6).all()   # acc excludes the sample at infinity
        assert (rgb >= 0).all() and (rgb <= 1 + 1e-3).all()          # convex-ish combination of sigmoids
        assert (depth >= 0).all() and (depth <= 1 + 1e-3).all()      # near = 0, far = 1


def test_gradient_is_linear_in_the_loss_scale():
    B = 2048
    model = _model()
    rays, rgbs = synthetic.train_rays(B, seed=9, device=DEV)
    draws = _draws(B, 64, 64, seed=4)
    fg = hn_train.FlatGrads(model.parameters())
    model.attach_flat_grads(fg)
    grads = []
    for scale in (1.0, 4.0):
        fg.zero()
        out = _forward(model, rays, draws)
        loss = torch.nn.functional.mse_loss(out['coarse']['rgb'], rgbs) + torch.nn.functional.mse_loss(out['fine']['rgb'], rgbs)
        (scale * loss).backward()
        grads.append(fg.flat.clone())
    rel = (4.0 * grads[0] - grads[1]).norm() / grads[1].norm()
    assert rel < 2e-2, rel    # dY is rounded to bf16 before it enters the UMMAs, so scaling is linear up to bf16 rounding


def test_training_with_fused_adam_tracks_torch_adam():
    """Three train steps (train_step + FusedAdam on the flat buffers) against the same steps with torch.optim.Adam:
    the loss trajectories agree, and the loss moves (the bf16 operand blobs are re-packed after every update)."""
    B = 2048
    rays, rgbs = synthetic.train_rays(B, seed=21, device=DEV)
    traj = []
    for fused in (True, False):
        model = _model()
        fg = hn_train.FlatGrads(model.parameters())
        model.attach_flat_grads(fg)
        opt = hn_train.FusedAdam(fg, lr=5e-4) if fused else torch.optim.Adam(model.parameters(), lr=5e-4, eps=1e-8)
        losses = []
        for it in range(3):
            with ref_loader._DrawTape(_draws(B, 64, 64, seed=100 + it)):
                losses.append(float(hn_train.train_step(model, rays, rgbs, fg, chunk=B, optimizer=opt)))
        traj.append(losses)
        assert set(model.state_dict().keys()) == set(synthetic.cfg1_state_dict_shapes().keys())
    a, b = traj
    assert abs(a[0] - b[0]) <= 1e-6 * abs(b[0])           # same weights, same draws (the loss sum uses atomics)
    assert abs(a[1] - a[0]) > 1e-6 and abs(a[2] - a[1]) > 1e-6
    for x, y in zip(a, b):
        assert abs(x - y) <= 2e-3 * abs(y), (a, b)


@pytest.mark.parametrize("cuda_graph", [False, True])
def test_fit_loop_learns_and_checkpoints_like_lightning(tmp_path, cuda_graph):
    """train.fit (stand-in for Trainer.fit(NeRFSystem), train.py:35-233): shuffled epochs, Adam + MultiStepLR per epoch,
    training_step's log entries, and a checkpoint that utils.load_ckpt(model, path, 'nerf') reads back; eagerly and with the
    full-size batches replayed from a CUDA graph."""
    from hypernerf_torch_b200 import utils as hn_utils
    P = 4096
    rays, _ = synthetic.train_rays(P, seed=31, device=DEV)
    rgbs = torch.tensor([0.8, 0.3, 0.1], device=DEV).expand(P, 3).contiguous()     # a learnable target
    model = _model()
    torch.manual_seed(7)
    log = hn_train.fit(model, rays, rgbs, num_epochs=3, batch_size=1000, lr=5e-4, decay_step=(2,), decay_gamma=0.1,
                       ckpt_path=str(tmp_path / "epoch={epoch}.ckpt"), cuda_graph=cuda_graph)
    assert len(log) == 3 * 5 and log[-1]['step'] == 15                   # 4 full batches + the short one per epoch
    assert [round(e['lr'], 8) for e in log[::5]] == [5e-4, 5e-4, 5e-5]   # MultiStepLR(milestones=[2], gamma=0.1)
    first, last = float(log[0]['train/loss']), float(log[-1]['train/loss'])
    assert last < 0.7 * first, (first, last)
    assert float(log[-1]['train/psnr']) > float(log[0]['train/psnr'])
    blob = torch.load(tmp_path / "epoch=2.ckpt")
    assert blob['epoch'] == 2 and blob['global_step'] == 15
    assert all(k.startswith('nerf.') for k in blob['state_dict'])
    fresh = _model()
    hn_utils.load_ckpt(fresh, str(tmp_path / "epoch=2.ckpt"), 'nerf')
    for (k, a), b in zip(model.state_dict().items(), fresh.state_dict().values()):
        assert torch.equal(a, b), k


def test_render_rays_driver_matches_model_forward():
    """train.render_rays (chunk-free eval driver; eval.py:77-103 batched_inference): the fine-level
    rgb / depth it returns are exactly what NerfModel.forward returns chunk by chunk with the same draws, whatever the
    chunk size, and nothing else is materialised."""
    model = _model(noise_std=None, n_fine=128).eval()
    rays = synthetic.frame_rays(image_id=1, seed=0, device=DEV, h=63, w=84)          # 5 292 rays
    N = rays.shape[0]
    g = torch.Generator(device=DEV).manual_seed(3)
    u_c, u_f = torch.rand(N, 64, device=DEV, generator=g), torch.rand(N, 128, device=DEV, generator=g)

    def draws_for(chunk):
        out = []
        for i in range(0, N, chunk):
            out += [u_c[i:i + chunk], u_f[i:i + chunk]]
        return out

    with ref_loader._DrawTape(draws_for(2000)):
        a = hn_train.render_rays(model, rays, chunk=2000)
    with ref_loader._DrawTape(draws_for(N)):
        b = hn_train.render_rays(model, rays, chunk=N, keys=('rgb', 'depth', 'acc'))
    with torch.no_grad(), ref_loader._DrawTape(draws_for(N)):
        full = model(mu.prepare_ray_dict(rays), dict(H.EXTRA))['fine']
    assert set(a) == {'rgb', 'depth'} and set(b) == {'rgb', 'depth', 'acc'}
    assert a['rgb'].shape == (N, 3) and a['depth'].shape == (N,)
    for k in ('rgb', 'depth'):
        assert torch.equal(a[k], full[k]) and torch.equal(b[k], full[k]), k
    assert not a['rgb'].requires_grad


def test_full_frame_render_psnr_matches_oracle():
    """BASELINE.json north_star / SURVEY.md §8(d) cfg3: one 1008 x 756 frame (762 048 rays, 64 + 128 samples, no_grad,
    noise off, reference-initialised weights).  The fp32 oracle runs on the GPU in torch eager (timing is not the point
    here); both sides get the same stratified / resampling draws.  PSNR against a fixed synthetic target image must
    agree within 0.05 dB; rgb within the north-star 2e-3 max-abs."""
    orc = H.orc
    sd = synthetic.make_state_dict(synthetic.cfg1_state_dict_shapes(), seed=0, boosted=False)
    model = H.make_model(n_fine=128, noise_std=None, sd=sd)
    sd_dev = H.to_dev(sd)
    cfg = orc.default_cfg(n_fine=128, noise_std=None)
    rays = synthetic.frame_rays(image_id=3, seed=0, device=DEV)
    N = rays.shape[0]
    assert N == 1008 * 756
    target = torch.rand(N, 3, generator=torch.Generator().manual_seed(5)).to(DEV)
    g = torch.Generator(device=DEV).manual_seed(17)
    sq_new = torch.zeros((), device=DEV, dtype=torch.float64)
    sq_ref = torch.zeros((), device=DEV, dtype=torch.float64)
    max_rgb = max_depth = 0.0
    chunk = 16384
    with torch.no_grad():
        for i in range(0, N, chunk):
            r = rays[i:i + chunk]
            B = r.shape[0]
            u_c = torch.rand(B, 64, device=DEV, generator=g)
            u_f = torch.rand(B, 128, device=DEV, generator=g)
            out = _forward(model, r, [u_c, u_f])['fine']
            ref = orc.forward(sd_dev, r[:, :3], r[:, 3:6], r[:, 8].long(), {'u_coarse': u_c, 'u_fine': u_f}, cfg)['fine']
            t = target[i:i + chunk]
            sq_new += ((out['rgb'] - t).double() ** 2).sum()
            sq_ref += ((ref['rgb'] - t).double() ** 2).sum()
            max_rgb = max(max_rgb, (out['rgb'] - ref['rgb']).abs().max().item())
            max_depth = max(max_depth, (out['depth'] - ref['depth']).abs().max().item())
    psnr_new = -10.0 * torch.log10(sq_new / (3 * N)).item()
    psnr_ref = -10.0 * torch.log10(sq_ref / (3 * N)).item()
    print(f"frame PSNR new {psnr_new:.4f} dB, oracle {psnr_ref:.4f} dB; max |rgb| diff {max_rgb:.2e}, depth {max_depth:.2e}")
    assert abs(psnr_new - psnr_ref) < 0.05
    assert max_rgb < 2e-3 and max_depth < 2e-3


@pytest.mark.parametrize("n_fine,noise_std", [(64, 1.0), (128, None)])
def test_reusing_the_coarse_warp_outputs_changes_nothing(n_fine, noise_std):
    """NerfModel.reuse_coarse_warp: the fine level runs the full network on its new depths only and the template NeRF
    alone (hn_mlp_fwd_trunk) on the depths it inherits from the coarse level, fed with the coarse pass's warped points.
    Every output is bit-identical to evaluating all depths through the full network (what the reference does); the
    gradient differs only by where the bf16 rounding of the summed warp / sheet upstream gradient happens."""
    B = 3000
    rays, rgbs = synthetic.train_rays(B, seed=13, device=DEV)
    d4 = _draws(B, 64, n_fine, seed=23)
    draws = d4 if noise_std is not None else [d4[0], d4[2]]
    res = []
    for reuse in (False, True):
        model = _model(noise_std=noise_std, n_fine=n_fine)
        model.reuse_coarse_warp = reuse
        fg = hn_train.FlatGrads(model.parameters())
        model.attach_flat_grads(fg)
        fg.zero()
        out = _forward(model, rays, draws)
        loss = torch.nn.functional.mse_loss(out['coarse']['rgb'], rgbs) + torch.nn.functional.mse_loss(out['fine']['rgb'], rgbs) \
            + 0.1 * out['fine']['warped_points'].square().mean()       # also exercises the gradient into warped_points
        loss.backward()
        res.append((out, fg.flat.clone(), [p.numel() for p in fg.params], fg.offsets))
        with torch.no_grad():
            res[-1] = res[-1] + (_forward(model, rays, draws),)      # inference instantiation of both paths
    (a, ga, numels, offs, ia), (b, gb, _, _, ib) = res
    for lvl in ('coarse', 'fine'):
        assert set(a[lvl]) == set(b[lvl])
        for k in a[lvl]:
            assert torch.equal(a[lvl][k], b[lvl][k]), (lvl, k)
            assert torch.equal(ia[lvl][k], ib[lvl][k]), ('inference', lvl, k)
    rel = (ga - gb).norm() / ga.norm()
    assert rel < 5e-3, rel
    names = [n for n, _ in _model().named_parameters()]
    for name, off, n in zip(names, offs, numels):
        x, y = ga[off:off + n], gb[off:off + n]
        assert (x - y).norm() <= 2e-2 * x.norm() + 1e-9, (name, ((x - y).norm() / x.norm()).item())


def test_graphed_train_step_matches_eager():
    """train.GraphedTrainStep (the chunk loop captured in a CUDA graph) against train.train_step on the same weights and
    rays: same loss and flat gradient up to the different random draws being excluded (noise off, deterministic u)."""
    from hypernerf_torch_b200 import synthetic
    from hypernerf_torch_b200 import train as hn_train
    from hypernerf_torch_b200.models import NerfModel
    from oracle import ref_loader
    kw = ref_loader.cfg1_kwargs(n_fine=64, noise_std=None)
    model = NerfModel(ref_loader.EMBEDDINGS, **kw)
    model.load_state_dict(synthetic.make_state_dict(model, seed=0, boosted=False))
    model = model.to(DEV)
    model.use_stratified_sampling = False        # no draws at all: linspace depths, deterministic resampling (models.py:146)
    fg = hn_train.FlatGrads(model.parameters())
    model.attach_flat_grads(fg)
    rays, rgbs = synthetic.train_rays(4096, seed=3, device=DEV)
    loss_e = hn_train.train_step(model, rays, rgbs, fg, chunk=2048)
    flat_e = fg.flat.clone()
    step = hn_train.GraphedTrainStep(model, fg, 4096, chunk=2048)
    for _ in range(3):                            # capture, then two replays
        loss_g = step(rays, rgbs)
        torch.cuda.synchronize()
        assert abs(float(loss_g) - float(loss_e)) < 1e-6
        rel = ((fg.flat - flat_e).norm() / flat_e.norm()).item()
        assert rel < 1e-4, rel                    # atomics order only
    assert step.launches > 0
    # new inputs go through the static buffers
    rays2, rgbs2 = synthetic.train_rays(4096, seed=4, device=DEV)
    loss2_g = float(step(rays2, rgbs2))
    loss2_e = float(hn_train.train_step(model, rays2, rgbs2, fg, chunk=2048))
    assert abs(loss2_g - loss2_e) < 1e-6
