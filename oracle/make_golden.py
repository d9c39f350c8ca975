"""TEST INFRASTRUCTURE — generates tests/golden/*.pt from the UNMODIFIED reference imported in place from
/root/reference (python oracle/make_golden.py).  The fixtures travel to the GPU box; the reference does not.

Each fixture: ray rows, target colours, the recorded random draws, the weight recipe (seed / boosted flag of
hypernerf_torch_b200.synthetic.make_state_dict + a checksum), every output of the reference forward, and the
gradients of the rgb-MSE loss (full tensors for the small parameters, norms for all).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hypernerf_torch_b200 import synthetic  # noqa: E402
from oracle import ref_loader  # noqa: E402

SMALL_GRADS = ["warp_embed.embed.weight", "warp_field.mlp.logit_layer.weight", "warp_field.mlp.logit_layer.bias",
               "hyper_sheet_mlp.mlp.logit_layer.weight", "hyper_sheet_mlp.mlp.linears.0.weight",
               "warp_field.mlp.linears.5.bias", "nerf_mlps_coarse.alpha_mlp.weight", "nerf_mlps_fine.alpha_mlp.weight",
               "nerf_mlps_fine.alpha_mlp.bias", "nerf_mlps_fine.rgb_mlp.logit_layer.weight",
               "nerf_mlps_coarse.trunk_mlp.linears.0.bias", "nerf_mlps_fine.trunk_mlp.linears.5.bias",
               "nerf_mlps_fine.bottleneck_mlp.bias", "nerf_mlps_fine.rgb_mlp.linears.0.bias"]


def make(name, n_rays, n_fine, noise_std, boosted, seed):
    torch.set_num_threads(8)
    model = ref_loader.build_reference_model(seed=0, n_samples_fine=n_fine, noise_std=noise_std)
    sd = synthetic.make_state_dict(model, seed=seed, boosted=boosted)
    model.load_state_dict(sd)
    rays, rgbs = synthetic.train_rays(n_rays, seed=seed + 10)
    torch.manual_seed(1234)
    taps = {}
    out, tape = ref_loader.run_reference(model, rays, taps=taps)
    loss = torch.nn.functional.mse_loss(out['coarse']['rgb'], rgbs) + torch.nn.functional.mse_loss(out['fine']['rgb'], rgbs)
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    fix = {
        'name': name, 'n_fine': n_fine, 'noise_std': noise_std, 'boosted': boosted, 'weight_seed': seed,
        'weight_checksum': float(sum(v.double().abs().sum() for v in sd.values())),
        'rays': rays, 'rgbs': rgbs, 'draws': tape, 'loss': float(loss.detach()), 'taps': taps,
        'out': {lvl: {k: v.detach().clone() for k, v in out[lvl].items()} for lvl in out},
        'grad_norms': {k: float(g.double().norm()) for k, g in grads.items()},
        'grad_small': {k: grads[k] for k in SMALL_GRADS},
    }
    path = os.path.join(ROOT, 'tests', 'golden', name + '.pt')
    torch.save(fix, path)
    print(name, 'loss', float(loss), os.path.getsize(path) // 1024, 'KiB')


# Configurations beyond cfg 1-3: constructor arguments exactly as the reference takes them.
CONFIGS = {
    # NeRFSystem's call (train.py:48-67) with the opt.py defaults: hyper_slice_out_dim 4, no template conditioning, 64+128
    'optdefault_b16': dict(kw=dict(near=0., far=1., n_samples_coarse=64, n_samples_fine=128, noise_std=1.0, use_warp=True,
                                   use_nerf_embed=False, use_alpha_cond=False, use_rgb_cond=False,
                                   hyper_slice_method='bendy_sheet', hyper_slice_out_dim=4, GLO_dim=8, share_GLO=True,
                                   xyz_fourier_dim=10, hyper_fourier_dim=6, view_fourier_dim=6), boosted=True, seed=40),
    # the constructor's own defaults where they run (models.py:111-127: nerf embed + alpha condition, view freqs 4, H 4) + rgb condition
    'cond_h4_vf4_b16': dict(kw=dict(n_samples_fine=64, noise_std=1.0, use_rgb_cond=True, hyper_slice_method='bendy_sheet'),
                            boosted='glo', seed=41),
    'alphacond_h8_b16': dict(kw=dict(n_samples_fine=64, noise_std=None, hyper_slice_method='bendy_sheet',
                                     hyper_slice_out_dim=8, view_fourier_dim=6), boosted=True, seed=42),
    # axis-aligned slicing: hyper point = GLO vector (models.py:533-534), hyper_slice_out_dim == GLO_dim
    'axis_h8_b16': dict(kw=dict(n_samples_fine=64, noise_std=1.0, use_nerf_embed=False, use_alpha_cond=False,
                                hyper_slice_method='axis_aligned_plane', hyper_slice_out_dim=8, view_fourier_dim=6),
                        boosted='glo', seed=43),
    # no warp: the template alone on the raw points (models.py:568-569)
    'nowarp_b16': dict(kw=dict(n_samples_fine=64, noise_std=1.0, use_warp=False, use_nerf_embed=False, use_alpha_cond=False),
                       boosted=False, seed=44),
    'nowarp_cond_b16': dict(kw=dict(n_samples_fine=128, noise_std=None, use_warp=False, use_rgb_cond=True, view_fourier_dim=6),
                            boosted='glo', seed=45),
}


def make_config(name, n_rays=16):
    """Fixture for one entry of CONFIGS: like make(), with the constructor arguments stored and every gradient tensor of
    at most 4 096 elements kept in full (norms for all)."""
    torch.set_num_threads(8)
    ref_models, _ = ref_loader.load_reference()
    c = CONFIGS[name]
    torch.manual_seed(0)
    model = ref_models.NerfModel(ref_loader.EMBEDDINGS, **c['kw'])
    sd = synthetic.make_state_dict(model, seed=c['seed'], boosted=c['boosted'])
    model.load_state_dict(sd)
    rays, rgbs = synthetic.train_rays(n_rays, seed=c['seed'] + 10)
    torch.manual_seed(1234)
    taps = {}
    out, tape = ref_loader.run_reference(model, rays, taps=taps)
    loss = torch.nn.functional.mse_loss(out['coarse']['rgb'], rgbs) + torch.nn.functional.mse_loss(out['fine']['rgb'], rgbs)
    loss.backward()
    grads = {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in model.named_parameters()}
    fix = {
        'name': name, 'kw': c['kw'], 'boosted': c['boosted'], 'weight_seed': c['seed'],
        'n_fine': c['kw'].get('n_samples_fine', 128), 'noise_std': c['kw'].get('noise_std'),
        'weight_checksum': float(sum(v.double().abs().sum() for v in sd.values())),
        'shapes': {k: tuple(v.shape) for k, v in sd.items()},
        'rays': rays, 'rgbs': rgbs, 'draws': tape, 'loss': float(loss.detach()), 'taps': taps,
        'out': {lvl: {k: v.detach().clone() for k, v in out[lvl].items()} for lvl in out},
        'grad_norms': {k: (float(g.double().norm()) if g is not None else None) for k, g in grads.items()},
        'grad_small': {k: g for k, g in grads.items() if g is not None and g.numel() <= 4096},
    }
    path = os.path.join(ROOT, 'tests', 'golden', name + '.pt')
    torch.save(fix, path)
    print(name, 'loss', float(loss), os.path.getsize(path) // 1024, 'KiB',
          'no-grad params:', [k for k, g in grads.items() if g is None])


STATIC_SMALL_GRADS = ["sigma.weight", "sigma.bias", "rgb.0.weight", "rgb.0.bias", "xyz_encoding_1.0.bias",
                      "xyz_encoding_5.0.bias", "dir_encoding.0.bias", "xyz_encoding_final.bias"]


def make_static(name, n_rays, n_samples, n_importance, perturb, noise_std, seed):
    """Static baseline (BASELINE.json configs[3]): models/nerf.py + models/rendering.py of the unmodified reference."""
    torch.set_num_threads(8)
    ref_loader.load_reference()
    from models.nerf import Embedding, NeRF          # the reference's modules (sys.path set by load_reference)
    from models.rendering import render_rays
    models = [NeRF(), NeRF()]
    sds = [synthetic.make_state_dict(synthetic.static_state_dict_shapes(), seed=seed + i) for i in range(2)]
    for m, sd in zip(models, sds):
        m.load_state_dict(sd)
    emb = [Embedding(3, 10), Embedding(3, 4)]
    rays9, rgbs = synthetic.train_rays(n_rays, seed=seed + 10)
    rays = rays9[:, :8].contiguous()
    torch.manual_seed(1234)
    with ref_loader._DrawTape() as tape:
        out = render_rays(models, emb, rays, N_samples=n_samples, perturb=perturb, noise_std=noise_std,
                          N_importance=n_importance, chunk=1 << 15)
    loss = torch.nn.functional.mse_loss(out['rgb_coarse'], rgbs) + torch.nn.functional.mse_loss(out['rgb_fine'], rgbs)
    loss.backward()
    fix = {'name': name, 'n_samples': n_samples, 'n_importance': n_importance, 'perturb': perturb, 'noise_std': noise_std,
           'weight_seed': seed, 'rays': rays, 'rgbs': rgbs, 'draws': tape.tape, 'loss': float(loss.detach()),
           'weight_checksum': [float(sum(v.double().abs().sum() for v in sd.values())) for sd in sds],
           'out': {k: v.detach().clone() for k, v in out.items()},
           'grad_norms': [{k: float(p.grad.double().norm()) for k, p in m.named_parameters()} for m in models],
           'grad_small': [{k: dict(m.named_parameters())[k].grad.detach().clone() for k in STATIC_SMALL_GRADS} for m in models]}
    path = os.path.join(ROOT, 'tests', 'golden', name + '.pt')
    torch.save(fix, path)
    print(name, 'loss', float(loss), os.path.getsize(path) // 1024, 'KiB')


def make_ndc_rays():
    """tests/golden/ndc_rays.pt: datasets/ray_utils.py of the reference (kornia.create_meshgrid, absent here, stubbed with
    the pixel-index grid it returns for normalized_coordinates=False, ray_utils.py:17-22) on two small frames."""
    import types
    k = types.ModuleType('kornia')

    def create_meshgrid(H, W, normalized_coordinates=False):
        ys, xs = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing='ij')
        return torch.stack([xs, ys], -1)[None]
    k.create_meshgrid = create_meshgrid
    sys.modules['kornia'] = k
    sys.path.insert(0, ref_loader.reference_dir())
    from datasets.ray_utils import get_ndc_rays, get_ray_directions, get_rays
    torch.manual_seed(0)
    cases = []
    for (H, W, focal) in ((12, 16, 13.7), (756 // 9, 1008 // 9, 815.13 / 9)):
        A = torch.linalg.qr(torch.randn(3, 3))[0]
        if torch.det(A) < 0:
            A[:, 0] = -A[:, 0]
        R = torch.linalg.qr(torch.eye(3) * 0.9 + 0.1 * A)[0]
        R = R * torch.sign(torch.diagonal(R))[None]      # mostly forward-facing camera
        c2w = torch.cat([R, (torch.rand(3, 1) - 0.5) * 0.6], 1)
        o, d = get_rays(get_ray_directions(H, W, focal), c2w)
        o, d = get_ndc_rays(H, W, focal, 1.0, o, d)
        rays = torch.cat([o, d, torch.zeros_like(o[:, :1]), torch.ones_like(o[:, :1])], 1)
        cases.append(dict(H=H, W=W, focal=focal, c2w=c2w, rays=rays))
    torch.save(cases, os.path.join(ROOT, 'tests', 'golden', 'ndc_rays.pt'))
    print('ndc_rays', [tuple(c['rays'].shape) for c in cases])


def make_se3_net():
    """What of config 5 the reference CAN run: SE3Field's networks.  warping.SE3Field(in_ch=3) is instantiated from the
    unmodified reference, loaded with synthetic weights, and its own posenc / trunk / w_net / v_net are evaluated on a batch
    of points (warping.py:212-227); only rigid.exp_se3 (rigid_body.py:59-83, single point, constants, no autograd) is left
    out.  Pins oracle.se3_field up to the exp map; the exp map itself is checked against torch.linalg.matrix_exp."""
    from hypernerf_torch_b200 import synthetic
    ref_loader.load_reference()
    from hypernerf import warping as ref_warping, model_utils as ref_mu
    torch.manual_seed(0)
    field = ref_warping.SE3Field(in_ch=3)
    shapes = {f"warp_field.{k}": tuple(v.shape) for k, v in field.state_dict().items()}
    sd = synthetic.make_state_dict(shapes, seed=21, boosted=True)
    field.load_state_dict({k[len("warp_field."):]: v for k, v in sd.items()})
    g = torch.Generator().manual_seed(5)
    points = torch.cat([(torch.rand(4, 64, 2, generator=g) - 0.5) * 3.0, -torch.rand(4, 64, 1, generator=g)], -1)
    with torch.no_grad():
        feat = ref_mu.posenc(points, min_deg=field.min_deg, max_deg=field.max_deg, use_identity=field.use_posenc_identity,
                             alpha=None)
        trunk = field.trunk(feat)
        w, v = field.w_net(trunk), field.v_net(trunk)
    fix = dict(shapes=shapes, weight_seed=21, points=points, feat=feat, trunk=trunk, w=w, v=v,
               weight_checksum=float(sum(t.double().abs().sum() for t in sd.values())))
    torch.save(fix, os.path.join(ROOT, "tests", "golden", "se3_net_ref.pt"))
    print("se3_net_ref", tuple(w.shape), float(w.norm(dim=-1).min()), float(w.norm(dim=-1).max()))


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'se3net':
        make_se3_net()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'configs':
        for name in (sys.argv[2:] or CONFIGS):
            make_config(name)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'rays':
        make_ndc_rays()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'static':
        make_static('static_train_b32', 32, 64, 64, 1.0, 1.0, 20)       # training shape: perturbed, sigma noise
        make_static('static_eval_b16', 16, 64, 128, 0, 0.0, 30)         # eval shape: deterministic sampling, no noise
        sys.exit(0)
    make('cfg1_refinit_b32', 32, 64, 1.0, False, 0)      # reference-style init, train config (noise on)
    make('cfg1_boosted_b32', 32, 64, 1.0, True, 1)       # warp / sheet branches carry signal
    make('cfg3_boosted_b16', 16, 128, None, True, 2)     # render config: 64+128, no noise
