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


if __name__ == '__main__':
    make('cfg1_refinit_b32', 32, 64, 1.0, False, 0)      # reference-style init, train config (noise on)
    make('cfg1_boosted_b32', 32, 64, 1.0, True, 1)       # warp / sheet branches carry signal
    make('cfg3_boosted_b16', 16, 128, None, True, 2)     # render config: 64+128, no noise
