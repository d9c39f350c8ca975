"""Synthetic workload of BASELINE.json / SURVEY.md §8(d): LLFF-shaped forward-facing NDC rays and deterministic
random-init weights.  Host-side data generation only (there is no network for datasets or checkpoints).

Ray geometry follows datasets/ray_utils.py:5-93 and datasets/llff.py:244-264 of the reference: H x W = 756 x 1008,
focal 815.13, identity rotation, camera centre t ~ U(-0.3, 0.3)^3 per image, near plane 1.0 -> NDC origins on
z = -1, unnormalised directions with d_z = 2; ray row = [o(3), d(3), near=0, far=1, image id].
"""
import math

import torch

H, W, FOCAL = 756, 1008, 815.13
NUM_IMAGES = 100


def ndc_rays_for_pixels(px, py, cam_t, near=1.0):
    """px, py: float pixel coordinates (N,); cam_t: (N,3) camera centres.  Returns (origins, directions) in NDC
    (get_ray_directions ray_utils.py:5-24 with identity c2w, get_ndc_rays ray_utils.py:53-93)."""
    dx = (px - W / 2) / FOCAL
    dy = -(py - H / 2) / FOCAL
    dz = -torch.ones_like(dx)
    rays_d = torch.stack([dx, dy, dz], -1)
    rays_o = cam_t
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    ox_oz = rays_o[..., 0] / rays_o[..., 2]
    oy_oz = rays_o[..., 1] / rays_o[..., 2]
    o0 = -1. / (W / (2. * FOCAL)) * ox_oz
    o1 = -1. / (H / (2. * FOCAL)) * oy_oz
    o2 = 1. + 2. * near / rays_o[..., 2]
    d0 = -1. / (W / (2. * FOCAL)) * (rays_d[..., 0] / rays_d[..., 2] - ox_oz)
    d1 = -1. / (H / (2. * FOCAL)) * (rays_d[..., 1] / rays_d[..., 2] - oy_oz)
    d2 = 1 - o2
    return torch.stack([o0, o1, o2], -1), torch.stack([d0, d1, d2], -1)


def train_rays(n, seed=0, device="cpu"):
    """n ray rows (n,9) drawn uniformly from the 100 x H x W pool + target colours (n,3) in U(0,1)."""
    g = torch.Generator().manual_seed(seed)
    cams = (torch.rand(NUM_IMAGES, 3, generator=g) - 0.5) * 0.6
    img = torch.randint(0, NUM_IMAGES, (n,), generator=g)
    px = torch.randint(0, W, (n,), generator=g).float()
    py = torch.randint(0, H, (n,), generator=g).float()
    o, d = ndc_rays_for_pixels(px, py, cams[img])
    rays = torch.cat([o, d, torch.zeros(n, 1), torch.ones(n, 1), img.float()[:, None]], 1)
    rgbs = torch.rand(n, 3, generator=torch.Generator().manual_seed(seed + 1))
    return rays.to(device), rgbs.to(device)


def frame_rays(image_id=0, seed=0, device="cpu", h=H, w=W):
    """All h*w rays of one frame (h*w, 9), row-major pixels, camera `image_id` of the seed's pool."""
    g = torch.Generator().manual_seed(seed)
    cams = (torch.rand(NUM_IMAGES, 3, generator=g) - 0.5) * 0.6
    ys, xs = torch.meshgrid(torch.arange(h).float() * (H / h), torch.arange(w).float() * (W / w), indexing="ij")
    px, py = xs.reshape(-1), ys.reshape(-1)
    n = px.numel()
    o, d = ndc_rays_for_pixels(px, py, cams[image_id].expand(n, 3))
    rays = torch.cat([o, d, torch.zeros(n, 1), torch.ones(n, 1), torch.full((n, 1), float(image_id))], 1)
    return rays.to(device)


def make_state_dict(module_or_shapes, seed=0, boosted=False):
    """Deterministic weights for a NerfModel-shaped state_dict (CPU generator, canonical key order).

    boosted=False follows the reference initialisers' scales (xavier-uniform hidden layers, warp output U(0,1e-4),
    sheet output N(0,1e-5), GLO N(0, 0.1/G), biases U(+-1/sqrt(fan_in))); boosted=True scales the warp / sheet output
    layers to trained-model magnitudes (warp offsets ~0.05, hyper coordinates ~0.3) and uses a 0.5-std GLO table so
    the warp and hyper-sheet branches carry signal in tests; boosted='glo' keeps the reference scales everywhere except the
    GLO tables (0.5 std), so that template conditioning and axis-aligned hyper points carry signal."""
    shapes = module_or_shapes if isinstance(module_or_shapes, dict) else \
        {k: tuple(v.shape) for k, v in module_or_shapes.state_dict().items()}
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in shapes.items():
        if name.endswith("embed.weight"):
            std = 0.5 if boosted else 0.1 / shape[1]   # (boosted == 'glo' is truthy)
            out[name] = torch.randn(shape, generator=g) * std
        elif name.endswith(".weight"):
            fan_out, fan_in = shape
            bound = math.sqrt(6.0 / (fan_in + fan_out))
            wt = (torch.rand(shape, generator=g) * 2 - 1) * bound
            if name == "warp_field.mlp.logit_layer.weight":
                wt = wt * 0.05 if boosted is True else torch.rand(shape, generator=g) * 1e-4
            if name in ("warp_field.w_net.logit_layer.weight", "warp_field.v_net.logit_layer.weight"):
                # SE3Field heads (warping.py:167-168: U(0, 1e-4)); boosted: rotations of ~0.1-1 rad (both branches of the
                # kernels' exp-map coefficients), translations ~0.05
                scale = 0.6 if ".w_net." in name else 0.1
                wt = wt * scale if boosted is True else torch.rand(shape, generator=g) * 1e-4
            if name == "hyper_sheet_mlp.mlp.logit_layer.weight":
                wt = wt * 0.3 if boosted is True else torch.randn(shape, generator=g) * 1e-5
            out[name] = wt
        else:  # bias: nn.Linear default U(-1/sqrt(fan_in), 1/sqrt(fan_in))
            fan_in = shapes[name[:-4] + "weight"][1]
            out[name] = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in)
    return out


def cfg1_state_dict_shapes():
    """state_dict names / shapes of the cfg-1 model without constructing a module."""
    s = {"warp_embed.embed.weight": (100, 8)}

    def mlp(prefix, in_ch, width, depth, out_ch, skip=4):
        for i in range(depth):
            fan_in = in_ch if i == 0 else (width + in_ch if (i - 1) == skip else width)
            s[f"{prefix}.linears.{i}.weight"] = (width, fan_in)
            s[f"{prefix}.linears.{i}.bias"] = (width,)
        s[f"{prefix}.logit_layer.weight"] = (out_ch, width)
        s[f"{prefix}.logit_layer.bias"] = (out_ch,)

    mlp("hyper_sheet_mlp.mlp", 53, 64, 6, 2)
    mlp("warp_field.mlp", 71, 128, 6, 3)
    for lvl in ("nerf_mlps_coarse", "nerf_mlps_fine"):
        mlp(f"{lvl}.trunk_mlp", 89, 256, 8, 256)
        s[f"{lvl}.bottleneck_mlp.weight"] = (128, 256)
        s[f"{lvl}.bottleneck_mlp.bias"] = (128,)
        mlp(f"{lvl}.rgb_mlp", 167, 128, 4, 3)
        s[f"{lvl}.alpha_mlp.weight"] = (1, 128)
        s[f"{lvl}.alpha_mlp.bias"] = (1,)
    return s


def static_state_dict_shapes():
    """state_dict names / shapes of one static NeRF (models/nerf.py: D=8, W=256, 63 / 27 input channels, skips=[4])."""
    s = {}
    for i in range(8):
        fan_in = 63 if i == 0 else (256 + 63 if i == 4 else 256)
        s[f"xyz_encoding_{i + 1}.0.weight"] = (256, fan_in)
        s[f"xyz_encoding_{i + 1}.0.bias"] = (256,)
    s["xyz_encoding_final.weight"] = (256, 256)
    s["xyz_encoding_final.bias"] = (256,)
    s["dir_encoding.0.weight"] = (128, 256 + 27)
    s["dir_encoding.0.bias"] = (128,)
    s["sigma.weight"] = (1, 256)
    s["sigma.bias"] = (1,)
    s["rgb.0.weight"] = (3, 128)
    s["rgb.0.bias"] = (3,)
    return s
