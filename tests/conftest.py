import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), map_location="cpu", weights_only=False)


def golden_state_dict(fix, shapes_from):
    from hypernerf_torch_b200 import synthetic
    sd = synthetic.make_state_dict(shapes_from, seed=fix['weight_seed'], boosted=fix['boosted'])
    chk = float(sum(v.double().abs().sum() for v in sd.values()))
    assert abs(chk - fix['weight_checksum']) <= 1e-6 * abs(chk), "weight recipe drifted from the golden fixture"
    return sd


# state_dict shapes of the cfg-1 model so CPU tests need neither the reference nor CUDA
def cfg1_shapes():
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
