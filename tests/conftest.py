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


@pytest.fixture(autouse=True)
def _seed_per_test(request):
    """Every test starts from its own fixed generator state (a hash of its node id): tests that draw with torch.rand /
    torch.randn without seeding are reproducible and independent of which tests ran before them."""
    import zlib
    seed = zlib.crc32(request.node.nodeid.encode())
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    yield


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), map_location="cpu", weights_only=False)


def golden_state_dict(fix, shapes_from):
    from hypernerf_torch_b200 import synthetic
    sd = synthetic.make_state_dict(shapes_from, seed=fix['weight_seed'], boosted=fix['boosted'])
    chk = float(sum(v.double().abs().sum() for v in sd.values()))
    assert abs(chk - fix['weight_checksum']) <= 1e-6 * abs(chk), "weight recipe drifted from the golden fixture"
    return sd


def cfg1_shapes():
    from hypernerf_torch_b200 import synthetic
    return synthetic.cfg1_state_dict_shapes()
