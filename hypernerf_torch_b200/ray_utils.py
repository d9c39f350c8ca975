"""Eval ray generation on the device (reference: datasets/ray_utils.py:5-93 and the test-split branch of
datasets/llff.py:316-332, which build the (H*W, 8|9) ray rows of a pose on the host)."""
import ctypes as C

import torch

from . import _lib
from ._lib import check, lib, ptr, stream


def frame_rays_ndc(H, W, focal, c2w, near_plane=1.0, image_id=None, device="cuda"):
    """All H*W rays of one pose in NDC as rows [o(3), d(3), 0, 1 (, image id)] (llff.py:244-248, 261-264): 9 columns
    when image_id is given, else 8.  c2w: (3,4) camera-to-world (any tensor / nested list)."""
    m = torch.as_tensor(c2w, dtype=torch.float32).reshape(12).cpu()
    arr = (C.c_float * 12)(*m.tolist())
    cols = 8 if image_id is None else 9
    rays = torch.empty(H * W, cols, device=device, dtype=torch.float32)
    check(lib().hn_make_ndc_rays(int(H), int(W), float(focal), arr, float(near_plane),
                                 float(0 if image_id is None else image_id), cols, ptr(rays), stream()), "hn_make_ndc_rays")
    _lib.count(1)
    return rays
