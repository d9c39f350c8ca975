"""CPU-only checks of the C-ABI boundary: the library loads, exports every symbol include/hypernerf_b200.h
declares, sizes are consistent with the reference model, and argument errors are reported (not crashed on)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, cfg1_shapes
from hypernerf_torch_b200 import _lib


def _desc(**over):
    kw = dict(glo_dim=8, hyper_dim=2, xyz_freqs=10, hyper_freqs=6, view_freqs=6, warp_freqs=10, sheet_freqs=7,
              num_embeddings=100, flags=3)
    kw.update(over)
    return _lib.ModelDesc(*[kw[k] for k in ("glo_dim", "hyper_dim", "xyz_freqs", "hyper_freqs", "view_freqs",
                                            "warp_freqs", "sheet_freqs", "num_embeddings", "flags")])


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "hypernerf_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|const char\*)\s+(hn_\w+)\s*\(", header, flags=re.M))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    L = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert _lib.lib().hn_abi_version() == 4
    # the product ABI carries no test / microbenchmark hooks: those live in libhypernerf_b200_probe.so with their own header
    assert not any(n.startswith(("hn_umma", "hn_epi", "hn_tmem", "hn_debug")) for n in declared)
    for n in ("hn_umma_probe", "hn_epi_rate", "hn_debug_set_timing_buffer"):
        assert not hasattr(L, n), n
    probe_h = open(os.path.join(ROOT, "include", "hypernerf_b200_probe.h")).read()
    probe_decl = set(re.findall(r"^\s*int\s+(hn_\w+)\s*\(", probe_h, flags=re.M)) - {"hn_debug_set_timing_buffer"}
    assert probe_decl == set(_lib.PROBE_EXPORTS), probe_decl ^ set(_lib.PROBE_EXPORTS)
    P = C.CDLL(_lib.PROBE_LIB_PATH)
    for name in probe_decl:
        assert hasattr(P, name), name


def test_query_matches_reference_parameter_count():
    import math
    n_ref = sum(math.prod(s) for s in cfg1_shapes().values())
    assert n_ref == 1483053                                    # SURVEY.md App. A.2 [probed]
    s = _lib.Sizes()
    assert _lib.lib().hn_query(C.byref(_desc()), 1024 * 64, C.byref(s)) == 0
    assert s.flat_param_floats == n_ref
    halves = 1024 * 64 // 64
    # 4320 bf16 columns per sample + 124 ReLU gate words per sample
    assert s.saved_bytes == halves * 540 * 1024 + halves * 124 * 64 * 4
    assert s.workspace_bytes == (1024 * 64 // 64) * 518 * 1024
    assert s.packed_bytes % 256 == 0 and s.packed_bytes > 2 * 2 * 700000


def test_errors_are_reported_not_crashed():
    L = _lib.lib()
    s = _lib.Sizes()
    assert L.hn_query(C.byref(_desc(flags=1)), 64, C.byref(s)) < 0
    assert b"bendy_sheet" in L.hn_last_error()
    assert L.hn_query(C.byref(_desc(hyper_dim=3)), 64, C.byref(s)) < 0
    assert b"instantiated" in L.hn_last_error()
    for h in (4, 8):                                  # opt.py default hyper_slice_out_dim and the axis-aligned shape
        assert L.hn_query(C.byref(_desc(hyper_dim=h)), 64, C.byref(s)) == 0
    assert L.hn_query(C.byref(_desc(hyper_dim=8, flags=1 | 8 | 16)), 64, C.byref(s)) == 0   # axis-aligned + alpha condition
    assert L.hn_query(C.byref(_desc(hyper_dim=0, flags=0)), 64, C.byref(s)) == 0            # no warp
    assert L.hn_query(None, 64, C.byref(s)) < 0
    # shape / null checks happen before any launch, so they are testable without a GPU
    assert L.hn_sample_pdf(None, None, None, 0, None, None, None, 4, 64, 62, 64, None, None, None, None) < 0
    assert L.hn_composite_fwd(None, None, None, None, 4, 1000, 0, 1e-5, 1e7, None, None, None, None, None, None, None) < 0
    assert L.hn_mlp_fwd(C.byref(_desc()), None, None, None, None, None, 0.0, 4, 64, None, 0, None, None, None, None, None, None) < 0
    with pytest.raises(_lib.NativeLibraryError):
        _lib.check(-1, "probe")


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "hypernerf_torch_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("no CPU or torch fallback", ""), fn


def test_checkpoint_prefix_stripping(tmp_path):
    """utils/__init__.py:66-88: Lightning checkpoints store the model under 'nerf.'; names must round-trip."""
    import torch
    from hypernerf_torch_b200 import synthetic, utils
    sd = synthetic.make_state_dict(synthetic.static_state_dict_shapes(), seed=1)
    ckpt = {'state_dict': {('nerf.' + k): v for k, v in sd.items()} | {'other.weight': torch.zeros(1)}}
    path = tmp_path / "ckpt.pt"
    torch.save(ckpt, path)
    got = utils.extract_model_state_dict(str(path), model_name='nerf')
    assert set(got) == set(sd) and all(torch.equal(got[k], sd[k]) for k in sd)

    from hypernerf_torch_b200.nerf import NeRF
    m = NeRF()
    utils.load_ckpt(m, str(path), model_name='nerf')
    assert all(torch.equal(m.state_dict()[k], sd[k]) for k in sd)
