"""CPU: the part of the reference NerfModel's public surface that needs no kernel (hypernerf/models.py:312-402) and the
list of what is not built.  The parameter containers construct without a GPU."""
import pytest
import torch

from oracle import ref_loader
from hypernerf_torch_b200.models import NerfModel

# get_condition_inputs returns the posenc'd view directions / embeddings as tensors for NerfMLP.forward; the fused kernels
# form them in-kernel and there is no stand-alone NerfMLP.forward to feed
NOT_BUILT = {"get_condition_inputs"}
REFERENCE_PUBLIC = {"num_nerf_embeds", "num_warp_embeds", "num_hyper_embeds", "nerf_embeds", "warp_embeds", "hyper_embeds",
                    "has_hyper", "has_hyper_embed", "has_embeds", "_encode_embed", "encode_hyper_embed", "encode_nerf_embed",
                    "encode_warp_embed", "query_template", "map_spatial_points", "map_hyper_points", "map_points", "apply_warp",
                    "render_samples", "forward"} | NOT_BUILT


def _model(**over):
    kw = ref_loader.cfg1_kwargs()
    kw.update(over)
    return NerfModel(ref_loader.EMBEDDINGS, **kw)


def test_public_surface_is_the_reference_one():
    m = _model()
    for name in REFERENCE_PUBLIC - NOT_BUILT:
        assert hasattr(m, name), name
    for name in NOT_BUILT:
        assert not hasattr(m, name), f"{name} exists now: take it off the not-built list"
    if ref_loader.reference_available():
        ref_models, _ = ref_loader.load_reference()
        ref_names = {n for n in vars(ref_models.NerfModel) if not n.startswith("__")}
        assert ref_names == REFERENCE_PUBLIC, ref_names ^ REFERENCE_PUBLIC


def test_embedding_helpers():
    m = _model(use_nerf_embed=True, use_alpha_cond=True)
    ids = torch.tensor([[3], [7], [99]])
    assert m.num_warp_embeds == 100 and m.num_nerf_embeds == 100 and m.num_hyper_embeds == 100
    assert torch.equal(m.warp_embeds, torch.arange(100)) and m.has_hyper and m.has_hyper_embed and m.has_embeds
    w = m.encode_warp_embed({'time': ids})
    assert w.shape == (3, 8) and torch.equal(w, m.warp_embed.embed.weight[ids[:, 0]])
    assert torch.equal(m.encode_hyper_embed({'time': ids}), w)                 # the bendy sheet shares the warp metadata
    assert torch.equal(m.encode_nerf_embed({'warp': ids}), m.nerf_embed.embed.weight[ids[:, 0]])
    # (left id, right id, progression): linear blend (models.py:352-374)
    mixed = m._encode_embed(torch.tensor([[3., 7., 0.25]]), m.warp_embed)
    torch.testing.assert_close(mixed, 0.75 * w[0:1] + 0.25 * w[1:2])
    with pytest.raises(RuntimeError):
        _model(hyper_slice_method=None).encode_hyper_embed({'time': ids})


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not mounted")
def test_embedding_helpers_match_live_reference():
    m = _model(use_nerf_embed=True, use_alpha_cond=True)
    ref = ref_loader.build_reference_model(use_nerf_embed=True, use_alpha_cond=True)
    ref.load_state_dict(m.state_dict())
    meta = {'time': torch.tensor([[1], [50]]), 'warp': torch.tensor([[2], [60]])}
    for fn in ("encode_warp_embed", "encode_hyper_embed", "encode_nerf_embed"):
        assert torch.equal(getattr(ref, fn)(meta), getattr(m, fn)(meta)), fn
    for prop in ("num_nerf_embeds", "num_warp_embeds", "num_hyper_embeds", "has_hyper", "has_hyper_embed", "has_embeds"):
        assert getattr(ref, prop) == getattr(m, prop), prop
    assert torch.equal(ref.warp_embeds, m.warp_embeds)


def test_empty_ray_batch_raises_like_the_reference():
    from hypernerf_torch_b200 import model_utils as mu
    with pytest.raises(IndexError):
        mu.prepare_ray_dict(torch.zeros(0, 9))
    if ref_loader.reference_available():
        _, ref_mu = ref_loader.load_reference()
        with pytest.raises(IndexError):
            ref_mu.prepare_ray_dict(torch.zeros(0, 9))
    m = _model()
    with pytest.raises(ValueError):
        m({'origins': torch.zeros(0, 3), 'directions': torch.zeros(0, 3), 'metadata': {'time': torch.zeros(0, dtype=torch.long)}}, {})
