"""GPU: pins the UMMA shared-memory descriptor conventions (hn_ptx.cuh) against a plain matmul."""
import pytest
import torch

from hypernerf_torch_b200 import _lib

pytestmark = pytest.mark.gpu

CASES = [(16, 16, 0, 0), (64, 64, 0, 0), (128, 128, 0, 0), (192, 80, 0, 0), (256, 256, 0, 0), (144, 176, 0, 0),
         (16, 192, 0, 0), (96, 256, 0, 0),
         (64, 64, 1, 1), (256, 64, 1, 1), (208, 64, 1, 1), (80, 128, 1, 1), (256, 256, 1, 1)]


@pytest.mark.parametrize("N,K,a_mn,b_mn", CASES)
def test_umma_probe(N, K, a_mn, b_mn):
    g = torch.Generator(device="cuda").manual_seed(N * 1000 + K)
    A = torch.randn(128, K, device="cuda", generator=g).bfloat16()
    B = torch.randn(N, K, device="cuda", generator=g).bfloat16()
    ref = A.float() @ B.float().t()
    Ain = A.t().contiguous() if a_mn else A.contiguous()
    Bin = B.t().contiguous() if b_mn else B.contiguous()
    D = torch.full((128, N), float("nan"), device="cuda")
    _lib.check(_lib.probe_lib().hn_umma_probe(_lib.ptr(Ain), _lib.ptr(Bin), _lib.ptr(D), N, K, a_mn, b_mn, _lib.stream()),
               "hn_umma_probe", _lib.probe_lib())
    torch.cuda.synchronize()
    err = (D - ref).abs().max().item()
    print(f"probe N={N} K={K} a_mn={a_mn} b_mn={b_mn} max_err={err:.3e}")
    assert err < 1e-2 * max(1.0, ref.abs().max().item() / 16)


@pytest.mark.parametrize("N,K", [(256, 256), (128, 128), (192, 80), (16, 192), (144, 176), (64, 64), (96, 256), (32, 16)])
def test_umma_pair_probe(N, K):
    """cta_group::2: M = 256 over a 2-CTA cluster, each CTA supplying its 128 A rows and N/2 rows of B."""
    g = torch.Generator(device="cuda").manual_seed(N * 1000 + K + 7)
    A = torch.randn(256, K, device="cuda", generator=g).bfloat16()
    B = torch.randn(N, K, device="cuda", generator=g).bfloat16()
    ref = A.float() @ B.float().t()
    D = torch.full((256, N), float("nan"), device="cuda")
    _lib.check(_lib.probe_lib().hn_umma_probe2(_lib.ptr(A.contiguous()), _lib.ptr(B.contiguous()), _lib.ptr(D), N, K, _lib.stream()),
               "hn_umma_probe2", _lib.probe_lib())
    torch.cuda.synchronize()
    err = (D - ref).abs().max().item()
    print(f"pair probe N={N} K={K} max_err={err:.3e}")
    assert err < 1e-2 * max(1.0, ref.abs().max().item() / 16)
