"""Static NeRF baseline (reference: models/nerf.py), B200 path.

`Embedding` and `NeRF` keep the reference's constructor arguments and `state_dict` names
(xyz_encoding_{1..8}.0.*, xyz_encoding_final.*, dir_encoding.0.*, sigma.*, rgb.0.*), so nerf_pl checkpoints load
unchanged.  The MLP itself runs in the fused tcgen05 kernels of libhypernerf_b200.so (HN_FLAG_STATIC_NERF): positional
encodings are computed in-kernel from raw points / directions, so `render_rays` (rendering.py) hands the kernels
points and directions, not embedded vectors.  There is no CPU or torch fallback.
"""
import ctypes as C

import torch
from torch import nn

from . import _lib
from ._lib import check, lib, ptr, stream
from ._packing import PackedWeights


class Embedding(nn.Module):
    """models/nerf.py:4-38.  Holds the frequency count; the encoding itself happens inside the fused kernels.  Calling
    it evaluates the same formula with torch ops on the tensor's device (utility for callers that want the vector)."""

    def __init__(self, in_channels, N_freqs, logscale=True):
        super().__init__()
        if not logscale:
            raise NotImplementedError("only logscale=True frequency bands (2**k) are built into the kernels")
        self.N_freqs = N_freqs
        self.in_channels = in_channels
        self.funcs = [torch.sin, torch.cos]
        self.out_channels = in_channels * (len(self.funcs) * N_freqs + 1)
        self.freq_bands = 2 ** torch.linspace(0, N_freqs - 1, N_freqs)

    def forward(self, x):
        out = [x]
        for freq in self.freq_bands:
            for func in self.funcs:
                out = out + [func(freq * x)]
        return torch.cat(out, -1)


class _FusedStatic(torch.autograd.Function):
    """hn_mlp_fwd / hn_mlp_bwd with HN_FLAG_STATIC_NERF as one autograd node.  Returns (sigma, rgb) with
    sigma = relu(raw + noise * noise_std) (rendering.py:150 folded into the kernel epilogue)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, model, points, dirs, noise, noise_std, *params):
        B, S = points.shape[0], points.shape[1]
        dev = points.device
        packed = model._packed_weights()
        pts = points.detach().to(torch.float32).contiguous()
        vd = dirs.detach().to(torch.float32).contiguous()
        sigma = torch.empty(B, S, device=dev, dtype=torch.float32)
        rgb = torch.empty(B, S, 3, device=dev, dtype=torch.float32)
        need_grad = any(ctx.needs_input_grad[5:])
        saved = None
        if need_grad:
            saved = torch.empty(model._sizes(B * S).saved_bytes, device=dev, dtype=torch.uint8)
        with _lib.timed("mlp_fwd", B * S):
            check(lib().hn_mlp_fwd(C.byref(model._desc), ptr(packed), ptr(pts), ptr(vd), None, ptr(noise),
                                   float(noise_std), B, S, None, 0, ptr(sigma), ptr(rgb), None, ptr(saved), None, stream()),
                  "hn_mlp_fwd")
        _lib.count(1)
        ctx.model, ctx.shape = model, (B, S)
        ctx.param_meta = [(p.shape, p.numel()) for p in params]
        ctx.save_for_backward(sigma, rgb, saved, packed)
        ctx.set_materialize_grads(False)
        return sigma, rgb

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, g_sigma, g_rgb):
        sigma, rgb, saved, packed = ctx.saved_tensors
        model = ctx.model
        B, S = ctx.shape
        if saved is None:
            raise RuntimeError("hn_mlp_bwd needs the activation stash; forward ran without grad enabled")
        dev = sigma.device
        g_sigma = torch.zeros_like(sigma) if g_sigma is None else g_sigma.to(torch.float32).contiguous()
        g_rgb = torch.zeros_like(rgb) if g_rgb is None else g_rgb.to(torch.float32).contiguous()
        offs, total = model._grad_offsets()
        flat_grad = torch.zeros(total, device=dev, dtype=torch.float32)
        work = torch.empty(model._sizes(B * S).workspace_bytes, device=dev, dtype=torch.uint8)
        with _lib.timed("mlp_dgrad", B * S):
            check(lib().hn_mlp_bwd_data(C.byref(model._desc), ptr(packed), None, ptr(sigma), ptr(rgb), None, ptr(saved),
                                        ptr(g_sigma), ptr(g_rgb), None, B, S, None, 0, 0, offs, ptr(flat_grad), ptr(work), None, None,
                                        stream()),
                  "hn_mlp_bwd_data")
        with _lib.timed("mlp_wgrad", B * S):
            check(lib().hn_mlp_bwd_weights(C.byref(model._desc), ptr(saved), B, S, 0, offs, ptr(flat_grad), ptr(work),
                                           stream()), "hn_mlp_bwd_weights")
        _lib.count(2)
        grads = [flat_grad[offs[i]:offs[i] + n].view(shape) if ctx.needs_input_grad[5 + i] else None
                 for i, (shape, n) in enumerate(ctx.param_meta)]
        return (None,) * 5 + tuple(grads)


class NeRF(PackedWeights, nn.Module):
    """models/nerf.py:41-123: D=8 x W=256 ReLU layers with the input concatenated in front of layer `skips`, sigma head,
    feature layer, view-direction layer and rgb head.  Only the default topology is built into the kernels."""

    def __init__(self, D=8, W=256, in_channels_xyz=63, in_channels_dir=27, skips=[4]):
        super().__init__()
        if D != 8 or W != 256 or in_channels_xyz != 63 or in_channels_dir != 27 or list(skips) != [4]:
            raise NotImplementedError("the sm_100a kernels are instantiated for NeRF(D=8, W=256, 63, 27, skips=[4])")
        self.D, self.W = D, W
        self.in_channels_xyz, self.in_channels_dir, self.skips = in_channels_xyz, in_channels_dir, skips
        for i in range(D):
            if i == 0:
                layer = nn.Linear(in_channels_xyz, W)
            elif i in skips:
                layer = nn.Linear(W + in_channels_xyz, W)
            else:
                layer = nn.Linear(W, W)
            setattr(self, f"xyz_encoding_{i + 1}", nn.Sequential(layer, nn.ReLU(True)))
        self.xyz_encoding_final = nn.Linear(W, W)
        self.dir_encoding = nn.Sequential(nn.Linear(W + in_channels_dir, W // 2), nn.ReLU(True))
        self.sigma = nn.Linear(W, 1)
        self.rgb = nn.Sequential(nn.Linear(W // 2, 3), nn.Sigmoid())
        self._desc = _lib.ModelDesc(glo_dim=0, hyper_dim=0, xyz_freqs=10, hyper_freqs=0, view_freqs=4, warp_freqs=0,
                                    sheet_freqs=0, num_embeddings=0, flags=_lib.HN_FLAG_STATIC_NERF)
        self._pack_levels = 1
        self._init_packing()
        self._size_cache = {}
        self._packed_bytes = self._sizes(0).packed_bytes
        self._grad_off_cache = None

    # -- native plumbing ---------------------------------------------------------------------------------------------
    def _canonical_params(self):
        """state_dict order of the reference module = HN_FLAG_STATIC_NERF parameter order (include/hypernerf_b200.h)."""
        out = []
        for i in range(self.D):
            lin = getattr(self, f"xyz_encoding_{i + 1}")[0]
            out += [lin.weight, lin.bias]
        for lin in (self.xyz_encoding_final, self.dir_encoding[0], self.sigma, self.rgb[0]):
            out += [lin.weight, lin.bias]
        return out

    def _sizes(self, n_samples):
        s = self._size_cache.get(n_samples)
        if s is None:
            s = _lib.Sizes()
            check(lib().hn_query(C.byref(self._desc), n_samples, C.byref(s)), "hn_query")
            self._size_cache[n_samples] = s
        return s

    def _grad_offsets(self):
        if self._grad_off_cache is None:
            offs, total = [], 0
            for p in self._canonical_params():
                offs.append(total)
                total += (p.numel() + 3) // 4 * 4
            self._grad_off_cache = ((C.c_int64 * len(offs))(*offs), total)
        return self._grad_off_cache

    def query(self, points, dirs, noise=None, noise_std=0.0):
        """Fused evaluation on raw sample points (B,S,3) and ray directions (B,3):
        returns (relu(sigma_raw + noise * noise_std) (B,S), rgb (B,S,3))."""
        params = self._canonical_params()
        if not torch.is_grad_enabled():
            params = [q.detach() for q in params]
        with self.packed_frozen():   # re-packs unless an enclosing block froze the blobs (_packing.py)
            return _FusedStatic.apply(self, points, dirs, noise, noise_std, *params)

    def forward(self, x, sigma_only=False):
        """models/nerf.py:84-123 on EMBEDDED inputs (B, 63 [+ 27]).  The first three channels of each embedding are the
        raw coordinates (nerf.py:33), which is what the fused kernels consume."""
        raise NotImplementedError("NeRF.forward on embedded vectors is not built: the fused kernels take raw points and "
                                  "directions; use rendering.render_rays (or NeRF.query)")
