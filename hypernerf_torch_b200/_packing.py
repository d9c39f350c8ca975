"""Life cycle of the packed bf16 weight blobs (hn_pack_weights) behind NerfModel / NeRF.

The kernels read kernel-layout bf16 copies of the fp32 master parameters.  No host-side key can tell reliably that a
parameter changed: optimizers that update through `p.data` (the reference's RAdam / Ranger, utils/optimizers.py:88,163,
242,396), EMA swaps and manual re-initialisation do not bump `Tensor._version`.  So the policy is not a cache key:

  * every top-level call (NerfModel.forward, NeRF.query) re-packs — one small launch per level, always correct;
  * callers that KNOW the weights cannot change between calls (the chunk loops of train.train_step and
    train.render_rays) wrap them in `with model.packed_frozen():` — packed once at entry, reused inside;
  * direct calls of the autograd functions outside any of those (tests, profiling scripts) fall back to a
    (`_version`, `data_ptr`) key, and `invalidate_packed()` drops the blobs explicitly.
"""
import contextlib
import ctypes as C

import torch

from . import _lib
from ._lib import check, lib, ptr, stream


class PackedWeights:
    """Mixin: needs `_canonical_params()`, `_desc`, `_packed_bytes` and `_pack_levels` (1 or 2)."""

    def _init_packing(self):
        self._pack_cache = {}
        self._pack_depth = 0

    def invalidate_packed(self):
        """Drop the packed blobs: the next call re-packs from the fp32 parameters."""
        self._pack_cache = {}

    def _pack_key(self, params):
        return (tuple(p._version for p in params), tuple(p.data_ptr() for p in params))

    def refresh_packed(self):
        """Pack every level now (hn_pack_weights), unconditionally."""
        params = self._canonical_params()
        dev = params[0].device
        if dev.type != 'cuda':
            raise _lib.NativeLibraryError("parameters must live on a CUDA device (no CPU path)")
        for p in params:
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise _lib.NativeLibraryError("parameters must be contiguous fp32 tensors")
        base = min(p.data_ptr() for p in params)
        if hasattr(self, "_param_offsets"):      # slot table with gaps (NerfModel): -1 for absent slots
            offs = self._param_offsets(base)
        else:
            offs = (C.c_int64 * len(params))(*[(p.data_ptr() - base) // 4 for p in params])
        key = self._pack_key(params)
        for level in range(self._pack_levels):
            hit = self._pack_cache.get(level)
            # the blob is reused as storage when nothing that was handed out can still be referenced by a pending
            # backward: a new tensor per re-pack keeps saved-for-backward blobs of earlier forwards intact
            packed = torch.empty(self._packed_bytes, device=dev, dtype=torch.uint8)
            check(lib().hn_pack_weights(C.byref(self._desc), C.c_void_p(base), offs, level, ptr(packed), stream()),
                  "hn_pack_weights")
            _lib.count(1)
            self._pack_cache[level] = (key, packed)
            del hit

    @contextlib.contextmanager
    def packed_frozen(self):
        """Pack once at entry of the outermost block; every forward inside reuses the blobs."""
        if self._pack_depth == 0:
            self.refresh_packed()
        self._pack_depth += 1
        try:
            yield self
        finally:
            self._pack_depth -= 1

    def _packed_weights(self, level=0):
        hit = self._pack_cache.get(level)
        if self._pack_depth > 0 and hit is not None:
            return hit[1]
        if hit is None or hit[0] != self._pack_key(self._canonical_params()):
            self.refresh_packed()
        return self._pack_cache[level][1]
