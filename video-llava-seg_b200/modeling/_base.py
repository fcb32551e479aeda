"""Shared plumbing of the host-side module mirrors: parameter containers keep the reference's
state_dict keys; forward() hands raw pointers to libvls_b200.so.  Packed (bf16 / re-laid-out)
weights are rebuilt lazily whenever the parameters are moved or reloaded."""
import ctypes

import torch
from torch import nn

from .._lib import VLS_DTYPE


class PackedModule(nn.Module):
    """nn.Module whose forward runs in the C library from a packed copy of its parameters."""

    def __init__(self):
        super().__init__()
        self._packed = None
        self._ws = None
        self.register_load_state_dict_post_hook(lambda m, keys: m.invalidate())

    def invalidate(self):
        self._packed = None

    def _apply(self, fn, *a, **kw):
        self._packed = None
        self._ws = None
        return super()._apply(fn, *a, **kw)

    def _device(self):
        return next(self.parameters()).device

    def _flat_sd(self):
        return {k: v for k, v in self.state_dict().items()}

    def _workspace(self, nbytes, device):
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != device:
            self._ws = torch.empty(max(int(nbytes), 256), device=device, dtype=torch.uint8)
        return self._ws


def dtype_code(t):
    try:
        return VLS_DTYPE[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported activation dtype {t.dtype}; expected float32 or bfloat16")


def require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("this path has no CPU implementation: inputs must be CUDA tensors")


def ctypes_ref(struct):
    """void* to a ctypes Structure (the struct must stay referenced by the caller during the call)."""
    return ctypes.cast(ctypes.pointer(struct), ctypes.c_void_p)


def batch_shared(x):
    """(tensor, batch stride) for a per-image tensor that may be an expand()ed view over the batch:
    collapses a zero-stride batch to one image so nothing is materialised."""
    if x.shape[0] > 1 and x.stride(0) == 0:
        x = x[:1]
    x = x.contiguous()
    return x, (x.stride(0) if x.shape[0] > 1 else 0)
