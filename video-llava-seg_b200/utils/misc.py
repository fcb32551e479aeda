"""Drop-in replacements for the connected-components helpers of sam2/utils/misc.py:47-63 and
:312-338, backed by libvls_b200.so.  Unlike the reference, a failure here RAISES instead of
silently skipping the post-processing."""
import torch

from .. import _lib
from .._lib import check, lib, ptr, stream


def get_connected_components(mask):
    """(N,1,H,W) binary mask -> (labels int32, counts int32), 8-connectivity.
    Same contract as `sam2._C.get_connected_componnets` (connected_components.cu:213-282)."""
    if not mask.is_cuda:
        raise RuntimeError("inputs must be a CUDA tensor")
    if mask.dim() != 4 or mask.shape[1] != 1:
        raise RuntimeError("inputs must be [N, 1, H, W] shape")
    m = mask.to(torch.uint8).contiguous()
    n, _, h, w = m.shape
    labels = torch.empty((n, 1, h, w), device=m.device, dtype=torch.int32)
    counts = torch.empty((n, 1, h, w), device=m.device, dtype=torch.int32)
    nbytes = lib().vls_cc_workspace_bytes(n, h, w)
    ws = torch.empty(max(nbytes, 1), device=m.device, dtype=torch.uint8)
    check(lib().vls_cc_label(ptr(m), n, h, w, ptr(labels), ptr(counts), ptr(ws), nbytes, stream()), "vls_cc_label")
    return labels, counts


def fill_holes_in_mask_scores(mask, max_area):
    """Fill background components (score <= 0) of area <= max_area with +0.1 (fused kernel:
    binarise, label, measure and patch in one launch; no label/area tensors are materialised)."""
    assert max_area > 0, "max_area must be positive"
    if not mask.is_cuda:
        raise RuntimeError("inputs must be a CUDA tensor")
    out = mask.to(torch.float32).clone(memory_format=torch.contiguous_format)
    n, c, h, w = out.shape
    assert c == 1
    nbytes = lib().vls_fill_holes_workspace_bytes(n, h, w)
    ws = torch.empty(max(nbytes, 1), device=out.device, dtype=torch.uint8)
    check(lib().vls_fill_holes(ptr(out), n, h, w, int(max_area), 0.1, ptr(ws), nbytes, stream()), "vls_fill_holes")
    return out
