"""Drop-in replacements for the connected-components helpers of sam2/utils/misc.py:47-63 and
:312-338, backed by libvls_b200.so.  Unlike the reference, a failure here RAISES instead of
silently skipping the post-processing."""
import torch

from .. import _lib
from .._lib import check, lib, ptr, stream


_WS = {}


def _workspace(nbytes, device):
    """Scratch buffer of the tiled path, kept per (device, stream) and grown on demand (r1 allocated one per call)."""
    if torch.cuda.is_current_stream_capturing():      # a captured graph keeps its own allocation alive (private pool)
        return torch.empty(max(nbytes, 1), device=device, dtype=torch.uint8)
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _WS.get(key)
    if ws is None or ws.numel() < max(nbytes, 1):
        ws = torch.empty(max(nbytes, 256), device=device, dtype=torch.uint8)
        _WS[key] = ws
    return ws


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
    ws = _workspace(nbytes, m.device)
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
    ws = _workspace(nbytes, out.device)
    check(lib().vls_fill_holes(ptr(out), n, h, w, int(max_area), 0.1, ptr(ws), nbytes, stream()), "vls_fill_holes")
    return out


def load_video_frames(video_path, image_size, offload_video_to_cpu, img_mean=(0.485, 0.456, 0.406),
                      img_std=(0.229, 0.224, 0.225), async_loading_frames=False, compute_device=torch.device("cuda")):
    """Frames -> ([T,3,S,S] normalised f32, video_height, video_width); host-side IO next to the hot path
    (sam2/utils/misc.py:172-309).  Accepts a JPEG folder (frames named by integer index, as the reference
    requires) or an already-normalised [T,3,H,W] tensor.  MP4 decoding needs `decord`, which this image lacks."""
    import os

    if isinstance(video_path, torch.Tensor):
        t = video_path
        assert t.dim() == 4 and t.shape[1] == 3, "expected a [T,3,H,W] tensor of normalised frames"
        h, w = int(t.shape[-2]), int(t.shape[-1])
        if (h, w) != (image_size, image_size):
            t = torch.nn.functional.interpolate(t.float(), size=(image_size, image_size), mode="bilinear",
                                                align_corners=False, antialias=True)
        return (t if offload_video_to_cpu else t.to(compute_device)), h, w
    if isinstance(video_path, str) and os.path.isdir(video_path):
        import numpy as np
        from PIL import Image

        names = [p for p in os.listdir(video_path) if os.path.splitext(p)[-1].lower() in (".jpg", ".jpeg")]
        names.sort(key=lambda p: int(os.path.splitext(p)[0]))
        if not names:
            raise RuntimeError(f"no images found in {video_path}")
        mean = torch.tensor(img_mean, dtype=torch.float32)[:, None, None]
        std = torch.tensor(img_std, dtype=torch.float32)[:, None, None]
        frames, h, w = [], None, None
        for n in names:
            im = Image.open(os.path.join(video_path, n))
            w, h = im.size
            a = np.array(im.convert("RGB").resize((image_size, image_size)))
            frames.append((torch.from_numpy(a).permute(2, 0, 1).float() / 255.0 - mean) / std)
        images = torch.stack(frames, 0)
        return (images if offload_video_to_cpu else images.to(compute_device)), h, w
    raise NotImplementedError("Only JPEG folders and frame tensors are supported (MP4 needs decord)")
