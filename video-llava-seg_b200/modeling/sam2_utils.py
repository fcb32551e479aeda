"""Parameter containers with the reference's layout (sam2/modeling/sam2_utils.py:112-153) plus the
two host-side helpers of that file that the tracking core needs."""
import torch
from torch import nn


class MLP(nn.Module):
    """`layers.{i}.weight/bias` container; the arithmetic runs inside libvls_b200 (small_linear kernels)."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers, activation=nn.ReLU, sigmoid_output=False):
        super().__init__()
        if activation is not nn.ReLU:
            raise NotImplementedError("the CUDA path implements ReLU MLP heads only")
        self.num_layers, self.sigmoid_output = num_layers, sigmoid_output
        dims = [input_dim] + [hidden_dim] * (num_layers - 1) + [output_dim]
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))


class LayerNorm2d(nn.Module):
    def __init__(self, num_channels, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(num_channels))
        self.bias = nn.Parameter(torch.zeros(num_channels))
        self.eps = eps


def get_1d_sine_pe(pos_inds, dim, temperature=10000):
    """Host-side (tiny) temporal encoding of object pointers, sam2_utils.py:64-74."""
    pe_dim = dim // 2
    dim_t = torch.arange(pe_dim, dtype=torch.float32, device=pos_inds.device)
    dim_t = temperature ** (2 * (dim_t // 2) / pe_dim)
    e = pos_inds.unsqueeze(-1) / dim_t
    return torch.cat([e.sin(), e.cos()], dim=-1)


def select_closest_cond_frames(frame_idx, cond_frame_outputs, max_cond_frame_num):
    """sam2_utils.py:19-61: keep the nearest conditioning frames (one before, one after, then by distance)."""
    if max_cond_frame_num == -1 or len(cond_frame_outputs) <= max_cond_frame_num:
        return cond_frame_outputs, {}
    assert max_cond_frame_num >= 2, "we should allow using 2+ conditioning frames"
    chosen = {}
    before = [t for t in cond_frame_outputs if t < frame_idx]
    after = [t for t in cond_frame_outputs if t >= frame_idx]
    if before:
        chosen[max(before)] = cond_frame_outputs[max(before)]
    if after:
        chosen[min(after)] = cond_frame_outputs[min(after)]
    rest = sorted((t for t in cond_frame_outputs if t not in chosen), key=lambda t: abs(t - frame_idx))
    for t in rest[: max_cond_frame_num - len(chosen)]:
        chosen[t] = cond_frame_outputs[t]
    return chosen, {t: v for t, v in cond_frame_outputs.items() if t not in chosen}
