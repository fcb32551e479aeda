"""Drop-in `MemoryAttention` (sam2/modeling/memory_attention.py:102-169): same constructor, same
state_dict keys, same forward signature; the whole 4-layer stack runs in one C call
(vls_mem_attn_forward: tcgen05 GEMMs with RoPE epilogues + tcgen05 flash attention)."""
import copy

import torch
from torch import nn

from .. import _pack
from .._lib import check, lib, ptr, stream
from ._base import PackedModule, ctypes_ref, dtype_code, require_cuda
from .sam.transformer import RoPEAttention


class MemoryAttentionLayer(nn.Module):
    """Parameter container mirroring memory_attention.py:17-56 (self_attn, cross_attn_image, linear1/2, norm1-3)."""

    def __init__(self, activation, cross_attention, d_model, dim_feedforward, dropout, pos_enc_at_attn,
                 pos_enc_at_cross_attn_keys, pos_enc_at_cross_attn_queries, self_attention):
        super().__init__()
        if activation != "relu" or pos_enc_at_attn or not pos_enc_at_cross_attn_keys or pos_enc_at_cross_attn_queries:
            raise NotImplementedError("CUDA path implements the SAM 2.1 layer: relu FFN, positional encoding added "
                                      "to cross-attention keys only")
        self.d_model, self.dim_feedforward, self.dropout_value = d_model, dim_feedforward, dropout
        self.self_attn, self.cross_attn_image = self_attention, cross_attention
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(d_model), nn.LayerNorm(d_model), nn.LayerNorm(d_model)
        self.activation_str = activation
        self.pos_enc_at_attn = pos_enc_at_attn
        self.pos_enc_at_cross_attn_queries = pos_enc_at_cross_attn_queries
        self.pos_enc_at_cross_attn_keys = pos_enc_at_cross_attn_keys


class MemoryAttention(PackedModule):
    def __init__(self, d_model, pos_enc_at_input, layer, num_layers, batch_first=True):
        super().__init__()
        if not batch_first:
            raise NotImplementedError("batch_first=False is not used by any SAM 2 config")
        self.d_model = d_model
        self.layers = nn.ModuleList([copy.deepcopy(layer) for _ in range(num_layers)])
        self.num_layers = num_layers
        self.norm = nn.LayerNorm(d_model)
        self.pos_enc_at_input = pos_enc_at_input
        self.batch_first = batch_first
        self._rope = {}
        # bumped by everything that runs on the module's workspace (eager calls and graph replays): a head computed ahead of
        # its frame (phase 1) is only valid while nobody else has used the workspace since
        self.ws_epoch = 0

    def _apply(self, fn, *a, **kw):
        self._rope = {}
        return super()._apply(fn, *a, **kw)

    def forward(self, curr, memory, curr_pos=None, memory_pos=None, num_obj_ptr_tokens=0, phase=0, keys_ahead=(0, 0, 0)):
        """curr / curr_pos: [Nq,B,256] (or 1-element lists of it); memory / memory_pos: [Nk,B,64];
        returns [Nq,B,256] in curr's dtype.  Same contract as memory_attention.py:119-169.
        phase (not in the reference): 1 = only the part that depends on `curr` alone (layer 0 up to its cross-attention
        queries, left in the workspace; returns None), 2 = the rest for the head run last, 0 = both (vls_b200.h).
        keys_ahead = (rows, shift_from, shift): the head also projects layer 0's keys of the first `rows` memory rows."""
        if isinstance(curr, list):
            assert isinstance(curr_pos, list) and len(curr) == len(curr_pos) == 1
            curr, curr_pos = curr[0], curr_pos[0]
        assert curr.shape[1] == memory.shape[1], "Batch size must be the same for curr and memory"
        require_cuda(curr, memory, curr_pos, memory_pos)
        dev = curr.device
        if self._packed is None:
            self._packed = _pack.pack_mem_attn(self._flat_sd(), "", dev)
        w = self._packed[0]
        nq, b, c = curr.shape
        nk = memory.shape[0]
        if nq not in self._rope:
            self._rope[nq] = _pack.axial_rope_tables(nq, dev)
        cos, sin = self._rope[nq]
        w.rope_cos, w.rope_sin, w.rope_len = cos.data_ptr(), sin.data_ptr(), nq

        def rows(t):  # (tensor, dtype, token stride, batch stride) with unit channel stride
            if t is None:
                return None, 0, 0, 0
            if t.stride(2) != 1:
                t = t.contiguous()
            return t, dtype_code(t), t.stride(0), t.stride(1)

        cu, cpos = rows(curr), rows(curr_pos if self.pos_enc_at_input else None)
        me, mpos = rows(memory), rows(memory_pos)
        out = torch.empty((nq, b, c), device=dev, dtype=curr.dtype) if phase in (0, 2) else None   # 1, 3, 4: head only
        nbytes = lib().vls_mem_attn_workspace_bytes(b, nq, nk)
        ws = self._workspace(nbytes, dev)
        self.ws_epoch += 1
        check(lib().vls_mem_attn_forward_phase(
            ctypes_ref(w), ptr(cu[0]), cu[1], cu[2], cu[3], ptr(cpos[0]), cpos[1], cpos[2], cpos[3],
            ptr(me[0]), me[1], me[2], me[3], ptr(mpos[0]), mpos[1], mpos[2], mpos[3], b, nq, nk,
            int(num_obj_ptr_tokens), ptr(out), 0 if out is None else dtype_code(out), 0 if out is None else out.stride(0),
            0 if out is None else out.stride(1), ptr(ws), ws.numel(), stream(), int(phase), int(keys_ahead[0]),
            int(keys_ahead[1]), int(keys_ahead[2])), "vls_mem_attn_forward_phase")
        return out

