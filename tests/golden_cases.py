"""Seeded inputs shared by tests/golden/make_golden.py (reference side, build container) and the
parity tests (oracle / CUDA side).  The draw ORDER is part of the fixture contract."""
import torch


def module_inputs():
    g = torch.Generator().manual_seed(1234)
    rn = lambda *s: torch.randn(*s, generator=g)
    d = {}
    nq, b = 256, 2          # 16x16 query grid, 2 memory frames + 2 pointers (8 tokens)
    nk = 2 * nq + 8
    d["curr"], d["curr_pos"] = rn(nq, b, 256) * 0.5, rn(nq, b, 256) * 0.5
    d["mem"], d["mem_pos"] = rn(nk, b, 64) * 0.5, rn(nk, b, 64) * 0.5
    d["n_ptr_tokens"] = 8
    d["emb"] = rn(2, 256, 64, 64) * 0.5
    d["s0"], d["s1"] = rn(2, 32, 256, 256) * 0.3, rn(2, 64, 128, 128) * 0.3
    d["sparse"] = rn(2, 2, 256)
    d["seg"] = rn(3, 1, 256)
    d["pix"] = rn(2, 256, 64, 64) * 0.5
    d["msk"] = torch.sigmoid(rn(2, 1, 1024, 1024) * 3) * 20 - 10
    return d
