"""GPU parity of the three drop-in modules (called through libvls_b200's C ABI) against
(a) the CPU oracle run live on the same seeded inputs and (b) the golden vectors the unmodified
reference produced for those inputs (tests/golden/modules.npz).

Tolerances: activations are bf16 with f32 accumulation (BASELINE.json north_star), so module outputs of
magnitude ~1 are compared at max-abs 4e-2 / mean-abs 4e-3; decoder mask logits at the north-star bound
1e-2 abs."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "modules.npz")


@pytest.fixture(scope="module")
def env(vls_lib):
    from tests import golden_cases
    from video_llava_seg_b200 import build_sam, synth

    assert torch.cuda.is_available()
    dev = torch.device("cuda:0")
    sd = synth.init_state_dict(0)
    return dict(dev=dev, sd=sd, inp=golden_cases.module_inputs(), gold=np.load(GOLD), build=build_sam)


def _stats(a, b):
    d = (a.float().cpu() - b.float().cpu()).abs()
    return d.max().item(), d.mean().item()


def test_memory_attention_small(env):
    from oracle import sam2_path as O

    dev, sd, inp = env["dev"], env["sd"], env["inp"]
    m = env["build"].load_prefixed(env["build"].build_memory_attention(), sd, "memory_attention.").to(dev).eval()
    out = m(curr=[inp["curr"].to(dev)], curr_pos=[inp["curr_pos"].to(dev)], memory=inp["mem"].to(dev),
            memory_pos=inp["mem_pos"].to(dev), num_obj_ptr_tokens=inp["n_ptr_tokens"])
    ref = O.memory_attention(sd, inp["curr"], inp["mem"], inp["curr_pos"], inp["mem_pos"], inp["n_ptr_tokens"])
    mx, mean = _stats(out, ref)
    gmx, _ = _stats(out, torch.from_numpy(env["gold"]["memattn_out"]))
    print(f"memory_attention small: max {mx:.3e} mean {mean:.3e} vs golden max {gmx:.3e}")
    assert out.shape == ref.shape and out.dtype == torch.float32
    assert mx < 4e-2 and mean < 4e-3 and gmx < 4e-2


def test_memory_attention_full_size_bf16_io(env):
    """Steady-state shape of BASELINE config 2: Nq=4096, Nk=7*4096+64, bf16 memory + f32 positions."""
    from oracle import sam2_path as O

    dev, sd = env["dev"], env["sd"]
    g = torch.Generator().manual_seed(77)
    nq, nk, b = 4096, 7 * 4096 + 64, 1
    curr = torch.randn(nq, b, 256, generator=g) * 0.5
    cpos = torch.randn(nq, b, 256, generator=g) * 0.5
    mem = (torch.randn(nk, b, 64, generator=g) * 0.5).bfloat16()
    mpos = torch.randn(nk, b, 64, generator=g) * 0.5
    m = env["build"].load_prefixed(env["build"].build_memory_attention(), sd, "memory_attention.").to(dev).eval()
    out = m(curr.to(dev), mem.to(dev), cpos.to(dev), mpos.to(dev), 64)
    ref = O.memory_attention(sd, curr, mem.float(), cpos, mpos, 64)
    mx, mean = _stats(out, ref)
    print(f"memory_attention full: max {mx:.3e} mean {mean:.3e}")
    assert mx < 5e-2 and mean < 4e-3


def test_memory_attention_mid_fused_matches_kernel_chain(env):
    """mid_fused.cu (self-attention out-proj + residual, LayerNorm2, cross-attention q-proj + RoPE in one cluster kernel)
    against the GEMM / LayerNorm / GEMM chain it replaces, at a ragged query count (2 objects, Nq = 4096) and the small shape."""
    from video_llava_seg_b200 import _lib

    dev, sd, inp = env["dev"], env["sd"], env["inp"]
    m = env["build"].load_prefixed(env["build"].build_memory_attention(), sd, "memory_attention.").to(dev).eval()
    g = torch.Generator().manual_seed(5)
    nq, nk, b = 4096, 2 * 4096 + 16, 2
    big = (torch.randn(nq, b, 256, generator=g) * 0.5, (torch.randn(nk, b, 64, generator=g) * 0.5).bfloat16(),
           torch.randn(nq, b, 256, generator=g) * 0.5, torch.randn(nk, b, 64, generator=g) * 0.5, 16)
    small = (inp["curr"], inp["mem"], inp["curr_pos"], inp["mem_pos"], inp["n_ptr_tokens"])
    lib = _lib.lib()
    try:
        for name, (curr, mem, cpos, mpos, nptr) in (("small", small), ("big", big)):
            outs = []
            for fused in (1, 0):
                lib.vls_set_tuning(b"mid_fused", fused)
                lib.vls_set_tuning(b"tail_quarter", fused)   # layer-tail prologue by column quarters vs full width per CTA
                outs.append(m(curr.to(dev), mem.to(dev), cpos.to(dev), mpos.to(dev), nptr).float().clone())
            mx, mean = _stats(outs[0], outs[1])
            print(f"mid_fused vs chain ({name}): max {mx:.3e} mean {mean:.3e}")
            assert mx < 2e-2 and mean < 1e-3, (name, mx, mean)
    finally:
        lib.vls_set_tuning(b"mid_fused", 1)
        lib.vls_set_tuning(b"tail_quarter", 1)


def test_memory_attention_head_and_rest_are_bit_identical_to_one_call(env):
    """vls_mem_attn_forward_phase: head (phase 1) + rest (phase 2) against the one-call form, bit for bit -- plain split, layer 0's
    keys of the first memories projected by the head, and the head run on a sliding-window bank BEFORE its shift (what the
    pipelined steady-state graph does: graphed._frame_body), for every tuning of the head, at 1 and 2 objects."""
    from video_llava_seg_b200 import _lib

    dev, sd = env["dev"], env["sd"]
    m = env["build"].load_prefixed(env["build"].build_memory_attention(), sd, "memory_attention.").to(dev).eval()
    lib = _lib.lib()
    g = torch.Generator().manual_seed(11)
    nq, slots, nptr = 1024, 4, 16                       # 32 x 32 tokens, 4 memories + 4 pointers of 4 tokens
    nk = slots * nq + nptr
    try:
        for b in (1, 2):
            curr = (torch.randn(nq, b, 256, generator=g) * 0.5).to(dev)
            cpos = (torch.randn(nq, b, 256, generator=g) * 0.5).to(dev)
            mpos = (torch.randn(nk, 1, 64, generator=g) * 0.5).expand(nk, b, 64).contiguous().to(dev)
            window = (torch.randn(nk + nq, b, 64, generator=g) * 0.5).bfloat16().to(dev)   # one memory more than the bank
            # the bank before the shift: [cond | m1 m2 m3 | ptrs];  after: [cond | m2 m3 m4 | ptrs']
            before = torch.cat([window[:slots * nq], window[(slots + 1) * nq:]]).contiguous()
            after = torch.cat([window[:nq], window[2 * nq:(slots + 1) * nq], window[(slots + 1) * nq:].flip(0)]).contiguous()
            whole = m(curr, after, cpos, mpos, nptr).clone()
            r0 = (slots - 1) * nq
            for short, inline, ahead_all in ((0, 1, 0), (1, 1, 0), (0, 0, 0), (0, 1, 1)):
                lib.vls_set_tuning(b"mem_attn_head_short", short)
                lib.vls_set_tuning(b"mem_attn_keys0_inline", inline)
                lib.vls_set_tuning(b"mem_attn_keys_ahead_all", ahead_all)
                # (a) plain split on the final bank, (b) keys ahead on the final bank, (c) keys ahead on the unshifted bank
                for name, head_mem, head_keys, rest_keys in (("split", after, (0, 0, 0), (0, 0, 0)),
                                                             ("keys ahead", after, (r0, r0, 0), (r0, 0, 0)),
                                                             ("keys ahead, shifted", before, (r0, nq, nq), (r0, 0, 0))):
                    m._ws.zero_()                      # nothing of the one-call run may survive in the workspace
                    if name == "split" or short:
                        assert m(curr, head_mem, cpos, mpos, nptr, phase=1, keys_ahead=head_keys) is None
                    else:                              # the head in its two halves
                        assert m(curr, head_mem, cpos, mpos, nptr, phase=3, keys_ahead=head_keys) is None
                        assert m(torch.zeros_like(curr), head_mem, cpos, mpos, nptr, phase=4, keys_ahead=head_keys) is None
                    out = m(torch.zeros_like(curr), after, cpos, mpos, nptr, phase=2, keys_ahead=rest_keys)
                    assert torch.equal(out, whole), (b, short, inline, ahead_all, name, (out - whole).abs().max().item())
        import pytest
        with pytest.raises(RuntimeError):               # keys ahead must be whole rotated blocks
            m(curr, after, cpos, mpos, nptr, phase=1, keys_ahead=(nq // 2, 0, 0))
    finally:
        lib.vls_set_tuning(b"mem_attn_head_short", 0)
        lib.vls_set_tuning(b"mem_attn_keys0_inline", 1)
        lib.vls_set_tuning(b"mem_attn_keys_ahead_all", 0)


def test_mask_decoder_video_and_llava(env):
    from oracle import sam2_path as O

    dev, sd, inp, gold = env["dev"], env["sd"], env["inp"], env["gold"]
    B = env["build"]
    dec = B.load_prefixed(B.build_mask_decoder(), sd, "sam_mask_decoder.").to(dev).eval()
    pe_mod = B.load_prefixed(B.build_prompt_encoder(), sd, "sam_prompt_encoder.").to(dev).eval()
    pe = pe_mod.get_dense_pe()
    assert (pe.cpu() - O.dense_pe(sd)).abs().max() < 1e-5
    dense = pe_mod.no_mask_embed.weight.reshape(1, -1, 1, 1).expand(2, -1, 64, 64)
    emb, s0, s1 = inp["emb"].to(dev), inp["s0"].to(dev), inp["s1"].to(dev)
    got = dec(image_embeddings=emb, image_pe=pe, sparse_prompt_embeddings=inp["sparse"].to(dev),
              dense_prompt_embeddings=dense, multimask_output=True, repeat_image=False, high_res_features=[s0, s1])
    ref = O.mask_decoder(sd, inp["emb"], O.dense_pe(sd), inp["sparse"], O.dense_no_mask(sd, 2), True, False,
                         [inp["s0"], inp["s1"]])
    names = ("masks", "iou", "tokens", "obj")
    for n, a, r in zip(names, got, ref):
        mx, mean = _stats(a, r)
        print(f"decoder(video) {n}: max {mx:.3e} mean {mean:.3e} ref-range [{r.min():.2f},{r.max():.2f}]")
        assert a.shape == r.shape
        assert mx < (1e-2 if n in ("masks", "iou", "obj") else 4e-2), (n, mx)
    assert _stats(got[0][:, :, ::4, ::4], torch.from_numpy(gold["dec_video_masks_s4"]))[0] < 1e-2
    assert _stats(got[1], torch.from_numpy(gold["dec_video_iou"]))[0] < 1e-2
    assert _stats(got[3], torch.from_numpy(gold["dec_video_obj"]))[0] < 1e-2
    # LLaVA seg-head flavour: one [SEG] embedding per object, image repeated, single mask (llava sam2.py:103-114)
    dense3 = dense[:1].expand(3, -1, -1, -1)
    got = dec(image_embeddings=emb[:1], image_pe=pe, sparse_prompt_embeddings=inp["seg"].to(dev),
              dense_prompt_embeddings=dense3, multimask_output=False, repeat_image=True,
              high_res_features=[s0[:1], s1[:1]])
    assert got[0].shape == (3, 1, 256, 256) and got[1].shape == (3, 1) and got[2].shape == (3, 1, 256)
    assert _stats(got[0][:, :, ::4, ::4], torch.from_numpy(gold["dec_llava_masks_s4"]))[0] < 1e-2
    assert _stats(got[1], torch.from_numpy(gold["dec_llava_iou"]))[0] < 1e-2
    # bf16 activations in (how the LLaVA head calls it after .to(bfloat16))
    got16 = dec(image_embeddings=emb[:1].bfloat16(), image_pe=pe, sparse_prompt_embeddings=inp["seg"].to(dev).bfloat16(),
                dense_prompt_embeddings=dense3.bfloat16(), multimask_output=False, repeat_image=True,
                high_res_features=[s0[:1].bfloat16(), s1[:1].bfloat16()])
    assert got16[0].dtype == torch.bfloat16
    assert _stats(got16[0][:, :, ::4, ::4], torch.from_numpy(gold["dec_llava_masks_s4"]))[0] < 3e-2


def test_mask_decoder_fused_paths_match_kernel_chain(env):
    """The cluster-kernel token side (dec_tok.cu) and image side (dec_img.cu) and the tcgen05 ConvT#2 (up2_masks_tc) against the chain of small kernels
    they replace (vls_set_tuning dec_fused / up2_tc = 0): same math, different summation order / operand rounding."""
    from video_llava_seg_b200 import _lib

    dev, sd, inp = env["dev"], env["sd"], env["inp"]
    B = env["build"]
    dec = B.load_prefixed(B.build_mask_decoder(), sd, "sam_mask_decoder.").to(dev).eval()
    pe_mod = B.load_prefixed(B.build_prompt_encoder(), sd, "sam_prompt_encoder.").to(dev).eval()
    pe = pe_mod.get_dense_pe()
    emb, s0, s1 = inp["emb"].to(dev), inp["s0"].to(dev), inp["s1"].to(dev)
    g = torch.Generator().manual_seed(3)
    lib = _lib.lib()
    try:
        for ns in (2, 3, 9):   # 8, 9 (two token tiles) and 15 token rows
            sparse = torch.randn(2, ns, 256, generator=g).to(dev)
            dense = pe_mod.no_mask_embed.weight.reshape(1, -1, 1, 1).expand(2, -1, 64, 64)
            outs = []
            for fused in (1, 0):
                lib.vls_set_tuning(b"dec_fused", fused)
                lib.vls_set_tuning(b"dec_img_fused", fused)
                lib.vls_set_tuning(b"up2_tc", fused)
                outs.append([o.float().clone() for o in dec(image_embeddings=emb, image_pe=pe, sparse_prompt_embeddings=sparse,
                                                           dense_prompt_embeddings=dense, multimask_output=True,
                                                           repeat_image=False, high_res_features=[s0, s1])])
            for n, a, r in zip(("masks", "iou", "tokens", "obj"), *outs):
                mx, mean = _stats(a, r)
                print(f"fused vs chain, {6 + ns} tokens, {n}: max {mx:.3e} mean {mean:.3e}")
                assert mx < 5e-3, (ns, n, mx)
    finally:
        lib.vls_set_tuning(b"dec_fused", 1)
        lib.vls_set_tuning(b"dec_img_fused", 1)
        lib.vls_set_tuning(b"up2_tc", 1)


def test_memory_encoder(env):
    from oracle import sam2_path as O

    dev, sd, inp, gold = env["dev"], env["sd"], env["inp"], env["gold"]
    B = env["build"]
    enc = B.load_prefixed(B.build_memory_encoder(), sd, "memory_encoder.").to(dev).eval()
    out = enc(inp["pix"].to(dev), inp["msk"].to(dev), skip_mask_sigmoid=True)
    ref = O.memory_encoder(sd, inp["pix"], inp["msk"], True)
    mx, mean = _stats(out["vision_features"], ref["vision_features"])
    print(f"memory_encoder: max {mx:.3e} mean {mean:.3e} ref-range [{ref['vision_features'].min():.2f},"
          f"{ref['vision_features'].max():.2f}]")
    assert mx < 4e-2 and mean < 4e-3
    assert _stats(out["vision_features"][:, :, ::2, ::2], torch.from_numpy(gold["memenc_feat_s2"]))[0] < 4e-2
    assert _stats(out["vision_pos_enc"][0][0], torch.from_numpy(gold["memenc_pos0"]))[0] < 1e-5
    # sigmoid inside (skip_mask_sigmoid=False)
    raw = torch.randn(2, 1, 1024, 1024, generator=torch.Generator().manual_seed(3))
    o2 = enc(inp["pix"].to(dev), raw.to(dev), skip_mask_sigmoid=False)
    r2 = O.memory_encoder(sd, inp["pix"], raw, False)
    assert _stats(o2["vision_features"], r2["vision_features"])[0] < 4e-2
    # fused low-res fast path == bilinear x4 + sigmoid*20-10 + encoder + occlusion embedding
    low = torch.randn(2, 1, 256, 256, generator=torch.Generator().manual_seed(4)) * 2
    vf = inp["pix"].flatten(2).permute(2, 0, 1).contiguous()
    gate = torch.tensor([0.0, 1.0])
    nchw, rows = enc.encode_from_low_res(vf.to(dev), low.to(dev), False, 20.0, -10.0, gate.to(dev),
                                         sd["no_obj_embed_spatial"])
    hi = torch.nn.functional.interpolate(low, size=(1024, 1024), mode="bilinear", align_corners=False)
    r3 = O.memory_encoder(sd, inp["pix"], torch.sigmoid(hi) * 20 - 10, True)["vision_features"]
    r3 = r3 + gate[:, None, None, None] * sd["no_obj_embed_spatial"][..., None, None]
    mx, mean = _stats(nchw, r3)
    print(f"memory_encoder fused low-res: max {mx:.3e} mean {mean:.3e}")
    assert mx < 5e-2 and mean < 5e-3
    assert _stats(rows.transpose(1, 2).reshape(2, 64, 64, 64), nchw)[0] == 0.0
