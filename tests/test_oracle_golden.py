"""CPU: pin the oracle restatement against the golden fixtures produced by the reference's own
pure-torch test references (tests/golden/make_golden.py)."""
import torch

from oracle import moe_oracle as O


def _aligned(w_ref, ids_ref, w, ids):
    """align (weights, ids) rows by id so unordered references compare as sets."""
    o1 = torch.argsort(ids_ref.long(), dim=1)
    o2 = torch.argsort(ids.long(), dim=1)
    return (torch.gather(w_ref, 1, o1), torch.gather(ids_ref.long(), 1, o1),
            torch.gather(w, 1, o2), torch.gather(ids.long(), 1, o2))


def test_fused_topk_matches_reference(golden):
    for c in golden["fused_topk"]:
        w, ids = O.topk_gating(c["logits"], c["k"], c["renorm"], c["scoring"], c["bias"])
        wr, ir, w2, i2 = _aligned(c["weights"], c["ids"], w, ids)
        assert torch.equal(ir, i2)
        torch.testing.assert_close(w2, wr, atol=2e-6, rtol=1e-5)
        # reference torch.topk is sorted descending by (biased) score: ordered lists must agree too
        assert torch.equal(ids.long(), c["ids"].long())


def test_grouped_topk_matches_reference_native(golden):
    for c in golden["grouped_topk_native"]:
        w, ids = O.grouped_topk(c["logits"], c["bias"], c["n_group"], c["topk_group"], c["k"],
                                c["renorm"], c["rsf"], c["scoring"])
        wr, ir, w2, i2 = _aligned(c["weights"], c["ids"], w, ids)
        assert torch.equal(ir, i2)  # sets per row (native order is implementation-defined)
        torch.testing.assert_close(w2, wr, atol=2e-5, rtol=1e-4)


def test_expert_map(golden):
    for c in golden["expert_map"]:
        local, emap = O.determine_expert_map(c["ep"], c["rank"], c["E"])
        assert local == c["local"]
        assert torch.equal(emap, c["emap"])
        ids = torch.tensor([[0, c["E"] - 1, -1], [3, 1, 2]], dtype=torch.int32)
        loc = O.global_to_local_expert_ids(ids, emap)
        assert loc[0, 2] == -1 and loc.dtype == torch.int32


def test_per_token_group_quant(golden):
    c = golden["ptg_quant"]
    q, s = O.per_token_group_quant_fp8(c["x"], 128)
    assert torch.equal(q.view(torch.uint8), c["q"].view(torch.uint8))
    assert torch.equal(s, c["s"])


def test_block_matmul(golden):
    c = golden["block_matmul"]
    out = O.w8a8_block_matmul(c["xq"], c["xs"], c["wq"], c["ws"])
    torch.testing.assert_close(out, c["out"], atol=1e-5, rtol=1e-5)


def test_experts_bf16(golden):
    c = golden["experts_bf16"]
    w = O.DequantExperts(c["w1"].float(), c["w2"].float())
    out = O.experts_forward(c["a"], w, c["topk_ids"], c["topk_weight"])
    outb = O.experts_forward_batched(c["a"], w, c["topk_ids"], c["topk_weight"])
    torch.testing.assert_close(out, outb, atol=1e-5, rtol=1e-5)
    # the reference computes in bf16 end to end; tolerance of reference tests/kernels/moe/test_moe.py:233
    torch.testing.assert_close(out, c["out"].float(), atol=2e-2, rtol=0)


def test_experts_fp8_block(golden):
    c = golden["experts_fp8_block"]
    out = O.experts_forward_w8a8_block(c["a"], c["w1q"], c["w1s"], c["w2q"], c["w2s"],
                                       c["topk_ids"], c["topk_weight"])
    # reference tolerance tests/kernels/moe/test_block_fp8.py:207-210 is 0.035; the restatement is much closer
    torch.testing.assert_close(out, c["out"].float(), atol=2e-3, rtol=2e-2)
    # the weight-only (W8A16) oracle must agree with the W8A8 reference within the block-fp8 tolerance
    w = O.DequantExperts(O.dequant_fp8_block(c["w1q"], c["w1s"]), O.dequant_fp8_block(c["w2q"], c["w2s"]))
    out16 = O.experts_forward_batched(c["a"], w, c["topk_ids"], c["topk_weight"])
    rel = (out16 - c["out"].float()).abs().mean() / c["out"].float().abs().mean()
    assert rel < 0.08  # A8 activation-quant noise; informational cross-check between the two oracle modes


def test_mxfp4_and_int4(golden):
    c = golden["mxfp4_dequant"]
    out = O.dequant_mxfp4(c["packed"], c["scale"], out_dtype=torch.float32)
    assert torch.equal(out, c["out"])
    c = golden["int4_pack"]
    # int32 packing along dim 0: nibble i of the word = element i (reference quant_utils.py:493-512)
    pk = c["packed"].T.contiguous().view(torch.uint8)  # [8 cols, 2 words*4 bytes]
    lo, hi = pk & 0xF, pk >> 4
    un = torch.stack([lo, hi], -1).reshape(pk.shape[0], -1).T.int()
    assert torch.equal(un, c["q"])
    assert torch.equal(c["unpacked"], c["q"])


def test_quantisers_roundtrip():
    g = torch.Generator().manual_seed(1)
    w = torch.randn(2, 64, 128, generator=g) / 10
    q, s = O.quant_fp8_block(w)
    assert (O.dequant_fp8_block(q, s) - w).abs().max() < 0.02
    p, sc = O.quant_int4_group(w, 32)
    assert (O.dequant_int4_group(p, sc, 32) - w).abs().max() < 0.04
    p, bs, gs = O.quant_nvfp4(w)
    assert (O.dequant_nvfp4(p, bs, gs) - w).abs().max() < 0.06
    p, e8 = O.quant_mxfp4(w)
    assert (O.dequant_mxfp4(p, e8) - w).abs().max() < 0.09
    # e2m1 grid codes decode to the grid
    codes = torch.arange(16, dtype=torch.uint8)
    vals = O.unpack_e2m1((codes | (codes << 4)).reshape(1, 16))
    assert vals[0, 0::2].tolist() == [0, .5, 1, 1.5, 2, 3, 4, 6, -0.0, -.5, -1, -1.5, -2, -3, -4, -6]


def test_gqa_decode(golden):
    c = golden["gqa_decode"]
    out, lse = O.gqa_decode(c["q"], c["k_cache"], c["v_cache"], torch.tensor(c["kv_lens"]),
                            c["block_tables"], c["scale"])
    torch.testing.assert_close(out, c["out"].float(), atol=1e-5, rtol=1e-5)


def test_mla_decode_self_consistency():
    g = torch.Generator().manual_seed(42)
    B, Hq, page, npg = 2, 16, 16, 6
    qn, qp = torch.randn(B, Hq, 512, generator=g), torch.randn(B, Hq, 64, generator=g)
    cache = torch.randn(npg * B, page, 576, generator=g)
    lens = torch.tensor([40, 96])
    pt = torch.arange(npg * B).reshape(B, npg).int()
    out, lse = O.mla_decode(qn, qp, cache, lens, pt, 0.1)
    # SDPA form of reference test_cutlass_mla_decode.py:150-196
    for b in range(B):
        kv = cache[pt[b].long()].reshape(-1, 576)[:lens[b]]
        att = torch.cat([qn[b], qp[b]], -1) @ kv.T * 0.1
        ref = torch.softmax(att, -1) @ kv[:, :512]
        torch.testing.assert_close(out[b], ref, atol=1e-5, rtol=1e-5)
        torch.testing.assert_close(lse[b], att.logsumexp(-1), atol=1e-5, rtol=1e-5)


def test_permute_is_stable_sort():
    ids = torch.tensor([[3, 1], [1, -1], [0, 3], [1, 2]], dtype=torch.int32)
    srt, off, inv = O.moe_permute(ids, 4)
    assert off.tolist() == [0, 1, 4, 5, 7]
    assert srt[:7].tolist() == [4, 1, 2, 6, 7, 0, 5]
    assert inv.tolist() == [5, 1, 2, -1, 0, 6, 3, 4]


def test_mxfp8_activation_quant(golden):
    """oracle.mx_quant_act (the activation quantisation of the native MXFP4 path) against the reference's own
    pure-torch MXFP8 quantiser (vllm/model_executor/layers/quantization/utils/mxfp8_utils.py:38-86): bit exact,
    including an all-zero block and rows spanning five orders of magnitude."""
    c = golden["mxfp8_quant"]
    x, q, s = c["x"], c["q"], c["scales"]
    ref = (q.float().reshape(*x.shape[:-1], -1, 32) * torch.pow(2.0, s.float() - 127).unsqueeze(-1)).reshape(x.shape)
    assert torch.equal(O.mx_quant_act(x), ref)


def test_nvfp4_dequant(golden):
    """oracle.dequant_nvfp4 against the reference's dequantize_nvfp4_to_dtype (tests/kernels/quantization/
    nvfp4_utils.py:38-62; scale factors converted from the swizzled 128x4 layout to linear by the reference's own
    convert_swizzled_to_linear).  The lk_moe boundary receives the reciprocal global scale (routed_experts.py:
    1686-1688), so a power-of-two global is bit exact and a generic one agrees to fp32 rounding (x * (1/g) vs x / g)."""
    c = golden["nvfp4_dequant"]
    sf = c["sf_linear"].view(torch.float8_e4m3fn)
    for case in c["cases"]:
        mine = O.dequant_nvfp4(c["packed"], sf, torch.tensor(1.0 / case["global_scale"]), out_dtype=torch.float32)
        if case["global_scale"] == 64.0:
            assert torch.equal(mine, case["out"])
        torch.testing.assert_close(mine, case["out"], rtol=1e-6, atol=0)


def test_int4_dequant_matches_reference_quantize_weights(golden):
    """oracle.dequant_int4_group ((q - 8) * s, uint4b8) against the dequantised reference returned by the reference's
    quantize_weights (quantization/utils/quant_utils.py:642-730) for the same codes and group scales: bit exact."""
    c = golden["int4_quantize_weights"]
    q = c["w_q"].T.contiguous()                       # [N, K] codes 0..15 (bias 8 already added by the reference)
    packed = (q[:, 0::2] | (q[:, 1::2] << 4)).to(torch.uint8)   # low nibble = even k (quant_utils.py:493-512)
    mine = O.dequant_int4_group(packed, c["w_s"].T.contiguous(), 32, out_dtype=torch.float32)
    assert torch.equal(mine, c["w_ref"].T.contiguous())


def test_moe_permute_matches_reference_torch_permute(golden):
    """oracle.moe_permute (+ global_to_local_expert_ids under EP) against the reference's torch_permute
    (tests/kernels/moe/test_moe_permute_unpermute.py:37-88): permutation, per-expert offsets, inverse map and the
    gathered rows are bit exact over the valid rows (EP 1 / 4 / 16)."""
    for c in golden["moe_permute"]:
        ids = c["topk_ids"]
        loc = ids if c["expert_map"] is None else O.global_to_local_expert_ids(ids, c["expert_map"])
        order, off, inv = O.moe_permute(loc, c["n_local"])
        nv = c["n_valid"]
        assert int(off[-1]) == nv
        assert torch.equal(order[:nv].long(), c["dst_row_id2src_row_id_map"][:nv].long())
        assert torch.equal(off.long(), c["expert_first_token_offset"].long())
        valid = inv >= 0
        assert torch.equal(inv[valid].long(), c["src_row_id2dst_row_id_map"].flatten()[valid].long())
        assert torch.equal(c["hidden"][order[:nv].long() // ids.shape[1]], c["permuted"][:nv])


def test_swigluoai_packed_activation(golden):
    """oracle.apply_activation(activation_type=1): packed halves, clamp, (up + 1) — bit exact against the reference's
    SiluAndMulWithClamp.forward_native (vllm/model_executor/layers/activation.py:242-246, beta = 1)."""
    c = golden["swigluoai_packed"]
    for case in c["cases"]:
        mine = O.apply_activation(c["x"], O.ACT_SWIGLUOAI, True, alpha=case["alpha"], limit=case["limit"])
        assert torch.equal(mine, case["out"])


def test_global_to_local_expert_ids(golden):
    """oracle.global_to_local_expert_ids against RoutedExperts.global_to_local_expert_ids (routed_experts.py:1332-1342),
    including padded (< 0) ids and ids beyond the map (clamped by the reference)."""
    c = golden["global_to_local"]
    assert torch.equal(O.global_to_local_expert_ids(c["topk_ids"], c["expert_map"]).long(), c["out"].long())


def test_rope_matches_reference_forward_static(golden):
    """oracle.rope_forward_static == the reference's RotaryEmbedding.forward_static (rotary_embedding/base.py:161-201),
    GPT-J (DeepSeek MLA) and NeoX styles, bit exact in bf16."""
    r = golden["rope"]
    for c in r["cases"]:
        q, k = O.rope_forward_static(r["positions"], r["q"].clone(), r["k"].clone(), 64, 64, r["cos_sin_cache"], c["neox"])
        assert torch.equal(q, c["q_out"]) and torch.equal(k, c["k_out"])


def test_fp8_ue8m0_matches_reference(golden):
    """oracle ue8m0 helpers == the reference's DeepGEMM-on-Blackwell functions (vllm/utils/deep_gemm.py:644-681
    per_block_cast_to_fp8 / _ceil_to_ue8m0, fp8_utils.py:986-1043 requant_weight_ue8m0_inplace), bit exact."""
    c = golden["fp8_ue8m0"]
    q, s = O.per_block_cast_to_fp8(c["x"], (128, 128), ue8m0=True)
    assert torch.equal(q.view(torch.uint8), c["q"].view(torch.uint8)) and torch.equal(s, c["s"])
    wq, ws = O.requant_weight_ue8m0(c["w_q"], c["w_s"])
    assert torch.equal(wq.view(torch.uint8), c["w_q_out"].view(torch.uint8)) and torch.equal(ws, c["w_s_out"])
    assert torch.equal(O.ceil_to_ue8m0(c["sc"]), c["sc_ceil"])
    # every requantised scale is a power of two and nothing saturates
    assert torch.equal(ws, torch.pow(2.0, torch.round(torch.log2(ws))))
    assert float(wq.float().abs().max()) <= 448.0


def test_rms_norm_matches_reference_native_ops():
    """oracle.rms_norm / fused_add_rms_norm == the reference's native ops (vllm/ir/ops/layernorm.py:10-21, :44-63, what
    RMSNorm.forward_native runs), bf16 / fp16 / fp32, with and without weight, bit exact (golden generated by
    tests/golden/make_golden_norm.py from the imported reference)."""
    import os
    gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_norm.pt"), weights_only=True)
    assert len(gold["cases"]) == 9
    for c in gold["cases"]:
        assert torch.equal(O.rms_norm(c["x"], c["weight"], c["eps"]), c["rms_norm"])
        y, r = O.fused_add_rms_norm(c["x"], c["residual"], c["weight"], c["eps"])
        assert torch.equal(y, c["fused_y"]) and torch.equal(r, c["fused_residual"])
