"""Generate golden fixtures from the reference's OWN pure-torch test references.

Run in the build container only (needs /root/reference):
    cd /tmp && PYTHONPATH=/root/reference python /root/repo/tests/golden/make_golden.py

It imports the reference's python helpers (no compiled ops, CPU only), feeds them seeded inputs and
stores inputs + outputs in tests/golden/golden_ref.pt.  The fixtures travel to the GPU box; the
reference tree does not.  Nothing here is imported by the product.
"""
import os
import sys

import torch

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_ref.pt")


def main():
    sys.path.insert(0, "/root/reference")
    from tests.kernels.moe.test_fused_topk import torch_topk  # tests/kernels/moe/test_fused_topk.py:19-45
    from tests.kernels.quant_utils import (native_per_token_group_quant_fp8,  # :157-180
                                           native_w8a8_block_matmul)  # :91-154
    from tests.kernels.utils import torch_experts  # tests/kernels/utils.py:855-994
    from tests.kernels.moe.test_ocp_mx_moe import mxfp4_dequantize  # :153-171
    from tests.kernels.attention.test_flashinfer import ref_paged_attn  # :29-80
    from vllm.model_executor.layers.fused_moe.expert_map_manager import determine_expert_map  # :22-113
    from vllm.model_executor.layers.fused_moe.router import grouped_topk_router as gtr  # :80-166
    from vllm.model_executor.layers.quantization.utils.quant_utils import (  # :493-539
        pack_quantized_values_into_int32, unpack_quantized_values_into_int32)
    from vllm.scalar_type import scalar_types

    g = {}
    gen = torch.Generator().manual_seed(0)

    # ---- fused_topk torch reference -----------------------------------------------------------
    cases = []
    for (M, E, k, renorm, scoring, use_bias) in [
        (33, 8, 2, True, "softmax", False), (7, 128, 8, True, "softmax", False),
        (5, 64, 6, False, "sigmoid", False), (9, 256, 8, True, "sigmoid", True),
        (1, 192, 4, False, "softmax", True),
    ]:
        logits = torch.randn(M, E, generator=gen)
        bias = torch.randn(E, generator=gen) if use_bias else None
        w, ids = torch_topk(logits, k, renorm, bias, scoring)
        cases.append(dict(logits=logits, k=k, renorm=renorm, scoring=scoring, bias=bias,
                          weights=w.float(), ids=ids.int()))
    g["fused_topk"] = cases

    # ---- grouped_topk native (un-compiled python body) -----------------------------------------
    fn = gtr.grouped_topk
    fn = getattr(fn, "_torchdynamo_orig_callable", fn)
    os.environ["VLLM_USE_FUSED_MOE_GROUPED_TOPK"] = "0"
    cases = []
    for (M, E, ng, tg, k, renorm, scoring, use_bias, rsf) in [
        (13, 256, 8, 4, 8, True, "sigmoid", True, 2.5),
        (4, 64, 4, 2, 6, True, "softmax", False, 1.0),
        (1, 128, 8, 3, 4, False, "sigmoid", True, 1.0),
    ]:
        logits = torch.randn(M, E, generator=gen)
        bias = torch.randn(E, generator=gen) if use_bias else None
        try:
            w, ids = fn(torch.zeros(M, 8), logits, k, renorm, ng, tg, scoring, rsf, bias)
        except Exception as ex:  # platform probing may fail on CPU: fall back is recorded, not hidden
            print("grouped_topk native failed:", repr(ex))
            raise
        cases.append(dict(logits=logits, bias=bias, n_group=ng, topk_group=tg, k=k, renorm=renorm,
                          scoring=scoring, rsf=rsf, weights=w.float(), ids=ids.int()))
    g["grouped_topk_native"] = cases

    # ---- expert map ------------------------------------------------------------------------------
    g["expert_map"] = [
        dict(ep=ep, rank=r, E=E, local=determine_expert_map(ep, r, E)[0], emap=determine_expert_map(ep, r, E)[1])
        for (ep, r, E) in [(8, 0, 256), (8, 7, 256), (4, 1, 10), (4, 3, 10), (2, 1, 128)]
    ]

    # ---- per-token-group fp8 quant + block matmul -----------------------------------------------
    x = torch.randn(6, 512, generator=gen) / 10
    xq, xs = native_per_token_group_quant_fp8(x.bfloat16(), 128)
    g["ptg_quant"] = dict(x=x.bfloat16(), q=xq, s=xs)
    wq = (torch.randn(256, 512, generator=gen)).clamp(-448, 448).to(torch.float8_e4m3fn)
    ws = torch.rand(2, 4, generator=gen) * 0.01
    g["block_matmul"] = dict(xq=xq, xs=xs, wq=wq, ws=ws,
                             out=native_w8a8_block_matmul(xq, wq, xs, ws, [128, 128], torch.float32))

    # ---- torch_experts: bf16 and block-fp8 --------------------------------------------------------
    # torch_experts instantiates the SiluAndMul CustomOp, which needs a current vLLM config; on this
    # CPU-only container we patch the op registry entry with its own ``forward_native`` body so the
    # reference arithmetic (silu(x[:d]) * x[d:], vllm/model_executor/layers/activation.py:140-143)
    # still comes from the reference file.
    import tests.kernels.utils as tku
    from vllm.model_executor.layers.activation import SiluAndMul

    class _NativeSilu:
        def __call__(self, x):
            return SiluAndMul.forward_native(x)

    tku.op_registry = dict(tku.op_registry)
    tku.op_registry["silu_and_mul"] = _NativeSilu
    tku.SiluAndMul = _NativeSilu
    # The block-fp8 branch quantises activations through a Triton kernel (unavailable on CPU); route
    # it to the reference's own torch restatement of the same op (tests/kernels/quant_utils.py:157-180).
    _orig_q = tku.moe_kernel_quantize_input

    def _quant(A, A_scale, quant_dtype, per_act_token_quant, block_shape=None, **kw):
        if quant_dtype == torch.float8_e4m3fn and block_shape is not None:
            return native_per_token_group_quant_fp8(A.contiguous(), block_shape[1])
        return _orig_q(A, A_scale, quant_dtype, per_act_token_quant, block_shape, **kw)

    tku.moe_kernel_quantize_input = _quant
    M, H, I, E, k = 9, 256, 128, 6, 2
    a = (torch.randn(M, H, generator=gen) / 10).bfloat16()
    w1 = (torch.randn(E, 2 * I, H, generator=gen) / 10).bfloat16()
    w2 = (torch.randn(E, H, I, generator=gen) / 10).bfloat16()
    score = torch.randn(M, E, generator=gen)
    tw, ti = torch.topk(torch.softmax(score, -1), k)
    out = torch_experts(a, w1, w2, tw, ti)
    g["experts_bf16"] = dict(a=a, w1=w1, w2=w2, topk_weight=tw, topk_ids=ti.int(), out=out)

    w1q = w1.float().clamp(-448, 448).to(torch.float8_e4m3fn)
    w2q = w2.float().clamp(-448, 448).to(torch.float8_e4m3fn)
    w1s = torch.rand(E, 2 * I // 128, H // 128, generator=gen) * 0.02 + 0.005
    w2s = torch.rand(E, H // 128, I // 128, generator=gen) * 0.02 + 0.005
    out = torch_experts(a, w1q, w2q, tw, ti, w1_scale=w1s, w2_scale=w2s,
                        quant_dtype=torch.float8_e4m3fn, per_act_token_quant=False, block_shape=[128, 128])
    g["experts_fp8_block"] = dict(a=a, w1q=w1q, w2q=w2q, w1s=w1s, w2s=w2s, topk_weight=tw,
                                  topk_ids=ti.int(), out=out)

    # ---- mxfp4 dequant / int4 pack -----------------------------------------------------------------
    xp = torch.randint(0, 256, (4, 64), generator=gen, dtype=torch.uint8)
    sc = torch.randint(118, 132, (4, 4), generator=gen, dtype=torch.uint8)
    g["mxfp4_dequant"] = dict(packed=xp, scale=sc, out=mxfp4_dequantize(xp, sc))
    q = torch.randint(0, 16, (16, 8), generator=gen, dtype=torch.int32)
    packed = pack_quantized_values_into_int32(q, scalar_types.uint4b8, packed_dim=0)
    g["int4_pack"] = dict(q=q, packed=packed,
                          unpacked=unpack_quantized_values_into_int32(packed, scalar_types.uint4b8, 0))

    # ---- paged GQA decode ----------------------------------------------------------------------------
    B, Hq, Hkv, D, page, npages = 3, 8, 2, 128, 16, 24
    qq = torch.randn(B, Hq, D, generator=gen)
    kc = torch.randn(npages, page, Hkv, D, generator=gen)
    vc = torch.randn(npages, page, Hkv, D, generator=gen)
    kv_lens = [5, 37, 64]
    bt = torch.randperm(npages, generator=gen)[: B * 4].reshape(B, 4).int()
    scale = D ** -0.5
    o = ref_paged_attn(qq.clone(), kc, vc, [1] * B, kv_lens, bt, scale)
    g["gqa_decode"] = dict(q=qq, k_cache=kc, v_cache=vc, kv_lens=kv_lens, block_tables=bt, scale=scale, out=o)

    # ---- MXFP8 activation quantisation (native MXFP4 path: W4A8-MX) ---------------------------------------
    # vllm/model_executor/layers/quantization/utils/mxfp8_utils.py:38-86 (_mxfp8_e4m3_quantize_torch)
    from vllm.model_executor.layers.quantization.utils.mxfp8_utils import _mxfp8_e4m3_quantize_torch
    gen2 = torch.Generator().manual_seed(1234)
    xm = torch.randn(16, 256, generator=gen2) * torch.logspace(-3, 2, 16).unsqueeze(1)   # rows of very different magnitude
    xm[3, 32:64] = 0.0                                                                   # an all-zero block
    xm = xm.bfloat16()
    qm, sm = _mxfp8_e4m3_quantize_torch(xm)
    g["mxfp8_quant"] = dict(x=xm, q=qm, scales=sm)

    # ---- NVFP4 dequantisation ---------------------------------------------------------------------------------
    # tests/kernels/quantization/nvfp4_utils.py:16-62 (swizzled 128x4 scale layout -> linear, value = e2m1 * sf / global)
    from tests.kernels.quantization.nvfp4_utils import convert_swizzled_to_linear, dequantize_nvfp4_to_dtype
    gen3 = torch.Generator().manual_seed(4321)
    m, kk = 128, 64
    fp4 = torch.randint(0, 256, (m, kk // 2), generator=gen3, dtype=torch.uint8)
    sf_sw = torch.randint(0x28, 0x58, (m * kk // 16,), generator=gen3, dtype=torch.uint8)   # finite positive e4m3 bytes
    cases = []
    for gs in (64.0, 2688.0 / 3.7):
        out = dequantize_nvfp4_to_dtype(fp4, sf_sw, torch.tensor(gs), torch.float32, "cpu")
        cases.append(dict(global_scale=gs, out=out))
    sf_lin = convert_swizzled_to_linear(sf_sw.view(torch.float8_e4m3fn), m, kk, 16).contiguous()
    g["nvfp4_dequant"] = dict(packed=fp4, sf_linear=sf_lin.view(torch.uint8), cases=cases)

    # ---- INT4 (uint4b8, group 32) quantise / dequantise recipe ------------------------------------------------
    # vllm/model_executor/layers/quantization/utils/quant_utils.py:642-730 (quantize_weights): w [K, N], groups along K
    from vllm.model_executor.layers.quantization.utils.quant_utils import quantize_weights
    gen4 = torch.Generator().manual_seed(777)
    wk = torch.randn(64, 16, generator=gen4) / 10
    w_ref, w_q, w_s, _ = quantize_weights(wk, scalar_types.uint4b8, 32)
    g["int4_quantize_weights"] = dict(w=wk, w_ref=w_ref, w_q=w_q.int(), w_s=w_s)

    # ---- moe_permute torch reference -------------------------------------------------------------------------
    # tests/kernels/moe/test_moe_permute_unpermute.py:37-88 (torch_permute); the function hard-codes device="cuda" for
    # two helper tensors, so torch.zeros / torch.arange are wrapped to build them on the CPU while it runs
    from vllm.platforms import current_platform as _cp
    try:
        type(_cp).manual_seed_all = lambda self, seed: None   # module-level set_random_seed(0): no device here
    except Exception:
        pass
    from tests.kernels.moe import test_moe_permute_unpermute as tpu
    _zeros, _arange = torch.zeros, torch.arange

    def _cpu(fn):
        def w(*a, **k):
            if k.get("device") == "cuda":
                k["device"] = "cpu"
            return fn(*a, **k)
        return w

    torch.zeros, torch.arange = _cpu(_zeros), _cpu(_arange)
    try:
        gen5 = torch.Generator().manual_seed(99)
        cases = []
        for (M, E, k, ep, rank) in [(17, 16, 2, 1, 0), (33, 64, 6, 4, 1), (5, 256, 8, 16, 3)]:
            hs = torch.randn(M, 32, generator=gen5)
            ids = torch.stack([torch.randperm(E, generator=gen5)[:k] for _ in range(M)]).int()
            n_local = E // ep
            emap = None
            if ep > 1:
                emap = torch.full((E,), -1, dtype=torch.int32)
                emap[rank * n_local:(rank + 1) * n_local] = torch.arange(n_local, dtype=torch.int32)
            ph, first_off, src2dst, dst2src, valid = tpu.torch_permute(hs, ids.long() if emap is None else ids.long(), k, E,
                                                                      n_local, rank * n_local, emap)
            cases.append(dict(hidden=hs, topk_ids=ids, n_expert=E, n_local=n_local, start=rank * n_local,
                              expert_map=emap, permuted=ph, expert_first_token_offset=first_off,
                              src_row_id2dst_row_id_map=src2dst, dst_row_id2src_row_id_map=dst2src,
                              n_valid=len(valid)))
        g["moe_permute"] = cases
    finally:
        torch.zeros, torch.arange = _zeros, _arange

    # ---- SwiGLU-OAI (packed halves + clamp) ---------------------------------------------------------------------
    # vllm/model_executor/layers/activation.py:205-246 (SiluAndMulWithClamp.forward_native; alpha = sigmoid scale,
    # beta = 1.0 up bias) — called unbound on an attribute bag: constructing the CustomOp needs a device-typed config
    import types
    from vllm.model_executor.layers.activation import SiluAndMulWithClamp
    gen6 = torch.Generator().manual_seed(2024)
    xa = torch.randn(9, 64, generator=gen6) * 4
    cases = []
    for (limit, alpha) in [(7.0, 1.702), (3.0, 1.0)]:
        ns = types.SimpleNamespace(swiglu_limit=limit, alpha=alpha, beta=1.0)
        cases.append(dict(limit=limit, alpha=alpha, out=SiluAndMulWithClamp.forward_native(ns, xa)))
    g["swigluoai_packed"] = dict(x=xa, cases=cases)

    # ---- LVLLM_* hybrid-scheduler predicates (vllm/envs.py:2292-2410) ----------------------------------------
    import vllm.envs as renvs
    names = ["model.layers.0.mlp.experts", "model.layers.1.mlp.experts", "model.layers.2.mlp.experts",
             "model.layers.7.mlp.experts", "model.layers.33.mlp.experts", "model.layers.40.mlp.experts",
             "mtp.0.mlp.experts", "layers.5.block_sparse_moe.experts"]
    env_cases = [
        {},
        {"LVLLM_MOE_NUMA_ENABLED": "1"},
        {"LVLLM_MOE_NUMA_ENABLED": "1", "LVLLM_GPU_RESIDENT_MOE_LAYERS": "0-1, 33-34,x,7"},
        {"LVLLM_MOE_NUMA_ENABLED": "1", "LVLLM_GPU_RESIDENT_MOE_LAYERS": "5,40-38,2-2", "LVLLM_GPU_PREFILL_MIN_BATCH_SIZE": "512"},
        {"LVLLM_MOE_NUMA_ENABLED": "0", "LVLLM_GPU_PREFILL_MIN_BATCH_SIZE": "2048"},
        {"LVLLM_MOE_NUMA_ENABLED": "1", "LVLLM_GPU_PREFILL_MIN_BATCH_SIZE": "64", "LVLLM_GPU_PREFETCH_WINDOW": "2"},
    ]
    keys = ["LVLLM_MOE_NUMA_ENABLED", "LVLLM_GPU_RESIDENT_MOE_LAYERS", "LVLLM_GPU_PREFILL_MIN_BATCH_SIZE",
            "LVLLM_GPU_PREFETCH_WINDOW", "LVLLM_ENABLE_MOE_LAYERWISE_LOAD"]
    saved = {k: os.environ.get(k) for k in keys}
    table = []
    try:
        for env in env_cases:
            for k in keys:
                os.environ.pop(k, None)
            os.environ.update(env)
            row = dict(env=dict(env), feature=bool(renvs.is_lk_moe_feature_enabled()),
                       use_gpu_prefill=bool(renvs.is_lk_moe_use_gpu_prefill()),
                       min_batch=int(renvs.get_gpu_prefill_min_batch_size()),
                       window=int(renvs.get_gpu_prefetch_window()), layers=[])
            for nm in names:
                row["layers"].append(dict(name=nm, resident=bool(renvs.is_lk_moe_gpu_resident_layer(nm)),
                                          gpu_prefill=bool(renvs.is_lk_moe_gpu_prefill_layer(nm)),
                                          cpu=bool(renvs.is_lk_moe_cpu_layer(nm)), mtp=bool(renvs.is_lk_moe_mtp_layer(nm))))
            table.append(row)
    finally:
        for k, v in saved.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v
    g["lvllm_env_predicates"] = table

    # ---- RoutedExperts.global_to_local_expert_ids (routed_experts.py:1332-1342), called unbound -----------------
    import types as _types
    from vllm.model_executor.layers.fused_moe.routed_experts import RoutedExperts
    gen7 = torch.Generator().manual_seed(31337)
    emap7 = torch.full((64,), -1, dtype=torch.int32)
    emap7[16:32] = torch.arange(16, dtype=torch.int32)
    ids7 = torch.randint(-2, 70, (11, 6), generator=gen7, dtype=torch.int64)   # includes padding (<0) and ids >= E
    res7 = RoutedExperts.global_to_local_expert_ids(_types.SimpleNamespace(_expert_map=emap7), ids7.clone())
    g["global_to_local"] = dict(expert_map=emap7, topk_ids=ids7, out=res7)

    # ---- RoPE: the reference's RotaryEmbedding.forward_static (rotary_embedding/base.py:161-201 ->
    # common.py:146-185), DeepSeek MLA shapes (rope dim 64; q_pe [T,Hq,64], k_pe [T,1,64]), both styles
    from vllm.model_executor.layers.rotary_embedding.base import RotaryEmbedding
    T, Hq, rot, max_pos = 5, 16, 64, 512
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, rot, 2, dtype=torch.float) / rot))
    fr = torch.einsum("i,j->ij", torch.arange(max_pos, dtype=torch.float), inv_freq)
    cos_sin_cache = torch.cat((fr.cos(), fr.sin()), dim=-1).bfloat16()      # layout of _compute_cos_sin_cache
    pos = torch.randint(0, max_pos, (T,), generator=gen)
    qpe = torch.randn(T, Hq, rot, generator=gen).bfloat16()
    kpe = torch.randn(T, 1, rot, generator=gen).bfloat16()
    cases = []
    for neox in (False, True):
        qo, ko = RotaryEmbedding.forward_static(pos, qpe.clone(), kpe.clone(), rot, rot, cos_sin_cache, neox)
        cases.append(dict(neox=neox, q_out=qo, k_out=ko))
    g["rope"] = dict(positions=pos, q=qpe, k=kpe, cos_sin_cache=cos_sin_cache, cases=cases)

    # ---- FP8 block scales in ue8m0 form: the reference's DeepGEMM-on-Blackwell helpers (vllm/utils/deep_gemm.py:644-681,
    # fp8_utils.py:986-1043).  per_block_cast_to_fp8 is wrapped in torch.compile: the undecorated function is called.
    import vllm.utils.deep_gemm as dg
    from vllm.model_executor.layers.quantization.utils.fp8_utils import requant_weight_ue8m0_inplace
    pbc = getattr(dg.per_block_cast_to_fp8, "__wrapped__", dg.per_block_cast_to_fp8)
    gen8 = torch.Generator().manual_seed(808)
    x8 = torch.randn(256, 384, generator=gen8) * torch.logspace(-3, 1, 384)[None, :]
    q8, s8 = pbc(x8, [128, 128], True)
    wq8 = (torch.randn(3, 256, 256, generator=gen8) * 40).to(torch.float8_e4m3fn)
    ws8 = torch.rand(3, 2, 2, generator=gen8) * 3e-3 + 2e-4
    wq8b, ws8b = wq8.clone(), ws8.clone()
    _orig = dg.per_block_cast_to_fp8
    dg.per_block_cast_to_fp8 = pbc            # requant imports the name at call time
    try:
        requant_weight_ue8m0_inplace(wq8b, ws8b)
    finally:
        dg.per_block_cast_to_fp8 = _orig
    sc8 = torch.rand(64, generator=gen8) * torch.logspace(-12, 2, 64)
    g["fp8_ue8m0"] = dict(x=x8, q=q8, s=s8, w_q=wq8, w_s=ws8, w_q_out=wq8b, w_s_out=ws8b, sc=sc8, sc_ceil=dg._ceil_to_ue8m0(sc8))

    torch.save(g, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
