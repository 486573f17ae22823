"""torch-tensor front ends of the operator-level C ABI (routing, permutation, decode attention).

Names and argument meaning follow the reference's python op wrappers (vllm/_custom_ops.py ``topk_softmax``,
``grouped_topk``, ``moe_permute`` / ``moe_unpermute``, ``sm100_cutlass_mla_decode``) so that parity tests read
like the reference's own.  All tensors must live on the current CUDA device; there is no CPU path.
"""
from __future__ import annotations

import torch

from . import _lib as L

_DT = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}
_OUT = {torch.bfloat16: 0, torch.float16: 1, torch.float32: 2}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _cuda(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (lvllm_b200 has no CPU path)")
    return t.contiguous()


def fused_topk(gating_output: torch.Tensor, topk: int, renormalize: bool, scoring_func: str = "softmax",
               e_score_correction_bias: torch.Tensor | None = None, routed_scaling_factor: float = 1.0,
               return_token_expert_indices: bool = False):
    """reference fused_topk_router.py:81-124 -> _moe_C.topk_softmax / topk_sigmoid."""
    g = _cuda(gating_output, "gating_output")
    M, E = g.shape
    w = torch.empty(M, topk, dtype=torch.float32, device=g.device)
    ids = torch.empty(M, topk, dtype=torch.int32, device=g.device)
    tei = torch.empty(M, topk, dtype=torch.int32, device=g.device) if return_token_expert_indices else None
    bias = _cuda(e_score_correction_bias.float(), "bias") if e_score_correction_bias is not None else None
    rc = L.lib().b200_topk_gating(_stream(), g.data_ptr(), _DT[g.dtype], bias.data_ptr() if bias is not None else None,
                                  M, E, topk, {"softmax": 0, "sigmoid": 1}[scoring_func], int(renormalize),
                                  float(routed_scaling_factor), w.data_ptr(), ids.data_ptr(),
                                  tei.data_ptr() if tei is not None else None)
    L.check(rc, "b200_topk_gating")
    return (w, ids, tei) if return_token_expert_indices else (w, ids)


def grouped_topk(gating_output: torch.Tensor, topk: int, renormalize: bool, num_expert_group: int, topk_group: int,
                 scoring_func: str = "sigmoid", routed_scaling_factor: float = 1.0,
                 e_score_correction_bias: torch.Tensor | None = None):
    """reference grouped_topk_router.py:28-166 (fused kernel semantics: ordered output)."""
    g = _cuda(gating_output, "gating_output")
    if scoring_func == "softmax":  # reference applies the softmax in python first (:58-70)
        g = torch.softmax(g.float(), dim=-1)
        scoring = 0
    else:
        scoring = 1
    M, E = g.shape
    w = torch.empty(M, topk, dtype=torch.float32, device=g.device)
    ids = torch.empty(M, topk, dtype=torch.int32, device=g.device)
    bias = _cuda(e_score_correction_bias.float(), "bias") if e_score_correction_bias is not None else None
    rc = L.lib().b200_grouped_topk(_stream(), g.data_ptr(), _DT[g.dtype], bias.data_ptr() if bias is not None else None,
                                   M, E, num_expert_group, topk_group, topk, scoring, int(renormalize),
                                   float(routed_scaling_factor), w.data_ptr(), ids.data_ptr())
    L.check(rc, "b200_grouped_topk")
    return w, ids


_router_ws: dict = {}


def router_topk(hidden_states: torch.Tensor, gate_weight: torch.Tensor, topk: int, renormalize: bool,
                scoring_func: str = "softmax", e_score_correction_bias: torch.Tensor | None = None,
                routed_scaling_factor: float = 1.0, num_expert_group: int = 0, topk_group: int = 0,
                expert_map: torch.Tensor | None = None, return_logits: bool = False, n_shared: int = 0,
                shared_local_base: int = -1, shared_weight: float = 1.0):
    """Fused router: GateLinear.forward (reference router/gate_linear.py:171-221) + fused_topk / grouped_topk
    (fused_topk_router.py:81-124, grouped_topk_router.py:28-166) + global_to_local_expert_ids in ONE kernel.
    hidden [M,H], gate_weight [E,H] (same 16-bit dtype).  ``num_expert_group > 0`` selects DeepSeek grouped routing.
    ``n_shared`` > 0 appends always-on columns for shared experts held at local ids shared_local_base + s (weight
    ``shared_weight``; reference runner/shared_experts.py) — local_ids are then always returned.
    Returns (topk_weights f32 [M,k'], topk_ids i32 [M,k'] global, local_ids i32 [M,k'] | None[, logits f32 [M,E]]),
    k' = k + n_shared."""
    h, wg = _cuda(hidden_states, "hidden_states"), _cuda(gate_weight, "gate_weight")
    assert h.dtype == wg.dtype and h.dtype in (torch.bfloat16, torch.float16)
    M, H = h.shape
    E = wg.shape[0]
    dev = h.device
    lib = L.lib()
    need = int(lib.b200_router_workspace_bytes(M, E, H))
    key = (dev.index, torch.cuda.current_stream().cuda_stream)
    ws = _router_ws.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(max(need, 1 << 20), dtype=torch.uint8, device=dev)   # counters must start at zero
        _router_ws[key] = ws
    kk = topk + n_shared
    w = torch.empty(M, kk, dtype=torch.float32, device=dev)
    ids = torch.empty(M, kk, dtype=torch.int32, device=dev)
    loc = torch.empty(M, kk, dtype=torch.int32, device=dev) if (expert_map is not None or n_shared) else None
    lg = torch.empty(M, E, dtype=torch.float32, device=dev) if return_logits else None
    bias = _cuda(e_score_correction_bias.float(), "bias") if e_score_correction_bias is not None else None
    em = _cuda(expert_map.to(torch.int32), "expert_map") if expert_map is not None else None
    if num_expert_group > 0:
        assert scoring_func in ("sigmoid", "none")
        mode, scoring = 2, 1 if scoring_func == "sigmoid" else 0
    else:
        mode, scoring = {"softmax": 0, "sigmoid": 1}[scoring_func], 0
    rc = lib.b200_router_topk(_stream(), h.data_ptr(), 1 if h.dtype == torch.float16 else 0, wg.data_ptr(), M, E, H,
                              bias.data_ptr() if bias is not None else None, mode, scoring, topk, int(renormalize),
                              int(num_expert_group), int(topk_group), float(routed_scaling_factor),
                              em.data_ptr() if em is not None else None, int(n_shared), int(shared_local_base),
                              float(shared_weight), ws.data_ptr(), ws.numel(), w.data_ptr(),
                              ids.data_ptr(), loc.data_ptr() if loc is not None else None,
                              lg.data_ptr() if lg is not None else None)
    L.check(rc, "b200_router_topk")
    return (w, ids, loc, lg) if return_logits else (w, ids, loc)


def global_to_local_expert_ids(topk_ids: torch.Tensor, expert_map: torch.Tensor) -> torch.Tensor:
    """reference routed_experts.py:1332-1342."""
    ids = _cuda(topk_ids.to(torch.int32), "topk_ids")
    em = _cuda(expert_map.to(torch.int32), "expert_map")
    out = torch.empty_like(ids)
    L.check(L.lib().b200_global_to_local_ids(_stream(), ids.data_ptr(), em.data_ptr(), em.numel(), ids.numel(),
                                             out.data_ptr()), "b200_global_to_local_ids")
    return out


def moe_permute(hidden_states: torch.Tensor | None, topk_ids: torch.Tensor, n_local_expert: int):
    """Stable sort by expert (reference moe_permute_unpermute.py:105-231).  Returns
    (permuted_hidden | None, sorted_slot i32 [M*k], expert_first_token_offset i64 [E+1], inv_permuted_idx i32 [M,k])."""
    ids = _cuda(topk_ids.to(torch.int32), "topk_ids")
    M, k = ids.shape
    dev = ids.device
    sorted_slot = torch.empty(M * k, dtype=torch.int32, device=dev)
    first = torch.empty(n_local_expert + 1, dtype=torch.int64, device=dev)
    inv = torch.empty(M, k, dtype=torch.int32, device=dev)
    perm = None
    hp, H = None, 0
    if hidden_states is not None:
        h = _cuda(hidden_states, "hidden_states")
        assert h.dtype in (torch.bfloat16, torch.float16)
        H = h.shape[1]
        perm = torch.zeros(M * k, H, dtype=h.dtype, device=dev)
        hp = h.data_ptr()
    rc = L.lib().b200_moe_permute(_stream(), hp, ids.data_ptr(), M, k, n_local_expert, H, sorted_slot.data_ptr(),
                                  first.data_ptr(), inv.data_ptr(), perm.data_ptr() if perm is not None else None)
    L.check(rc, "b200_moe_permute")
    return perm, sorted_slot, first, inv


def moe_unpermute(permuted: torch.Tensor, topk_weights: torch.Tensor, inv_permuted_idx: torch.Tensor,
                  out_dtype: torch.dtype | None = None) -> torch.Tensor:
    """out[t] = sum_j w[t,j] * permuted[inv[t,j]] in fp32 (reference moe_permute_unpermute.py:234-277)."""
    p = _cuda(permuted, "permuted")
    w = _cuda(topk_weights.float(), "topk_weights")
    inv = _cuda(inv_permuted_idx.to(torch.int32), "inv_permuted_idx")
    M, k = inv.shape
    H = p.shape[1]
    od = out_dtype or p.dtype
    out = torch.empty(M, H, dtype=od, device=p.device)
    rc = L.lib().b200_moe_unpermute(_stream(), p.data_ptr(), 1 if p.dtype == torch.float16 else 0, w.data_ptr(),
                                    inv.data_ptr(), M, k, H, out.data_ptr(), _OUT[od])
    L.check(rc, "b200_moe_unpermute")
    return out


def rmsnorm_cast(x_f32: torch.Tensor, out: torch.Tensor, gain: float = 1.0, eps: float = 1e-6) -> torch.Tensor:
    """out = cast(x * gain * rsqrt(mean(x^2)+eps)) — the cast Lvllm applies to lk_moe's fp32 output
    (reference routed_experts.py:1855) fused with the RMS normalisation that follows in the layer."""
    x = _cuda(x_f32, "x")
    assert x.dtype == torch.float32 and out.is_cuda and out.is_contiguous()
    M, H = x.shape
    L.check(L.lib().b200_rmsnorm_cast(_stream(), x.data_ptr(), out.data_ptr(), M, H, float(gain), float(eps),
                                      1 if out.dtype == torch.float16 else 0), "b200_rmsnorm_cast")
    return out


def mla_decode(q_nope: torch.Tensor, q_pe: torch.Tensor, kv_c_and_k_pe_cache: torch.Tensor, seq_lens: torch.Tensor,
               page_table: torch.Tensor, sm_scale: float, num_kv_splits: int = 0, max_seq_len: int | None = None,
               q_scale: float = 1.0, k_scale: float = 1.0, w_uv: torch.Tensor | None = None):
    """Paged MLA decode; mirrors ops.sm100_cutlass_mla_decode (reference vllm/_custom_ops.py:3212,
    backends/mla/cutlass_mla.py:176-257).  bf16 q + bf16 cache, or an e4m3 cache (``torch.float8_e4m3fn``, 576 B / token)
    with bf16 or e4m3 queries (the reference's fp8 mode; q_scale / k_scale = per-tensor dequantisation scales).
    ``max_seq_len`` (host-side bound on seq_lens) sizes the split count / workspace instead of the page-table width.
    With ``w_uv`` [Hq,512,128] the split merge is fused with the v up-projection and (out_v [B,Hq,128], out, lse) is returned.
    Returns (out bf16 [B,Hq,512], lse f32 [B,Hq])."""
    qn, qp = _cuda(q_nope, "q_nope"), _cuda(q_pe, "q_pe")
    kv = _cuda(kv_c_and_k_pe_cache, "kv_cache")
    f8 = torch.float8_e4m3fn
    assert kv.dtype in (torch.bfloat16, f8) and kv.shape[-1] == 576
    assert qn.dtype == qp.dtype and qn.dtype in (torch.bfloat16, f8) and (qn.dtype != f8 or kv.dtype == f8)
    B, Hq, _ = qn.shape
    page = kv.shape[1]
    sl = _cuda(seq_lens.to(torch.int32), "seq_lens")
    pt = _cuda(page_table.to(torch.int32), "page_table")
    max_pages = pt.shape[1]
    if max_seq_len is not None:
        # the kernel only walks pages below the bound: a narrower view of the table keeps the split count (and the
        # B*Hq*splits*514*4-byte workspace) proportional to the real context, not to max_model_len
        max_pages = min(max_pages, -(-int(max_seq_len) // page))
        if max_pages != pt.shape[1]:
            pt = pt[:, :max_pages].contiguous()
    if num_kv_splits <= 0:
        # a CTA = (request, split, value-dim half) walks ceil(tiles / splits) 128-token tiles with an online softmax:
        # enough splits to fill ~2 waves of the 148 SMs, no more (each split costs an fp32 partial of the output)
        tiles = max(1, -(-(max_pages * page) // 128))
        num_kv_splits = max(1, min(tiles, 296 // (2 * B)))
    ws = torch.empty(L.lib().b200_mla_decode_workspace_bytes(B, Hq, num_kv_splits), dtype=torch.uint8, device=qn.device)
    out = torch.empty(B, Hq, 512, dtype=torch.bfloat16, device=qn.device)
    lse = torch.empty(B, Hq, dtype=torch.float32, device=qn.device)
    if w_uv is not None:
        # split merge fused with the v up-projection (reference _v_up_proj, mla_attention.py:1154-1176)
        wv = _cuda(w_uv, "w_uv")
        assert wv.dtype == torch.bfloat16 and wv.shape == (Hq, 512, 128)
        out_v = torch.empty(B, Hq, 128, dtype=torch.bfloat16, device=qn.device)
        rc = L.lib().b200_mla_decode_vup(_stream(), qn.data_ptr(), qp.data_ptr(), int(qn.dtype == f8), kv.data_ptr(),
                                         int(kv.dtype == f8), float(q_scale), float(k_scale), sl.data_ptr(), pt.data_ptr(),
                                         B, Hq, page, max_pages, float(sm_scale), num_kv_splits, ws.data_ptr(),
                                         wv.data_ptr(), out_v.data_ptr(), out.data_ptr(), lse.data_ptr())
        L.check(rc, "b200_mla_decode_vup")
        return out_v, out, lse
    rc = L.lib().b200_mla_decode_ex(_stream(), qn.data_ptr(), qp.data_ptr(), int(qn.dtype == f8), kv.data_ptr(),
                                    int(kv.dtype == f8), float(q_scale), float(k_scale), sl.data_ptr(), pt.data_ptr(),
                                    B, Hq, page, max_pages, float(sm_scale), num_kv_splits, ws.data_ptr(),
                                    out.data_ptr(), lse.data_ptr())
    L.check(rc, "b200_mla_decode")
    return out, lse


def mla_rope_cache_write(q_pe: torch.Tensor, k_pe: torch.Tensor, kv_c: torch.Tensor, positions: torch.Tensor,
                         cos_sin_cache: torch.Tensor, slot_mapping: torch.Tensor, kv_cache: torch.Tensor,
                         is_neox_style: bool = False, kv_scale: float = 1.0) -> None:
    """RoPE of q_pe [T,Hq,64] / k_pe [T,64] IN PLACE (reference RotaryEmbedding.forward_static) fused with
    concat_and_cache_mla: kv_cache [blocks, block_size, 576] (bf16 or e4m3) row slot_mapping[t] <- [kv_c[t] | rope(k_pe[t])]."""
    for t, n in ((q_pe, "q_pe"), (k_pe, "k_pe"), (kv_c, "kv_c"), (cos_sin_cache, "cos_sin_cache")):
        assert t.is_cuda and t.is_contiguous() and t.dtype == torch.bfloat16, n
    assert kv_cache.is_cuda and kv_cache.is_contiguous() and kv_cache.shape[-1] == 576 and cos_sin_cache.shape[-1] == 64
    T, Hq, _ = q_pe.shape
    pos = _cuda(positions.to(torch.int64), "positions")
    sm = _cuda(slot_mapping.to(torch.int64), "slot_mapping")
    rc = L.lib().b200_mla_rope_cache_write(_stream(), q_pe.data_ptr(), k_pe.data_ptr(), kv_c.data_ptr(), pos.data_ptr(),
                                           cos_sin_cache.data_ptr(), int(is_neox_style), sm.data_ptr(), kv_cache.data_ptr(),
                                           int(kv_cache.dtype == torch.float8_e4m3fn), float(kv_scale), T, Hq)
    L.check(rc, "b200_mla_rope_cache_write")


def mla_q_absorb(q_nope: torch.Tensor, w_uk_t: torch.Tensor) -> torch.Tensor:
    """ql_nope [T,Hq,512] = q_nope [T,Hq,128] x W_UK_T [Hq,128,512] (reference mla_attention.py:875-893)."""
    q, w = _cuda(q_nope, "q_nope"), _cuda(w_uk_t, "w_uk_t")
    T, Hq, P = q.shape
    assert q.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and P == 128 and w.shape == (Hq, 128, 512)
    out = torch.empty(T, Hq, 512, dtype=torch.bfloat16, device=q.device)
    L.check(L.lib().b200_mla_q_absorb(_stream(), q.data_ptr(), w.data_ptr(), out.data_ptr(), T, Hq), "b200_mla_q_absorb")
    return out


def gqa_decode(q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, seq_lens: torch.Tensor,
               page_table: torch.Tensor, sm_scale: float, num_kv_splits: int = 0, max_seq_len: int | None = None):
    """Paged GQA decode (q [B,Hq,128] bf16; caches [pages,page,Hkv,128] bf16).  Returns (out, lse).
    ``max_seq_len`` (host-side bound on seq_lens, as the reference's attention metadata carries it) sizes the split count
    and the B*Hq*splits*130*4-byte workspace from the real context instead of the page-table width (= max_model_len)."""
    qq, kc, vc = _cuda(q, "q"), _cuda(k_cache, "k_cache"), _cuda(v_cache, "v_cache")
    assert qq.dtype == torch.bfloat16 and kc.dtype == torch.bfloat16
    B, Hq, D = qq.shape
    page, Hkv = kc.shape[1], kc.shape[2]
    sl = _cuda(seq_lens.to(torch.int32), "seq_lens")
    pt = _cuda(page_table.to(torch.int32), "page_table")
    if max_seq_len is not None:
        max_pages = max(1, min(pt.shape[1], -(-int(max_seq_len) // page)))
        if max_pages != pt.shape[1]:
            pt = pt[:, :max_pages].contiguous()   # the kernels only walk pages below the bound
    if num_kv_splits <= 0:
        # tensor-core path: one 128-token split per work item, enough splits to cover the page table
        num_kv_splits = max(1, -(-(pt.shape[1] * page) // 128))
        if num_kv_splits > 1024:
            num_kv_splits = max(1, min(32, 592 // max(1, B * Hkv)))
    ws = torch.empty(L.lib().b200_gqa_decode_workspace_bytes(B, Hq, D, num_kv_splits), dtype=torch.uint8, device=qq.device)
    out = torch.empty(B, Hq, D, dtype=torch.bfloat16, device=qq.device)
    lse = torch.empty(B, Hq, dtype=torch.float32, device=qq.device)
    rc = L.lib().b200_gqa_decode(_stream(), qq.data_ptr(), kc.data_ptr(), vc.data_ptr(), sl.data_ptr(), pt.data_ptr(),
                                 B, Hq, Hkv, D, page, pt.shape[1], float(sm_scale), num_kv_splits, ws.data_ptr(),
                                 out.data_ptr(), lse.data_ptr())
    L.check(rc, "b200_gqa_decode")
    return out, lse
