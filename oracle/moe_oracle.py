"""CPU oracle for the Lvllm MoE / decode-attention hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  The product path (``lvllm_b200`` / ``lk_moe``) never does and fails
loudly when the CUDA library is missing.

Pinning.  The arithmetic being replaced lives in the closed third-party wheel ``lk_moe==2.3.3`` (reference
``requirements/cuda.txt:37``); its source is not under /root/reference and the reference holds no test or
golden vector for it: at THAT boundary parity stays unpinned.  This file restates the
*upstream-vLLM* semantics the lk_moe call site is embedded in, function by function, each citing the
reference file:line it follows, and is pinned two ways:
  * against the reference tree's own COMPILED CPU fused MoE (``csrc/cpu/cpu_fused_moe.cpp``, built from where it
    lies by ``oracle/build_ref.py`` into ``oracle/_ref/libref_moe.so``; ``tests/test_c_port.py``) for the bf16
    expert forward, on both of its ISA paths, and its COMPILED CPU paged MLA decode (``csrc/cpu/mla_decode.cpp``)
    for ``mla_decode``;
  * against outputs of the reference's own pure-torch references and helpers generated in the build container
    (``tests/golden/make_golden.py`` -> ``tests/golden/golden_ref.pt``; ``tests/test_oracle_golden.py``) for
    routing, permutation, every quantised format, activations, GQA attention and the LVLLM_* predicates.

Everything is float32 torch-on-CPU / numpy; no CUDA, no vLLM import.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

F32 = torch.float32
FP8 = torch.float8_e4m3fn
FP8_MAX = 448.0

# --------------------------------------------------------------------------------------------
# Routing
# --------------------------------------------------------------------------------------------


def topk_gating(logits: torch.Tensor, k: int, renormalize: bool, scoring: str = "softmax",
                bias: torch.Tensor | None = None, routed_scaling_factor: float = 1.0):
    """softmax / sigmoid top-k with lower-index tie-break.

    Follows ``topkGating`` in reference csrc/libtorch_stable/moe/topk_softmax_kernels.cu:408-592:
    fp32 scores (softmax uses expf :429, sigmoid 1/(1+exp(-x)) :456), NaN/Inf -> 0 (:466-471),
    ``bias`` is added for *selection only* (:476-493), k rounds of arg-max where the lower expert index
    wins ties (:515-543), the winner is knocked out with -10000 (:576), renormalise with
    ``denom = sum>0 ? sum : 1`` and scale by routed_scaling_factor (:582-592).
    Returns (weights f32 [M,k], ids i32 [M,k]) in descending selection order.
    """
    x = logits.detach().to(F32).cpu().numpy().astype(np.float32)
    M, E = x.shape
    if scoring == "softmax":
        m = x.max(axis=1, keepdims=True)
        e = np.exp((x - m).astype(np.float32)).astype(np.float32)
        s = e.sum(axis=1, keepdims=True, dtype=np.float32)
        p = (e * (np.float32(1.0) / s)).astype(np.float32)
    elif scoring == "sigmoid":
        p = (np.float32(1.0) / (np.float32(1.0) + np.exp(-x).astype(np.float32))).astype(np.float32)
    else:
        raise ValueError(scoring)
    p = np.where(np.isfinite(p), p, np.float32(0.0)).astype(np.float32)
    choice = p.copy()
    if bias is not None:
        choice = (choice + bias.detach().to(F32).cpu().numpy().astype(np.float32)[None, :]).astype(np.float32)
    w = np.zeros((M, k), np.float32)
    ids = np.zeros((M, k), np.int32)
    rows = np.arange(M)
    for j in range(k):
        idx = choice.argmax(axis=1)  # first occurrence == lower index wins ties
        ids[:, j] = idx
        w[:, j] = p[rows, idx]
        choice[rows, idx] = np.float32(-10000.0)
    scale = np.full((M, 1), np.float32(routed_scaling_factor), np.float32)
    if renormalize:
        ssum = np.zeros((M,), np.float32)
        for j in range(k):  # sequential fp32 accumulation like the kernel (:573)
            ssum = (ssum + w[:, j]).astype(np.float32)
        denom = np.where(ssum > 0, ssum, np.float32(1.0)).astype(np.float32)
        scale = (scale[:, 0] / denom).astype(np.float32)[:, None]
    w = (w * scale).astype(np.float32)
    return torch.from_numpy(w), torch.from_numpy(ids)


def sigmoid_tanh_form(x: np.ndarray) -> np.ndarray:
    """reference grouped_topk_kernels.cu:454-456 (sigmoid_accurate)."""
    return (np.float32(0.5) * np.tanh(np.float32(0.5) * x).astype(np.float32) + np.float32(0.5)).astype(np.float32)


def grouped_topk(logits: torch.Tensor, bias: torch.Tensor | None, n_group: int, topk_group: int, k: int,
                 renormalize: bool, routed_scaling_factor: float = 1.0, scoring: str = "sigmoid"):
    """DeepSeek group-limited (no-aux) routing, ordered-output semantics of the fused kernel.

    Follows reference csrc/libtorch_stable/moe/grouped_topk_kernels.cu:477-521 (group score = sum of the
    two largest biased scores of the group), :590-618 (stable selection of ``topk_group`` groups, lower
    group id wins ties; a -inf k-th group makes the row degenerate: ids 0..k-1, weights 1/k), :620-649
    (stable top-k over the finite candidates of the selected groups, lower expert id wins ties),
    :651-675 (weights from the UNBIASED scores, ``scale = rsf / (sum + 1e-20)``).
    With ``bias is None`` the group score is the group max (torch-native path, reference
    vllm/model_executor/layers/fused_moe/router/grouped_topk_router.py:128-131) and selection uses the
    unbiased scores.  ``scoring='softmax'`` applies a row softmax first (:58-70 of the same file).
    Returns (weights f32 [M,k], ids i32 [M,k]) in descending (biased score, then lower id) order.
    """
    x = logits.detach().to(F32).cpu().numpy().astype(np.float32)
    M, E = x.shape
    epg = E // n_group
    if scoring == "sigmoid":
        sc = sigmoid_tanh_form(x)
    elif scoring == "softmax":
        m = x.max(axis=1, keepdims=True)
        e = np.exp((x - m).astype(np.float32)).astype(np.float32)
        sc = (e / e.sum(axis=1, keepdims=True, dtype=np.float32)).astype(np.float32)
    elif scoring == "none":
        sc = x.copy()
    else:
        raise ValueError(scoring)
    has_bias = bias is not None
    b = bias.detach().to(F32).cpu().numpy().astype(np.float32) if has_bias else np.zeros((E,), np.float32)
    biased = (sc + b[None, :]).astype(np.float32)
    w = np.zeros((M, k), np.float32)
    ids = np.zeros((M, k), np.int32)
    NEG = np.float32(-np.inf)
    for t in range(M):
        g = biased[t].reshape(n_group, epg)
        if has_bias:
            srt = np.sort(g, axis=1)[:, ::-1]
            gscore = (srt[:, 0] + srt[:, 1]).astype(np.float32) if epg > 1 else (srt[:, 0] * 2).astype(np.float32)
        else:
            gscore = g.max(axis=1)
        gscore = np.where(np.isnan(gscore), NEG, gscore)
        order = np.lexsort((np.arange(n_group), -gscore.astype(np.float64)))  # desc score, asc id
        sel = order[:topk_group]
        if not (gscore[sel[-1]] > NEG):
            ids[t] = np.arange(k, dtype=np.int32)
            w[t] = np.float32(1.0) / np.float32(k)
            continue
        cand = np.full((E,), NEG, np.float32)
        for gid in sel:
            lo = gid * epg
            fin = np.isfinite(x[t, lo:lo + epg])
            cand[lo:lo + epg] = np.where(fin, biased[t, lo:lo + epg], NEG)
        order = np.lexsort((np.arange(E), -cand.astype(np.float64)))
        top = order[:k]
        ids[t] = top.astype(np.int32)
        unb = sc[t, top].astype(np.float32)
        scale = np.float32(routed_scaling_factor)
        if renormalize:
            ssum = np.float32(1e-20)
            for v in unb:
                ssum = np.float32(ssum + v)
            scale = np.float32(scale / ssum)
        w[t] = (unb * scale).astype(np.float32)
    return torch.from_numpy(w), torch.from_numpy(ids)


def global_to_local_expert_ids(topk_ids: torch.Tensor, expert_map: torch.Tensor) -> torch.Tensor:
    """reference vllm/model_executor/layers/fused_moe/routed_experts.py:1332-1342."""
    clamped = torch.clamp(topk_ids.long(), 0, expert_map.numel() - 1)
    out = expert_map[clamped].to(torch.int32)
    out[topk_ids < 0] = -1
    return out


def determine_expert_map(ep_size: int, ep_rank: int, global_num_experts: int):
    """Linear expert placement.  reference .../fused_moe/expert_map_manager.py:65-90: the first
    ``E % ep`` ranks get one extra expert; non-local entries are -1."""
    base, rem = divmod(global_num_experts, ep_size)
    local = base + (1 if ep_rank < rem else 0)
    start = ep_rank * base + min(ep_rank, rem)
    emap = torch.full((global_num_experts,), -1, dtype=torch.int32)
    emap[start:start + local] = torch.arange(local, dtype=torch.int32)
    return local, emap


# --------------------------------------------------------------------------------------------
# Permutation (stable sort by expert)
# --------------------------------------------------------------------------------------------


def moe_permute(topk_ids: torch.Tensor, num_local_experts: int):
    """Stable sort of the flattened (token,k) slots by local expert id.

    Follows the torch reference of reference tests/kernels/moe/test_moe_permute_unpermute.py:37-88
    (``torch.sort(stable=True)`` on ``topk_ids.flatten()``); slots with id < 0 or >= E are invalid and
    sort to the end (reference moe_permute_unpermute_kernel.cu:135-162 maps them to E).
    Returns:
      sorted_slot        i32 [M*k]  source slot (t*k+j) of each permuted row, valid rows first
      expert_first_off   i64 [E+1]  first permuted row of each expert
      inv_perm           i32 [M*k]  permuted row of slot (t*k+j), -1 for invalid slots
    """
    flat = topk_ids.reshape(-1).to(torch.int64)
    E = num_local_experts
    key = torch.where((flat < 0) | (flat >= E), torch.full_like(flat, E), flat)
    _, order = torch.sort(key, stable=True)
    counts = torch.bincount(key, minlength=E + 1)[:E]
    off = torch.zeros(E + 1, dtype=torch.int64)
    off[1:] = torch.cumsum(counts, 0)
    n_valid = int(off[-1])
    inv = torch.full((flat.numel(),), -1, dtype=torch.int32)
    inv[order[:n_valid]] = torch.arange(n_valid, dtype=torch.int32)
    return order.to(torch.int32), off, inv


# --------------------------------------------------------------------------------------------
# Weight formats: synthetic quantisers (recipes of the reference's test factories) + dequantisers
# --------------------------------------------------------------------------------------------

E2M1_VALUES = torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0], dtype=F32)


def quant_fp8_block(w: torch.Tensor, block=(128, 128)):
    """absmax/448 per [block_n, block_k] tile.  Recipe of reference tests/kernels/moe/utils.py
    (per_block_cast_to_fp8 style) / vllm fp8.py block quant: w [..., N, K] -> (e4m3, f32 scales)."""
    *lead, N, K = w.shape
    bn, bk = block
    nb, kb = -(-N // bn), -(-K // bk)
    wp = torch.zeros(*lead, nb * bn, kb * bk, dtype=F32)
    wp[..., :N, :K] = w.to(F32)
    t = wp.reshape(*lead, nb, bn, kb, bk)
    amax = t.abs().amax(dim=(-3, -1), keepdim=True).clamp(min=1e-4)
    scale = amax / FP8_MAX
    q = (t / scale).clamp(-FP8_MAX, FP8_MAX).to(FP8)
    q = q.reshape(*lead, nb * bn, kb * bk)[..., :N, :K].contiguous()
    return q, scale.reshape(*lead, nb, kb).contiguous().to(F32)


def dequant_fp8_block(q: torch.Tensor, scale: torch.Tensor, block=(128, 128)) -> torch.Tensor:
    """w = e4m3 * scale[n/128, k/128]  (reference tests/kernels/quant_utils.py:91-154 applies the same
    per-tile scale to partial products).  Per-tensor scales ([E] / [E,2]) are handled by the caller."""
    *lead, N, K = q.shape
    bn, bk = block
    s = scale.repeat_interleave(bn, dim=-2)[..., :N, :].repeat_interleave(bk, dim=-1)[..., :K]
    return q.to(F32) * s


def ceil_to_ue8m0(x: torch.Tensor) -> torch.Tensor:
    """reference vllm/utils/deep_gemm.py:644-645 (_ceil_to_ue8m0): 2^ceil(log2(|x|))."""
    return torch.pow(2.0, torch.ceil(torch.log2(x.abs())))


def per_token_group_quant_fp8(x: torch.Tensor, group: int = 128, eps: float = 1e-10, ue8m0: bool = False):
    """reference tests/kernels/quant_utils.py:157-180 (native_per_token_group_quant_fp8); ``ue8m0``: the scale is rounded
    up to a power of two, reference fp8_utils.py:100-113 (``y_s = exp2(ceil(log2(scale_raw))) if use_ue8m0``)."""
    shp = x.shape
    x_ = x.to(F32).reshape(-1, group)
    amax = x_.abs().amax(dim=-1, keepdim=True).clamp(min=eps)
    s = amax / FP8_MAX
    if ue8m0:
        s = ceil_to_ue8m0(s)
    q = (x_ / s).clamp(-FP8_MAX, FP8_MAX).to(FP8)
    return q.reshape(shp), s.reshape(*shp[:-1], shp[-1] // group)


def per_block_cast_to_fp8(x: torch.Tensor, block=(128, 128), ue8m0: bool = False):
    """reference vllm/utils/deep_gemm.py:662-681: [m, n] fp32 -> (fp8 [m, n], scales [m/bm, n/bn]); amax clamped at 1e-4,
    sf = amax / 448 (rounded up to a power of two with ``ue8m0``), x * (1 / sf) cast to e4m3."""
    m, n = x.shape
    bm, bn = block
    mp, np_ = -(-m // bm) * bm, -(-n // bn) * bn
    xp = torch.zeros(mp, np_, dtype=x.dtype)
    xp[:m, :n] = x
    v = xp.view(-1, bm, np_ // bn, bn)
    amax = v.abs().float().amax(dim=(1, 3), keepdim=True).clamp(1e-4)
    sf = amax / FP8_MAX
    if ue8m0:
        sf = ceil_to_ue8m0(sf)
    q = (v * (1.0 / sf)).to(FP8)
    return q.view_as(xp)[:m, :n].contiguous(), sf.view(v.size(0), v.size(2))


def requant_weight_ue8m0(w_q: torch.Tensor, w_s: torch.Tensor, block=(128, 128)):
    """reference fp8_utils.py:986-1043 (requant_weight_ue8m0_inplace), out of place: de-quantise with the fp32 block scales,
    re-quantise with power-of-two scales.  w_q fp8 [..., N, K], w_s f32 [..., N/bm, K/bk] -> (fp8, f32 power-of-two scales)."""
    lead = w_q.shape[:-2]
    wq = w_q.reshape(-1, *w_q.shape[-2:])
    ws = w_s.reshape(-1, *w_s.shape[-2:])
    oq, os_ = torch.empty_like(wq), torch.empty_like(ws, dtype=F32)
    for i in range(wq.shape[0]):
        n, k = wq[i].shape
        s_exp = ws[i].to(F32).repeat_interleave(block[0], dim=0).repeat_interleave(block[1], dim=1)[:n, :k]
        oq[i], os_[i] = per_block_cast_to_fp8(wq[i].to(F32) * s_exp, block, ue8m0=True)
    return oq.reshape(*lead, *w_q.shape[-2:]), os_.reshape(*lead, *w_s.shape[-2:])


def quant_int4_group(w: torch.Tensor, group: int = 32, scale_dtype=torch.bfloat16):
    """Symmetric uint4b8 group quantisation along K (compressed-tensors WNA16; recipe of reference
    vllm/model_executor/layers/quantization/utils/quant_utils.py ``quantize_weights``: s = absmax/7,
    q = clamp(round(w/s), -8, 7) + 8).  w [..., N, K] -> (uint8 [..., N, K/2] low nibble = even k
    (reference pack order quant_utils.py:493-512), scales [..., N, K/group])."""
    *lead, N, K = w.shape
    t = w.to(F32).reshape(*lead, N, K // group, group)
    s = (t.abs().amax(dim=-1, keepdim=True) / 7.0).clamp(min=1e-5).to(scale_dtype)
    q = torch.round(t / s.to(F32)).clamp(-8, 7).to(torch.int32) + 8
    q = q.reshape(*lead, N, K)
    packed = (q[..., 0::2] | (q[..., 1::2] << 4)).to(torch.uint8)
    return packed.contiguous(), s.reshape(*lead, N, K // group).contiguous()


def dequant_int4_group(packed: torch.Tensor, scales: torch.Tensor, group: int, out_dtype=torch.bfloat16):
    """w = (q - 8) * s rounded to the activation dtype (reference quant_utils.py:929-949 subtracts the
    uint4b8 bias; reference ``quantize_weights`` returns the dequantised reference in the weight dtype)."""
    lo = (packed & 0xF).to(torch.int32)
    hi = (packed >> 4).to(torch.int32)
    q = torch.stack([lo, hi], dim=-1).reshape(*packed.shape[:-1], packed.shape[-1] * 2)
    s = scales.to(F32).repeat_interleave(group, dim=-1)
    return ((q - 8).to(F32) * s).to(out_dtype).to(F32)


def _e2m1_quant_index(v: torch.Tensor) -> torch.Tensor:
    """round-to-nearest-even onto the e2m1 grid; returns 4-bit codes (sign in bit 3)."""
    a = v.abs().clamp(max=6.0)
    # thresholds between grid points with ties-to-even on the mantissa bit
    bounds = torch.tensor([0.25, 0.75, 1.25, 1.75, 2.5, 3.5, 5.0], dtype=F32)
    idx = torch.bucketize(a, bounds, right=False)  # a == bound goes to the lower bucket
    # ties to even: at 0.25->0 (idx0), 0.75->1.0 (idx2), 1.25->1.0(idx2), 1.75->2.0 (idx4),
    # 2.5->2.0 (idx4), 3.5->4.0 (idx6), 5.0->4.0 (idx6)
    tie_up = torch.tensor([0, 1, 0, 1, 0, 1, 0], dtype=torch.bool)
    for i, bnd in enumerate(bounds.tolist()):
        if tie_up[i]:
            idx = torch.where(a == bnd, torch.full_like(idx, i + 1), idx)
    code = idx.to(torch.int32) | ((v < 0).to(torch.int32) << 3)
    return code


def _pack_nibbles(code: torch.Tensor) -> torch.Tensor:
    return (code[..., 0::2] | (code[..., 1::2] << 4)).to(torch.uint8).contiguous()


_E2M1_BYTE_LUT = None


def unpack_e2m1(packed: torch.Tensor) -> torch.Tensor:
    """reference tests/kernels/quantization/nvfp4_utils.py:64-88 (break_fp4_bytes): low nibble first.
    Implemented as one gather from a 256-entry (low value, high value) table: same result, model-sized layers
    dequantise in fractions of a second."""
    global _E2M1_BYTE_LUT
    if _E2M1_BYTE_LUT is None:
        c = torch.arange(256, dtype=torch.int64)
        both = torch.stack([c & 0xF, c >> 4], dim=-1)
        mag = E2M1_VALUES[both & 7]
        _E2M1_BYTE_LUT = torch.where((both & 8) != 0, -mag, mag).contiguous()
    v = _E2M1_BYTE_LUT[packed.reshape(-1).to(torch.int64)]
    return v.reshape(*packed.shape[:-1], packed.shape[-1] * 2)


def quant_nvfp4(w: torch.Tensor):
    """NVFP4: global = 448*6/absmax, block-16 e4m3 scales (reference tests/kernels/moe/utils.py nvfp4
    recipe / modelopt).  Returns (uint8 [...,N,K/2], e4m3 scales [...,N,K/16] LINEAR layout,
    dequant global scale f32 [...] = 1/global)."""
    *lead, N, K = w.shape
    wf = w.to(F32)
    amax = wf.abs().amax(dim=(-2, -1)).clamp(min=1e-6)
    gs = (FP8_MAX * 6.0) / amax  # quant-time global scale
    t = wf.reshape(*lead, N, K // 16, 16)
    bmax = t.abs().amax(dim=-1, keepdim=True)
    bs = (bmax / 6.0 * gs[..., None, None, None]).clamp(max=FP8_MAX).to(FP8)
    bsf = bs.to(F32) / gs[..., None, None, None]
    bsf = torch.where(bsf == 0, torch.ones_like(bsf), bsf)
    code = _e2m1_quant_index((t / bsf).reshape(*lead, N, K))
    return _pack_nibbles(code), bs.reshape(*lead, N, K // 16).contiguous(), (1.0 / gs).to(F32)


def dequant_nvfp4(packed: torch.Tensor, block_scale: torch.Tensor, global_scale: torch.Tensor,
                  out_dtype=torch.bfloat16) -> torch.Tensor:
    """w = e2m1 * e4m3_block_scale * global (reference tests/kernels/quantization/nvfp4_utils.py:38-62
    with ``global`` already the multiplicative dequant factor as passed at the lk_moe boundary,
    reference routed_experts.py:1686-1688)."""
    v = unpack_e2m1(packed)
    s = block_scale.to(F32).repeat_interleave(16, dim=-1)
    g = global_scale.to(F32)
    while g.dim() < v.dim():
        g = g.unsqueeze(-1)
    return (v * s * g).to(out_dtype).to(F32)


def quant_mxfp4(w: torch.Tensor):
    """MXFP4: e8m0 block-32 scales = 2^ceil(log2(absmax/6)) (OCP MX recipe used by reference
    tests/kernels/moe/test_ocp_mx_moe.py).  Returns (uint8 [...,N,K/2], uint8 e8m0 [...,N,K/32])."""
    *lead, N, K = w.shape
    t = w.to(F32).reshape(*lead, N, K // 32, 32)
    bmax = t.abs().amax(dim=-1, keepdim=True).clamp(min=2.0 ** -120)
    e = torch.ceil(torch.log2(bmax / 6.0)).clamp(-127, 127)
    sc = torch.pow(2.0, e)
    code = _e2m1_quant_index((t / sc).reshape(*lead, N, K))
    e8 = (e + 127).to(torch.uint8).reshape(*lead, N, K // 32)
    return _pack_nibbles(code), e8.contiguous()


def dequant_mxfp4(packed: torch.Tensor, e8m0: torch.Tensor, out_dtype=torch.bfloat16) -> torch.Tensor:
    """reference tests/kernels/moe/test_ocp_mx_moe.py:153-171 (scale = bits<<23 viewed as f32)."""
    v = unpack_e2m1(packed)
    s = (e8m0.to(torch.int32) << 23).view(F32).repeat_interleave(32, dim=-1)
    return (v * s).to(out_dtype).to(F32)


# --------------------------------------------------------------------------------------------
# Expert forward
# --------------------------------------------------------------------------------------------

ACT_SILU, ACT_SWIGLUOAI, ACT_RELU2 = 0, 1, 2


def apply_activation(h: torch.Tensor, activation_type: int, has_gate: bool = True,
                     alpha: float = 1.702, limit: float = 7.0, interleaved: bool = False) -> torch.Tensor:
    """reference vllm/model_executor/layers/fused_moe/activation.py:128-213.  Gated layout is packed
    halves: rows [0:I] gate, [I:2I] up (reference routed_experts.py:564-570); ``interleaved=True`` is the
    GPT-OSS layout (gate = even columns, up = odd columns) of SWIGLUOAI proper
    (reference csrc/cpu/cpu_fused_moe_activations.hpp:36-77, layers/activation.py:495-503)."""
    if not has_gate:
        if activation_type == ACT_RELU2:
            return torch.relu(h) ** 2
        raise ValueError("non-gated experts use relu2 (activation_type=2)")
    I = h.shape[-1] // 2
    g, u = (h[..., 0::2], h[..., 1::2]) if interleaved else (h[..., :I], h[..., I:])
    if activation_type == ACT_SILU:
        return torch.nn.functional.silu(g) * u
    if activation_type == ACT_SWIGLUOAI:  # packed-halves variant (SWIGLUOAI_UNINTERLEAVE), clamp
        g = g.clamp(max=limit)
        u = u.clamp(min=-limit, max=limit)
        return g * torch.sigmoid(alpha * g) * (u + 1.0)   # operation order of the reference's forward_native
    raise ValueError(activation_type)


@dataclass
class DequantExperts:
    """fp32 expert weights already dequantised: w13 [E,2I,H] (or [E,I,H] non-gated), w2 [E,H,I]."""
    w13: torch.Tensor
    w2: torch.Tensor


def experts_forward(hidden: torch.Tensor, w: DequantExperts, topk_ids: torch.Tensor,
                    topk_weights: torch.Tensor, activation_type: int = ACT_SILU, has_gate: bool = True,
                    act_dtype=torch.bfloat16, round_intermediate: bool = True) -> torch.Tensor:
    """Weight-only oracle: out[t] = sum_j w[t,j] * W2_e( act(W13_e x[t]) ), e = ids[t,j]; ids<0 skipped.

    Follows reference tests/kernels/utils.py:855-994 (torch_experts, quant_dtype None branch) with fp32
    accumulation; the intermediate is rounded to the activation dtype like the reference's
    ``tmp2 = act()(tmp1)`` in ``a.dtype`` (set round_intermediate=False for a pure fp32 chain).
    Output fp32 [M,H] (lk_moe cpu_decode contract, reference routed_experts.py:1833-1855).
    """
    x = hidden.to(F32)
    M, H = x.shape
    out = torch.zeros(M, H, dtype=F32)
    k = topk_ids.shape[1]
    for t in range(M):
        for j in range(k):
            e = int(topk_ids[t, j])
            if e < 0 or e >= w.w13.shape[0]:
                continue
            h1 = w.w13[e] @ x[t]
            a = apply_activation(h1, activation_type, has_gate)
            if round_intermediate:
                a = a.to(act_dtype).to(F32)
            out[t] += float(topk_weights[t, j]) * (w.w2[e] @ a)
    return out


def experts_forward_batched(hidden, w: DequantExperts, topk_ids, topk_weights, activation_type=ACT_SILU,
                            has_gate=True, act_dtype=torch.bfloat16, round_intermediate=True):
    """Same math as experts_forward, grouped per expert so larger M finishes in seconds."""
    x = hidden.to(F32)
    M, H = x.shape
    k = topk_ids.shape[1]
    out = torch.zeros(M, H, dtype=F32)
    flat = topk_ids.reshape(-1)
    tok = torch.arange(M).repeat_interleave(k)
    wts = topk_weights.reshape(-1).to(F32)
    for e in range(w.w13.shape[0]):
        sel = (flat == e).nonzero().flatten()
        if sel.numel() == 0:
            continue
        xs = x[tok[sel]]
        a = apply_activation(xs @ w.w13[e].T, activation_type, has_gate)
        if round_intermediate:
            a = a.to(act_dtype).to(F32)
        y = (a @ w.w2[e].T) * wts[sel, None]
        out.index_add_(0, tok[sel], y)
    return out


def mx_quant_act(x: torch.Tensor) -> torch.Tensor:
    """MXFP8 activation quantisation of the native MXFP4 path (W4A8-MX): e4m3 values with one ue8m0 scale per 32
    channels, scale = 2^ceil(log2(absmax / 448)) computed on the fp32 bit pattern exactly as the kernel does
    (lvllm_b200/csrc/moe_fused.cu::mx_scale_byte); returns the dequantised fp32 tensor.  OCP MX block format as in
    reference tests/kernels/moe/test_ocp_mx_moe.py:150-171 (e8m0 scale = bits << 23)."""
    xf = x.to(F32)
    *lead, K = xf.shape
    g = xf.reshape(*lead, K // 32, 32)
    am = g.abs().amax(dim=-1, keepdim=True).clamp(min=1e-30)
    t = (am * torch.tensor(1.0 / 448.0, dtype=F32)).contiguous()
    bits = t.view(torch.int32)
    eb = (bits >> 23) + ((bits & 0x7FFFFF) != 0).to(torch.int32)
    eb = eb.clamp(1, 253)
    scale = (eb << 23).view(F32)
    q = (g / scale).to(FP8).to(F32)
    return (q * scale).reshape(*lead, K)


def experts_forward_w4a8_mx(hidden, w: DequantExperts, topk_ids, topk_weights, activation_type=ACT_SILU,
                            has_gate=True):
    """Oracle of the native block-scaled MXFP4 path: weights dequantised exactly (``w``), activations and the
    intermediate quantised to MXFP8 per 32 channels, gate / up sums and the activation rounded to fp16 as the
    kernel's epilogue does.  Same expert semantics as experts_forward."""
    x = mx_quant_act(hidden.to(F32))
    M, H = x.shape
    k = topk_ids.shape[1]
    out = torch.zeros(M, H, dtype=F32)
    flat = topk_ids.reshape(-1)
    tok = torch.arange(M).repeat_interleave(k)
    wts = topk_weights.reshape(-1).to(F32)
    for e in range(w.w13.shape[0]):
        sel = (flat == e).nonzero().flatten()
        if sel.numel() == 0:
            continue
        h1 = (x[tok[sel]] @ w.w13[e].T).to(torch.float16).to(F32)
        a = apply_activation(h1, activation_type, has_gate).clamp(-65504, 65504).to(torch.float16).to(F32)
        y = (mx_quant_act(a) @ w.w2[e].T) * wts[sel, None]
        out.index_add_(0, tok[sel], y)
    return out


def experts_forward_lazy(hidden, num_experts: int, weights_of, topk_ids, topk_weights, mode: str = "weight_only",
                         activation_type=ACT_SILU, has_gate=True, act_dtype=torch.bfloat16, interleaved=False):
    """experts_forward_batched (mode "weight_only") / experts_forward_w4a8_mx (mode "w4a8_mx") with expert e's
    dequantised fp32 weights produced on demand by ``weights_of(e) -> (w13_e [2I,H], w2_e [H,I])`` and dropped
    again, so that model-sized layers (the shapes bench.py times: 128 experts x 18.9 M weights) fit in host memory.
    Only experts that receive tokens are materialised.  Same per-expert arithmetic, same citations."""
    assert mode in ("weight_only", "w4a8_mx")
    x = hidden.to(F32)
    if mode == "w4a8_mx":
        x = mx_quant_act(x)
    M, H = x.shape
    k = topk_ids.shape[1]
    out = torch.zeros(M, H, dtype=F32)
    flat = topk_ids.reshape(-1)
    tok = torch.arange(M).repeat_interleave(k)
    wts = topk_weights.reshape(-1).to(F32)
    for e in range(num_experts):
        sel = (flat == e).nonzero().flatten()
        if sel.numel() == 0:
            continue
        w13e, w2e = weights_of(e)
        h1 = x[tok[sel]] @ w13e.to(F32).T
        if mode == "w4a8_mx":
            h1 = h1.to(torch.float16).to(F32)
            a = apply_activation(h1, activation_type, has_gate, interleaved=interleaved)
            a = mx_quant_act(a.clamp(-65504, 65504).to(torch.float16).to(F32))
        else:
            a = apply_activation(h1, activation_type, has_gate, interleaved=interleaved).to(act_dtype).to(F32)
        out.index_add_(0, tok[sel], (a @ w2e.to(F32).T) * wts[sel, None])
    return out


def w8a8_block_matmul(xq: torch.Tensor, xs: torch.Tensor, wq: torch.Tensor, ws: torch.Tensor,
                      block=(128, 128)) -> torch.Tensor:
    """reference tests/kernels/quant_utils.py:91-154 (native_w8a8_block_matmul): per K-tile partial
    products scaled by a_scale[m,kt]*w_scale[nt,kt], fp32."""
    A = xq.to(F32)
    B = wq.to(F32)
    M, K = A.shape
    N = B.shape[0]
    bn, bk = block
    C = torch.zeros(M, N, dtype=F32)
    for kt in range(-(-K // bk)):
        a = A[:, kt * bk:(kt + 1) * bk]
        part = a @ B[:, kt * bk:(kt + 1) * bk].T
        sw = ws[:, kt].repeat_interleave(bn)[:N]
        C += part * xs[:, kt:kt + 1] * sw[None, :]
    return C


def experts_forward_w8a8_block(hidden, w13_q, w13_s, w2_q, w2_s, topk_ids, topk_weights,
                               act_dtype=torch.bfloat16, block=(128, 128), activation_type=ACT_SILU, has_gate=True,
                               ue8m0=False):
    """Block-FP8 W8A8 oracle (DeepSeek-V3 numerics): activations quantised per token per 128 group
    before each GEMM.  Follows reference tests/kernels/utils.py:929-950 (block_shape branch of
    torch_experts) == tests/kernels/moe/test_block_fp8.py:112-137; GEMM outputs rounded to act dtype,
    final weighted sum in fp32 (:976-980).  ``ue8m0``: the DeepGEMM-on-Blackwell variant of the same chain — the caller
    passes weights from requant_weight_ue8m0 and both activation quantisations use power-of-two scales."""
    x = hidden.to(act_dtype)
    M, H = x.shape
    k = topk_ids.shape[1]
    out = torch.zeros(M, H, dtype=F32)
    xq, xs = per_token_group_quant_fp8(x, block[1], ue8m0=ue8m0)
    flat = topk_ids.reshape(-1)
    tok = torch.arange(M).repeat_interleave(k)
    wts = topk_weights.reshape(-1).to(F32)
    for e in range(w13_q.shape[0]):
        sel = (flat == e).nonzero().flatten()
        if sel.numel() == 0:
            continue
        h1 = w8a8_block_matmul(xq[tok[sel]], xs[tok[sel]], w13_q[e], w13_s[e], block).to(act_dtype)
        a = apply_activation(h1.to(F32), activation_type, has_gate).to(act_dtype)
        aq, as_ = per_token_group_quant_fp8(a, block[1], ue8m0=ue8m0)
        y = w8a8_block_matmul(aq, as_, w2_q[e], w2_s[e], block).to(act_dtype).to(F32)
        out.index_add_(0, tok[sel], y * wts[sel, None])
    return out


# --------------------------------------------------------------------------------------------
# Decode attention
# --------------------------------------------------------------------------------------------


def mla_decode(q_nope: torch.Tensor, q_pe: torch.Tensor, kv_cache: torch.Tensor, seq_lens: torch.Tensor,
               page_table: torch.Tensor, sm_scale: float, kv_lora: int = 512):
    """Paged MLA decode in the absorbed form.

    Follows reference tests/kernels/attention/test_cutlass_mla_decode.py:150-196 (per-request SDPA over
    the gathered latent cache, K = all 576 columns, V = first 512) and the op contract
    csrc/libtorch_stable/attention/mla/sm100_cutlass_mla_kernel.cu:225-262.
    q_nope [B,Hq,512], q_pe [B,Hq,64], kv_cache [pages,page,576], page_table i32 [B,max_pages].
    Returns out f32 [B,Hq,512], lse f32 [B,Hq] (natural log, scaled logits)."""
    B, Hq, _ = q_nope.shape
    page = kv_cache.shape[1]
    out = torch.zeros(B, Hq, kv_lora, dtype=F32)
    lse = torch.zeros(B, Hq, dtype=F32)
    q = torch.cat([q_nope, q_pe], dim=-1).to(F32)
    for b in range(B):
        S = int(seq_lens[b])
        if S == 0:
            lse[b] = -math.inf
            continue
        npg = -(-S // page)
        kv = kv_cache[page_table[b, :npg].long()].reshape(-1, kv_cache.shape[-1])[:S].to(F32)
        logits = (q[b] @ kv.T) * sm_scale
        lse[b] = torch.logsumexp(logits, dim=-1)
        p = torch.softmax(logits, dim=-1)
        out[b] = p @ kv[:, :kv_lora]
    return out, lse


def gqa_decode(q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, seq_lens: torch.Tensor,
               page_table: torch.Tensor, sm_scale: float):
    """Paged GQA decode.  Follows reference tests/kernels/attention/test_flashinfer.py:29-80
    (ref_paged_attn with query_len 1): q [B,Hq,D], k_cache/v_cache [pages,page,Hkv,D].
    Returns out f32 [B,Hq,D], lse f32 [B,Hq]."""
    B, Hq, D = q.shape
    page, Hkv = k_cache.shape[1], k_cache.shape[2]
    grp = Hq // Hkv
    out = torch.zeros(B, Hq, D, dtype=F32)
    lse = torch.zeros(B, Hq, dtype=F32)
    for b in range(B):
        S = int(seq_lens[b])
        if S == 0:
            lse[b] = -math.inf
            continue
        npg = -(-S // page)
        idx = page_table[b, :npg].long()
        kk = k_cache[idx].reshape(-1, Hkv, D)[:S].to(F32)
        vv = v_cache[idx].reshape(-1, Hkv, D)[:S].to(F32)
        kk = kk.repeat_interleave(grp, dim=1)
        vv = vv.repeat_interleave(grp, dim=1)
        logits = torch.einsum("hd,shd->hs", q[b].to(F32), kk) * sm_scale
        lse[b] = torch.logsumexp(logits, dim=-1)
        p = torch.softmax(logits, dim=-1)
        out[b] = torch.einsum("hs,shd->hd", p, vv)
    return out, lse


# --------------------------------------------------------------------------------------------
# MLA decode neighbours: RoPE + latent-cache write, q absorb, v up-projection
# --------------------------------------------------------------------------------------------
def apply_rotary_emb(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, is_neox_style: bool) -> torch.Tensor:
    """reference vllm/model_executor/layers/rotary_embedding/common.py:146-185 (ApplyRotaryEmb.forward_static):
    x [T, heads, rot], cos / sin [T, rot/2]; arithmetic in x.dtype exactly as the reference writes it."""
    cos = cos.unsqueeze(-2).to(x.dtype)
    sin = sin.unsqueeze(-2).to(x.dtype)
    if is_neox_style:
        x1, x2 = torch.chunk(x, 2, dim=-1)
    else:
        x1, x2 = x[..., ::2], x[..., 1::2]
    o1 = x1 * cos - x2 * sin
    o2 = x2 * cos + x1 * sin
    if is_neox_style:
        return torch.cat((o1, o2), dim=-1)
    return torch.stack((o1, o2), dim=-1).flatten(-2)


def rope_forward_static(positions, query, key, head_size, rotary_dim, cos_sin_cache, is_neox_style):
    """reference rotary_embedding/base.py:161-201 (RotaryEmbedding.forward_static)."""
    positions = positions.flatten()
    T = positions.shape[0]
    cos, sin = cos_sin_cache.index_select(0, positions).chunk(2, dim=-1)
    out = []
    for x in (query, key):
        shp = x.shape
        x = x.reshape(T, -1, head_size)
        rot = apply_rotary_emb(x[..., :rotary_dim], cos, sin, is_neox_style)
        out.append(torch.cat((rot, x[..., rotary_dim:]), dim=-1).reshape(shp))
    return out[0], out[1]


def concat_and_cache_mla(kv_c, k_pe, kv_cache, slot_mapping, scale: float = 1.0):
    """reference csrc/libtorch_stable/cache_kernels.cu:403-444: row slot of the paged latent cache
    [blocks, block_size, 576] <- [kv_c (512) | k_pe (64)]; slot < 0 = padded token; an fp8 cache stores value / scale."""
    blocks, bs, D = kv_cache.shape
    flat = kv_cache.view(blocks * bs, D)
    for t in range(kv_c.shape[0]):
        sl = int(slot_mapping[t])
        if sl < 0:
            continue
        row = torch.cat([kv_c[t], k_pe[t].reshape(-1)]).to(F32)
        flat[sl] = (row / scale).to(kv_cache.dtype) if kv_cache.dtype == FP8 else row.to(kv_cache.dtype)
    return kv_cache


def mla_q_absorb(q_nope: torch.Tensor, w_uk_t: torch.Tensor) -> torch.Tensor:
    """reference mla_attention.py:875-893: (N,B,P) x (N,P,L) -> (N,B,L) -> (B,N,L); q_nope [B,N,P], W_UK_T [N,P,L]."""
    return torch.bmm(q_nope.transpose(0, 1).to(F32), w_uk_t.to(F32)).transpose(0, 1).to(q_nope.dtype)


def mla_v_up(o: torch.Tensor, w_uv: torch.Tensor) -> torch.Tensor:
    """reference mla_attention.py:1154-1176 (_v_up_proj): (N,B,L) x (N,L,V) -> (B,N,V); o [B,N,L], W_UV [N,L,V]."""
    return torch.bmm(o.transpose(0, 1).to(F32), w_uv.to(F32)).transpose(0, 1).to(o.dtype)


# --------------------------------------------------------------------------------------------
# RMSNorm around the MoE (post-MoE all-reduce / combine fused with residual add + RMSNorm)
# --------------------------------------------------------------------------------------------
def rms_norm(x: torch.Tensor, weight: torch.Tensor | None, eps: float) -> torch.Tensor:
    """reference vllm/ir/ops/layernorm.py:10-21 (what RMSNorm.forward_native calls, layers/layernorm.py:74-94):
    fp32 statistics, the product with the weight in the WEIGHT's dtype, result in x's dtype."""
    orig = x.dtype
    x = x.to(F32)
    x = x * torch.rsqrt(x.pow(2).mean(dim=-1, keepdim=True) + eps)
    if weight is not None:
        x = x.to(weight.dtype) * weight
    return x.to(orig)


def fused_add_rms_norm(x: torch.Tensor, residual: torch.Tensor, weight: torch.Tensor | None, eps: float):
    """reference vllm/ir/ops/layernorm.py:44-63: s = x + residual in fp32; new residual = s in x's dtype; y = rms_norm
    of the UNROUNDED fp32 sum.  Returns (y, new_residual)."""
    orig = x.dtype
    s = x.to(F32) + residual.to(F32)
    new_res = s.to(orig)
    y = s * torch.rsqrt(s.pow(2).mean(dim=-1, keepdim=True) + eps)
    if weight is not None:
        y = y.to(weight.dtype) * weight
    return y.to(orig), new_res


def moe_sum_add_rms_norm(partial_sum_f32: torch.Tensor, residual: torch.Tensor | None, weight: torch.Tensor | None,
                         gain: float, eps: float, act_dtype=torch.bfloat16):
    """What b200_ep_allreduce_norm / b200_ep_combine_norm / b200_rmsnorm_cast compute after the reduction over the ranks:
    the reference's fused_add_rms_norm applied to the fp32 MoE output WITHOUT first rounding it to the activation dtype
    (the reference rounds the all-reduced MoE output to bf16 before the next layer's norm, moe_runner.py:488-494; the
    fused kernel keeps the fp32 sum, which is at least as precise).  weight None: scalar `gain` instead (bench layers).
    Returns (y, new_residual, fp32 sum)."""
    s = partial_sum_f32.to(F32) + (residual.to(F32) if residual is not None else 0)
    new_res = s.to(act_dtype)
    y = s * torch.rsqrt(s.pow(2).mean(dim=-1, keepdim=True) + eps)
    y = (y.to(act_dtype) * weight) if weight is not None else (y * gain).to(act_dtype)
    return y, new_res, s

