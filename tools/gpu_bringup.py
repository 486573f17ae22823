"""GPU bring-up: staged, verbose checks of the CUDA path against the oracle (run on the B200 box).
Each stage runs in its own subprocess with a timeout so that a trapped kernel cannot hide later stages.
    python tools/gpu_bringup.py [stage ...]
"""
import ctypes
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _moe_case(fmt, M, E, k, H, I, seed=0, dump=True):
    import torch
    import lk_moe
    from oracle import moe_oracle as O
    from lvllm_b200 import _lib
    g = torch.Generator().manual_seed(seed)
    hidden = (torch.randn(M, H, generator=g) / 10).bfloat16()
    score = torch.randn(M, E, generator=g)
    w, ids = torch.topk(torch.softmax(score, -1), k)
    w = w.float().contiguous()
    ids = ids.int().contiguous()
    cfg = lk_moe.MOEConfigV2()
    cfg.expert_num, cfg.top_k, cfg.hidden_size, cfg.intermediate_size = E, k, H, I
    cfg.max_batch_size, cfg.max_num_seqs = 4096, 64
    if fmt == "bf16":
        w13 = (torch.randn(E, 2 * I, H, generator=g) / 10).bfloat16()
        w2 = (torch.randn(E, H, I, generator=g) / 10).bfloat16()
        ref = O.experts_forward_batched(hidden, O.DequantExperts(w13.float(), w2.float()), ids, w)
        moe = lk_moe.MOE_BF16(cfg, w13.data_ptr(), w2.data_ptr(), 0, 0, 0, 0)
    else:
        cfg.groupN = cfg.groupK = 128
        w13, s13 = O.quant_fp8_block(torch.randn(E, 2 * I, H, generator=g) / 10)
        w2, s2 = O.quant_fp8_block(torch.randn(E, H, I, generator=g) / 10)
        ref = O.experts_forward_w8a8_block(hidden, w13, s13, w2, s2, ids, w)
        moe = lk_moe.MOE_FP8(cfg, w13.data_ptr(), w2.data_ptr(), s13.data_ptr(), s2.data_ptr(), 0, 0)
    out = torch.empty(M, H, dtype=torch.float32)
    moe.cpu_prefill(M, k, ids.data_ptr(), w.data_ptr(), hidden.data_ptr(), out.data_ptr())
    err = (out - ref).abs().max().item()
    rel = ((out - ref).abs().mean() / ref.abs().mean()).item()
    print(f"[{fmt}] M={M} E={E} k={k} H={H} I={I}: max_abs_err={err:.4e} rel={rel:.4e} ref_absmean={ref.abs().mean():.4e}")
    if rel > 0.02 and dump:
        st = torch.zeros(8, dtype=torch.int32)
        _lib.lib().b200moe_debug_read(5, st.data_ptr(), 32)
        print("  route state:", st.tolist())
        ch = torch.zeros(4 * 8, dtype=torch.int32)
        _lib.lib().b200moe_debug_read(6, ch.data_ptr(), 4 * 8 * 4)
        print("  chunks:", ch.view(-1, 4)[: max(1, st[0])].tolist())
        ros = torch.zeros(M * k, dtype=torch.int32)
        _lib.lib().b200moe_debug_read(7, ros.data_ptr(), M * k * 4)
        print("  row_of_slot:", ros.tolist()[:32])
        nrows = int(st[1])
        y = torch.zeros(nrows, H, dtype=torch.float32)
        _lib.lib().b200moe_debug_read(4, y.data_ptr(), nrows * H * 4)
        print("  y[row0,:8]    =", y[0, :8].tolist())
        # expected y for slot 0
        e0 = int(ids[0, 0])
        if fmt == "bf16":
            h1 = w13[e0].float() @ hidden[0].float()
            a = (torch.nn.functional.silu(h1[:I]) * h1[I:]).bfloat16().float()
            yexp = w2[e0].float() @ a
            print("  y_exp[row0,:8]=", yexp[:8].tolist())
            # intermediate (tiled bf16): decode row 0
            KB2 = I // 64
            raw = torch.zeros(max(1, nrows // 8) * KB2 * 1024, dtype=torch.uint8)
            _lib.lib().b200moe_debug_read(2, raw.data_ptr(), raw.numel())
            row = []
            for f in range(16):
                kb, b = f // 64, (f % 64) * 2
                off = kb * 1024 + 0 * 128 + (((b >> 4) ^ 0) << 4) + (b & 15)
                row.append(raw[off:off + 2].view(torch.bfloat16).item())
            print("  inter[row0,:16]    =", row)
            print("  inter_exp[row0,:16]=", a[:16].tolist())
    moe.close()
    return rel


def stage_routing():
    import torch
    from oracle import moe_oracle as O
    from lvllm_b200 import ops
    logits = torch.randn(9, 256)
    bias = torch.randn(256)
    w, ids = ops.fused_topk(logits.cuda(), 8, True, "sigmoid", bias.cuda(), 2.5)
    wr, ir = O.topk_gating(logits, 8, True, "sigmoid", bias, 2.5)
    print("fused_topk ids equal:", torch.equal(ids.cpu(), ir), "max w err", (w.cpu() - wr).abs().max().item())
    w, ids = ops.grouped_topk(logits.cuda(), 8, True, 8, 4, "sigmoid", 2.5, bias.cuda())
    wr, ir = O.grouped_topk(logits, bias, 8, 4, 8, True, 2.5)
    print("grouped_topk ids equal:", torch.equal(ids.cpu(), ir), "max w err", (w.cpu() - wr).abs().max().item())


def stage_bf16():
    _moe_case("bf16", 1, 2, 1, 256, 128)
    _moe_case("bf16", 4, 4, 2, 512, 256, seed=1)
    _moe_case("bf16", 33, 8, 2, 1024, 512, seed=2, dump=False)


def stage_fp8():
    _moe_case("fp8", 1, 2, 1, 256, 128)
    _moe_case("fp8", 4, 4, 2, 512, 256, seed=1)
    _moe_case("fp8", 33, 8, 2, 1024, 512, seed=2, dump=False)


def stage_bw():
    """first bandwidth read: DeepSeek-V3 expert shapes, FP8, 32 local experts, M=1, k=8 under a CUDA graph"""
    import torch
    import lk_moe
    E, k, H, I = 32, 8, 7168, 2048
    dev = torch.device("cuda")
    w13 = (torch.randn(E, 2 * I, H, device=dev, dtype=torch.bfloat16) / 10).to(torch.float8_e4m3fn)
    w2 = (torch.randn(E, H, I, device=dev, dtype=torch.bfloat16) / 10).to(torch.float8_e4m3fn)
    s13 = torch.rand(E, 2 * I // 128, H // 128, device=dev) * 0.01 + 0.001
    s2 = torch.rand(E, H // 128, I // 128, device=dev) * 0.01 + 0.001
    cfg = lk_moe.MOEConfigV2()
    cfg.expert_num, cfg.top_k, cfg.hidden_size, cfg.intermediate_size = E, k, H, I
    cfg.max_batch_size, cfg.max_num_seqs, cfg.groupN, cfg.groupK = 4096, 64, 128, 128
    moe = lk_moe.MOE_FP8(cfg, w13.data_ptr(), w2.data_ptr(), s13.data_ptr(), s2.data_ptr(), 0, 0, weights_on_device=True)
    del w13, w2
    from lvllm_b200 import _lib
    _dbg = torch.zeros(160 * 16, dtype=torch.int64)
    _lib.lib().b200moe_debug_read(9, _dbg.data_ptr(), _dbg.numel() * 8)   # arm the in-kernel stamps before capture
    for M in ([1] if os.environ.get('B200MOE_DBG_MODE') else (1, 4, 16, 64)):
        hidden = (torch.randn(M, H, device=dev) / 10).bfloat16()
        ids = torch.stack([torch.randperm(E, device=dev)[:k] for _ in range(M)]).int().contiguous()
        w = torch.rand(M, k, device=dev).float()
        out = torch.zeros(M, H, device=dev)
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            moe.cpu_decode(st.cuda_stream, M, k, hidden.data_ptr(), ids.data_ptr(), w.data_ptr(), out.data_ptr())
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=st):
                moe.cpu_decode(torch.cuda.current_stream().cuda_stream, M, k, hidden.data_ptr(), ids.data_ptr(),
                               w.data_ptr(), out.data_ptr())
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            gr.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        # in-kernel timeline of the fused kernel (globaltimer stamps per CTA)
        from lvllm_b200 import _lib
        dbg = torch.zeros(160 * 16, dtype=torch.int64)
        gr.replay()
        torch.cuda.synchronize()
        _lib.lib().b200moe_debug_read(9, dbg.data_ptr(), dbg.numel() * 8)
        d = dbg.view(160, 16)[:148].double()
        t0 = d[:, 0][d[:, 0] > 0].min()
        names = ["entry", "table", "gather", "x_ok", "A1 issued", "g1 first ok", "all issued", "epi ph1", "epi ph2", "exit", "F done", "F comb start"]
        line = []
        for i, nm in enumerate(names):
            col = d[:, i]
            col = col[col > 0]
            if col.numel():
                line.append(f"{nm}: {((col.min()-t0)/1e3):.1f}/{((col.median()-t0)/1e3):.1f}/{((col.max()-t0)/1e3):.1f}")
        print(f"   timeline us (min/med/max over CTAs) M={M}: " + " | ".join(line))
        for i, nm in zip(range(12, 16), ["F cnt_inc+bars", "F partial loads", "F act+quant+stores", "F proxy fence"]):
            col = d[:, i]; col = col[col > 0]
            if col.numel():
                print(f"   {nm} duration us: min {col.min()/1e3:.2f} med {col.median()/1e3:.2f} max {col.max()/1e3:.2f} (n={col.numel()})")
        ne = len(torch.unique(ids))
        by = ne * 44.051e6
        print(f"M={M}: distinct experts={ne} median {ts[len(ts)//2]*1e3:.1f} us  min {ts[0]*1e3:.1f} us -> "
              f"{by/ts[len(ts)//2]/1e6:.0f} GB/s (algorithmic expert bytes)")


def stage_mla():
    import math
    import torch
    from oracle import moe_oracle as O
    from lvllm_b200 import ops
    g = torch.Generator().manual_seed(0)
    for (B, S, page, Hq) in [(1, 100, 64, 128), (1, 4096, 64, 128), (2, 300, 16, 16)]:
        npg = -(-S // page)
        cache = torch.randn(B * npg, page, 576, generator=g).bfloat16()
        pt = torch.randperm(B * npg, generator=g).reshape(B, npg).int()
        lens = torch.tensor([S - 7 * b for b in range(B)], dtype=torch.int32)
        qn = torch.randn(B, Hq, 512, generator=g).bfloat16()
        qp = torch.randn(B, Hq, 64, generator=g).bfloat16()
        sc = 1 / math.sqrt(576)
        ref, lse_ref = O.mla_decode(qn, qp, cache, lens, pt, sc)
        out, lse = ops.mla_decode(qn.cuda(), qp.cuda(), cache.cuda(), lens.cuda(), pt.cuda(), sc)
        o = out.cpu().float()
        cos = 1 - 2 * (o.double() * ref.double()).sum() / ((o.double() ** 2 + ref.double() ** 2).sum())
        print(f"[mla] B={B} S={S} page={page} Hq={Hq}: max_abs_err={(o-ref).abs().max():.4e} cos_diff={cos:.3e} "
              f"lse_err={(lse.cpu()-lse_ref).abs().max():.3e} ref_absmean={ref.abs().mean():.3e}")
        sys.stdout.flush()
        if (o - ref).abs().max() > 0.05:
            print("   out[0,0,:8] =", o[0, 0, :8].tolist())
            print("   ref[0,0,:8] =", ref[0, 0, :8].tolist())
            print("   out[0,0,256:264] =", o[0, 0, 256:264].tolist())
            print("   ref[0,0,256:264] =", ref[0, 0, 256:264].tolist())
    # timing at the DeepSeek-V3 decode shape
    print("[mla] timing section", flush=True)
    B, S, page, Hq = 1, 4096, 64, 128
    dev = torch.device("cuda")
    cache = torch.randn(B * S // page, page, 576, device=dev).bfloat16()
    pt = torch.arange(B * S // page, device=dev, dtype=torch.int32).reshape(B, -1)
    lens = torch.full((B,), S, device=dev, dtype=torch.int32)
    qn = torch.randn(B, Hq, 512, device=dev).bfloat16()
    qp = torch.randn(B, Hq, 64, device=dev).bfloat16()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3):
            ops.mla_decode(qn, qp, cache, lens, pt, 0.04)
        torch.cuda.synchronize()
        print("[mla] warmup done", flush=True)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st):
            for _ in range(20):
                ops.mla_decode(qn, qp, cache, lens, pt, 0.04)
    print("[mla] captured", flush=True)
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 20 * 1e3)
    print(f"[mla] B=1 S=4096 Hq=128: {sorted(ts)[2]:.1f} us per call (20 calls per graph)")


def stage_gqa():
    import torch
    from oracle import moe_oracle as O
    from lvllm_b200 import ops
    g = torch.Generator().manual_seed(0)
    for (B, S, page, Hq, Hkv) in [(2, 100, 16, 8, 2), (3, 700, 16, 64, 4), (5, 513, 64, 32, 8)]:
        npg = -(-S // page)
        kc = torch.randn(B * npg, page, Hkv, 128, generator=g).bfloat16()
        vc = torch.randn(B * npg, page, Hkv, 128, generator=g).bfloat16()
        pt = torch.randperm(B * npg, generator=g).reshape(B, npg).int()
        lens = torch.tensor([S - 7 * b for b in range(B)], dtype=torch.int32)
        q = torch.randn(B, Hq, 128, generator=g).bfloat16()
        ref, lse_ref = O.gqa_decode(q, kc, vc, lens, pt, 128 ** -0.5)
        out, lse = ops.gqa_decode(q.cuda(), kc.cuda(), vc.cuda(), lens.cuda(), pt.cuda(), 128 ** -0.5)
        o = out.cpu().float()
        cos = 1 - 2 * (o.double() * ref.double()).sum() / ((o.double() ** 2 + ref.double() ** 2).sum())
        print(f"[gqa] B={B} S={S} page={page} Hq={Hq} Hkv={Hkv}: max_abs_err={(o-ref).abs().max():.4e} cos_diff={cos:.3e} "
              f"lse_err={(lse.cpu()-lse_ref).abs().max():.3e}", flush=True)
    dev = torch.device("cuda")
    for (B, S, page, Hq, Hkv) in [(256, 512, 16, 64, 4), (64, 2048, 16, 32, 8), (1, 8192, 16, 64, 4)]:
        npg = S // page
        kc = torch.randn(B * npg, page, Hkv, 128, device=dev).bfloat16()
        vc = torch.randn(B * npg, page, Hkv, 128, device=dev).bfloat16()
        pt = torch.randperm(B * npg, device=dev).reshape(B, npg).int()
        lens = torch.full((B,), S, device=dev, dtype=torch.int32)
        q = torch.randn(B, Hq, 128, device=dev).bfloat16()
        for splits in (0, 8 if B == 1 else 2):
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                for _ in range(3):
                    ops.gqa_decode(q, kc, vc, lens, pt, 0.088, num_kv_splits=splits)
                torch.cuda.synchronize()
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=st):
                    for _ in range(10):
                        ops.gqa_decode(q, kc, vc, lens, pt, 0.088, num_kv_splits=splits)
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) / 10 * 1e3)
            us = sorted(ts)[2]
            byts = B * S * Hkv * 128 * 2 * 2
            print(f"[gqa] B={B} S={S} Hq={Hq} Hkv={Hkv} splits={splits or 'tc'}: {us:.1f} us per call, "
                  f"{byts / us / 1e3:.0f} GB/s of KV", flush=True)


def stage_mixed():
    """does tcgen05 kind::f16 accept A = fp16 (weights) with B = bf16 (activations)?"""
    import torch
    import lk_moe
    from oracle import moe_oracle as O
    os.environ["B200MOE_MIXED_TEST"] = "1"
    g = torch.Generator().manual_seed(3)
    M, E, k, H, I = 4, 4, 2, 512, 256
    hidden = (torch.randn(M, H, generator=g) / 10).bfloat16()
    w13 = (torch.randn(E, 2 * I, H, generator=g) / 10).half()
    w2 = (torch.randn(E, H, I, generator=g) / 10).half()
    tw, ids = torch.topk(torch.softmax(torch.randn(M, E, generator=g), -1), k)
    tw, ids = tw.float().contiguous(), ids.int().contiguous()
    ref = O.experts_forward_batched(hidden, O.DequantExperts(w13.float(), w2.float()), ids, tw)
    cfg = lk_moe.MOEConfigV2()
    cfg.expert_num, cfg.top_k, cfg.hidden_size, cfg.intermediate_size = E, k, H, I
    cfg.max_batch_size, cfg.max_num_seqs = 64, 16
    moe = lk_moe.MOE_BF16(cfg, w13.data_ptr(), w2.data_ptr(), 0, 0, 0, 0)   # fp16 bits, bf16 activations
    out = torch.empty(M, H, dtype=torch.float32)
    moe.cpu_prefill(M, k, ids.data_ptr(), tw.data_ptr(), hidden.data_ptr(), out.data_ptr())
    print(f"[mixed f16 x bf16] max_abs_err={(out-ref).abs().max():.4e} rel={((out-ref).abs().mean()/ref.abs().mean()):.4e}")


def stage_w4():
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_gpu_parity as T
    for fmt in ("int4", "nvfp4", "mxfp4"):
        for M in (1, 16):
            moe, hidden, ids, w, ref = T._w4_case(fmt, M, 2, 1, 256, 128, 11)
            out = torch.empty(M, 256, dtype=torch.float32)
            moe.cpu_prefill(M, 1, ids.data_ptr(), w.data_ptr(), hidden.data_ptr(), out.data_ptr())
            rel = ((out - ref).abs().mean() / ref.abs().mean()).item()
            print(f"[{fmt}] M={M}: max_abs_err={(out-ref).abs().max():.4e} rel={rel:.4e} ref_absmean={ref.abs().mean():.3e}", flush=True)
            if rel > 0.05:
                print("   out[0,:8] =", out[0, :8].tolist())
                print("   ref[0,:8] =", ref[0, :8].tolist())
            moe.close()


def stage_mx():
    """native block-scaled MXFP4 path (opt-in, B200MOE_MX_NATIVE=1): correctness against both oracle modes + timing"""
    import torch
    os.environ["B200MOE_MX_NATIVE"] = "1"
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_gpu_parity as T
    from oracle import moe_oracle as O
    for M in (1, 16, 40):
        E, k, H, I, seed = 2, 1, 256, 128, 11
        moe, hidden, ids, w, ref16 = T._w4_case("mxfp4", M, E, k, H, I, seed)
        g = torch.Generator().manual_seed(seed)
        _ = torch.randn(M, H, generator=g)
        p13, s13 = O.quant_mxfp4(torch.randn(E, 2 * I, H, generator=g) / 10)
        p2, s2 = O.quant_mxfp4(torch.randn(E, H, I, generator=g) / 10)
        ref8 = O.experts_forward_w4a8_mx(hidden, O.DequantExperts(O.dequant_mxfp4(p13, s13), O.dequant_mxfp4(p2, s2)), ids, w)
        out = torch.empty(M, H, dtype=torch.float32)
        moe.cpu_prefill(M, k, ids.data_ptr(), w.data_ptr(), hidden.data_ptr(), out.data_ptr())
        print(f"[mx native] M={M}: rel vs W4A8-MX oracle {((out-ref8).abs().mean()/ref8.abs().mean()).item():.4e} "
              f"max_abs {(out-ref8).abs().max():.4e} | rel vs W4A16 oracle {((out-ref16).abs().mean()/ref16.abs().mean()).item():.4e} "
              f"ref_absmean={ref8.abs().mean():.3e}", flush=True)
        if (out - ref8).abs().mean() / ref8.abs().mean() > 0.02:
            print("   out[0,:8] =", out[0, :8].tolist())
            print("   ref[0,:8] =", ref8[0, :8].tolist())
        moe.close()


def stage_bw4():
    """bandwidth + in-kernel timeline of the 4-bit path at Qwen3-235B expert shapes (MXFP4, 128 experts)"""
    import torch
    import lk_moe
    from lvllm_b200 import _lib
    E, k, H, I = 128, 8, 4096, 1536
    EP = int(os.environ.get("BW4_EP", "1"))     # >1: this GPU holds E/EP experts of an EP job (ids of the others = -1)
    E_global, E = E, E // EP
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(0)
    p13 = torch.randint(0, 256, (E, 2 * I, H // 2), device=dev, dtype=torch.uint8, generator=g)
    p2 = torch.randint(0, 256, (E, H, I // 2), device=dev, dtype=torch.uint8, generator=g)
    s13 = torch.randint(117, 122, (E, 2 * I, H // 32), device=dev, dtype=torch.uint8, generator=g)
    s2 = torch.randint(117, 122, (E, H, I // 32), device=dev, dtype=torch.uint8, generator=g)
    cfg = lk_moe.MOEConfigV2()
    cfg.expert_num, cfg.top_k, cfg.hidden_size, cfg.intermediate_size = E, k, H, I
    cfg.max_batch_size, cfg.max_num_seqs, cfg.groupN, cfg.groupK = 4096, 256, 1, 32
    moe = lk_moe.MOE_MXFP4(cfg, p13.data_ptr(), p2.data_ptr(), s13.data_ptr(), s2.data_ptr(), 0, 0, weights_on_device=True)
    del p13, p2
    _dbg = torch.zeros(160 * 16, dtype=torch.int64)
    _lib.lib().b200moe_debug_read(9, _dbg.data_ptr(), _dbg.numel() * 8)
    bpe = 3 * H * I * (0.5 + 1 / 32)
    for M in ([16] if os.environ.get('B200MOE_DBG_MODE') else (1, 16, 64, 256)):
        hidden = (torch.randn(M, H, device=dev) / 10).bfloat16()
        ids = torch.stack([torch.randperm(E_global, device=dev)[:k] for _ in range(M)]).int()
        ids = torch.where(ids < E, ids, torch.full_like(ids, -1)).contiguous()
        w = torch.rand(M, k, device=dev).float()
        out = torch.zeros(M, H, device=dev)
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            moe.cpu_decode(st.cuda_stream, M, k, hidden.data_ptr(), ids.data_ptr(), w.data_ptr(), out.data_ptr())
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=st):
                moe.cpu_decode(torch.cuda.current_stream().cuda_stream, M, k, hidden.data_ptr(), ids.data_ptr(),
                               w.data_ptr(), out.data_ptr())
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        dbg = torch.zeros(160 * 16, dtype=torch.int64)
        _lib.lib().b200moe_debug_read(9, dbg.data_ptr(), dbg.numel() * 8)
        d = dbg.view(160, 16)[:148].double()
        t0 = d[:, 0][d[:, 0] > 0].min()
        names = ["entry", "table", "gather", "x_ok", "A1 issued", "g1 first ok", "all issued", "epi ph1", "epi ph2", "exit", "F done", "F comb start", "fixup done", "g_rd0 issued", "g_rd1 issued", "g_rd2 issued"]
        line = []
        for i, nm in enumerate(names):
            col = d[:, i]; col = col[col > 0]
            if col.numel():
                line.append(f"{nm}: {((col.min()-t0)/1e3):.0f}/{((col.median()-t0)/1e3):.0f}/{((col.max()-t0)/1e3):.0f}")
        ne = len(torch.unique(ids[ids >= 0]))
        print(f"M={M}: experts={ne} {ts[2]*1e3:.0f} us -> {ne*bpe/ts[2]/1e6:.0f} GB/s | " + " | ".join(line), flush=True)
        cyc = d[:, 12:16] * 0   # slot 12 is the fix-up-done stamp now (probes need -DF_PROBE=1)
        print("   dequant warp 11 cycles per k-block (median over CTAs): wait raw-full %.0f | wait dq-slot %.0f | convert %.0f | fence+arrive %.0f"
              % tuple(cyc[:, i][cyc[:, i] > 0].median().item() if (cyc[:, i] > 0).any() else 0 for i in range(4)), flush=True)


STAGES = {"mx": stage_mx, "gqa": stage_gqa, "bw4": stage_bw4, "w4": stage_w4, "mixed": stage_mixed, "mla": stage_mla, "routing": stage_routing, "bf16": stage_bf16, "fp8": stage_fp8, "bw": stage_bw}

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        STAGES[sys.argv[2]]()
        sys.exit(0)
    names = sys.argv[1:] or list(STAGES)
    for n in names:
        t = time.time()
        print(f"===== stage {n}", flush=True)
        try:
            r = subprocess.run([sys.executable, "-u", os.path.abspath(__file__), "--child", n], timeout=int(os.environ.get("STAGE_TIMEOUT", "240")),
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            print(r.stdout[-6000:])
            print(f"===== stage {n} rc={r.returncode} ({time.time()-t:.0f}s)", flush=True)
        except subprocess.TimeoutExpired as ex:
            print((ex.stdout or b"")[-3000:] if isinstance(ex.stdout, (bytes, str)) else "")
            print(f"===== stage {n} TIMEOUT", flush=True)
