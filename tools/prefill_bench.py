"""Prefill-class expert GEMM (SURVEY.md 8 row a11 / BASELINE config 4 batch shape): gpu_prefill of M tokens through one
MoE layer of an EP shard, CUDA-event timed.  Prints achieved dense TFLOP/s (2*3*H*I flops per routed row) next to the
measured cuBLAS bf16 peak of MEASURED_PEAKS.json; run under ncu for the tensor-pipe figure:
    ncu --set full -k regex:moe_gemm_kernel -s 2 -c 2 -o gpurun_out/prof_prefill python tools/prefill_bench.py
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import lk_moe
    fmt = sys.argv[1] if len(sys.argv) > 1 else "fp8"
    M = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    E, k, H, I = 32, 8, 7168, 2048          # DeepSeek-V3 EP8 shard
    n_global = 256
    if fmt == "mxfp4":
        E, k, H, I, n_global = 128, 8, 2048, 768, 128      # Qwen3-30B-A3B layer (BASELINE configs[1]), all experts local
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0)
    cfg = lk_moe.MOEConfigV2()
    cfg.expert_num, cfg.top_k, cfg.hidden_size, cfg.intermediate_size = E, k, H, I
    cfg.max_batch_size, cfg.max_num_seqs = M, 256
    if fmt in ("fp8", "fp8e8m0"):
        if fmt == "fp8e8m0":
            os.environ["B200MOE_FP8_E8M0"] = "1"      # DeepGEMM-on-Blackwell numerics: block-scaled tcgen05.mma, no promotion
        cfg.groupN = cfg.groupK = 128
        w13 = (torch.randn(E, 2 * I, H, device=dev, dtype=torch.bfloat16, generator=g) / 10).to(torch.float8_e4m3fn)
        w2 = (torch.randn(E, H, I, device=dev, dtype=torch.bfloat16, generator=g) / 10).to(torch.float8_e4m3fn)
        s13 = torch.rand(E, 2 * I // 128, H // 128, device=dev, generator=g) * 4e-3 + 1e-3
        s2 = torch.rand(E, H // 128, I // 128, device=dev, generator=g) * 4e-3 + 1e-3
        moe = lk_moe.MOE_FP8(cfg, w13.data_ptr(), w2.data_ptr(), s13.data_ptr(), s2.data_ptr(), 0, 0, weights_on_device=True)
    elif fmt == "mxfp4":
        cfg.groupN, cfg.groupK = 1, 32
        w13 = torch.randint(0, 256, (E, 2 * I, H // 2), device=dev, dtype=torch.uint8, generator=g)   # random e2m1 pairs
        w2 = torch.randint(0, 256, (E, H, I // 2), device=dev, dtype=torch.uint8, generator=g)
        s13 = torch.randint(118, 123, (E, 2 * I, H // 32), device=dev, dtype=torch.uint8, generator=g)   # ue8m0 ~ 2^-7
        s2 = torch.randint(118, 123, (E, H, I // 32), device=dev, dtype=torch.uint8, generator=g)
        moe = lk_moe.MOE_MXFP4(cfg, w13.data_ptr(), w2.data_ptr(), s13.data_ptr(), s2.data_ptr(), 0, 0, weights_on_device=True)
    else:
        w13 = torch.randn(E, 2 * I, H, device=dev, dtype=torch.bfloat16, generator=g) / 10
        w2 = torch.randn(E, H, I, device=dev, dtype=torch.bfloat16, generator=g) / 10
        moe = lk_moe.MOE_BF16(cfg, w13.data_ptr(), w2.data_ptr(), 0, 0, 0, 0, weights_on_device=True)
    del w13, w2
    hid = (torch.randn(M, H, device=dev, generator=g) / 10).bfloat16()
    gids = torch.stack([torch.randperm(n_global, device=dev, generator=g)[:k] for _ in range(M)]).int()
    ids = torch.where(gids < E, gids, torch.full_like(gids, -1)).contiguous()      # rank 0's view of an EP8 job
    w = torch.rand(M, k, device=dev, generator=g).float()
    out = torch.empty(M, H, dtype=torch.bfloat16, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    rows = int((ids >= 0).sum())
    ts = []
    for it in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        moe.gpu_prefill(hid.data_ptr(), out.data_ptr(), ids.data_ptr(), w.data_ptr(), M, k, st)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[1:])
    ms = ts[len(ts) // 2]
    flops = rows * 2.0 * 3 * H * I
    wb = {"fp8": 1, "fp8e8m0": 1, "mxfp4": 0.5 + 1 / 32}.get(fmt, 2)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    print(json.dumps({"workload": f"gpu_prefill {fmt} M={M} k={k} E_local={E} H={H} I={I} ",
                      "routed_rows": rows, "rows_per_expert": rows / E, "ms_per_layer": ms, "tflops": flops / ms / 1e9,
                      "bf16_cublas_peak_tflops_measured": peaks.get("bf16_tflops"),
                      "weights_gb": E * 3 * H * I * wb / 1e9, "weights_gbs_if_read_once": E * 3 * H * I * wb / ms / 1e6,
                      "w4_prefill_min": os.environ.get("B200MOE_W4_PREFILL_MIN"),
                      "finite": bool(torch.isfinite(out.float()).all())}))


if __name__ == "__main__":
    main()
