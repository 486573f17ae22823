"""Bring-up of the opt-in chunk-PAIR form of the 16-bit grouped GEMM (B200MOE_GEMM_PAIR=1, csrc/moe_gemm.cu): one process,
results appended to the output file step by step so that a cut-off call still leaves what was reached.
  1. parity: pair form against one-chunk-per-unit on small layers (even and odd chunk counts per expert) — bit-identical
  2. timing: gpu_prefill of 8192 tokens through a DeepSeek-V3 EP8 shard layer in bf16, both forms
usage: python tools/pair_check.py gpurun_out/r2/pair_check.jsonl
"""
import json
import os
import sys
import time

t_start = time.time()
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/pair_check.jsonl"
os.makedirs(os.path.dirname(OUT) or ".", exist_ok=True)


def emit(**kw):
    kw["t"] = round(time.time() - t_start, 1)
    with open(OUT, "a") as f:
        f.write(json.dumps(kw) + "\n")
        f.flush()
        os.fsync(f.fileno())
    print(json.dumps(kw), flush=True)


def layer(E, k, H, I, M, dev, g):
    import lk_moe
    cfg = lk_moe.MOEConfigV2()
    cfg.expert_num, cfg.top_k, cfg.hidden_size, cfg.intermediate_size = E, k, H, I
    cfg.max_batch_size, cfg.max_num_seqs = M, 256
    w13 = torch.randn(E, 2 * I, H, device=dev, dtype=torch.bfloat16, generator=g) / 10
    w2 = torch.randn(E, H, I, device=dev, dtype=torch.bfloat16, generator=g) / 10
    return lk_moe.MOE_BF16(cfg, w13.data_ptr(), w2.data_ptr(), 0, 0, 0, 0, weights_on_device=True)


def run(moe, hid, ids, w, M, k, pair, iters=1):
    os.environ["B200MOE_GEMM_PAIR"] = "1" if pair else "0"
    out = torch.empty(M, hid.shape[1], dtype=torch.bfloat16, device=hid.device)
    st = torch.cuda.current_stream().cuda_stream
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        moe.gpu_prefill(hid.data_ptr(), out.data_ptr(), ids.data_ptr(), w.data_ptr(), M, k, st)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return out, ts


def main():
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0)
    emit(step="torch imported")
    for (E, k, H, I, M) in ((8, 2, 512, 256, 1024), (8, 2, 512, 256, 1500), (4, 2, 1024, 512, 777)):
        moe = layer(E, k, H, I, M, dev, g)
        hid = (torch.randn(M, H, device=dev, generator=g) / 10).bfloat16()
        ids = torch.randint(0, E, (M, k), device=dev, generator=g, dtype=torch.int32)
        w = torch.rand(M, k, device=dev, generator=g).float()
        o0, _ = run(moe, hid, ids, w, M, k, False)
        o1, _ = run(moe, hid, ids, w, M, k, True)
        emit(step="parity", shape=[E, k, H, I, M], bit_identical=bool(torch.equal(o0, o1)),
             max_abs_diff=float((o0.float() - o1.float()).abs().max()), finite=bool(torch.isfinite(o1.float()).all()))
        moe.close()
    E, k, H, I, M = 32, 8, 7168, 2048, 8192
    moe = layer(E, k, H, I, M, dev, g)
    hid = (torch.randn(M, H, device=dev, generator=g) / 10).bfloat16()
    gids = torch.stack([torch.randperm(256, device=dev, generator=g)[:k] for _ in range(M)]).int()
    ids = torch.where(gids < E, gids, torch.full_like(gids, -1)).contiguous()
    w = torch.rand(M, k, device=dev, generator=g).float()
    rows = int((ids >= 0).sum())
    res = {}
    for pair in (False, True, False, True):
        o, ts = run(moe, hid, ids, w, M, k, pair, iters=6)
        ms = sorted(ts[1:])[2]
        res.setdefault(pair, o)
        emit(step="timing", pair=pair, ms_per_layer=ms, tflops=rows * 6.0 * H * I / ms / 1e9, routed_rows=rows)
    emit(step="parity_8192", bit_identical=bool(torch.equal(res[False], res[True])))


if __name__ == "__main__":
    main()
