"""Bring-up of an opt-in, bit-identical variant of the large-batch path that an environment variable selects per call
(B200MOE_GEMM_PAIR=1: chunk-pair form of the 16-bit grouped GEMM; B200MOE_COMBINE=2: compacted combine).  One process,
results appended to the output file step by step so that a cut-off call still leaves what was reached.
  1. parity: variant against the default on small layers (gpu_prefill bf16 output, cpu_prefill fp32 output) — bit-identical
  2. timing: gpu_prefill of 8192 tokens through a DeepSeek-V3 EP8 shard layer (sparse local ids) and through a layer whose
     experts are all local (every slot valid), bf16, both forms
usage: python tools/variant_check.py OUT.jsonl [ENVVAR OFF ON]      (default: B200MOE_GEMM_PAIR 0 1)
"""
import json
import os
import sys
import time

t_start = time.time()
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/variant_check.jsonl"
VAR, OFF, ON = (sys.argv[2:5] if len(sys.argv) >= 5 else ("B200MOE_GEMM_PAIR", "0", "1"))
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
    os.environ[VAR] = ON if pair else OFF
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
    for (E, k, H, I, M, neg) in ((8, 2, 512, 256, 1024, 0.0), (8, 2, 512, 256, 1500, 0.3), (4, 2, 1024, 512, 777, 0.0),
                                (8, 8, 4096, 256, 600, 0.5), (6, 5, 5120, 256, 515, 0.8)):
        moe = layer(E, k, H, I, M, dev, g)
        hid = (torch.randn(M, H, device=dev, generator=g) / 10).bfloat16()
        ids = torch.stack([torch.randperm(E, device=dev, generator=g)[:k] for _ in range(M)]).int() if k <= E else None
        drop = torch.rand(M, k, device=dev, generator=g) < neg
        ids = torch.where(drop, torch.full_like(ids, -1), ids).contiguous()
        w = torch.rand(M, k, device=dev, generator=g).float()
        o0, _ = run(moe, hid, ids, w, M, k, False)
        o1, _ = run(moe, hid, ids, w, M, k, True)
        # host-pointer entry point: fp32 output
        hh, ih, wh = hid.cpu(), ids.cpu(), w.cpu()
        f0, f1 = torch.empty(M, H), torch.empty(M, H)
        os.environ[VAR] = OFF
        moe.cpu_prefill(M, k, ih.data_ptr(), wh.data_ptr(), hh.data_ptr(), f0.data_ptr())
        os.environ[VAR] = ON
        moe.cpu_prefill(M, k, ih.data_ptr(), wh.data_ptr(), hh.data_ptr(), f1.data_ptr())
        emit(step="parity", var=VAR, shape=[E, k, H, I, M], dropped=neg, bit_identical=bool(torch.equal(o0, o1)),
             bit_identical_fp32=bool(torch.equal(f0, f1)), max_abs_diff=float((o0.float() - o1.float()).abs().max()),
             finite=bool(torch.isfinite(o1.float()).all()), nonzero=bool((f1 != 0).any()))
        moe.close()
    for name, (E, k, H, I, M, n_global) in (("dsv3 EP8 shard", (32, 8, 7168, 2048, 8192, 256)),
                                            ("all experts local", (32, 8, 4096, 768, 8192, 32))):
        moe = layer(E, k, H, I, M, dev, g)
        hid = (torch.randn(M, H, device=dev, generator=g) / 10).bfloat16()
        gids = torch.stack([torch.randperm(n_global, device=dev, generator=g)[:k] for _ in range(M)]).int()
        ids = torch.where(gids < E, gids, torch.full_like(gids, -1)).contiguous()
        w = torch.rand(M, k, device=dev, generator=g).float()
        rows = int((ids >= 0).sum())
        res = {}
        for pair in (False, True, False, True):
            o, ts = run(moe, hid, ids, w, M, k, pair, iters=6)
            ms = sorted(ts[1:])[2]
            res.setdefault(pair, o)
            emit(step="timing", var=VAR, layer=name, variant=pair, ms_per_layer=ms, tflops=rows * 6.0 * H * I / ms / 1e9,
                 routed_rows=rows)
        emit(step="parity_8192", layer=name, bit_identical=bool(torch.equal(res[False], res[True])))
        moe.close()
        del moe


if __name__ == "__main__":
    main()
