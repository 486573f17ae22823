"""Paged MLA decode: this repo's kernel next to the Blackwell CUTLASS MLA shipped in the image's vLLM wheel
(`torch.ops._C.sm100_cutlass_mla_decode`, the kernel SURVEY.md 2b names as the one to beat; call sequence of the
reference's tests/kernels/attention/test_cutlass_mla_decode.py:112-148).  Same box, same shapes, CUDA-event timed
(median of 20 after 5 warm-ups, a 256 MB L2 flush between iterations).  Prints one JSON line per shape.
    python tools/mla_vs_ref.py
"""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timeit(fn, flush):
    ts = []
    for i in range(25):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= 5:
            ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    from lvllm_b200 import ops
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    try:
        import vllm._custom_ops as vops   # noqa: F401
        have_ref = hasattr(torch.ops._C, "sm100_cutlass_mla_decode")
        why = None if have_ref else "torch.ops._C.sm100_cutlass_mla_decode not registered"
    except Exception as ex:  # the wheel's compiled ops may not load on every box
        have_ref, why = False, repr(ex)
    g = torch.Generator(device=dev).manual_seed(42)
    for (B, S, page) in [(1, 4096, 64), (64, 2048, 64), (128, 4096, 64)]:
        for kv_dtype in (torch.bfloat16, torch.float8_e4m3fn):
            Hq, d, dv = 128, 576, 512
            npg = -(-S // page)
            cache = torch.randn(B * npg, page, d, device=dev, dtype=torch.bfloat16, generator=g).to(kv_dtype)
            q = torch.randn(B, Hq, d, device=dev, dtype=torch.bfloat16, generator=g).to(kv_dtype)
            qn, qp = q[..., :dv].contiguous(), q[..., dv:].contiguous()
            lens = torch.full((B,), S, dtype=torch.int32, device=dev)
            pt = torch.arange(B * npg, dtype=torch.int32, device=dev).view(B, npg)
            scale = 1.0 / math.sqrt(d)
            ours = timeit(lambda: ops.mla_decode(qn, qp, cache, lens, pt, scale, max_seq_len=S), flush)
            out_o, lse_o = ops.mla_decode(qn, qp, cache, lens, pt, scale, max_seq_len=S)
            line = {"B": B, "S": S, "page": page, "kv_dtype": str(kv_dtype).split(".")[-1], "b200_mla_decode_us": ours,
                    "kv_bytes": B * S * d * cache.element_size(),
                    "b200_kv_gbs": B * S * d * cache.element_size() / ours / 1e3}
            if have_ref:
                try:
                    sm = torch.cuda.get_device_properties(dev).multi_processor_count
                    wsz = vops.sm100_cutlass_mla_get_workspace_size(S * page, B, sm, num_kv_splits=1)
                    ws = torch.empty(wsz, device=dev, dtype=torch.uint8)
                    out = torch.empty(B, Hq, dv, dtype=torch.bfloat16, device=dev)
                    lse = torch.empty(B, Hq, dtype=torch.float32, device=dev)
                    ref_fn = lambda: vops.sm100_cutlass_mla_decode(out, lse, qn, qp, cache, lens, pt, ws, scale, 1)
                    line["vllm_sm100_cutlass_mla_us"] = timeit(ref_fn, flush)
                    ref_fn()
                    torch.cuda.synchronize()
                    a, b = out.double().flatten(), out_o.double().flatten()
                    line["cos_diff_vs_vllm"] = float(1 - 2 * (a * b).sum() / max((a * a + b * b).sum(), 1e-12))
                    line["speedup_vs_vllm"] = line["vllm_sm100_cutlass_mla_us"] / ours
                except Exception as ex:
                    line["vllm_error"] = repr(ex)[:300]
            else:
                line["vllm_unavailable"] = why[:300]
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
