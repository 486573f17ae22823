"""Golden fixture for oracle.rms_norm / fused_add_rms_norm: the reference's OWN native ops.

Run in the build container only (needs /root/reference):
    cd /tmp && PYTHONPATH=/root/reference python /root/repo/tests/golden/make_golden_norm.py

vllm.ir.ops.layernorm.rms_norm / fused_add_rms_norm (reference vllm/ir/ops/layernorm.py:10-21, :44-63 — what
RMSNorm.forward_native dispatches to, layers/layernorm.py:74-94) are called on seeded CPU inputs; inputs and outputs go to
tests/golden/golden_norm.pt.  Nothing here is imported by the product.
"""
import os
import sys

import torch

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_norm.pt")


def main():
    sys.path.insert(0, "/root/reference")
    from vllm.ir.ops import layernorm as LN
    gen = torch.Generator().manual_seed(0)
    cases = []
    for dt in (torch.bfloat16, torch.float16, torch.float32):
        for (M, H, use_w) in ((1, 4096, True), (5, 512, True), (3, 1024, False)):
            x = (torch.randn(M, H, generator=gen) * 3).to(dt)
            r = torch.randn(M, H, generator=gen).to(dt)
            w = (torch.rand(H, generator=gen) + 0.5).to(dt) if use_w else None
            y = LN.rms_norm(x.clone(), w, 1e-6)
            y2, r2 = LN.fused_add_rms_norm(x.clone(), r.clone(), w, 1e-6)
            cases.append(dict(x=x, residual=r, weight=w, eps=1e-6, rms_norm=y, fused_y=y2, fused_residual=r2))
    torch.save(dict(cases=cases), OUT)
    print("wrote", OUT, len(cases), "cases")


if __name__ == "__main__":
    main()
