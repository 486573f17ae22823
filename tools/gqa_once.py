"""One paged-GQA decode call at the bench shape (target of ncu captures: ncu --set full -k regex:gqa_decode python tools/gqa_once.py [B S page Hq Hkv])."""
import sys, torch
sys.path.insert(0, ".")
from lvllm_b200 import ops
dev = torch.device("cuda")
B, S, page, Hq, Hkv = [int(x) for x in sys.argv[1:6]] if len(sys.argv) > 5 else (256, 512, 16, 64, 4)
npg = S // page
kc = torch.randn(B * npg, page, Hkv, 128, device=dev).bfloat16()
vc = torch.randn(B * npg, page, Hkv, 128, device=dev).bfloat16()
pt = torch.randperm(B * npg, device=dev).reshape(B, npg).int()
lens = torch.full((B,), S, device=dev, dtype=torch.int32)
q = torch.randn(B, Hq, 128, device=dev).bfloat16()
for _ in range(3):
    ops.gqa_decode(q, kc, vc, lens, pt, 0.088)
torch.cuda.synchronize()
