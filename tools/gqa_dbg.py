import sys, torch
sys.path.insert(0, ".")
from lvllm_b200 import ops
dev = torch.device("cuda")
torch.manual_seed(0)
shapes = [(64, 2048, 16, 32, 8), (256, 512, 16, 64, 4), (1, 8192, 16, 64, 4), (7, 1300, 16, 64, 4), (128, 1024, 16, 32, 8)]
data = []
for (B, S, page, Hq, Hkv) in shapes:
    npg = -(-S // page)
    kc = torch.randn(B * npg, page, Hkv, 128, device=dev).bfloat16()
    vc = torch.randn(B * npg, page, Hkv, 128, device=dev).bfloat16()
    pt = torch.randperm(B * npg, device=dev).reshape(B, npg).int()
    lens = torch.randint(1, S + 1, (B,), device=dev, dtype=torch.int32)
    q = torch.randn(B, Hq, 128, device=dev).bfloat16()
    data.append((q, kc, vc, lens, pt))
ref = [None] * len(data)
for it in range(int(sys.argv[1])):
    for i, d in enumerate(data):
        o, _ = ops.gqa_decode(*d, 0.088)
        if ref[i] is None:
            ref[i] = o.clone()
        elif not torch.equal(ref[i], o):
            print("MISMATCH", it, i, (ref[i].float() - o.float()).abs().max().item(), flush=True)
    if it % 50 == 0:
        torch.cuda.synchronize()
        print("iter", it, flush=True)
torch.cuda.synchronize()
print("done", flush=True)
