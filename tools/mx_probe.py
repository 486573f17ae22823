"""Driver of tools/mx_probe.cu (round-2 groundwork): runs one block-scaled tcgen05 MMA tile under every operand-layout
hypothesis and reports which one reproduces the CPU result.  Needs a B200:  gpurun -- python tools/mx_probe.py"""
import ctypes as C
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tools", "libmx_probe.so")
E2M1 = torch.tensor([0, .5, 1, 1.5, 2, 3, 4, 6, -0., -.5, -1, -1.5, -2, -3, -4, -6])


def build():
    src = os.path.join(ROOT, "tools", "mx_probe.cu")
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-shared", "-Xcompiler",
                               "-fPIC", "-I", os.path.join(ROOT, "lvllm_b200", "csrc"), "-o", SO, src])
    return C.CDLL(SO)


def main():
    lib = build()
    if not torch.cuda.is_available():
        print("built", SO, "(no GPU here: nothing run)")
        return
    g = torch.Generator().manual_seed(0)
    a_packed = torch.randint(0, 256, (128, 64), dtype=torch.uint8, generator=g)
    sa = torch.randint(124, 131, (128, 4), dtype=torch.uint8, generator=g)       # 2^-3 .. 2^3
    b = (torch.randn(32, 128, generator=g)).to(torch.float8_e4m3fn)
    sb = torch.randint(124, 131, (32, 4), dtype=torch.uint8, generator=g)
    lo, hi = (a_packed & 15).long(), (a_packed >> 4).long()
    a = E2M1[torch.stack([lo, hi], -1).reshape(128, 128)]
    a = a * torch.pow(2.0, sa.float() - 127).repeat_interleave(32, 1)
    bf = b.float() * torch.pow(2.0, sb.float() - 127).repeat_interleave(32, 1)
    ref = a.double() @ bf.double().t()
    dev = torch.device("cuda")
    d = [t.to(dev) for t in (a_packed, sa, b.view(torch.uint8), sb)]
    for av in (0, 1, 2):
        for sv in (0, 1):
            out = torch.full((128, 32), float("nan"), device=dev)
            rc = lib.mx_probe_run(C.c_void_p(d[0].data_ptr()), C.c_void_p(d[1].data_ptr()), C.c_void_p(d[2].data_ptr()),
                                  C.c_void_p(d[3].data_ptr()), C.c_void_p(out.data_ptr()), av, sv, C.c_void_p(0))
            try:
                torch.cuda.synchronize()
            except RuntimeError as ex:
                print(f"a_variant {av} sf_variant {sv}: launch rc {rc}, CUDA error {ex}")
                sys.exit(1)
            err = (out.cpu().double() - ref).abs().max().item()
            print(f"a_variant {av} sf_variant {sv}: max abs err {err:.4e} (ref absmax {ref.abs().max():.3e}) "
                  f"{'MATCH' if err < 1e-2 * ref.abs().max() else ''}", flush=True)


if __name__ == "__main__":
    main()
