"""Bring-up probe of the block-scaled FP8 prefill kernel (moe_gemm_kernel MODE 2): relative error against the oracle's ue8m0
chain for (a) the promotion kernel on power-of-two scales, (b) MODE 2 with each candidate TMEM placement of the token scale
words (B200MOE_SFB_VARIANT).  Run on a B200:  python tools/e8m0_probe.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["B200MOE_FP8_E8M0"] = "1"


def main():
    import lk_moe
    from oracle import moe_oracle as O   # checker only (tools/, like tests/)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_shapes import _cfg, _fp8_e8m0_case, _rel
    for (M, E, H, I) in [(1100, 4, 512, 256), (1100, 4, 2048, 1024)]:
        k = 2
        hid, ids, w, (w13, s13, w2, s2), ref, ref32 = _fp8_e8m0_case(E, k, H, I, M, 4400 + M + H)
        moe = lk_moe.MOE_FP8(_cfg(E, k, H, I, gN=128, gK=128, max_seqs=256), w13.data_ptr(), w2.data_ptr(), s13.data_ptr(),
                             s2.data_ptr(), 0, 0)
        for name, env in (("promotion kernel", {"B200MOE_E8M0_PROMO": "1"}), ("MODE 2 sfb_variant 0", {"B200MOE_SFB_VARIANT": "0"}),
                          ("MODE 2 sfb_variant 1", {"B200MOE_SFB_VARIANT": "1"})):
            for kk in ("B200MOE_E8M0_PROMO", "B200MOE_SFB_VARIANT"):
                os.environ.pop(kk, None)
            os.environ.update(env)
            out = torch.empty(M, H, dtype=torch.float32)
            moe.cpu_prefill(M, k, ids.data_ptr(), w.data_ptr(), hid.data_ptr(), out.data_ptr())
            print(f"M={M} E={E} H={H} I={I} {name}: rel vs ue8m0 oracle {_rel(out, ref):.3e}  vs fp32-scale oracle {_rel(out, ref32):.3e}",
                  flush=True)
        moe.close()


if __name__ == "__main__":
    main()
