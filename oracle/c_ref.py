"""ctypes front end of oracle/moe_ref.c (TEST INFRASTRUCTURE / CPU BASELINE ONLY — never imported by the
product).  Build with lvllm_b200.build.build_oracle_c()."""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libmoe_ref.so")
        if not os.path.exists(path):
            from lvllm_b200.build import build_oracle_c
            path = build_oracle_c()
        os.environ.setdefault("OMP_WAIT_POLICY", "passive")   # idle workers sleep instead of spinning
        _LIB = C.CDLL(path)
        _LIB.moe_ref_num_threads.restype = C.c_int
    return _LIB


def forward_bf16(hidden, w13, w2, ids, tw):
    M, H = hidden.shape
    E, N1, _ = w13.shape
    out = torch.empty(M, H, dtype=torch.float32)
    lib().moe_ref_forward_bf16(C.c_void_p(hidden.data_ptr()), C.c_void_p(w13.data_ptr()), C.c_void_p(w2.data_ptr()),
                               C.c_void_p(ids.data_ptr()), C.c_void_p(tw.data_ptr()), C.c_void_p(out.data_ptr()),
                               M, ids.shape[1], E, H, N1 // 2)
    return out


def forward_fp8_block(hidden, w13, s13, w2, s2, ids, tw):
    M, H = hidden.shape
    E, N1, _ = w13.shape
    out = torch.empty(M, H, dtype=torch.float32)
    lib().moe_ref_forward_fp8_block(C.c_void_p(hidden.data_ptr()), C.c_void_p(w13.data_ptr()),
                                    C.c_void_p(s13.data_ptr()), C.c_void_p(w2.data_ptr()), C.c_void_p(s2.data_ptr()),
                                    C.c_void_p(ids.data_ptr()), C.c_void_p(tw.data_ptr()), C.c_void_p(out.data_ptr()),
                                    M, ids.shape[1], E, H, N1 // 2)
    return out


def forward_w4(hidden, w13, s13, w2, s2, ids, tw, fmt: str, g13=None, g2=None, exact: bool = True):
    """4-bit weight-only experts: fmt in {"int4", "nvfp4", "mxfp4"} (checkpoint layouts, see moe_ref.c).
    exact=False: group scales factored out of the inner sums (the form the CPU baseline times)."""
    M, H = hidden.shape
    E, N1, _ = w13.shape
    code = {"int4": 1, "nvfp4": 2, "mxfp4": 3}[fmt] + (0 if exact else 16)
    out = torch.empty(M, H, dtype=torch.float32)
    gp13 = C.c_void_p(g13.data_ptr()) if g13 is not None else C.c_void_p(0)
    gp2 = C.c_void_p(g2.data_ptr()) if g2 is not None else C.c_void_p(0)
    lib().moe_ref_forward_w4(C.c_void_p(hidden.data_ptr()), C.c_void_p(w13.data_ptr()), C.c_void_p(s13.data_ptr()),
                             C.c_void_p(w2.data_ptr()), C.c_void_p(s2.data_ptr()), gp13, gp2,
                             C.c_void_p(ids.data_ptr()), C.c_void_p(tw.data_ptr()), C.c_void_p(out.data_ptr()),
                             M, ids.shape[1], E, H, N1 // 2, code)
    return out


def forward_w4_batched(hidden, w13, s13, w2, s2, ids, tw, fmt: str, g13=None, g2=None):
    """Expert-major batched form (moe_ref_forward_w4_batched): tokens of an expert processed together, rows dequantised
    once, AVX-512 dot products — the form bench.py's CPU arm times at the real decode batch."""
    M, H = hidden.shape
    E, N1, _ = w13.shape
    code = {"int4": 1, "nvfp4": 2, "mxfp4": 3}[fmt]
    out = torch.empty(M, H, dtype=torch.float32)
    gp13 = C.c_void_p(g13.data_ptr()) if g13 is not None else C.c_void_p(0)
    gp2 = C.c_void_p(g2.data_ptr()) if g2 is not None else C.c_void_p(0)
    lib().moe_ref_forward_w4_batched(C.c_void_p(hidden.data_ptr()), C.c_void_p(w13.data_ptr()), C.c_void_p(s13.data_ptr()),
                                     C.c_void_p(w2.data_ptr()), C.c_void_p(s2.data_ptr()), gp13, gp2,
                                     C.c_void_p(ids.data_ptr()), C.c_void_p(tw.data_ptr()), C.c_void_p(out.data_ptr()),
                                     M, ids.shape[1], E, H, N1 // 2, code)
    return out
