"""ctypes front end of oracle/_ref/libref_moe.so — the reference's own CPU fused MoE (csrc/cpu/cpu_fused_moe.cpp),
compiled by oracle/build_ref.py.  TEST INFRASTRUCTURE / CPU BASELINE ONLY — never imported by the product."""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def cpu_flags() -> set:
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("flags"):
                return set(ln.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def available() -> bool:
    """The library exists (built here, travels with the snapshot) and this host can execute it (AVX-512 + bf16)."""
    f = cpu_flags()
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_moe.so")) and {"avx512f", "avx512bw", "avx512vl", "avx512dq", "avx512_bf16", "avx512_vnni", "f16c", "fma", "avx2"} <= f


def isa() -> str:
    return "amx" if {"amx_bf16", "amx_tile"} <= cpu_flags() else "vec"


def lib():
    global _LIB
    if _LIB is None:
        os.environ.setdefault("OMP_WAIT_POLICY", "passive")
        _LIB = C.CDLL(os.path.join(_HERE, "_ref", "libref_moe.so"))
        _LIB.ref_moe_create.restype = C.c_void_p
        _LIB.ref_moe_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_char_p]
        _LIB.ref_moe_forward.restype = C.c_int
        _LIB.ref_moe_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_char_p]
        _LIB.ref_moe_destroy.argtypes = [C.c_void_p]
        _LIB.ref_moe_last_error.restype = C.c_char_p
        _LIB.ref_mla_decode.restype = C.c_int
        _LIB.ref_mla_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_int,
                                        C.c_int, C.c_int, C.c_int]
        _LIB.ref_rms_norm.restype = C.c_int
        _LIB.ref_rms_norm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int]
        _LIB.ref_fused_add_rms_norm.restype = C.c_int
        _LIB.ref_fused_add_rms_norm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int]
        _LIB.ref_rotary_embedding.restype = C.c_int
        _LIB.ref_rotary_embedding.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        _LIB.ref_silu_and_mul.restype = C.c_int
        _LIB.ref_silu_and_mul.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    return _LIB


class RefMoe:
    """bf16 experts in checkpoint layout (w13 [E,2I,H], w2 [E,H,I]) re-packed by the reference's prepack_moe_weight."""

    def __init__(self, w13: torch.Tensor, w2: torch.Tensor, isa_name: str | None = None):
        assert w13.dtype == torch.bfloat16 and w2.dtype == torch.bfloat16 and w13.is_contiguous() and w2.is_contiguous()
        E, N1, H = w13.shape
        self.H = H
        self._h = lib().ref_moe_create(w13.data_ptr(), w2.data_ptr(), E, H, N1 // 2, (isa_name or isa()).encode())
        if not self._h:
            raise RuntimeError("reference cpu_fused_moe: " + lib().ref_moe_last_error().decode()[:500])

    def forward(self, hidden: torch.Tensor, ids: torch.Tensor, weights: torch.Tensor, act: str = "silu") -> torch.Tensor:
        assert hidden.dtype == torch.bfloat16 and ids.dtype == torch.int32 and weights.dtype == torch.float32
        M, k = ids.shape
        out = torch.empty(M, self.H, dtype=torch.bfloat16)
        rc = lib().ref_moe_forward(self._h, hidden.contiguous().data_ptr(), ids.contiguous().data_ptr(),
                                   weights.contiguous().data_ptr(), out.data_ptr(), M, k, act.encode())
        if rc:
            raise RuntimeError("reference cpu_fused_moe: " + lib().ref_moe_last_error().decode()[:500])
        return out

    def close(self):
        if self._h:
            lib().ref_moe_destroy(self._h)
            self._h = None


def mla_decode(q_nope: torch.Tensor, q_pe: torch.Tensor, kv_cache: torch.Tensor, seq_lens: torch.Tensor,
               block_tables: torch.Tensor, scale: float) -> torch.Tensor:
    """The reference's CPU paged MLA decode (csrc/cpu/mla_decode.cpp:356-383; page size 16 only): bf16 in, bf16 out."""
    assert kv_cache.shape[1] == 16 and kv_cache.shape[2] == 576
    q = torch.cat([q_nope, q_pe], dim=-1).bfloat16().contiguous()
    kv = kv_cache.bfloat16().contiguous()
    bt = block_tables.to(torch.int32).contiguous()
    sl = seq_lens.to(torch.int32).contiguous()
    B, Hq = q.shape[0], q.shape[1]
    out = torch.empty(B, Hq, 512, dtype=torch.bfloat16)
    rc = lib().ref_mla_decode(out.data_ptr(), q.data_ptr(), kv.data_ptr(), float(scale), bt.data_ptr(), sl.data_ptr(),
                              B, Hq, kv.shape[0], bt.shape[1])
    if rc:
        raise RuntimeError("reference mla_decode_kvcache: " + lib().ref_moe_last_error().decode()[:500])
    return out


_DT = {torch.bfloat16: 0, torch.float16: 1, torch.float32: 2}


def _err(what: str):
    raise RuntimeError(f"reference {what}: " + lib().ref_moe_last_error().decode()[:500])


def rms_norm(x: torch.Tensor, weight: torch.Tensor | None, eps: float) -> torch.Tensor:
    """The reference's CPU RMSNorm kernel (csrc/cpu/layernorm.cpp:99-115) on [M, H]."""
    x = x.contiguous()
    M, H = x.shape
    out = torch.empty_like(x)
    w = weight.to(x.dtype).contiguous() if weight is not None else None
    if lib().ref_rms_norm(out.data_ptr(), x.data_ptr(), w.data_ptr() if w is not None else None, float(eps), M, H, _DT[x.dtype]):
        _err("rms_norm")
    return out


def fused_add_rms_norm(x: torch.Tensor, residual: torch.Tensor, weight: torch.Tensor | None, eps: float):
    """The reference's CPU residual-add + RMSNorm kernel (csrc/cpu/layernorm.cpp:117-136); returns (y, new_residual)."""
    y, r = x.contiguous().clone(), residual.to(x.dtype).contiguous().clone()
    M, H = y.shape
    w = weight.to(x.dtype).contiguous() if weight is not None else None
    if lib().ref_fused_add_rms_norm(y.data_ptr(), r.data_ptr(), w.data_ptr() if w is not None else None, float(eps), M, H,
                                    _DT[x.dtype]):
        _err("fused_add_rms_norm")
    return y, r


def rotary_embedding(positions: torch.Tensor, query: torch.Tensor, key: torch.Tensor | None, head_size: int,
                     cos_sin_cache: torch.Tensor, is_neox: bool):
    """The reference's CPU rotary embedding (csrc/cpu/pos_encoding.cpp:332-366): query [T, Hq*head], key [T, Hk*head]."""
    pos = positions.to(torch.int64).contiguous()
    q = query.contiguous().clone()
    k = key.contiguous().clone() if key is not None else None
    cs = cos_sin_cache.to(q.dtype).contiguous()
    T = pos.numel()
    if lib().ref_rotary_embedding(pos.data_ptr(), q.data_ptr(), k.data_ptr() if k is not None else None, T,
                                  q.shape[-1] // head_size, (k.shape[-1] // head_size) if k is not None else 0, head_size,
                                  cs.data_ptr(), cs.shape[0], cs.shape[1], int(is_neox), _DT[q.dtype]):
        _err("rotary_embedding")
    return q, k


def silu_and_mul(x: torch.Tensor) -> torch.Tensor:
    """The reference's CPU silu_and_mul (csrc/cpu/activation.cpp:87-97): [M, 2d] -> [M, d]."""
    x = x.contiguous()
    M, d2 = x.shape
    out = torch.empty(M, d2 // 2, dtype=x.dtype)
    if lib().ref_silu_and_mul(out.data_ptr(), x.data_ptr(), M, d2 // 2, _DT[x.dtype]):
        _err("silu_and_mul")
    return out

