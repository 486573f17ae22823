"""The plain-C/OpenMP port (oracle/moe_ref.c — the CPU baseline bench.py times) against the numpy/torch oracle on
the same seeded inputs: bf16, block-FP8 (weight-only) and the three 4-bit checkpoint formats."""
import pytest
import torch

from oracle import c_ref
from oracle import moe_oracle as O

E, H, I, M, K = 4, 256, 128, 3, 2


@pytest.fixture(scope="module")
def case():
    g = torch.Generator().manual_seed(0)
    w13 = torch.randn(E, 2 * I, H, generator=g) / 10
    w2 = torch.randn(E, H, I, generator=g) / 10
    hid = (torch.randn(M, H, generator=g) / 10).bfloat16()
    ids = torch.tensor([[0, 3], [1, -1], [2, 0]], dtype=torch.int32)   # one skipped slot (id < 0)
    tw = torch.rand(M, K, generator=g)
    return w13, w2, hid, ids, tw


def _check(out, ref):
    # bf16-rounded intermediates on both sides; accumulation order differs
    torch.testing.assert_close(out, ref, atol=1e-3, rtol=2e-2)


def test_c_port_bf16(case):
    w13, w2, hid, ids, tw = case
    ref = O.experts_forward(hid, O.DequantExperts(w13.bfloat16().float(), w2.bfloat16().float()), ids, tw)
    out = c_ref.forward_bf16(hid, w13.bfloat16().contiguous(), w2.bfloat16().contiguous(), ids, tw)
    _check(out, ref)


def test_c_port_fp8_block(case):
    w13, w2, hid, ids, tw = case
    q13, s13 = O.quant_fp8_block(w13)
    q2, s2 = O.quant_fp8_block(w2)
    ref = O.experts_forward(hid, O.DequantExperts(O.dequant_fp8_block(q13, s13), O.dequant_fp8_block(q2, s2)), ids, tw)
    out = c_ref.forward_fp8_block(hid, q13.contiguous(), s13.contiguous(), q2.contiguous(), s2.contiguous(), ids, tw)
    torch.testing.assert_close(out, ref, atol=2e-3, rtol=3e-2)   # the port keeps fp32 weights (no bf16 rounding)


@pytest.mark.parametrize("fmt", ["int4", "nvfp4", "mxfp4"])
def test_c_port_w4(case, fmt):
    w13, w2, hid, ids, tw = case
    g13 = g2 = None
    if fmt == "int4":
        p13, s13 = O.quant_int4_group(w13)
        p2, s2 = O.quant_int4_group(w2)
        d13, d2 = O.dequant_int4_group(p13, s13, 32), O.dequant_int4_group(p2, s2, 32)
    elif fmt == "nvfp4":
        pa, sa, ga = O.quant_nvfp4(w13[:, :I])
        pb, sb, gb = O.quant_nvfp4(w13[:, I:])
        p13, s13 = torch.cat([pa, pb], 1), torch.cat([sa, sb], 1)
        g13 = torch.stack([ga, gb], 1).contiguous()
        p2, s2, g2 = O.quant_nvfp4(w2)
        d13 = torch.cat([O.dequant_nvfp4(pa, sa, ga), O.dequant_nvfp4(pb, sb, gb)], 1)
        d2 = O.dequant_nvfp4(p2, s2, g2)
    else:
        p13, s13 = O.quant_mxfp4(w13)
        p2, s2 = O.quant_mxfp4(w2)
        d13, d2 = O.dequant_mxfp4(p13, s13), O.dequant_mxfp4(p2, s2)
    ref = O.experts_forward(hid, O.DequantExperts(d13, d2), ids, tw)
    args = (hid, p13.contiguous(), s13.contiguous(), p2.contiguous(), s2.contiguous(), ids, tw, fmt, g13, g2)
    _check(c_ref.forward_w4(*args), ref)
    # the throughput form the CPU baseline times (group scale factored out): same result up to rounding
    torch.testing.assert_close(c_ref.forward_w4(*args, exact=False), ref, atol=2e-3, rtol=3e-2)
    # the expert-major batched form bench.py's CPU arm times at the real decode batch
    torch.testing.assert_close(c_ref.forward_w4_batched(*args), ref, atol=2e-3, rtol=3e-2)


def test_oracle_matches_compiled_reference_cpu_moe():
    """The restated oracle (bf16 experts) against the REFERENCE's own compiled CPU fused MoE
    (csrc/cpu/cpu_fused_moe.cpp, built by oracle/build_ref.py into oracle/_ref/): both ISA paths the host supports.
    Output of the reference kernel is bf16; tolerance = bf16 rounding of the result."""
    from oracle import build_ref, ref_moe
    build_ref.build()          # no-op unless /root/reference is present and the library is stale / missing
    if not ref_moe.available():
        pytest.skip("oracle/_ref/libref_moe.so absent or host CPU without AVX-512 bf16")
    g = torch.Generator().manual_seed(0)
    E_, H_, I_, M_, k_ = 8, 512, 256, 9, 2
    w13 = (torch.randn(E_, 2 * I_, H_, generator=g) / 10).bfloat16()
    w2 = (torch.randn(E_, H_, I_, generator=g) / 10).bfloat16()
    hid = (torch.randn(M_, H_, generator=g) / 10).bfloat16()
    tw, ids = torch.topk(torch.softmax(torch.randn(M_, E_, generator=g), -1), k_)
    ids, tw = ids.int().contiguous(), tw.float().contiguous()
    ref = O.experts_forward_batched(hid, O.DequantExperts(w13.float(), w2.float()), ids, tw)
    isas = ["vec"] + (["amx"] if ref_moe.isa() == "amx" else [])
    for isa in isas:
        m = ref_moe.RefMoe(w13, w2, isa)
        out = m.forward(hid, ids, tw).float()
        m.close()
        torch.testing.assert_close(out, ref, atol=1e-3, rtol=2e-2)
        # and the plain-C port against the same compiled reference
        port = c_ref.forward_bf16(hid, w13, w2, ids, tw)
        torch.testing.assert_close(port, out, atol=1e-3, rtol=2e-2)


def test_mla_oracle_matches_compiled_reference_cpu_mla():
    """oracle.mla_decode against the REFERENCE's own compiled CPU paged MLA decode (csrc/cpu/mla_decode.cpp:356-383,
    page size 16) on shuffled pages and ragged lengths; metric / threshold of the reference's MLA test
    (tests/kernels/attention/test_cutlass_mla_decode.py:15-32: cos_diff < 1e-5)."""
    import math
    from oracle import build_ref, ref_moe
    build_ref.build()
    if not ref_moe.available():
        pytest.skip("oracle/_ref/libref_moe.so absent or host CPU without AVX-512 bf16")
    g = torch.Generator().manual_seed(42)
    for (B, S, Hq) in [(3, 300, 16), (2, 1, 128), (1, 1000, 128)]:
        page = 16
        lens = torch.tensor([max(1, S - 37 * b) for b in range(B)], dtype=torch.int32)
        npg = -(-S // page)
        cache = torch.randn(B * npg + 3, page, 576, generator=g).bfloat16()
        pt = torch.randperm(B * npg + 3, generator=g)[:B * npg].reshape(B, npg).int()
        qn = torch.randn(B, Hq, 512, generator=g).bfloat16()
        qp = torch.randn(B, Hq, 64, generator=g).bfloat16()
        scale = 1 / math.sqrt(576)
        ref, _ = O.mla_decode(qn, qp, cache, lens, pt, scale)
        out = ref_moe.mla_decode(qn, qp, cache, lens, pt, scale).float()
        a, b = out.double().flatten(), ref.double().flatten()
        assert float(1 - 2 * (a * b).sum() / (a * a + b * b).sum()) < 1e-5


def test_swigluoai_interleaved_matches_compiled_reference():
    """activation_type 1 is ambiguous at the lk_moe call site (SURVEY §8 notes): the interleaved GPT-OSS form
    (gate = even rows of w13, up = odd rows, clamp 7, alpha 1.702) of the oracle against the reference's compiled CPU
    fused MoE run with act="swigluoai"."""
    from oracle import build_ref, ref_moe
    build_ref.build()
    if not ref_moe.available():
        pytest.skip("oracle/_ref/libref_moe.so absent or host CPU without AVX-512 bf16")
    g = torch.Generator().manual_seed(5)
    E_, H_, I_, M_, k_ = 4, 256, 128, 7, 2
    w13 = (torch.randn(E_, 2 * I_, H_, generator=g) / 4).bfloat16()
    w2 = (torch.randn(E_, H_, I_, generator=g) / 10).bfloat16()
    hid = torch.randn(M_, H_, generator=g).bfloat16()          # large enough to exercise the clamps
    tw, ids = torch.topk(torch.softmax(torch.randn(M_, E_, generator=g), -1), k_)
    ids, tw = ids.int().contiguous(), tw.float().contiguous()
    x = hid.float()
    ref = torch.zeros(M_, H_)
    for t in range(M_):
        for j in range(k_):
            e = int(ids[t, j])
            a = O.apply_activation(w13[e].float() @ x[t], O.ACT_SWIGLUOAI, True, interleaved=True)
            ref[t] += float(tw[t, j]) * (w2[e].float() @ a.bfloat16().float())
    m = ref_moe.RefMoe(w13, w2, "vec")
    out = m.forward(hid, ids, tw, act="swigluoai").float()
    m.close()
    torch.testing.assert_close(out, ref, atol=3e-2, rtol=3e-2)   # bf16 output + the kernel's fast exp


def _need_ref():
    from oracle import build_ref, ref_moe
    build_ref.build()
    if not ref_moe.available():
        pytest.skip("oracle/_ref/libref_moe.so absent or host CPU without AVX-512 bf16")
    return ref_moe


_ULP = {torch.bfloat16: 2.0 ** -7, torch.float16: 2.0 ** -10, torch.float32: 2.0 ** -22}


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16, torch.float32])
def test_rmsnorm_oracle_matches_compiled_reference(dt):
    """oracle.rms_norm / fused_add_rms_norm (restating vllm/ir/ops/layernorm.py:10-21, 44-63, what RMSNorm.forward_native
    runs) against the REFERENCE's compiled CPU kernels (csrc/cpu/layernorm.cpp:99-136).  The new residual is bit exact;
    the normalised output agrees to one rounding of the output dtype (the C++ kernel multiplies by the weight in fp32 and
    rounds once, the Python form rounds before and after the weight)."""
    ref_moe = _need_ref()
    g = torch.Generator().manual_seed(11)
    for M_, H_ in ((1, 7168), (5, 512), (33, 4096)):
        x = (torch.randn(M_, H_, generator=g) * 3).to(dt)
        r = torch.randn(M_, H_, generator=g).to(dt)
        w = (torch.rand(H_, generator=g) + 0.5).to(dt)
        tol = dict(rtol=_ULP[dt], atol=1e-6)
        torch.testing.assert_close(O.rms_norm(x, w, 1e-6).float(), ref_moe.rms_norm(x, w, 1e-6).float(), **tol)
        torch.testing.assert_close(O.rms_norm(x, None, 1e-6).float(), ref_moe.rms_norm(x, None, 1e-6).float(), **tol)
        y, nr = O.fused_add_rms_norm(x, r, w, 1e-6)
        y_ref, nr_ref = ref_moe.fused_add_rms_norm(x, r, w, 1e-6)
        assert torch.equal(nr, nr_ref)
        # the C++ kernel (like the reference's CUDA one) normalises the ROUNDED sum it has just stored as the new residual
        # (layernorm.cpp:83-90), the Python form the unrounded fp32 sum: one rounding apart on the input, one on the output
        torch.testing.assert_close(O.rms_norm(nr, w, 1e-6).float(), y_ref.float(), **tol)
        torch.testing.assert_close(y.float(), y_ref.float(), rtol=3 * _ULP[dt], atol=1e-6)
        # the form the fused EP kernels compute (fp32 MoE sum, optional residual, weight or scalar gain) is the same
        # function whenever the fp32 sum is representable in the activation dtype
        if dt != torch.float32:
            y2, nr2, s2 = O.moe_sum_add_rms_norm(x.float(), r, w, 1.0, 1e-6, dt)
            assert torch.equal(y2, y) and torch.equal(nr2, nr) and torch.equal(s2, x.float() + r.float())
            y3, _, _ = O.moe_sum_add_rms_norm(x.float(), None, None, 0.1, 1e-6, dt)
            torch.testing.assert_close(y3.float(), O.rms_norm(x, None, 1e-6).float() * 0.1, rtol=2 * _ULP[dt], atol=1e-6)


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16, torch.float32])
def test_rope_oracle_matches_compiled_reference(dt):
    """oracle.rope_forward_static (restating RotaryEmbedding.forward_static, bit-pinned to the Python reference by the golden
    `rope`) against the REFERENCE's compiled CPU rotary embedding (csrc/cpu/pos_encoding.cpp:332-366), NeoX and GPT-J (the
    DeepSeek MLA rope head) styles: agreement to one rounding of the dtype (the C++ kernel computes in fp32)."""
    ref_moe = _need_ref()
    g = torch.Generator().manual_seed(12)
    T, Hq, Hk, hs, max_pos = 9, 16, 1, 64, 4096
    inv = 1.0 / (10000 ** (torch.arange(0, hs, 2).float() / hs))
    fr = torch.einsum("i,j->ij", torch.arange(max_pos).float(), inv)
    cs = torch.cat([fr.cos(), fr.sin()], -1).to(dt)
    pos = torch.randint(0, max_pos, (T,), generator=g)
    q = torch.randn(T, Hq * hs, generator=g).to(dt)
    k = torch.randn(T, Hk * hs, generator=g).to(dt)
    for neox in (True, False):
        q_ref, k_ref = ref_moe.rotary_embedding(pos, q, k, hs, cs, neox)
        q_o, k_o = O.rope_forward_static(pos, q.clone(), k.clone(), hs, hs, cs, neox)
        # |x| <= ~4.5 here and every output is a two-term sum of products: two roundings of the dtype at that magnitude
        atol = 2 * 8 * _ULP[dt]
        torch.testing.assert_close(q_o.float(), q_ref.float(), rtol=0, atol=atol)
        torch.testing.assert_close(k_o.float(), k_ref.float(), rtol=0, atol=atol)


def test_silu_and_mul_oracle_matches_compiled_reference():
    """oracle.apply_activation(SiLU, packed halves: gate = first half, up = second half) against the REFERENCE's compiled
    CPU silu_and_mul (csrc/cpu/activation.cpp:87-97)."""
    ref_moe = _need_ref()
    g = torch.Generator().manual_seed(13)
    for dt in (torch.bfloat16, torch.float16):
        h = (torch.randn(7, 2 * 384, generator=g) * 2).to(dt)
        ref = ref_moe.silu_and_mul(h)
        out = O.apply_activation(h.float(), O.ACT_SILU).to(dt)
        torch.testing.assert_close(out.float(), ref.float(), rtol=_ULP[dt], atol=1e-6)
