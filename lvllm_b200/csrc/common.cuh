// Common device helpers for the sm_100a kernels: mbarrier, 1-D bulk async copies (UBLKCP), tcgen05
// (UMMA + TMEM) wrappers and small utilities.  Hand-written inline PTX; no CUTLASS dependency.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

#define B200_DEVICE __device__ __forceinline__

B200_DEVICE uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

B200_DEVICE bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t.reg .b32 R;\n\t"
      "elect.sync R|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------- mbarrier
B200_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
B200_DEVICE void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
B200_DEVICE void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

B200_DEVICE void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
B200_DEVICE void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
B200_DEVICE bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
B200_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ----------------------------------------------------------------------------- bulk copy (TMA 1-D)
// global -> shared, completion signalled on an mbarrier as transaction bytes.  size % 16 == 0,
// both addresses 16-B aligned.  SASS: UBLKCP.
B200_DEVICE void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// same with an L2 eviction-priority hint (created with createpolicy)
B200_DEVICE void bulk_g2s_hint(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
          "r"(smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}
// 2-D tensor-map copy (TMA tile mode, SASS: UTMALDG): box at element coordinate (c0, c1) -> shared memory, completion
// as transaction bytes on an mbarrier.  `tmap` points at a CUtensorMap in the kernel parameter space (__grid_constant__).
B200_DEVICE void tma_load_2d_hint(void* smem_dst, const void* tmap, int c0, int c1, uint64_t* bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}
B200_DEVICE uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
B200_DEVICE uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}

// ----------------------------------------------------------------------------- cp.async (LDGSTS)
// 16-byte global -> shared copy without register staging; src_size 0 zero-fills the destination.
B200_DEVICE void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t sz = valid ? 16u : 0u;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz)
               : "memory");
}
// 8-byte variant (LDGSTS.64): used to spread packed 4-bit data over 16-byte chunks
B200_DEVICE void cp_async8(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
// the mbarrier receives one (pre-counted) arrival when all prior cp.async of this thread have landed
B200_DEVICE void cp_async_mbar_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ----------------------------------------------------------------------------- tcgen05 / TMEM
B200_DEVICE void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // one full warp, ncols pow2 >= 32
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
B200_DEVICE void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
B200_DEVICE void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
B200_DEVICE void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// MMA completion -> mbarrier arrive (implies fence::before_thread_sync)
B200_DEVICE void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// Shared-memory matrix descriptor, K-major operand with 128-byte swizzle.  Rows are 128 B, groups of
// 8 rows form a 1024-B swizzle atom; SBO = distance between consecutive 8-row groups.
// Bit layout (sm_100): [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1,
// [49,52) base offset, [61,64) layout (2 = SWIZZLE_128B).
B200_DEVICE uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Instruction descriptor (upper 32 bits of the 64-bit idesc operand).  fmt: kind::f16 -> 0 f16, 1 bf16;
// kind::f8f6f4 -> 0 e4m3, 1 e5m2.  D is fp32, both operands K-major, dense.
B200_DEVICE uint32_t umma_idesc(uint32_t a_fmt, uint32_t b_fmt, uint32_t M, uint32_t N) {
  uint32_t d = 0;
  d |= 1u << 4;               // c_format = F32
  d |= (a_fmt & 7u) << 7;     // a_format
  d |= (b_fmt & 7u) << 10;    // b_format
  d |= ((N >> 3) & 63u) << 17;  // n_dim
  d |= ((M >> 4) & 31u) << 24;  // m_dim
  return d;
}

B200_DEVICE void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
B200_DEVICE void umma_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// kind::mxf8f6f4.block_scale: instruction descriptor (CUTLASS InstrDescriptorBlockScaled: a_format / b_format 0 = E4M3,
// 5 = E2M1; ue8m0 scales; M = 128; a_sf_id / b_sf_id are OR-ed in per instruction at bits 29 / 4) and the MMA itself:
// A / B from shared-memory descriptors, D and both scale-factor operands in TMEM
B200_DEVICE uint32_t umma_idesc_mx(uint32_t a_fmt, uint32_t b_fmt, uint32_t n) {
  uint32_t d = 0;
  d |= a_fmt << 7;
  d |= b_fmt << 10;
  d |= ((n >> 3) & 63u) << 17;     // n_dim
  d |= 1u << 23;                   // scale_format = E8M0
  d |= ((128u >> 4) & 31u) << 24;  // m_dim
  return d;
}
B200_DEVICE void umma_mx(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                         uint32_t tmem_sfa, uint32_t tmem_sfb) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
      : "memory");
}

// per-token-group FP8 activation scale (reference fp8_utils.py:100-113): absmax / 448, rounded UP to a power of two in
// ue8m0 mode (exp2(ceil(log2(.))), the DeepGEMM-on-Blackwell convention)
B200_DEVICE float fp8_group_scale(float absmax, int e8m0) {
  float sc = fmaxf(absmax, 1e-10f) / 448.0f;
  if (e8m0) {
    const uint32_t b = __float_as_uint(sc);
    sc = __uint_as_float(((b >> 23) + ((b & 0x7FFFFFu) ? 1u : 0u)) << 23);
  }
  return sc;
}

// TMEM -> registers: 32 lanes x 16 consecutive fp32 columns (warp w reads lanes 32*(w%4)..+31).
B200_DEVICE void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 8 consecutive fp32 columns
B200_DEVICE void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
B200_DEVICE void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// registers -> TMEM: each thread writes 16 consecutive 32-bit columns of its own lane
B200_DEVICE void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// the same 32-bit word into 4 consecutive columns of the thread's lane
B200_DEVICE void tmem_st4(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %1, %1, %1};" ::"r"(taddr), "r"(v) : "memory");
}
// four words into 4 consecutive columns of the thread's lane
B200_DEVICE void tmem_st4v(uint32_t taddr, uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v0), "r"(v1), "r"(v2), "r"(v3)
               : "memory");
}
B200_DEVICE void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// A operand from TMEM (lane = row, two 16-bit K elements per 32-bit column), B from shared memory
B200_DEVICE void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Guard-predicated forms: executed by every lane of a converged warp with warp-uniform operands, issued by the
// lane(s) whose `issue` flag is set.  Keeping the call site in uniform control flow lets ptxas hold descriptors
// and TMEM addresses in uniform registers (one UTCHMMA per MMA instead of an ELECT / R2UR.BROADCAST loop).
B200_DEVICE void umma_f16_ts_if(uint32_t issue, uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(issue)
      : "memory");
}
B200_DEVICE void umma_f16_if(uint32_t issue, uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(issue)
      : "memory");
}
B200_DEVICE void umma_f8_if(uint32_t issue, uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(issue)
      : "memory");
}
B200_DEVICE void umma_commit_if(uint32_t issue, uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(issue)
      : "memory");
}

// ----------------------------------------------------------------------------- misc
B200_DEVICE float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
B200_DEVICE float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

B200_DEVICE int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
B200_DEVICE void red_release_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Byte offset of logical element (row r, byte b of a 128-B K segment) inside a [rows x 128 B] K-major
// tile stored with the 128-byte swizzle (16-B chunk index XOR row%8), 8-row groups of 1024 B.
__host__ __device__ inline uint32_t sw128_offset(uint32_t r, uint32_t b) {
  return (r >> 3) * 1024u + (r & 7u) * 128u + ((((b >> 4) ^ (r & 7u)) & 7u) << 4) + (b & 15u);
}

}  // namespace b200
