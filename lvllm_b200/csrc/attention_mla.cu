// Paged MLA decode (absorbed form) on tcgen05 tensor cores.
//
//   S[128 heads, 128 tokens] = Q[128, 576] * K^T        (9 k-blocks of 64 dims, kind::f16, accumulator in TMEM)
//   P = exp2(S*scale - m) (bf16, shared memory), l = sum P
//   O[128 heads, 256 dims]   = P[128, 128 tokens] * V[128 tokens, 256]     (V = first 512 latent dims; this
//                                                                          CTA owns one half of them)
//
// One CTA = (request, 128-token KV split, half of the 512 value dims).  The KV tile is loaded ONCE into
// shared memory as nine [128 tokens x 128 B] blocks with the 128-byte swizzle: block kb is the K-major B
// operand of the QK^T MMA for dims [64kb, 64kb+64) and — the same bytes, read through an MN-major
// descriptor — the B operand of the PV MMA for value dims [64kb, 64kb+64).  128 tokens per CTA means the
// whole split's scores sit in TMEM (128 columns) next to O (256 columns): softmax needs no online rescale.
// Partial (O, m, l) per split are merged by decode_merge_kernel (attention.cu).
//
// Warp roles (288 threads): warps 0-3 softmax + epilogue (TMEM lane = head), warps 4-7 loaders
// (16-byte gathers of q rows / paged KV rows -> swizzled smem, generic->async proxy fence), warp 8 MMA issuer.
// The op contract is reference csrc/libtorch_stable/attention/mla/sm100_cutlass_mla_kernel.cu:225-262.
// Roofline: KV bytes (HBM / L2) — at B=1 the kernel is latency bound; the tensor cores remove the
// 1.1 GFLOP/layer of CUDA-core FMAs the v1 kernel spent 200 us on.
#include <math_constants.h>

#include "common.cuh"
#include "moe_internal.cuh"

namespace b200 {

constexpr int MLA_TOK = 128;            // tokens per CTA
constexpr int MLA_KB = 9;               // 576 / 64
constexpr int MLA_THREADS = 288;
constexpr int MLA_QSTAGES = 3;
constexpr int MLA_K_AREA = MLA_KB * TILE_BYTES;          // 144 KB
constexpr int MLA_Q_RING = MLA_QSTAGES * TILE_BYTES;     // 32 KB
constexpr int MLA_P_TILE = 2 * TILE_BYTES;               // 128 heads x 128 tokens bf16 = 2 k-blocks
constexpr int MLA_SMEM = MLA_K_AREA + MLA_Q_RING + MLA_P_TILE + 1024 + 1024;

struct MlaBars {
  uint64_t full[MLA_KB], empty[MLA_QSTAGES];
  uint64_t s_full, p_full, o_full;
  uint32_t tmem_base;
};

B200_DEVICE void mla_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}

// instruction descriptor with an MN-major B operand (bit 16)
B200_DEVICE uint32_t umma_idesc_bmn(uint32_t fmt, uint32_t M, uint32_t N) { return umma_idesc(fmt, fmt, M, N) | (1u << 16); }

// MN-major, 128B-swizzled operand: atoms of 64 elements (128 B) along MN x 8 rows along K.
// LBO = byte distance between 64-element atoms along MN, SBO = byte distance between 8-row groups along K.
B200_DEVICE uint64_t umma_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// eight e4m3 bytes -> eight bf16 (exact: e4m3 has 4 exponent / 3 mantissa bits)
B200_DEVICE uint4 fp8x8_to_bf16x8(uint2 v) {
  const uint32_t w[2] = {v.x, v.y};
  uint32_t o[4];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const __half2_raw hr = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)((w[i] >> (16 * h)) & 0xFFFFu), __NV_E4M3);
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hr));
      const __nv_bfloat162 b = __floats2bfloat162_rn(f.x, f.y);
      o[i * 2 + h] = *reinterpret_cast<const uint32_t*>(&b);
    }
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}

// KV8: the latent cache is e4m3 (kv_cache_dtype "fp8": 576 bytes per token, half the HBM traffic of the bf16 cache;
// reference fp8 dtypes backends/mla/cutlass_mla.py:44-45).  The loaders widen it to bf16 on the way into shared memory
// (exact), so the tensor-core pipeline and its numerics are those of the bf16 kernel — P is NOT re-quantised to fp8.
// q8: the queries are e4m3 too (the reference's fp8 test feeds fp8 q, tests/kernels/attention/test_cutlass_mla_decode.py:101-110).
template <bool KV8>
__global__ void __launch_bounds__(MLA_THREADS, 1)
    mla_decode_tc_kernel(const void* __restrict__ q_nope_v, const void* __restrict__ q_pe_v,
                         const void* __restrict__ kv_v, const int32_t* __restrict__ seq_lens,
                         const int32_t* __restrict__ page_table, int Hq, int page_size, int max_pages,
                         float scale_log2, int num_splits, float* __restrict__ part_o, float* __restrict__ part_ml,
                         int q8, float out_scale, int tiles_per_cta) {
  const __nv_bfloat16* q_nope = reinterpret_cast<const __nv_bfloat16*>(q_nope_v);
  const __nv_bfloat16* q_pe = reinterpret_cast<const __nv_bfloat16*>(q_pe_v);
  const __nv_bfloat16* kv = reinterpret_cast<const __nv_bfloat16*>(kv_v);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* k_area = smem;
  uint8_t* q_ring = smem + MLA_K_AREA;
  uint8_t* p_tile = q_ring + MLA_Q_RING;
  MlaBars* bars = reinterpret_cast<MlaBars*>(p_tile + MLA_P_TILE);

  const int split = blockIdx.x, half = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = seq_lens[b];
  const int t_first = split * tiles_per_cta * MLA_TOK;
  // this CTA walks n_tiles consecutive 128-token tiles of the request with an online softmax: O stays in TMEM, partial
  // (O, m, l) leave the CTA once.  (Round 1 wrote one fp32 partial per 128-token tile: at batch 64 x 2048 tokens that was
  // 3.5x the KV bytes in partial traffic and the kernel ran at 0.5 TB/s of KV.)
  const int n_tiles = min(tiles_per_cta, (S - t_first + MLA_TOK - 1) / MLA_TOK);
  const size_t pbase = ((size_t)b * Hq) * num_splits;

  if (n_tiles <= 0) {
    // empty split: publish (m = -inf, l = 0) so that the merge ignores it
    if (half == 0 && tid < Hq) {
      part_ml[(pbase + (size_t)tid * num_splits + split) * 2] = -CUDART_INF_F;
      part_ml[(pbase + (size_t)tid * num_splits + split) * 2 + 1] = 0.f;
    }
    return;
  }

  if (tid == 0) {
    for (int i = 0; i < MLA_KB; ++i) mbar_init(&bars->full[i], 128);
    for (int i = 0; i < MLA_QSTAGES; ++i) mbar_init(&bars->empty[i], 1);
    mbar_init(&bars->s_full, 1);
    mbar_init(&bars->p_full, 128);
    mbar_init(&bars->o_full, 1);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(&bars->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const uint32_t tm_s = tmem_base;            // S: columns [0,128)
  const uint32_t tm_o = tmem_base + 128;      // O: columns [128,384)
  // k-block order of a tile: the five blocks the PV MMA of THIS half does not read (the other half's value dims and the
  // rope block) come first — they can be refilled as soon as QK^T of the previous tile has completed, the four value
  // blocks only after its PV
  auto kb_of = [&](int j) { return j < 4 ? (1 - half) * 4 + j : (j == 4 ? 8 : half * 4 + (j - 5)); };

  if (warp >= 4 && warp < 8) {
    // ======================================================================= loaders
    const int lt = tid - 128;                 // 0..127
    const int c = lt & 7;                     // 16-byte chunk within the 128-byte row segment
    const int r0 = lt >> 3;                   // rows r0 + 16u, u = 0..7
    const uint8_t* qn8 = reinterpret_cast<const uint8_t*>(q_nope_v);
    const uint8_t* qp8 = reinterpret_cast<const uint8_t*>(q_pe_v);
    uint32_t qit = 0;                         // q-ring slot counter, continuous over the tiles
    for (int ti = 0; ti < n_tiles; ++ti) {
      const int t0 = t_first + ti * MLA_TOK;
      const int nt = min(MLA_TOK, S - t0);
      // base pointers of this thread's 8 KV rows (paged)
      const __nv_bfloat16* krow[8];
      const uint8_t* krow8[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int tt = r0 + 16 * u;
        if (tt < nt) {
          const int tok = t0 + tt;
          const int page = page_table[(size_t)b * max_pages + tok / page_size];
          krow[u] = kv + ((size_t)page * page_size + tok % page_size) * 576;
          krow8[u] = reinterpret_cast<const uint8_t*>(kv_v) + ((size_t)page * page_size + tok % page_size) * 576;
        } else {
          krow[u] = nullptr;
          krow8[u] = nullptr;
        }
      }
      if (KV8) {
        // e4m3 cache (and, with q8, e4m3 queries): 8-byte loads, widened to bf16 in registers, 16-byte stores into the
        // swizzled tiles; MLA_PF k-blocks of loads are kept in flight per thread (register ring)
        constexpr int MLA_PF = 3;
        uint2 kreg[MLA_PF][8], qreg[MLA_PF][8];
        auto issue = [&](int j, int buf) {
          const int kb = kb_of(j);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int r = r0 + 16 * u;
            kreg[buf][u] = krow8[u] ? __ldg(reinterpret_cast<const uint2*>(krow8[u] + kb * 64 + c * 8)) : make_uint2(0u, 0u);
            qreg[buf][u] = make_uint2(0u, 0u);
            if (q8 && r < Hq)
              qreg[buf][u] = __ldg(reinterpret_cast<const uint2*>(kb < 8 ? qn8 + ((size_t)b * Hq + r) * 512 + kb * 64 + c * 8
                                                                           : qp8 + ((size_t)b * Hq + r) * 64 + c * 8));
          }
        };
#pragma unroll
        for (int j = 0; j < MLA_PF - 1; ++j) issue(j, j);
#pragma unroll
        for (int j = 0; j < MLA_KB; ++j, ++qit) {
          const int kb = kb_of(j);
          const int s = qit % MLA_QSTAGES, buf = j % MLA_PF;
          if (j + MLA_PF - 1 < MLA_KB) issue(j + MLA_PF - 1, (j + MLA_PF - 1) % MLA_PF);
          if (ti > 0) mla_wait(j < 5 ? &bars->s_full : &bars->o_full, (ti - 1) & 1);   // the block's previous readers are done
          mla_wait(&bars->empty[s], ((qit / MLA_QSTAGES) & 1) ^ 1);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int r = r0 + 16 * u;
            if (q8) {
              *reinterpret_cast<uint4*>(q_ring + s * TILE_BYTES + sw128_offset(r, c * 16)) = fp8x8_to_bf16x8(qreg[buf][u]);
            } else {
              const bool qok = r < Hq;
              const __nv_bfloat16* qsrc = !qok ? q_nope
                                               : (kb < 8 ? q_nope + ((size_t)b * Hq + r) * 512 + kb * 64 + c * 8
                                                         : q_pe + ((size_t)b * Hq + r) * 64 + c * 8);
              cp_async16(q_ring + s * TILE_BYTES + sw128_offset(r, c * 16), qsrc, qok);
            }
            *reinterpret_cast<uint4*>(k_area + kb * TILE_BYTES + sw128_offset(r, c * 16)) = fp8x8_to_bf16x8(kreg[buf][u]);
          }
          if (!q8) asm volatile("cp.async.wait_all;" ::: "memory");
          fence_proxy_async();   // generic-proxy stores -> UMMA (async proxy) reads
          mbar_arrive(&bars->full[kb]);
        }
      } else {
        // cp.async (LDGSTS) straight into the swizzled tiles: no register staging; completion is signalled per k-block
        for (int j = 0; j < MLA_KB; ++j, ++qit) {
          const int kb = kb_of(j);
          const int s = qit % MLA_QSTAGES;
          if (ti > 0) mla_wait(j < 5 ? &bars->s_full : &bars->o_full, (ti - 1) & 1);
          mla_wait(&bars->empty[s], ((qit / MLA_QSTAGES) & 1) ^ 1);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int r = r0 + 16 * u;
            const bool qok = r < Hq;
            const __nv_bfloat16* qsrc = !qok ? q_nope
                                             : (kb < 8 ? q_nope + ((size_t)b * Hq + r) * 512 + kb * 64 + c * 8
                                                       : q_pe + ((size_t)b * Hq + r) * 64 + c * 8);
            cp_async16(q_ring + s * TILE_BYTES + sw128_offset(r, c * 16), qsrc, qok);
            const bool kok = krow[u] != nullptr;
            cp_async16(k_area + kb * TILE_BYTES + sw128_offset(r, c * 16), kok ? krow[u] + kb * 64 + c * 8 : kv, kok);
          }
          cp_async_mbar_arrive_noinc(&bars->full[kb]);
        }
      }
    }
  } else if (warp == 8) {
    // ======================================================================= MMA issuer
    if (elect_one()) {
      const uint32_t idesc_qk = umma_idesc(1, 1, 128, 128);       // bf16 x bf16, M=128, N=128 tokens
      const uint32_t idesc_pv = umma_idesc_bmn(1, 128, 256);
      uint32_t qit = 0;
      for (int ti = 0; ti < n_tiles; ++ti) {
        for (int j = 0; j < MLA_KB; ++j, ++qit) {
          const int kb = kb_of(j);
          const int s = qit % MLA_QSTAGES;
          mla_wait(&bars->full[kb], ti & 1);
          fence_proxy_async();   // LDGSTS / st.shared (generic proxy) writes -> UMMA (async proxy) reads
          tc_fence_after();
          const uint32_t qa = smem_u32(q_ring + s * TILE_BYTES);
          const uint32_t ka = smem_u32(k_area + kb * TILE_BYTES);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_f16(tm_s, umma_desc_sw128(qa + ks * 32, 1024), umma_desc_sw128(ka + ks * 32, 1024), idesc_qk,
                     (j > 0 || ks > 0) ? 1u : 0u);
          umma_commit(&bars->empty[s]);
        }
        umma_commit(&bars->s_full);
        // O (+)= P * V: A = P (K-major over tokens, 2 k-blocks of 64 tokens), B = V (MN-major view of K blocks)
        mla_wait(&bars->p_full, ti & 1);
        tc_fence_after();
        const uint32_t pa = smem_u32(p_tile);
        const uint32_t va = smem_u32(k_area + (half * 4) * TILE_BYTES);
#pragma unroll
        for (int k16 = 0; k16 < MLA_TOK / 16; ++k16) {
          const uint64_t ad = umma_desc_sw128(pa + (k16 >> 2) * TILE_BYTES + (k16 & 3) * 32, 1024);
          const uint64_t bd = umma_desc_sw128_mn(va + k16 * 2048, TILE_BYTES, 1024);
          umma_f16(tm_o, ad, bd, idesc_pv, (ti > 0 || k16 > 0) ? 1u : 0u);
        }
        umma_commit(&bars->o_full);
      }
    }
  } else {
    // ======================================================================= softmax + epilogue (thread = head)
    const int h = tid;                        // 0..127
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    float m_run = -CUDART_INF_F;              // running row maximum (raw logits)
    float l = 0.f;
    for (int ti = 0; ti < n_tiles; ++ti) {
      const int nt = min(MLA_TOK, S - (t_first + ti * MLA_TOK));
      mla_wait(&bars->s_full, ti & 1);
      tc_fence_after();
      float m = -CUDART_INF_F;
      // pass 1: row max over the valid tokens of the tile
#pragma unroll
      for (int c16 = 0; c16 < MLA_TOK / 16; ++c16) {
        float v[16];
        tmem_ld16(tm_s + lane_off + c16 * 16, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c16 * 16 + i < nt) m = fmaxf(m, v[i]);
      }
      // online softmax with a LAZY rescale: the running maximum moves (and O, l are rescaled) only when the tile's
      // maximum exceeds it by more than 2^8 in the exponent domain; otherwise P = exp2(s - m_run) <= 256 is accumulated as is
      float factor = 1.f;
      bool need = false;
      if (ti == 0) {
        m_run = m;
      } else if ((m - m_run) * scale_log2 > 8.f) {
        factor = exp2f((m_run - m) * scale_log2);
        m_run = m;
        need = true;
      }
      if (ti > 0) {
        // P (shared memory) and O (TMEM) are still read by the previous tile's PV MMAs until o_full flips
        mla_wait(&bars->o_full, (ti - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, need)) {
#pragma unroll 4
          for (int c16 = 0; c16 < 256 / 16; ++c16) {
            float v[16];
            tmem_ld16(tm_o + lane_off + c16 * 16, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] *= factor;
            tmem_st16(tm_o + lane_off + c16 * 16, reinterpret_cast<const uint32_t*>(v));
          }
          tmem_st_wait();
        }
        l *= factor;
      }
      const float ms = m_run * scale_log2;
      // pass 2: P = exp2(s*scale - m*scale) -> bf16 into the K-major P tile (row = head, 128 B = 64 tokens)
#pragma unroll
      for (int c16 = 0; c16 < MLA_TOK / 16; ++c16) {
        float v[16];
        tmem_ld16(tm_s + lane_off + c16 * 16, v);
        tmem_ld_wait();
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          const int t = c16 * 16 + i;
          float p0 = (t < nt) ? exp2f(fmaf(v[i], scale_log2, -ms)) : 0.f;
          float p1 = (t + 1 < nt) ? exp2f(fmaf(v[i + 1], scale_log2, -ms)) : 0.f;
          const __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
          // accumulate l from the ROUNDED probabilities so that O / l is consistent with the bf16 P used by the MMA
          l += __low2float(pb) + __high2float(pb);
          pk[i >> 1] = *reinterpret_cast<const uint32_t*>(&pb);
        }
        // 16 tokens = 32 bytes = two 16-byte chunks of k-block (c16 >> 2)
        uint8_t* base = p_tile + (c16 >> 2) * TILE_BYTES;
        const int boff = (c16 & 3) * 32;
        *reinterpret_cast<uint4*>(base + sw128_offset(h, boff)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(base + sw128_offset(h, boff + 16)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&bars->p_full);
    }
    // epilogue: partial O (un-normalised) + (m, l) in the exp2 domain expected by the merge kernels
    mla_wait(&bars->o_full, (n_tiles - 1) & 1);
    tc_fence_after();
    // NOTE: tcgen05.ld is .sync.aligned — every lane of the warp must execute it; only the stores are predicated
    {
      const int hh = h < Hq ? h : 0;
      float* po = part_o + (pbase + (size_t)hh * num_splits + split) * 512 + half * 256;
#pragma unroll 4
      for (int c16 = 0; c16 < 256 / 16; ++c16) {
        float v[16];
        tmem_ld16(tm_o + lane_off + c16 * 16, v);
        tmem_ld_wait();
        if (h < Hq) {
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            *reinterpret_cast<float4*>(po + c16 * 16 + i) =
                make_float4(v[i] * out_scale, v[i + 1] * out_scale, v[i + 2] * out_scale, v[i + 3] * out_scale);
        }
      }
      if (half == 0 && h < Hq) {
        part_ml[(pbase + (size_t)h * num_splits + split) * 2] = m_run * scale_log2;
        part_ml[(pbase + (size_t)h * num_splits + split) * 2 + 1] = l;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 512);
}

int launch_mla_tc(cudaStream_t st, const void* q_nope, const void* q_pe, const void* kv, const int32_t* seq_lens,
                  const int32_t* page_table, int batch, int Hq, int page_size, int max_pages, float sm_scale,
                  int num_splits, float* part_o, float* part_ml, int kv_fp8, int q_fp8, float descale_q, float descale_k) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(mla_decode_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MLA_SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(mla_decode_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MLA_SMEM);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mla_decode_tc)");
    attr = true;
  }
  if (q_fp8 && !kv_fp8) {
    set_error("b200_mla_decode: e4m3 queries need the e4m3 cache");
    return B200_ERR_INVALID;
  }
  dim3 grid(num_splits, 2, batch);
  // each CTA walks tpc consecutive 128-token tiles (online softmax); the splits together cover the page table
  const int tiles = (int)(((int64_t)max_pages * page_size + MLA_TOK - 1) / MLA_TOK);
  const int tpc = (tiles + num_splits - 1) / num_splits;
  // descales (reference cutlass_mla.py: q_scale * k_scale folded into the softmax scale, k_scale into the output)
  const float sl2 = sm_scale * 1.4426950408889634f * (kv_fp8 ? descale_k * (q_fp8 ? descale_q : 1.f) : 1.f);
  const float osc = kv_fp8 ? descale_k : 1.f;
  if (kv_fp8)
    mla_decode_tc_kernel<true><<<grid, MLA_THREADS, MLA_SMEM, st>>>(q_nope, q_pe, kv, seq_lens, page_table, Hq, page_size,
                                                                   max_pages, sl2, num_splits, part_o, part_ml, q_fp8, osc, tpc);
  else
    mla_decode_tc_kernel<false><<<grid, MLA_THREADS, MLA_SMEM, st>>>(q_nope, q_pe, kv, seq_lens, page_table, Hq, page_size,
                                                                    max_pages, sl2, num_splits, part_o, part_ml, 0, 1.f, tpc);
  ++g_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "mla_decode_tc launch");
  return 0;
}

}  // namespace b200
