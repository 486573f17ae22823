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

__global__ void __launch_bounds__(MLA_THREADS, 1)
    mla_decode_tc_kernel(const __nv_bfloat16* __restrict__ q_nope, const __nv_bfloat16* __restrict__ q_pe,
                         const __nv_bfloat16* __restrict__ kv, const int32_t* __restrict__ seq_lens,
                         const int32_t* __restrict__ page_table, int Hq, int page_size, int max_pages,
                         float scale_log2, int num_splits, float* __restrict__ part_o, float* __restrict__ part_ml) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* k_area = smem;
  uint8_t* q_ring = smem + MLA_K_AREA;
  uint8_t* p_tile = q_ring + MLA_Q_RING;
  MlaBars* bars = reinterpret_cast<MlaBars*>(p_tile + MLA_P_TILE);

  const int split = blockIdx.x, half = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = seq_lens[b];
  const int t0 = split * MLA_TOK;
  const int nt = min(MLA_TOK, S - t0);     // valid tokens of this split (<= 0: nothing to do)
  const size_t pbase = ((size_t)b * Hq) * num_splits;

  if (nt <= 0) {
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

  if (warp >= 4 && warp < 8) {
    // ======================================================================= loaders
    const int lt = tid - 128;                 // 0..127
    const int c = lt & 7;                     // 16-byte chunk within the 128-byte row segment
    const int r0 = lt >> 3;                   // rows r0 + 16u, u = 0..7
    // base pointers of this thread's 8 KV rows (paged) — looked up once
    const __nv_bfloat16* krow[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int tt = r0 + 16 * u;
      if (tt < nt) {
        const int tok = t0 + tt;
        const int page = page_table[(size_t)b * max_pages + tok / page_size];
        krow[u] = kv + ((size_t)page * page_size + tok % page_size) * 576;
      } else {
        krow[u] = nullptr;
      }
    }
    // cp.async (LDGSTS) straight into the swizzled tiles: no register staging, every k-block's loads are in
    // flight at once (K) or as deep as the Q ring allows; completion is signalled per k-block on full[kb]
    for (int kb = 0; kb < MLA_KB; ++kb) {
      const int s = kb % MLA_QSTAGES;
      mla_wait(&bars->empty[s], ((kb / MLA_QSTAGES) & 1) ^ 1);
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
  } else if (warp == 8) {
    // ======================================================================= MMA issuer
    if (elect_one()) {
      const uint32_t idesc_qk = umma_idesc(1, 1, 128, 128);       // bf16 x bf16, M=128, N=128 tokens
      for (int kb = 0; kb < MLA_KB; ++kb) {
        const int s = kb % MLA_QSTAGES;
        mla_wait(&bars->full[kb], 0);
        fence_proxy_async();   // LDGSTS (generic proxy) writes -> UMMA (async proxy) reads
        tc_fence_after();
        const uint32_t qa = smem_u32(q_ring + s * TILE_BYTES);
        const uint32_t ka = smem_u32(k_area + kb * TILE_BYTES);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_f16(tm_s, umma_desc_sw128(qa + ks * 32, 1024), umma_desc_sw128(ka + ks * 32, 1024), idesc_qk,
                   (kb > 0 || ks > 0) ? 1u : 0u);
        umma_commit(&bars->empty[s]);
      }
      umma_commit(&bars->s_full);
      // O = P * V: A = P (K-major over tokens, 2 k-blocks of 64 tokens), B = V (MN-major view of K blocks)
      mla_wait(&bars->p_full, 0);
      tc_fence_after();
      const uint32_t idesc_pv = umma_idesc_bmn(1, 128, 256);
      const uint32_t pa = smem_u32(p_tile);
      const uint32_t va = smem_u32(k_area + (half * 4) * TILE_BYTES);
#pragma unroll
      for (int k16 = 0; k16 < MLA_TOK / 16; ++k16) {
        const uint64_t ad = umma_desc_sw128(pa + (k16 >> 2) * TILE_BYTES + (k16 & 3) * 32, 1024);
        const uint64_t bd = umma_desc_sw128_mn(va + k16 * 2048, TILE_BYTES, 1024);
        umma_f16(tm_o, ad, bd, idesc_pv, k16 > 0 ? 1u : 0u);
      }
      umma_commit(&bars->o_full);
    }
  } else {
    // ======================================================================= softmax + epilogue (thread = head)
    const int h = tid;                        // 0..127
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    mla_wait(&bars->s_full, 0);
    tc_fence_after();
    float m = -CUDART_INF_F;
    // pass 1: row max over the valid tokens
#pragma unroll
    for (int c16 = 0; c16 < MLA_TOK / 16; ++c16) {
      float v[16];
      tmem_ld16(tm_s + lane_off + c16 * 16, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (c16 * 16 + i < nt) m = fmaxf(m, v[i]);
    }
    const float ms = m * scale_log2;
    float l = 0.f;
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
    // epilogue: partial O (un-normalised) + (m, l) in the exp2 domain expected by decode_merge_kernel
    mla_wait(&bars->o_full, 0);
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
            *reinterpret_cast<float4*>(po + c16 * 16 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
      }
      if (half == 0 && h < Hq) {
        part_ml[(pbase + (size_t)h * num_splits + split) * 2] = ms;
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
                  int num_splits, float* part_o, float* part_ml) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(mla_decode_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MLA_SMEM);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mla_decode_tc)");
    attr = true;
  }
  dim3 grid(num_splits, 2, batch);
  mla_decode_tc_kernel<<<grid, MLA_THREADS, MLA_SMEM, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(q_nope), reinterpret_cast<const __nv_bfloat16*>(q_pe),
      reinterpret_cast<const __nv_bfloat16*>(kv), seq_lens, page_table, Hq, page_size, max_pages,
      sm_scale * 1.4426950408889634f, num_splits, part_o, part_ml);
  ++g_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "mla_decode_tc launch");
  return 0;
}

}  // namespace b200
