// Paged GQA decode on tcgen05 tensor cores (head_dim 128), persistent CTAs.
//
// Work item = (request b, kv head, 128-token split).  The G = Hq/Hkv query heads that share the kv head are
// the M rows of the MMAs (rows >= G are zero padding — the kernel is KV-bandwidth bound, tensor time is noise):
//   S[128, 128 tokens] = Q[128, 128] K^T     P = exp2(S*scale - m) (bf16)     O[128, 128] = P V
// K and V tiles are gathered with cp.async straight into 128B-swizzled [64-dim block][128 tokens][128 B]
// blocks: K is the K-major B operand of QK^T, V — same layout — the MN-major B operand of PV.
// A CTA loops over items with single-buffered tiles: the K/Q loads of item i+1 overlap softmax + PV of item i.
// Partial (O, m, l) per split are merged by attn_merge_kernel.  Reference semantics:
// tests/kernels/attention/test_flashinfer.py:29-80 (ref_paged_attn, decode case).
// Roofline: HBM — algorithmic bytes = sum_b S_b * Hkv * 128 * 2 (K and V) * 2 B.
#include <math_constants.h>

#include "common.cuh"
#include "moe_internal.cuh"

namespace b200 {

constexpr int GQA_TOK = 128;
constexpr int GQA_THREADS = 288;
constexpr int GQA_SMEM = 4 * 2 * TILE_BYTES + 1024 + 1024;   // Q, K, V, P: two 16 KB blocks each

struct GqaBars {
  uint64_t qk_full, v_full, s_full, p_full, o_full, k_free, v_free;
  uint32_t tmem_base;
};

B200_DEVICE void gqa_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}
B200_DEVICE uint64_t gqa_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(GQA_THREADS, 1)
    gqa_decode_tc_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kc,
                         const __nv_bfloat16* __restrict__ vc, const int32_t* __restrict__ seq_lens,
                         const int32_t* __restrict__ page_table, int B, int Hq, int Hkv, int page_size, int max_pages,
                         float scale_log2, int num_splits, float* __restrict__ part_o, float* __restrict__ part_ml) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_t = smem;
  uint8_t* k_t = q_t + 2 * TILE_BYTES;
  uint8_t* v_t = k_t + 2 * TILE_BYTES;
  uint8_t* p_t = v_t + 2 * TILE_BYTES;
  GqaBars* bars = reinterpret_cast<GqaBars*>(p_t + 2 * TILE_BYTES);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = Hq / Hkv;
  const int total = B * Hkv * num_splits;

  if (tid == 0) {
    mbar_init(&bars->qk_full, 128);
    mbar_init(&bars->v_full, 128);
    mbar_init(&bars->s_full, 1);
    mbar_init(&bars->p_full, 128);
    mbar_init(&bars->o_full, 1);
    mbar_init(&bars->k_free, 1);
    mbar_init(&bars->v_free, 1);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(&bars->tmem_base, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm_s = bars->tmem_base, tm_o = bars->tmem_base + 128;

  // every role walks the same item list; `n` counts the items that are actually processed (phase parity)
  if (warp >= 4 && warp < 8) {
    // ======================================================================= loaders
    const int lt = tid - 128, c = lt & 7, r0 = lt >> 3;
    uint32_t n = 0;
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
      const int split = item % num_splits, kvh = (item / num_splits) % Hkv, b = item / (num_splits * Hkv);
      const int S = seq_lens[b], t0 = split * GQA_TOK;
      const int nt = min(GQA_TOK, S - t0);
      if (nt <= 0) continue;
      size_t rowoff[8];
      bool rok[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int tt = r0 + 16 * u;
        rok[u] = tt < nt;
        rowoff[u] = 0;
        if (rok[u]) {
          const int tok = t0 + tt;
          const int page = page_table[(size_t)b * max_pages + tok / page_size];
          rowoff[u] = (((size_t)page * page_size + tok % page_size) * Hkv + kvh) * 128;
        }
      }
      gqa_wait(&bars->k_free, (n & 1) ^ 1);          // QK^T of the previous item has consumed Q and K
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int r = r0 + 16 * u;
          const bool qok = r < G;
          cp_async16(q_t + kb * TILE_BYTES + sw128_offset(r, c * 16),
                     qok ? q + ((size_t)b * Hq + kvh * G + r) * 128 + kb * 64 + c * 8 : q, qok);
          cp_async16(k_t + kb * TILE_BYTES + sw128_offset(r, c * 16), rok[u] ? kc + rowoff[u] + kb * 64 + c * 8 : kc, rok[u]);
        }
      }
      cp_async_mbar_arrive_noinc(&bars->qk_full);
      gqa_wait(&bars->v_free, (n & 1) ^ 1);          // PV of the previous item has consumed V
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int r = r0 + 16 * u;
          cp_async16(v_t + kb * TILE_BYTES + sw128_offset(r, c * 16), rok[u] ? vc + rowoff[u] + kb * 64 + c * 8 : vc, rok[u]);
        }
      }
      cp_async_mbar_arrive_noinc(&bars->v_full);
      ++n;
    }
  } else if (warp == 8) {
    // ======================================================================= MMA issuer
    if (lane == 0) {
      const uint32_t idesc_qk = umma_idesc(1, 1, 128, 128);
      const uint32_t idesc_pv = umma_idesc(1, 1, 128, 128) | (1u << 16);   // B (= V) is MN-major
      uint32_t n = 0;
      for (int item = blockIdx.x; item < total; item += gridDim.x) {
        const int split = item % num_splits, b = item / (num_splits * Hkv);
        if (seq_lens[b] - split * GQA_TOK <= 0) continue;
        const uint32_t ph = n & 1;
        gqa_wait(&bars->qk_full, ph);
        fence_proxy_async();
        tc_fence_after();
        const uint32_t qa = smem_u32(q_t), ka = smem_u32(k_t);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_f16(tm_s, umma_desc_sw128(qa + kb * TILE_BYTES + ks * 32, 1024),
                     umma_desc_sw128(ka + kb * TILE_BYTES + ks * 32, 1024), idesc_qk, (kb > 0 || ks > 0) ? 1u : 0u);
        umma_commit(&bars->k_free);
        umma_commit(&bars->s_full);
        gqa_wait(&bars->p_full, ph);
        gqa_wait(&bars->v_full, ph);
        fence_proxy_async();
        tc_fence_after();
        const uint32_t pa = smem_u32(p_t), va = smem_u32(v_t);
#pragma unroll
        for (int k16 = 0; k16 < GQA_TOK / 16; ++k16)
          umma_f16(tm_o, umma_desc_sw128(pa + (k16 >> 2) * TILE_BYTES + (k16 & 3) * 32, 1024),
                   gqa_desc_mn(va + k16 * 2048, TILE_BYTES, 1024), idesc_pv, k16 > 0 ? 1u : 0u);
        umma_commit(&bars->v_free);
        umma_commit(&bars->o_full);
        ++n;
      }
    }
  } else if (warp < 4) {
    // ======================================================================= softmax + epilogue (thread = q head of the group)
    const int h = tid;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    uint32_t n = 0;
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
      const int split = item % num_splits, kvh = (item / num_splits) % Hkv, b = item / (num_splits * Hkv);
      const int S = seq_lens[b], t0 = split * GQA_TOK;
      const int nt = min(GQA_TOK, S - t0);
      const size_t pidx = ((size_t)b * Hq + kvh * G + (h < G ? h : 0)) * num_splits + split;
      if (nt <= 0) {
        if (h < G) {
          part_ml[pidx * 2] = -CUDART_INF_F;
          part_ml[pidx * 2 + 1] = 0.f;
        }
        continue;
      }
      const uint32_t ph = n & 1;
      gqa_wait(&bars->s_full, ph);
      tc_fence_after();
      float m = -CUDART_INF_F;
#pragma unroll
      for (int c16 = 0; c16 < GQA_TOK / 16; ++c16) {
        float v[16];
        tmem_ld16(tm_s + lane_off + c16 * 16, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c16 * 16 + i < nt) m = fmaxf(m, v[i]);
      }
      const float ms = m * scale_log2;
      float l = 0.f;
#pragma unroll
      for (int c16 = 0; c16 < GQA_TOK / 16; ++c16) {
        float v[16];
        tmem_ld16(tm_s + lane_off + c16 * 16, v);
        tmem_ld_wait();
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          const int t = c16 * 16 + i;
          const float p0 = (t < nt) ? exp2f(fmaf(v[i], scale_log2, -ms)) : 0.f;
          const float p1 = (t + 1 < nt) ? exp2f(fmaf(v[i + 1], scale_log2, -ms)) : 0.f;
          const __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
          l += __low2float(pb) + __high2float(pb);
          pk[i >> 1] = *reinterpret_cast<const uint32_t*>(&pb);
        }
        uint8_t* base = p_t + (c16 >> 2) * TILE_BYTES;
        const int boff = (c16 & 3) * 32;
        *reinterpret_cast<uint4*>(base + sw128_offset(h, boff)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(base + sw128_offset(h, boff + 16)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&bars->p_full);
      gqa_wait(&bars->o_full, ph);
      tc_fence_after();
      float* po = part_o + pidx * 128;
#pragma unroll 2
      for (int c16 = 0; c16 < 128 / 16; ++c16) {
        float v[16];
        tmem_ld16(tm_o + lane_off + c16 * 16, v);   // .sync.aligned: executed by every lane, stores predicated
        tmem_ld_wait();
        if (h < G) {
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            *reinterpret_cast<float4*>(po + c16 * 16 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
      }
      if (h < G) {
        part_ml[pidx * 2] = ms;
        part_ml[pidx * 2 + 1] = l;
      }
      tc_fence_before();
      ++n;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(bars->tmem_base, 256);
}

// merge of the split partials: one CTA per (b, head), DV/4 threads (thread = float4 of the output row)
template <int DV>
__global__ void __launch_bounds__(DV / 4) attn_merge_kernel(const float* __restrict__ part_o,
                                                           const float* __restrict__ part_ml, int num_splits,
                                                           __nv_bfloat16* __restrict__ out, float* __restrict__ lse) {
  constexpr int NT = DV / 4;
  __shared__ float wgt[1024];
  __shared__ float red[NT / 32 > 0 ? NT / 32 : 1];
  const int w = blockIdx.x, tid = threadIdx.x;
  const float* ml = part_ml + (size_t)w * num_splits * 2;
  float mx = -CUDART_INF_F;
  for (int s = tid; s < num_splits; s += NT) mx = fmaxf(mx, ml[s * 2]);
  mx = warp_max(mx);
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int i = 1; i < NT / 32; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float ls = 0.f;
  for (int s = tid; s < num_splits; s += NT) {
    const float msv = ml[s * 2];
    const float f = (msv == -CUDART_INF_F) ? 0.f : exp2f(msv - mx);
    wgt[s] = f;
    ls += f * ml[s * 2 + 1];
  }
  ls = warp_sum(ls);
  if ((tid & 31) == 0) red[tid >> 5] = ls;
  __syncthreads();
  float lsum = 0.f;
  for (int i = 0; i < NT / 32; ++i) lsum += red[i];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* po = part_o + (size_t)w * num_splits * DV + tid * 4;
  for (int s0 = 0; s0 < num_splits; s0 += 8) {
    float4 v[8];
    float f[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int s = s0 + u;
      f[u] = (s < num_splits) ? wgt[s] : 0.f;
      v[u] = (f[u] != 0.f) ? *reinterpret_cast<const float4*>(po + (size_t)s * DV) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      acc.x = fmaf(f[u], v[u].x, acc.x);
      acc.y = fmaf(f[u], v[u].y, acc.y);
      acc.z = fmaf(f[u], v[u].z, acc.z);
      acc.w = fmaf(f[u], v[u].w, acc.w);
    }
  }
  const float inv = lsum > 0.f ? 1.f / lsum : 0.f;
  __nv_bfloat162 a = __floats2bfloat162_rn(acc.x * inv, acc.y * inv), c = __floats2bfloat162_rn(acc.z * inv, acc.w * inv);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&a);
  pk.y = *reinterpret_cast<uint32_t*>(&c);
  *reinterpret_cast<uint2*>(out + (size_t)w * DV + tid * 4) = pk;
  if (lse && tid == 0) lse[w] = (lsum > 0.f) ? (mx * 0.6931471805599453f + logf(lsum)) : -CUDART_INF_F;
}

int launch_gqa_tc(cudaStream_t st, const void* q, const void* kc, const void* vc, const int32_t* seq_lens,
                  const int32_t* page_table, int batch, int Hq, int Hkv, int page_size, int max_pages, float sm_scale,
                  int num_splits, float* part_o, float* part_ml, void* out, float* lse) {
  static bool attr = false;
  static int num_sms = 0;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gqa_decode_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GQA_SMEM);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(gqa_decode_tc)");
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    attr = true;
  }
  const int total = batch * Hkv * num_splits;
  const int grid = total < num_sms ? total : num_sms;
  gqa_decode_tc_kernel<<<grid, GQA_THREADS, GQA_SMEM, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(q), reinterpret_cast<const __nv_bfloat16*>(kc),
      reinterpret_cast<const __nv_bfloat16*>(vc), seq_lens, page_table, batch, Hq, Hkv, page_size, max_pages,
      sm_scale * 1.4426950408889634f, num_splits, part_o, part_ml);
  attn_merge_kernel<128><<<batch * Hq, 32, 0, st>>>(part_o, part_ml, num_splits, reinterpret_cast<__nv_bfloat16*>(out), lse);
  g_launches += 2;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "gqa_decode_tc launch");
  return 0;
}

}  // namespace b200
