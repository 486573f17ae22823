// Paged GQA decode on tcgen05 tensor cores (head_dim 128), persistent CTAs, multi-stage KV pipeline.
//
// Work item = (request b, kv head, 128-token split).  The kernel is KV-bandwidth bound, so the tiles are laid
// out to make every thread useful: tokens are the M rows of the MMAs and the G = Hq/Hkv query heads sharing the
// kv head are the N columns (N = G rounded up to 16):
//   S^T[128 tokens, N]  = K[128 tokens, 128 dims] . Q^T           (A = K tile, K-major; B = Q rows, K-major)
//   P^T = exp2(S^T * scale - m)                                    (thread = token; m via redux.sync.max.f32)
//   O^T[128 dims, N]    = V^T[128 dims, 128 tokens] . P^T          (A = V tile read MN-major; B = P rows, K-major)
//   L[128, N]           = 1[128, 128 tokens] . P^T                 (row sums of P on the tensor core: every
//                                                                   epilogue thread reads l[h] from its own lane)
// K and V tiles are gathered with cp.async straight into 128B-swizzled [64-dim block][128 tokens][128 B] blocks;
// NST stages of (Q, K, V) let the loaders run ahead of the MMAs; S^T is double-buffered in TMEM so that QK^T of
// item n+1 overlaps softmax of item n.  Partial (O, m, l) per split are merged by attn_merge_kernel.
// Reference semantics: tests/kernels/attention/test_flashinfer.py:29-80 (ref_paged_attn, decode case).
// Roofline: HBM — algorithmic bytes = sum_b S_b * Hkv * 128 * 2 (K and V) * 2 B.
#include <math_constants.h>
#include <stdio.h>

#include "common.cuh"
#include "moe_internal.cuh"

namespace b200 {

constexpr int GQA_TOK = 128;
constexpr int GQA_THREADS = 416;          // warps 0-3 / 4-7: softmax+epilogue groups, 8-11 loaders, 12 MMA issuer
constexpr int GQA_MAX_STAGES = 3;
constexpr int GQA_SMEM_MAX = 232448;
constexpr int GQA_MAX_ITEMS = 1024;       // items per CTA the token-count table holds
constexpr int GQA_MAX_N = 80;             // 2 groups x (S^T, O^T, L) x N columns must fit the 512 TMEM columns

struct GqaBars {
  uint64_t qk_full[GQA_MAX_STAGES], v_full[GQA_MAX_STAGES], kv_free[GQA_MAX_STAGES];
  uint64_t s_full[2], p_full[2], o_full[2];
  uint32_t tmem_base;
  alignas(16) float wmax[2][4][GQA_MAX_N];
  int16_t nt_tab[GQA_MAX_ITEMS];   // valid tokens of this CTA's i-th item (<= 0: empty split)
};

// intra-CTA mbarrier wait (producer / MMA / softmax warps of ONE CTA: no other CTA or GPU is waited for).  A protocol bug
// must not hang the box: back off after a while, trap after seconds.
B200_DEVICE void gqa_wait(uint64_t* bar, uint32_t parity, int tag = 0, uint32_t n = 0) {
  (void)tag;
  (void)n;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 20)) __nanosleep(128);
    if (spins > (1u << 26)) __trap();
  }
}
// MN-major, 128B-swizzled operand: atoms of 64 elements (128 B) along MN x 8 rows along K.
B200_DEVICE uint64_t gqa_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
B200_DEVICE float redux_max(float v) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}
B200_DEVICE float ex2_approx(float v) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
B200_DEVICE void bar_sync_group(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }

// smem: [ones 1 KB][P group 0 | P group 1: 2 x N x 128 B each][stage: Q (2 x N x 128 B) | K 32 KB | V 32 KB] x NST [bars]
__global__ void __launch_bounds__(GQA_THREADS, 1)
    gqa_decode_tc_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kc,
                         const __nv_bfloat16* __restrict__ vc, const int32_t* __restrict__ seq_lens,
                         const int32_t* __restrict__ page_table, int B, int Hq, int Hkv, int page_size, int max_pages,
                         float scale_log2, int num_splits, int N, int NST, float* __restrict__ part_o,
                         float* __restrict__ part_ml) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int qbytes = N * 128;                       // one 64-dim block of the Q (or P) rows
  const int pbytes = (2 * qbytes + 1023) & ~1023;   // one group's P tile
  const int stage_bytes = 2 * qbytes + 4 * TILE_BYTES;
  uint8_t* ones = smem;
  uint8_t* p_base = smem + 1024;
  uint8_t* stages = p_base + 2 * pbytes;
  GqaBars* bars = reinterpret_cast<GqaBars*>(stages + (size_t)NST * stage_bytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = Hq / Hkv;
  const int total = B * Hkv * num_splits;

  if (tid == 0) {
    for (int i = 0; i < GQA_MAX_STAGES; ++i) {
      mbar_init(&bars->qk_full[i], 128);
      mbar_init(&bars->v_full[i], 128);
      mbar_init(&bars->kv_free[i], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&bars->s_full[g], 1);
      mbar_init(&bars->p_full[g], 128);
      mbar_init(&bars->o_full[g], 1);
    }
    fence_barrier_init();
  }
  const int my_items = blockIdx.x < total ? (total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  for (int i = tid; i < my_items; i += GQA_THREADS) {
    const int item = blockIdx.x + i * gridDim.x;
    const int left = seq_lens[item / (num_splits * Hkv)] - (item % num_splits) * GQA_TOK;
    bars->nt_tab[i] = (int16_t)max(0, min(GQA_TOK, left));
  }
  if (tid < 64) reinterpret_cast<uint4*>(ones)[tid] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  if (warp == 12) tmem_alloc(&bars->tmem_base, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = bars->tmem_base;   // group g: S^T at g*256, O^T at g*256 + N, L at g*256 + 2N

  if (warp >= 8 && warp < 12) {
    // ======================================================================= loaders
    const int lt = tid - 256, c = lt & 7, r0 = lt >> 3;
    // page-table entries are looked up one item ahead so that their latency is off the issue path
    int pg[8], pg_next[8];
    auto lookup = [&](int i, int* out) {
      if (i >= my_items) return;
      const int item = blockIdx.x + i * gridDim.x;
      const int split = item % num_splits, b = item / (num_splits * Hkv);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int pi = min((split * GQA_TOK + r0 + 16 * u) / page_size, max_pages - 1);
        out[u] = page_table[(size_t)b * max_pages + pi];
      }
    };
    lookup(0, pg);
    uint32_t n = 0;
    for (int i = 0; i < my_items; ++i) {
      lookup(i + 1, pg_next);
      const int nt = bars->nt_tab[i];
      if (nt > 0) {
        const int item = blockIdx.x + i * gridDim.x;
        const int split = item % num_splits, kvh = (item / num_splits) % Hkv, b = item / (num_splits * Hkv);
        const int s = n % NST;
        uint8_t* q_t = stages + (size_t)s * stage_bytes;
        uint8_t* k_t = q_t + 2 * qbytes;
        uint8_t* v_t = k_t + 2 * TILE_BYTES;
        gqa_wait(&bars->kv_free[s], ((n / NST) & 1) ^ 1, 1, n);
        for (int r = r0; r < N; r += 16) {
          const bool qok = r < G;
          const __nv_bfloat16* src = q + ((size_t)b * Hq + kvh * G + (qok ? r : 0)) * 128 + c * 8;
          cp_async16(q_t + sw128_offset(r, c * 16), src, qok);
          cp_async16(q_t + qbytes + sw128_offset(r, c * 16), src + 64, qok);
        }
        size_t rowoff[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int tok = split * GQA_TOK + r0 + 16 * u;
          rowoff[u] = (((size_t)pg[u] * page_size + tok % page_size) * Hkv + kvh) * 128 + c * 8;
        }
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const bool ok = r0 + 16 * u < nt;
            cp_async16(k_t + kb * TILE_BYTES + sw128_offset(r0 + 16 * u, c * 16), ok ? kc + rowoff[u] + kb * 64 : kc, ok);
          }
        }
        cp_async_mbar_arrive_noinc(&bars->qk_full[s]);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const bool ok = r0 + 16 * u < nt;
            cp_async16(v_t + kb * TILE_BYTES + sw128_offset(r0 + 16 * u, c * 16), ok ? vc + rowoff[u] + kb * 64 : vc, ok);
          }
        }
        cp_async_mbar_arrive_noinc(&bars->v_full[s]);
        ++n;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) pg[u] = pg_next[u];
    }
  } else if (warp == 12) {
    // ======================================================================= MMA issuer
    if (elect_one()) {   // elect.sync: single-lane region, operands stay in uniform registers
      const uint32_t idesc_qk = umma_idesc(1, 1, 128, N);                 // A = K (K-major), B = Q (K-major)
      const uint32_t idesc_pv = umma_idesc(1, 1, 128, N) | (1u << 15);    // A = V read MN-major
      const uint32_t idesc_l = umma_idesc(1, 1, 128, N);
      const uint32_t oa = smem_u32(ones);
      auto issue_qk = [&](uint32_t n) {
        const int s = n % NST, g = n & 1;
        gqa_wait(&bars->qk_full[s], (n / NST) & 1, 2, n);
        fence_proxy_async();
        tc_fence_after();
        const uint32_t qa = smem_u32(stages + (size_t)s * stage_bytes), ka = qa + 2 * qbytes;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_f16(tm + g * 256, umma_desc_sw128(ka + kb * TILE_BYTES + ks * 32, 1024),
                     umma_desc_sw128(qa + kb * qbytes + ks * 32, 1024), idesc_qk, (kb > 0 || ks > 0) ? 1u : 0u);
        umma_commit(&bars->s_full[g]);
      };
      auto issue_pv = [&](uint32_t n) {
        const int s = n % NST, g = n & 1;
        gqa_wait(&bars->p_full[g], (n >> 1) & 1, 3, n);
        gqa_wait(&bars->v_full[s], (n / NST) & 1, 4, n);
        fence_proxy_async();
        tc_fence_after();
        const uint32_t va = smem_u32(stages + (size_t)s * stage_bytes) + 2 * qbytes + 2 * TILE_BYTES;
        const uint32_t pa = smem_u32(p_base + g * pbytes);
        const uint32_t tm_o = tm + g * 256 + N, tm_l = tm + g * 256 + 2 * N;
#pragma unroll
        for (int k16 = 0; k16 < GQA_TOK / 16; ++k16) {
          const uint64_t pd = umma_desc_sw128(pa + (k16 >> 2) * qbytes + (k16 & 3) * 32, 1024);
          umma_f16(tm_o, gqa_desc_mn(va + k16 * 2048, TILE_BYTES, 1024), pd, idesc_pv, k16 > 0 ? 1u : 0u);
          umma_f16(tm_l, umma_desc_sw128(oa + (k16 & 3) * 32, 0), pd, idesc_l, k16 > 0 ? 1u : 0u);
        }
        umma_commit(&bars->kv_free[s]);
        umma_commit(&bars->o_full[g]);
      };
      // order: QK(0), QK(1), PV(0), QK(2), PV(1), ...  — item n belongs to softmax group n & 1
      uint32_t issued = 0;
      for (int i = 0; i < my_items; ++i) {
        if (bars->nt_tab[i] <= 0) continue;
        issue_qk(issued++);
        if (issued > 1) issue_pv(issued - 2);
      }
      if (issued > 0) issue_pv(issued - 1);
    }
  } else if (warp < 8) {
    // ======================================================================= softmax (thread = token) + epilogue (thread = dim)
    const int g = warp >> 2, t = tid & 127;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tm_s = tm + g * 256 + lane_off, tm_o = tm_s + N, tm_l = tm_s + 2 * N;
    uint8_t* prow = p_base + g * pbytes + (t >> 6) * qbytes;
    const uint32_t boff = (t & 63) * 2, bch = boff >> 4, blo = boff & 15;
    float(*wmax)[GQA_MAX_N] = bars->wmax[g];
    uint32_t n = 0;
    for (int i = 0; i < my_items; ++i) {
      const int nt = bars->nt_tab[i];
      const int item = blockIdx.x + i * gridDim.x;
      const int split = item % num_splits, kvh = (item / num_splits) % Hkv, b = item / (num_splits * Hkv);
      const size_t pbase = ((size_t)b * Hq + kvh * G) * num_splits + split;   // + h * num_splits
      if (nt <= 0) {
        if (g == 0 && t < G) {
          part_ml[(pbase + (size_t)t * num_splits) * 2] = -CUDART_INF_F;
          part_ml[(pbase + (size_t)t * num_splits) * 2 + 1] = 0.f;
        }
        continue;
      }
      if ((int)(n & 1) != g) {
        ++n;
        continue;
      }
      const uint32_t ph = (n >> 1) & 1;
      const bool tv = t < nt;
      gqa_wait(&bars->s_full[g], ph, 5, n);
      tc_fence_after();
      float my_ms = 0.f;                          // thread t == h keeps m[h] * scale for the partial record
      for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        __syncwarp();
        tmem_ld16(tm_s + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          v[j] = tv ? v[j] : -CUDART_INF_F;
          const float wm = redux_max(v[j]);
          if (lane == j) wmax[warp & 3][c0 + j] = wm;
        }
        bar_sync_group(g);
        float m[16];
#pragma unroll
        for (int j4 = 0; j4 < 16; j4 += 4) {
          const float4 a = *reinterpret_cast<const float4*>(&wmax[0][c0 + j4]);
          const float4 bq = *reinterpret_cast<const float4*>(&wmax[1][c0 + j4]);
          const float4 cq = *reinterpret_cast<const float4*>(&wmax[2][c0 + j4]);
          const float4 d = *reinterpret_cast<const float4*>(&wmax[3][c0 + j4]);
          m[j4 + 0] = fmaxf(fmaxf(a.x, bq.x), fmaxf(cq.x, d.x));
          m[j4 + 1] = fmaxf(fmaxf(a.y, bq.y), fmaxf(cq.y, d.y));
          m[j4 + 2] = fmaxf(fmaxf(a.z, bq.z), fmaxf(cq.z, d.z));
          m[j4 + 3] = fmaxf(fmaxf(a.w, bq.w), fmaxf(cq.w, d.w));
        }
        uint8_t* pr = prow + (c0 >> 3) * 1024;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float ms = m[j] * scale_log2;
          my_ms = (t == c0 + j) ? ms : my_ms;
          const float p = ex2_approx(fmaf(v[j], scale_log2, -ms));      // masked tokens: ex2(-inf) = 0
          *reinterpret_cast<__nv_bfloat16*>(pr + (j >> 3) * 1024 + (j & 7) * 128 + ((bch ^ (j & 7)) << 4) + blo) =
              __float2bfloat16_rn(p);
        }
        if (c0 + 16 < N) bar_sync_group(g);       // wmax is reused by the next 16 heads
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&bars->p_full[g]);
      gqa_wait(&bars->o_full[g], ph, 6, n);
      tc_fence_after();
      // epilogue: thread = value dim d (TMEM lane), columns = heads
      float* po = part_o + pbase * 128 + t;
      const size_t hstride = (size_t)num_splits * 128;
      for (int c0 = 0; c0 < N; c0 += 16) {
        float o[16], l[16];
        __syncwarp();
        tmem_ld16(tm_o + c0, o);
        tmem_ld16(tm_l + c0, l);
        tmem_ld_wait();
        float lm = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (c0 + j < G) po[(size_t)(c0 + j) * hstride] = o[j];
          lm = (t == c0 + j) ? l[j] : lm;
        }
        if (t >= c0 && t < c0 + 16 && t < G) {
          part_ml[(pbase + (size_t)t * num_splits) * 2] = my_ms;
          part_ml[(pbase + (size_t)t * num_splits) * 2 + 1] = lm;
        }
      }
      tc_fence_before();
      ++n;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) tmem_dealloc(bars->tmem_base, 512);
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
    cudaError_t e = cudaFuncSetAttribute(gqa_decode_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GQA_SMEM_MAX);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(gqa_decode_tc)");
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    attr = true;
  }
  const int G = Hq / Hkv;
  const int N = (G + 15) / 16 * 16;
  const int qbytes = N * 128;
  if (N > GQA_MAX_N) return 1;   // not handled here
  const int fixed = 1024 /*align*/ + 1024 /*ones*/ + 2 * ((2 * qbytes + 1023) & ~1023) + (int)sizeof(GqaBars) + 64;
  const int stage_bytes = 2 * qbytes + 4 * TILE_BYTES;
  int nst = (GQA_SMEM_MAX - fixed) / stage_bytes;
  if (nst > GQA_MAX_STAGES) nst = GQA_MAX_STAGES;
  if (nst < 1) {
    set_error("gqa_decode_tc: group size does not fit shared memory");
    return B200_ERR_INVALID;
  }
  const int smem = fixed + nst * stage_bytes;
  const int total = batch * Hkv * num_splits;
  const int grid = total < num_sms ? total : num_sms;
  if ((total + grid - 1) / grid > GQA_MAX_ITEMS) return 1;   // not handled here: caller uses the split-KV CUDA-core path
  gqa_decode_tc_kernel<<<grid, GQA_THREADS, smem, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(q), reinterpret_cast<const __nv_bfloat16*>(kc),
      reinterpret_cast<const __nv_bfloat16*>(vc), seq_lens, page_table, batch, Hq, Hkv, page_size, max_pages,
      sm_scale * 1.4426950408889634f, num_splits, N, nst, part_o, part_ml);
  attn_merge_kernel<128><<<batch * Hq, 32, 0, st>>>(part_o, part_ml, num_splits, reinterpret_cast<__nv_bfloat16*>(out), lse);
  g_launches += 2;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "gqa_decode_tc launch");
  return 0;
}

}  // namespace b200
