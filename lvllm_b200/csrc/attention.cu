// Paged decode attention (one query token per request), split-KV + LSE merge.
//   MLA (absorbed): q = [q_nope(512) | q_pe(64)], one shared latent KV head, K = all 576 cols, V = first 512
//                   (op contract of reference csrc/libtorch_stable/attention/mla/sm100_cutlass_mla_kernel.cu:225-262)
//   GQA           : q [B,Hq,D], separate K / V caches [pages,page,Hkv,D]
//                   (reference tests/kernels/attention/test_flashinfer.py:29-80)
//
// v1 kernel (CUDA cores): a CTA owns (request, kv-head, head-group, split); KV tiles of TOK tokens are staged
// ONCE in shared memory with 16-byte coalesced loads and shared by all heads of the group (one warp per q
// head), so HBM traffic == algorithmic KV bytes; the warp does the dot product over its lane-strided dims,
// online softmax in fp32 (exp2 domain), and accumulates V.  Partial (o, m, l) per split are merged by
// a second tiny kernel that also emits the natural-log LSE.
// Roofline: HBM (KV bytes) for GQA at batch; latency/FMA bound for MLA at B=1 (tensor-core version is the
// follow-up).
#include <math_constants.h>

#include <cstdlib>

#include "common.cuh"
#include "moe_internal.cuh"

namespace b200 {

constexpr int ATT_TOK = 16;  // tokens per smem tile

template <int DK, int DV, int WARPS, bool MLA>
__global__ void __launch_bounds__(WARPS * 32)
    decode_attn_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ q2,
                       const __nv_bfloat16* __restrict__ kc, const __nv_bfloat16* __restrict__ vc,
                       const int32_t* __restrict__ seq_lens, const int32_t* __restrict__ page_table, int Hq, int Hkv,
                       int page_size, int max_pages, float scale_log2, int num_splits, float* __restrict__ part_o,
                       float* __restrict__ part_ml) {
  // blockIdx.x = ((b * Hkv + hkv) * groups_per_kv + hg), blockIdx.y = split
  const int G = Hq / Hkv;
  const int groups_per_kv = (G + WARPS - 1) / WARPS;
  int bx = blockIdx.x;
  const int hg = bx % groups_per_kv;
  bx /= groups_per_kv;
  const int hkv = bx % Hkv;
  const int b = bx / Hkv;
  const int split = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head_in_group = hg * WARPS + warp;
  const bool head_ok = head_in_group < G;
  const int hq = hkv * G + (head_ok ? head_in_group : 0);

  const int S = seq_lens[b];
  const int per = ((S + num_splits - 1) / num_splits + ATT_TOK - 1) / ATT_TOK * ATT_TOK;
  const int s0 = split * per;
  const int s1 = min(S, s0 + per);

  __shared__ __align__(16) __nv_bfloat16 ks[ATT_TOK][DK];
  __shared__ __align__(16) __nv_bfloat16 vs[MLA ? 1 : ATT_TOK][MLA ? 8 : DV];

  constexpr int NK = DK / 32, NV = DV / 32;
  float qr[NK];
#pragma unroll
  for (int i = 0; i < NK; ++i) {
    const int d = lane + 32 * i;
    float v;
    if (MLA) {
      v = (d < 512) ? __bfloat162float(q[((size_t)b * Hq + hq) * 512 + d])
                    : __bfloat162float(q2[((size_t)b * Hq + hq) * 64 + (d - 512)]);
    } else {
      v = __bfloat162float(q[((size_t)b * Hq + hq) * DK + d]);
    }
    qr[i] = v * scale_log2;
  }
  float o[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) o[i] = 0.f;
  float m = -CUDART_INF_F, l = 0.f;

  for (int t0 = s0; t0 < s1; t0 += ATT_TOK) {
    const int nt = min(ATT_TOK, s1 - t0);
    __syncthreads();
    // cooperative 16-byte loads of the K (and V) rows of this tile
    constexpr int KCH = DK / 8;
    for (int i = threadIdx.x; i < ATT_TOK * KCH; i += WARPS * 32) {
      const int tt = i / KCH, c = i % KCH;
      if (tt < nt) {
        const int tok = t0 + tt;
        const int page = page_table[(size_t)b * max_pages + tok / page_size];
        const size_t row = (size_t)page * page_size + tok % page_size;
        const __nv_bfloat16* src = MLA ? kc + row * DK : kc + (row * Hkv + hkv) * DK;
        *reinterpret_cast<uint4*>(&ks[tt][c * 8]) = *reinterpret_cast<const uint4*>(src + c * 8);
      }
    }
    if (!MLA) {
      constexpr int VCH = DV / 8;
      for (int i = threadIdx.x; i < ATT_TOK * VCH; i += WARPS * 32) {
        const int tt = i / VCH, c = i % VCH;
        if (tt < nt) {
          const int tok = t0 + tt;
          const int page = page_table[(size_t)b * max_pages + tok / page_size];
          const size_t row = (size_t)page * page_size + tok % page_size;
          *reinterpret_cast<uint4*>(&vs[MLA ? 0 : tt][c * 8]) =
              *reinterpret_cast<const uint4*>(vc + (row * Hkv + hkv) * DV + c * 8);
        }
      }
    }
    __syncthreads();
    if (head_ok) {
      float sc[ATT_TOK];
#pragma unroll
      for (int tt = 0; tt < ATT_TOK; ++tt) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < NK; ++i) acc = fmaf(qr[i], __bfloat162float(ks[tt][lane + 32 * i]), acc);
        sc[tt] = acc;
      }
#pragma unroll
      for (int tt = 0; tt < ATT_TOK; ++tt) {
        sc[tt] = warp_sum(sc[tt]);
        if (tt >= nt) sc[tt] = -CUDART_INF_F;
      }
      float mt = m;
#pragma unroll
      for (int tt = 0; tt < ATT_TOK; ++tt) mt = fmaxf(mt, sc[tt]);
      const float corr = (m == -CUDART_INF_F) ? 0.f : exp2f(m - mt);
      l *= corr;
#pragma unroll
      for (int i = 0; i < NV; ++i) o[i] *= corr;
#pragma unroll
      for (int tt = 0; tt < ATT_TOK; ++tt) {
        const float p = (tt < nt) ? exp2f(sc[tt] - mt) : 0.f;
        l += p;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const float v = MLA ? __bfloat162float(ks[tt][lane + 32 * i]) : __bfloat162float(vs[MLA ? 0 : tt][(lane + 32 * i) % (MLA ? 8 : DV)]);
          o[i] = fmaf(p, v, o[i]);
        }
      }
      m = mt;
    }
  }
  if (head_ok) {
    const size_t base = ((size_t)b * Hq + hq) * num_splits + split;
#pragma unroll
    for (int i = 0; i < NV; ++i) part_o[base * DV + lane + 32 * i] = o[i];
    if (lane == 0) {
      part_ml[base * 2] = m;
      part_ml[base * 2 + 1] = l;
    }
  }
}

// merge splits: one warp per (b, head)
template <int DV>
__global__ void __launch_bounds__(128) decode_merge_kernel(const float* __restrict__ part_o,
                                                          const float* __restrict__ part_ml, int BH, int num_splits,
                                                          __nv_bfloat16* __restrict__ out, float* __restrict__ lse) {
  const int w = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (w >= BH) return;
  float mx = -CUDART_INF_F;
  for (int s = 0; s < num_splits; ++s) mx = fmaxf(mx, part_ml[((size_t)w * num_splits + s) * 2]);
  float lsum = 0.f;
  float acc[DV / 32];
#pragma unroll
  for (int i = 0; i < DV / 32; ++i) acc[i] = 0.f;
  for (int s = 0; s < num_splits; ++s) {
    const size_t base = (size_t)w * num_splits + s;
    const float ms = part_ml[base * 2], ls = part_ml[base * 2 + 1];
    if (ms == -CUDART_INF_F) continue;
    const float f = exp2f(ms - mx);
    lsum += ls * f;
#pragma unroll
    for (int i = 0; i < DV / 32; ++i) acc[i] = fmaf(f, part_o[base * DV + lane + 32 * i], acc[i]);
  }
  const float inv = lsum > 0.f ? 1.f / lsum : 0.f;
#pragma unroll
  for (int i = 0; i < DV / 32; ++i) out[(size_t)w * DV + lane + 32 * i] = __float2bfloat16_rn(acc[i] * inv);
  if (lse && lane == 0) lse[w] = (lsum > 0.f) ? (mx * 0.6931471805599453f + logf(lsum)) : -CUDART_INF_F;
}

// merge for many splits (tensor-core MLA): one CTA per (b, head), thread = float4 of the 512 dims; the split
// weights are computed once in shared memory, partial rows are read 8 splits at a time (independent loads)
__global__ void __launch_bounds__(128) mla_merge_kernel(const float* __restrict__ part_o,
                                                       const float* __restrict__ part_ml, int num_splits,
                                                       __nv_bfloat16* __restrict__ out, float* __restrict__ lse) {
  __shared__ float wgt[512];
  __shared__ float red[4];
  const int w = blockIdx.x, tid = threadIdx.x;
  const float* ml = part_ml + (size_t)w * num_splits * 2;
  float mx = -CUDART_INF_F;
  for (int s = tid; s < num_splits; s += 128) mx = fmaxf(mx, ml[s * 2]);
  mx = warp_max(mx);
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float ls = 0.f;
  for (int s = tid; s < num_splits; s += 128) {
    const float ms = ml[s * 2];
    const float f = (ms == -CUDART_INF_F) ? 0.f : exp2f(ms - mx);
    wgt[s] = f;
    ls += f * ml[s * 2 + 1];
  }
  ls = warp_sum(ls);
  if ((tid & 31) == 0) red[tid >> 5] = ls;
  __syncthreads();
  const float lsum = red[0] + red[1] + red[2] + red[3];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* po = part_o + (size_t)w * num_splits * 512 + tid * 4;
  for (int s0 = 0; s0 < num_splits; s0 += 16) {   // 16 independent 16-byte loads in flight per thread
    float4 v[16];
    float f[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int s = s0 + u;
      f[u] = (s < num_splits) ? wgt[s] : 0.f;
      v[u] = (f[u] != 0.f) ? __ldcs(reinterpret_cast<const float4*>(po + (size_t)s * 512)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
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
  *reinterpret_cast<uint2*>(out + (size_t)w * 512 + tid * 4) = pk;
  if (lse && tid == 0) lse[w] = (lsum > 0.f) ? (mx * 0.6931471805599453f + logf(lsum)) : -CUDART_INF_F;
}

}  // namespace b200

using namespace b200;

extern "C" {

int64_t b200_mla_decode_workspace_bytes(int batch, int num_heads, int num_splits) {
  return (int64_t)batch * num_heads * num_splits * (512 + 2) * 4;
}
int64_t b200_gqa_decode_workspace_bytes(int batch, int num_q_heads, int head_dim, int num_splits) {
  return (int64_t)batch * num_q_heads * num_splits * (head_dim + 2) * 4;
}

int b200_mla_decode(void* stream, const void* q_nope, const void* q_pe, const void* kv_cache,
                    const int32_t* seq_lens, const int32_t* page_table, int batch, int num_heads, int page_size,
                    int max_pages, float sm_scale, int num_splits, void* workspace, void* out, float* lse) {
  return b200_mla_decode_ex(stream, q_nope, q_pe, 0, kv_cache, 0, 1.f, 1.f, seq_lens, page_table, batch, num_heads,
                            page_size, max_pages, sm_scale, num_splits, workspace, out, lse);
}

int b200_mla_decode_ex(void* stream, const void* q_nope, const void* q_pe, int q_dtype, const void* kv_cache,
                       int kv_dtype, float descale_q, float descale_k, const int32_t* seq_lens,
                       const int32_t* page_table, int batch, int num_heads, int page_size, int max_pages, float sm_scale,
                       int num_splits, void* workspace, void* out, float* lse) {
  if (!q_nope || !q_pe || !kv_cache || !seq_lens || !page_table || !workspace || !out || batch <= 0 ||
      num_heads <= 0 || page_size <= 0 || num_splits <= 0 || q_dtype < 0 || q_dtype > 1 || kv_dtype < 0 || kv_dtype > 1) {
    set_error("b200_mla_decode: bad argument");
    return B200_ERR_INVALID;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  constexpr int WARPS = 8;
  float* po = reinterpret_cast<float*>(workspace);
  float* pml = po + (size_t)batch * num_heads * num_splits * 512;
  // tensor-core path: one CTA per (request, 128-token split, value-dim half); needs the split count to cover
  // the longest sequence the page table can describe (seq_lens live on the device)
  static const bool use_tc = []() {
    const char* v = getenv("B200_MLA_DISABLE_TC");
    return !(v && v[0] == '1');
  }();
  if (use_tc && num_heads <= 128) {   // each CTA walks ceil(tiles / num_splits) 128-token tiles (online softmax)
    int rc = launch_mla_tc(st, q_nope, q_pe, kv_cache, seq_lens, page_table, batch, num_heads, page_size, max_pages,
                           sm_scale, num_splits, po, pml, kv_dtype, q_dtype, descale_q, descale_k);
    if (rc) return rc;
    if (num_splits <= 512)
      mla_merge_kernel<<<batch * num_heads, 128, 0, st>>>(po, pml, num_splits, reinterpret_cast<__nv_bfloat16*>(out), lse);
    else
      decode_merge_kernel<512><<<(batch * num_heads + 3) / 4, 128, 0, st>>>(po, pml, batch * num_heads, num_splits,
                                                                           reinterpret_cast<__nv_bfloat16*>(out), lse);
    ++g_launches;
    cudaError_t e2 = cudaGetLastError();
    if (e2 != cudaSuccess) return cuda_fail(e2, "mla merge launch");
    return 0;
  }
  if (kv_dtype || q_dtype) {
    set_error("b200_mla_decode: the e4m3 cache is served by the tensor-core kernel only (num_heads <= 128)");
    return B200_ERR_INVALID;
  }
  const int groups = (num_heads + WARPS - 1) / WARPS;
  dim3 grid(batch * groups, num_splits);
  decode_attn_kernel<576, 512, WARPS, true><<<grid, WARPS * 32, 0, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(q_nope), reinterpret_cast<const __nv_bfloat16*>(q_pe),
      reinterpret_cast<const __nv_bfloat16*>(kv_cache), nullptr, seq_lens, page_table, num_heads, 1, page_size,
      max_pages, sm_scale * 1.4426950408889634f, num_splits, po, pml);
  decode_merge_kernel<512><<<(batch * num_heads + 3) / 4, 128, 0, st>>>(po, pml, batch * num_heads, num_splits,
                                                                       reinterpret_cast<__nv_bfloat16*>(out), lse);
  g_launches += 2;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "mla_decode launch");
  return 0;
}

int b200_gqa_decode(void* stream, const void* q, const void* k_cache, const void* v_cache, const int32_t* seq_lens,
                    const int32_t* page_table, int batch, int num_q_heads, int num_kv_heads, int head_dim,
                    int page_size, int max_pages, float sm_scale, int num_splits, void* workspace, void* out,
                    float* lse) {
  if (!q || !k_cache || !v_cache || !seq_lens || !page_table || !workspace || !out || batch <= 0 ||
      num_q_heads <= 0 || num_kv_heads <= 0 || num_q_heads % num_kv_heads || page_size <= 0 || num_splits <= 0) {
    set_error("b200_gqa_decode: bad argument");
    return B200_ERR_INVALID;
  }
  if (head_dim != 128) {
    set_error("b200_gqa_decode: head_dim must be 128");
    return B200_ERR_INVALID;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int G = num_q_heads / num_kv_heads;
  float* po = reinterpret_cast<float*>(workspace);
  float* pml = po + (size_t)batch * num_q_heads * num_splits * 128;
  const float sl2 = sm_scale * 1.4426950408889634f;
  // tensor-core path (persistent CTAs over (request, kv head, 128-token split) items): needs the split count
  // to cover the longest sequence the page table can describe
  static const bool use_tc = []() {
    const char* v = getenv("B200_GQA_DISABLE_TC");
    return !(v && v[0] == '1');
  }();
  if (use_tc && G <= 128 && num_splits <= 1024 && (int64_t)num_splits * 128 >= (int64_t)max_pages * page_size) {
    const int rc = launch_gqa_tc(st, q, k_cache, v_cache, seq_lens, page_table, batch, num_q_heads, num_kv_heads,
                                 page_size, max_pages, sm_scale, num_splits, po, pml, out, lse);
    if (rc <= 0) return rc;   // rc == 1: more work items per CTA than the kernel's table holds
  }
#define LAUNCH_GQA(W)                                                                                          \
  {                                                                                                            \
    const int groups = (G + W - 1) / W;                                                                        \
    dim3 grid(batch * num_kv_heads * groups, num_splits);                                                      \
    decode_attn_kernel<128, 128, W, false><<<grid, W * 32, 0, st>>>(                                           \
        reinterpret_cast<const __nv_bfloat16*>(q), nullptr, reinterpret_cast<const __nv_bfloat16*>(k_cache),   \
        reinterpret_cast<const __nv_bfloat16*>(v_cache), seq_lens, page_table, num_q_heads, num_kv_heads,      \
        page_size, max_pages, sl2, num_splits, po, pml);                                                       \
  }
  if (G <= 4)
    LAUNCH_GQA(4)
  else if (G <= 8)
    LAUNCH_GQA(8)
  else
    LAUNCH_GQA(16)
#undef LAUNCH_GQA
  decode_merge_kernel<128><<<(batch * num_q_heads + 3) / 4, 128, 0, st>>>(
      po, pml, batch * num_q_heads, num_splits, reinterpret_cast<__nv_bfloat16*>(out), lse);
  g_launches += 2;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "gqa_decode launch");
  return 0;
}

}  // extern "C"
