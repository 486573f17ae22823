// Expert-parallel combine over NVLink peer memory (CUDA IPC), hand-written: the lk_moe EP/TP contract is
// "replicated tokens, local experts, sum over ranks" (reference moe_runner.py:488-494 ->
// tensor_model_parallel_all_reduce; P2P pattern of reference csrc/custom_all_reduce.cuh:7-60).
//
// One-shot all-reduce for decode-sized payloads: every rank publishes its fp32 partial in its own
// IPC-exported buffer, raises a per-(CTA, source) flag in every peer with st.release.sys, waits for the
// peers' flags with ld.acquire.sys, then pulls the peers' slices over NVLink (ld.cv, peer data must not be
// served from a stale L1 line) and sums them in FIXED rank order, so all ranks produce bit-identical
// results.  Data slots and flags are double-buffered by epoch parity; the epoch lives in device memory so
// the kernel is CUDA-graph replayable.
#include "common.cuh"
#include "moe_internal.cuh"

namespace b200 {

constexpr int EP_MAX_WORLD = 8;
constexpr int EP_MAX_CTAS = 64;
constexpr int EP_FLAG_INTS = 2 * EP_MAX_WORLD * EP_MAX_CTAS + EP_MAX_CTAS;  // flags + per-CTA epoch counters
constexpr int EP_KINDS = 4;   // independent flag/epoch sets: 0 all-reduce, 1 dispatch, 2 combine, 3 push all-reduce + norm

struct EpPeers {
  float* data[EP_MAX_WORLD];
  int32_t* flags[EP_MAX_WORLD];
};

B200_DEVICE void st_release_sys(int32_t* p, int32_t v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
B200_DEVICE int32_t ld_acquire_sys(const int32_t* p) {
  int32_t v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(512, 1)
    ep_allreduce_kernel(EpPeers peers, int world, int rank, const float* __restrict__ local_in, int64_t numel,
                        int64_t slot_elems, void* __restrict__ out, int out_dtype) {
  const int c = blockIdx.x;
  int32_t* my_flags = peers.flags[rank];
  int32_t* epoch_ctr = my_flags + 2 * EP_MAX_WORLD * EP_MAX_CTAS;
  __shared__ int32_t s_epoch;
  if (threadIdx.x == 0) s_epoch = epoch_ctr[c] + 1;
  __syncthreads();
  const int32_t epoch = s_epoch;
  const int par = epoch & 1;
  const int64_t per = ((numel + gridDim.x - 1) / gridDim.x + 3) & ~int64_t(3);
  const int64_t lo = (int64_t)c * per;
  const int64_t hi = (lo + per < numel) ? lo + per : numel;
  float* mine = peers.data[rank] + (int64_t)par * slot_elems;
  for (int64_t i = lo + threadIdx.x * 4; i < hi; i += blockDim.x * 4)
    *reinterpret_cast<float4*>(mine + i) = *reinterpret_cast<const float4*>(local_in + i);
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < world) {
    const int j = threadIdx.x;
    st_release_sys(peers.flags[j] + (par * EP_MAX_WORLD + rank) * EP_MAX_CTAS + c, epoch);
    const int32_t* f = my_flags + (par * EP_MAX_WORLD + j) * EP_MAX_CTAS + c;
    unsigned spins = 0;
    while (ld_acquire_sys(f) < epoch) {
      if (++spins > 4096) __nanosleep(256);   // a slow peer is legal: back off, never kill the context
    }
  }
  __syncthreads();
  for (int64_t i = lo + threadIdx.x * 4; i < hi; i += blockDim.x * 4) {
    // every peer's load is issued before the first add (a load-add chain per peer would serialise the NVLink round trips)
    float4 v[EP_MAX_WORLD];
#pragma unroll
    for (int r = 0; r < EP_MAX_WORLD; ++r)
      v[r] = r < world ? __ldcv(reinterpret_cast<const float4*>(peers.data[r] + (int64_t)par * slot_elems + i))
                       : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < EP_MAX_WORLD; ++r) {
      acc.x += v[r].x;
      acc.y += v[r].y;
      acc.z += v[r].z;
      acc.w += v[r].w;
    }
    if (out_dtype == 2) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + i) = acc;
    } else if (out_dtype == 0) {
      __nv_bfloat162 a = __floats2bfloat162_rn(acc.x, acc.y), b = __floats2bfloat162_rn(acc.z, acc.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&a);
      pk.y = *reinterpret_cast<uint32_t*>(&b);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + i) = pk;
    } else {
      __half2 a = __floats2half2_rn(acc.x, acc.y), b = __floats2half2_rn(acc.z, acc.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&a);
      pk.y = *reinterpret_cast<uint32_t*>(&b);
      *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(out) + i) = pk;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) epoch_ctr[c] = epoch;
}

// ------------------------------------------------------------------------------------------------------------
// Push all-reduce fused with residual add + RMSNorm (reference moe_runner.py:462-496 followed by
// the layer's fused_add_rms_norm; vllm/distributed/device_communicators/flashinfer_all_reduce.py is the kernel to beat).
// Decode-sized payloads are latency bound, so the partial is PUSHED: every rank stores its fp32 rows straight into
// slot [parity][rank] of every peer's buffer (posted NVLink writes, no request/response round trip), raises the flags,
// and after the flag wait reduces from LOCAL memory in fixed rank order (bit-identical on all ranks).  Each CTA owns
// whole token rows, so the row's RMS statistics never leave the CTA:
//     x = sum_r partial_r (+ residual);  residual_out = cast(x);  out = cast(x * rsqrt(mean(x^2) + eps)) * gamma (| * gain)
// (operation order of the reference's RMSNorm.forward_native).  Buffer layout: 2 slots of the pull all-reduce, then
// [2 parities][world][slot_elems] fp32.
B200_DEVICE float ep_to_f32(const void* p, int fp16, size_t i) {
  return fp16 ? __half2float(reinterpret_cast<const __half*>(p)[i]) : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}
B200_DEVICE void ep_store16(void* p, int fp16, size_t i, float v) {
  if (fp16)
    reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
  else
    reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
}

__global__ void __launch_bounds__(512, 1)
    ep_allreduce_norm_kernel(EpPeers peers, int world, int rank, const float* __restrict__ local_in, int M, int H,
                             int64_t slot_elems, void* __restrict__ residual, const void* __restrict__ gamma, float gain,
                             float eps, void* __restrict__ out, float* __restrict__ sum_out, int fp16) {
  const int c = blockIdx.x;
  int32_t* my_flags = peers.flags[rank] + 3 * EP_FLAG_INTS;
  int32_t* epoch_ctr = my_flags + 2 * EP_MAX_WORLD * EP_MAX_CTAS;
  __shared__ int32_t s_epoch;
  __shared__ float s_red[16];
  if (threadIdx.x == 0) s_epoch = epoch_ctr[c] + 1;
  __syncthreads();
  const int32_t epoch = s_epoch;
  const int par = epoch & 1;
  const int H4 = H >> 2;
  // ---- push this rank's rows of the CTA's tokens to every rank (including itself)
  for (int t = c; t < M; t += gridDim.x) {
    const float4* src = reinterpret_cast<const float4*>(local_in + (size_t)t * H);
    for (int i = threadIdx.x; i < H4; i += blockDim.x) {
      const float4 v = src[i];
      for (int d = 0; d < world; ++d)
        reinterpret_cast<float4*>(peers.data[d] + (2 + (int64_t)par * world + rank) * slot_elems + (size_t)t * H)[i] = v;
    }
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < world) {
    const int j = threadIdx.x;
    st_release_sys(peers.flags[j] + 3 * EP_FLAG_INTS + (par * EP_MAX_WORLD + rank) * EP_MAX_CTAS + c, epoch);
    const int32_t* f = my_flags + (par * EP_MAX_WORLD + j) * EP_MAX_CTAS + c;
    unsigned long long t0 = 0;
    unsigned spins = 0;
    while (ld_acquire_sys(f) < epoch) {
      // a slow peer (another process building its model, a preempted context) is legal: back off instead of trapping
      if (++spins > 4096) __nanosleep(256);
      (void)t0;
    }
  }
  __syncthreads();
  // ---- reduce from local memory + residual + RMSNorm, one token row at a time
  const float* mine = peers.data[rank] + (2 + (int64_t)par * world) * slot_elems;   // slots 0, 1 belong to the pull all-reduce
  for (int t = c; t < M; t += gridDim.x) {
    float4 x[4];
    float ss = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = threadIdx.x + u * blockDim.x;
      x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < H4) {
        for (int r = 0; r < world; ++r) {
          const float4 v = __ldcv(reinterpret_cast<const float4*>(mine + (int64_t)r * slot_elems + (size_t)t * H) + i);
          x[u].x += v.x;
          x[u].y += v.y;
          x[u].z += v.z;
          x[u].w += v.w;
        }
        const size_t o = (size_t)t * H + (size_t)i * 4;
        if (residual) {
          x[u].x += ep_to_f32(residual, fp16, o);
          x[u].y += ep_to_f32(residual, fp16, o + 1);
          x[u].z += ep_to_f32(residual, fp16, o + 2);
          x[u].w += ep_to_f32(residual, fp16, o + 3);
          ep_store16(residual, fp16, o, x[u].x);
          ep_store16(residual, fp16, o + 1, x[u].y);
          ep_store16(residual, fp16, o + 2, x[u].z);
          ep_store16(residual, fp16, o + 3, x[u].w);
        }
        if (sum_out) *reinterpret_cast<float4*>(sum_out + o) = x[u];
        ss += x[u].x * x[u].x + x[u].y * x[u].y + x[u].z * x[u].z + x[u].w * x[u].w;
      }
    }
    ss = warp_sum(ss);
    __syncthreads();   // s_red of the previous row is consumed
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) tot += s_red[i];
    const float rs = rsqrtf(tot / (float)H + eps);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = threadIdx.x + u * blockDim.x;
      if (i < H4) {
        const size_t o = (size_t)t * H + (size_t)i * 4;
        const size_t h = (size_t)i * 4;
        const float v[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          // reference forward_native: x.to(orig_dtype) * weight
          float y = v[q] * rs;
          if (gamma) {
            const float yr = fp16 ? __half2float(__float2half_rn(y)) : __bfloat162float(__float2bfloat16_rn(y));
            y = yr * ep_to_f32(gamma, fp16, h + q);
          } else {
            y *= gain;
          }
          ep_store16(out, fp16, o + q, y);
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) epoch_ctr[c] = epoch;
}

// ------------------------------------------------------------------------------------------------------------
// Dispatch / combine all-to-all for token-sharded callers (DP attention + EP experts): rank r owns
// tokens [r*m_local, (r+1)*m_local) of the global batch M = world*m_local and experts [r*epr, (r+1)*epr).
// Static slots: every rank's data buffer holds X [M][H] 16-bit, IDS [M][k] i32 (expert ids local to that rank,
// -1 = not here), W [M][k] f32 and Y [M][H] f32.
//   dispatch: the token owner pushes each row over NVLink to the <= k ranks that own one of its experts
//             (plus the remapped ids / weights to every rank), then raises per-(CTA, source) flags
//             (st.release.sys) and waits for the peers' flags: when the kernel retires the local X/IDS/W are
//             complete and the local MoE forward can run on them unchanged (ids < 0 are skipped).
//   combine:  after the local MoE wrote Y, the token owner pulls the partial rows from exactly those ranks and
//             sums them in ascending rank order (deterministic), writing [m_local][H] in the caller's dtype.
// Two barriers per layer make single buffering safe: a peer can only push layer L+1 rows after its combine(L)
// saw this rank's Y-ready flag (local MoE(L) retired), and the local MoE(L+1) can only overwrite Y after the
// local dispatch(L+1) saw every peer's flag (their combine(L) pulls retired).
struct EpA2APeers {
  uint8_t* data[EP_MAX_WORLD];
  int32_t* flags[EP_MAX_WORLD];
};
struct EpA2ALayout {
  int64_t off_x, off_ids, off_w, off_y, total;
};
__host__ __device__ inline EpA2ALayout ep_a2a_layout(int M, int H, int k) {
  EpA2ALayout l;
  auto up = [](int64_t v) { return (v + 255) & ~int64_t(255); };
  l.off_x = 0;
  l.off_ids = up((int64_t)M * H * 2);
  l.off_w = l.off_ids + up((int64_t)M * k * 4);
  l.off_y = l.off_w + up((int64_t)M * k * 4);
  l.total = l.off_y + up((int64_t)M * H * 4);
  return l;
}

B200_DEVICE int32_t ep_epoch_begin(int32_t* epoch_ctr, int c) {
  __shared__ int32_t s_epoch;
  if (threadIdx.x == 0) s_epoch = epoch_ctr[c] + 1;
  __syncthreads();
  return s_epoch;
}
// publish `epoch` to every peer's flag (kind block already applied to the pointers) and wait for theirs
B200_DEVICE void ep_flag_exchange(const EpA2APeers& peers, int kind, int world, int rank, int c, int32_t epoch) {
  if ((int)threadIdx.x < world) {
    const int j = threadIdx.x;
    st_release_sys(peers.flags[j] + kind * EP_FLAG_INTS + rank * EP_MAX_CTAS + c, epoch);
    const int32_t* f = peers.flags[rank] + kind * EP_FLAG_INTS + j * EP_MAX_CTAS + c;
    unsigned spins = 0;
    while (ld_acquire_sys(f) < epoch) {
      if (++spins > 4096) __nanosleep(256);   // a slow peer is legal: back off, never kill the context
    }
  }
}

__global__ void __launch_bounds__(256, 1)
    ep_dispatch_kernel(EpA2APeers peers, int world, int rank, const uint16_t* __restrict__ hidden,
                       const int32_t* __restrict__ ids, const float* __restrict__ weights, int m_local, int k, int H,
                       int epr) {
  const int c = blockIdx.x;
  int32_t* epoch_ctr = peers.flags[rank] + 1 * EP_FLAG_INTS + 2 * EP_MAX_WORLD * EP_MAX_CTAS;
  const int32_t epoch = ep_epoch_begin(epoch_ctr, c);
  const EpA2ALayout lay = ep_a2a_layout(world * m_local, H, k);
  for (int t = c; t < m_local; t += gridDim.x) {
    const int64_t row = (int64_t)rank * m_local + t;
    uint32_t mask = 0;
    for (int j = 0; j < k; ++j) {
      const int id = ids[(int64_t)t * k + j];
      if (id >= 0 && id / epr < world) mask |= 1u << (id / epr);
    }
    const uint4* src = reinterpret_cast<const uint4*>(hidden + (int64_t)t * H);
    for (int d = 0; d < world; ++d) {
      if (mask >> d & 1) {
        uint4* dst = reinterpret_cast<uint4*>(peers.data[d] + lay.off_x + row * H * 2);
        for (int i = threadIdx.x; i < H / 8; i += blockDim.x) dst[i] = src[i];
      }
      if ((int)threadIdx.x < k) {
        const int id = ids[(int64_t)t * k + threadIdx.x];
        const bool here = id >= 0 && id / epr == d;
        reinterpret_cast<int32_t*>(peers.data[d] + lay.off_ids)[row * k + threadIdx.x] = here ? id - d * epr : -1;
        reinterpret_cast<float*>(peers.data[d] + lay.off_w)[row * k + threadIdx.x] = weights[(int64_t)t * k + threadIdx.x];
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  ep_flag_exchange(peers, 1, world, rank, c, epoch);
  __syncthreads();
  if (threadIdx.x == 0) epoch_ctr[c] = epoch;
}

__global__ void __launch_bounds__(256, 1)
    ep_combine_kernel(EpA2APeers peers, int world, int rank, const int32_t* __restrict__ ids, int m_local, int k, int H,
                      int epr, void* __restrict__ out, int out_dtype) {
  const int c = blockIdx.x;
  int32_t* epoch_ctr = peers.flags[rank] + 2 * EP_FLAG_INTS + 2 * EP_MAX_WORLD * EP_MAX_CTAS;
  const int32_t epoch = ep_epoch_begin(epoch_ctr, c);
  const EpA2ALayout lay = ep_a2a_layout(world * m_local, H, k);
  __threadfence_system();      // the local MoE's Y (earlier kernels of this stream) before the Y-ready flag
  ep_flag_exchange(peers, 2, world, rank, c, epoch);
  __syncthreads();
  for (int t = c; t < m_local; t += gridDim.x) {
    const int64_t row = (int64_t)rank * m_local + t;
    uint32_t mask = 0;
    for (int j = 0; j < k; ++j) {
      const int id = ids[(int64_t)t * k + j];
      if (id >= 0 && id / epr < world) mask |= 1u << (id / epr);
    }
    for (int i = threadIdx.x; i < H / 4; i += blockDim.x) {
      float4 v[EP_MAX_WORLD];
#pragma unroll
      for (int d = 0; d < EP_MAX_WORLD; ++d)
        v[d] = (d < world && (mask >> d & 1)) ? __ldcv(reinterpret_cast<const float4*>(peers.data[d] + lay.off_y + row * H * 4) + i)
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int d = 0; d < EP_MAX_WORLD; ++d) {   // ascending rank order; skipped ranks contribute +0
        acc.x += v[d].x;
        acc.y += v[d].y;
        acc.z += v[d].z;
        acc.w += v[d].w;
      }
      const int64_t o = (int64_t)t * H + i * 4;
      if (out_dtype == 2) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + o) = acc;
      } else if (out_dtype == 0) {
        __nv_bfloat162 a = __floats2bfloat162_rn(acc.x, acc.y), b = __floats2bfloat162_rn(acc.z, acc.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&a);
        pk.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + o) = pk;
      } else {
        __half2 a = __floats2half2_rn(acc.x, acc.y), b = __floats2half2_rn(acc.z, acc.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&a);
        pk.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(out) + o) = pk;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) epoch_ctr[c] = epoch;
}

// combine fused with residual add + RMSNorm (the op pair that follows the MoE block): the token owner pulls the partial
// rows of its tokens, sums them in ascending rank order and normalises the row it already holds in registers — the
// fp32 [m_local, H] combine output never goes to HBM.  Same flag protocol / epochs as ep_combine_kernel (kind 2).
__global__ void __launch_bounds__(256, 1)
    ep_combine_norm_kernel(EpA2APeers peers, int world, int rank, const int32_t* __restrict__ ids, int m_local, int k, int H,
                           int epr, void* __restrict__ residual, const void* __restrict__ gamma, float gain, float eps,
                           void* __restrict__ out, int fp16) {
  const int c = blockIdx.x;
  int32_t* epoch_ctr = peers.flags[rank] + 2 * EP_FLAG_INTS + 2 * EP_MAX_WORLD * EP_MAX_CTAS;
  const int32_t epoch = ep_epoch_begin(epoch_ctr, c);
  const EpA2ALayout lay = ep_a2a_layout(world * m_local, H, k);
  __shared__ float s_red[8];
  __threadfence_system();      // the local MoE's Y (earlier kernels of this stream) before the Y-ready flag
  ep_flag_exchange(peers, 2, world, rank, c, epoch);
  __syncthreads();
  const int H4 = H >> 2;
  for (int t = c; t < m_local; t += gridDim.x) {
    const int64_t row = (int64_t)rank * m_local + t;
    uint32_t mask = 0;
    for (int j = 0; j < k; ++j) {
      const int id = ids[(int64_t)t * k + j];
      if (id >= 0 && id / epr < world) mask |= 1u << (id / epr);
    }
    float4 x[8];
    float ss = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = threadIdx.x + u * 256;
      x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < H4) {
        float4 v[EP_MAX_WORLD];
#pragma unroll
        for (int d = 0; d < EP_MAX_WORLD; ++d)
          v[d] = (d < world && (mask >> d & 1)) ? __ldcv(reinterpret_cast<const float4*>(peers.data[d] + lay.off_y + row * H * 4) + i)
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int d = 0; d < EP_MAX_WORLD; ++d) {
          x[u].x += v[d].x;
          x[u].y += v[d].y;
          x[u].z += v[d].z;
          x[u].w += v[d].w;
        }
        const size_t o = (size_t)t * H + (size_t)i * 4;
        if (residual) {
          x[u].x += ep_to_f32(residual, fp16, o);
          x[u].y += ep_to_f32(residual, fp16, o + 1);
          x[u].z += ep_to_f32(residual, fp16, o + 2);
          x[u].w += ep_to_f32(residual, fp16, o + 3);
          ep_store16(residual, fp16, o, x[u].x);
          ep_store16(residual, fp16, o + 1, x[u].y);
          ep_store16(residual, fp16, o + 2, x[u].z);
          ep_store16(residual, fp16, o + 3, x[u].w);
        }
        ss += x[u].x * x[u].x + x[u].y * x[u].y + x[u].z * x[u].z + x[u].w * x[u].w;
      }
    }
    ss = warp_sum(ss);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += s_red[i];
    const float rs = rsqrtf(tot / (float)H + eps);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = threadIdx.x + u * 256;
      if (i < H4) {
        const size_t o = (size_t)t * H + (size_t)i * 4, h = (size_t)i * 4;
        const float v[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float y = v[q] * rs;
          if (gamma) {
            const float yr = fp16 ? __half2float(__float2half_rn(y)) : __bfloat162float(__float2bfloat16_rn(y));
            y = yr * ep_to_f32(gamma, fp16, h + q);
          } else {
            y *= gain;
          }
          ep_store16(out, fp16, o + q, y);
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) epoch_ctr[c] = epoch;
}

}  // namespace b200

using namespace b200;

extern "C" {

int64_t b200_ep_flag_bytes(void) { return (int64_t)EP_KINDS * EP_FLAG_INTS * 4; }

int b200_ep_buffer_create(int64_t bytes, void** dev_ptr, void* ipc_handle_64B) {
  if (bytes <= 0 || !dev_ptr || !ipc_handle_64B) {
    set_error("b200_ep_buffer_create: bad argument");
    return B200_ERR_INVALID;
  }
  cudaError_t e = cudaMalloc(dev_ptr, bytes);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(ep buffer)");
  if ((e = cudaMemset(*dev_ptr, 0, bytes)) != cudaSuccess) return cuda_fail(e, "cudaMemset(ep buffer)");
  cudaIpcMemHandle_t h;
  if ((e = cudaIpcGetMemHandle(&h, *dev_ptr)) != cudaSuccess) return cuda_fail(e, "cudaIpcGetMemHandle");
  static_assert(sizeof(h) == 64, "ipc handle size");
  memcpy(ipc_handle_64B, &h, 64);
  return 0;
}

int b200_ep_buffer_open(const void* ipc_handle_64B, void** dev_ptr) {
  if (!ipc_handle_64B || !dev_ptr) {
    set_error("b200_ep_buffer_open: bad argument");
    return B200_ERR_INVALID;
  }
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle_64B, 64);
  cudaError_t e = cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return cuda_fail(e, "cudaIpcOpenMemHandle");
  return 0;
}

int b200_ep_buffer_close(void* dev_ptr, int is_owner) {
  if (!dev_ptr) return 0;
  cudaError_t e = is_owner ? cudaFree(dev_ptr) : cudaIpcCloseMemHandle(dev_ptr);
  if (e != cudaSuccess) return cuda_fail(e, "ep buffer close");
  return 0;
}

int b200_ep_allreduce(void* stream, void* const* peer_bufs, int32_t* const* peer_flags, int world, int rank,
                      const float* local_in, int64_t numel, int64_t slot_elems, void* out, int out_dtype) {
  if (!peer_bufs || !peer_flags || world < 1 || world > EP_MAX_WORLD || rank < 0 || rank >= world || !local_in ||
      !out || numel <= 0 || numel % 4 || slot_elems < numel || out_dtype < 0 || out_dtype > 2) {
    set_error("b200_ep_allreduce: bad argument (world <= 8, numel % 4 == 0, slot_elems >= numel)");
    return B200_ERR_INVALID;
  }
  EpPeers p;
  for (int r = 0; r < EP_MAX_WORLD; ++r) {
    p.data[r] = r < world ? reinterpret_cast<float*>(peer_bufs[r]) : nullptr;
    p.flags[r] = r < world ? peer_flags[r] : nullptr;
  }
  int ctas = (int)((numel + 8191) / 8192);
  if (ctas < 1) ctas = 1;
  if (ctas > EP_MAX_CTAS) ctas = EP_MAX_CTAS;
  ep_allreduce_kernel<<<ctas, 512, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p, world, rank, local_in, numel,
                                                                               slot_elems, out, out_dtype);
  ++g_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "ep_allreduce launch");
  return 0;
}

int b200_ep_allreduce_norm(void* stream, void* const* peer_bufs, int32_t* const* peer_flags, int world, int rank,
                           const float* local_in, int num_tokens, int hidden, int64_t slot_elems, void* residual,
                           const void* gamma, float gain, float eps, void* out, float* sum_out, int act_dtype) {
  if (!peer_bufs || !peer_flags || world < 1 || world > EP_MAX_WORLD || rank < 0 || rank >= world || !local_in || !out ||
      num_tokens <= 0 || hidden <= 0 || hidden % 4 || hidden > 8192 || (int64_t)num_tokens * hidden > slot_elems ||
      (act_dtype != B200_ACT_BF16 && act_dtype != B200_ACT_FP16)) {
    set_error("b200_ep_allreduce_norm: bad argument (world <= 8, hidden % 4 == 0, hidden <= 8192, tokens*hidden <= slot_elems)");
    return B200_ERR_INVALID;
  }
  EpPeers p;
  for (int r = 0; r < EP_MAX_WORLD; ++r) {
    p.data[r] = r < world ? reinterpret_cast<float*>(peer_bufs[r]) : nullptr;
    p.flags[r] = r < world ? peer_flags[r] : nullptr;
  }
  const int ctas = num_tokens < EP_MAX_CTAS ? num_tokens : EP_MAX_CTAS;
  ep_allreduce_norm_kernel<<<ctas, 512, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      p, world, rank, local_in, num_tokens, hidden, slot_elems, residual, gamma, gain, eps, out, sum_out,
      act_dtype == B200_ACT_FP16);
  ++g_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "ep_allreduce_norm launch");
  return 0;
}

int64_t b200_ep_a2a_layout(int global_tokens, int hidden, int top_k, int64_t* off_x, int64_t* off_ids, int64_t* off_w,
                           int64_t* off_y) {
  if (global_tokens <= 0 || hidden <= 0 || top_k <= 0) return 0;
  const EpA2ALayout l = ep_a2a_layout(global_tokens, hidden, top_k);
  if (off_x) *off_x = l.off_x;
  if (off_ids) *off_ids = l.off_ids;
  if (off_w) *off_w = l.off_w;
  if (off_y) *off_y = l.off_y;
  return l.total;
}

static int ep_a2a_check(const char* what, void* const* peer_bufs, int32_t* const* peer_flags, int world, int rank,
                        int m_local, int k, int H, int epr, EpA2APeers* p) {
  if (!peer_bufs || !peer_flags || world < 1 || world > EP_MAX_WORLD || rank < 0 || rank >= world || m_local <= 0 ||
      k <= 0 || k > 32 || H <= 0 || H % 8 || epr <= 0) {
    set_error(std::string(what) + ": bad argument (world <= 8, top_k <= 32, hidden % 8 == 0)");
    return B200_ERR_INVALID;
  }
  for (int r = 0; r < EP_MAX_WORLD; ++r) {
    p->data[r] = r < world ? reinterpret_cast<uint8_t*>(peer_bufs[r]) : nullptr;
    p->flags[r] = r < world ? peer_flags[r] : nullptr;
  }
  return 0;
}

int b200_ep_dispatch(void* stream, void* const* peer_bufs, int32_t* const* peer_flags, int world, int rank,
                     const void* hidden_local, const int32_t* ids_global, const float* weights, int m_local, int top_k,
                     int hidden, int experts_per_rank) {
  EpA2APeers p;
  int rc = ep_a2a_check("b200_ep_dispatch", peer_bufs, peer_flags, world, rank, m_local, top_k, hidden,
                        experts_per_rank, &p);
  if (rc) return rc;
  if (!hidden_local || !ids_global || !weights) {
    set_error("b200_ep_dispatch: null pointer");
    return B200_ERR_INVALID;
  }
  const int ctas = m_local < EP_MAX_CTAS ? m_local : EP_MAX_CTAS;
  ep_dispatch_kernel<<<ctas, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      p, world, rank, reinterpret_cast<const uint16_t*>(hidden_local), ids_global, weights, m_local, top_k, hidden,
      experts_per_rank);
  ++g_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "ep_dispatch launch");
  return 0;
}

int b200_ep_combine(void* stream, void* const* peer_bufs, int32_t* const* peer_flags, int world, int rank,
                    const int32_t* ids_global, int m_local, int top_k, int hidden, int experts_per_rank, void* out,
                    int out_dtype) {
  EpA2APeers p;
  int rc = ep_a2a_check("b200_ep_combine", peer_bufs, peer_flags, world, rank, m_local, top_k, hidden,
                        experts_per_rank, &p);
  if (rc) return rc;
  if (!ids_global || !out || out_dtype < 0 || out_dtype > 2) {
    set_error("b200_ep_combine: bad argument");
    return B200_ERR_INVALID;
  }
  const int ctas = m_local < EP_MAX_CTAS ? m_local : EP_MAX_CTAS;
  ep_combine_kernel<<<ctas, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p, world, rank, ids_global, m_local, top_k,
                                                                             hidden, experts_per_rank, out, out_dtype);
  ++g_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "ep_combine launch");
  return 0;
}

}  // extern "C"

extern "C" int b200_ep_combine_norm(void* stream, void* const* peer_bufs, int32_t* const* peer_flags, int world, int rank,
                                    const int32_t* ids_global, int m_local, int top_k, int hidden, int experts_per_rank,
                                    void* residual, const void* gamma, float gain, float eps, void* out, int act_dtype) {
  EpA2APeers p;
  int rc = ep_a2a_check("b200_ep_combine_norm", peer_bufs, peer_flags, world, rank, m_local, top_k, hidden, experts_per_rank, &p);
  if (rc) return rc;
  if (!ids_global || !out || hidden > 8192 || (act_dtype != B200_ACT_BF16 && act_dtype != B200_ACT_FP16)) {
    set_error("b200_ep_combine_norm: bad argument (hidden <= 8192)");
    return B200_ERR_INVALID;
  }
  const int ctas = m_local < EP_MAX_CTAS ? m_local : EP_MAX_CTAS;
  ep_combine_norm_kernel<<<ctas, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      p, world, rank, ids_global, m_local, top_k, hidden, experts_per_rank, residual, gamma, gain, eps, out,
      act_dtype == B200_ACT_FP16);
  ++g_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "ep_combine_norm launch");
  return 0;
}
