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
      if (++spins > (1u << 28)) __trap();
    }
  }
  __syncthreads();
  for (int64_t i = lo + threadIdx.x * 4; i < hi; i += blockDim.x * 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < world; ++r) {
      const float4 v = __ldcv(reinterpret_cast<const float4*>(peers.data[r] + (int64_t)par * slot_elems + i));
      acc.x += v.x;
      acc.y += v.y;
      acc.z += v.z;
      acc.w += v.w;
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

}  // namespace b200

using namespace b200;

extern "C" {

int64_t b200_ep_flag_bytes(void) { return (int64_t)EP_FLAG_INTS * 4; }

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

}  // extern "C"
