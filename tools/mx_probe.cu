// Bring-up probe for the native block-scaled path (round-2 groundwork, not part of the product library):
// ONE tcgen05.mma.kind::mxf8f6f4.block_scale tile  D[128, 32] = A[128, 128 (e2m1, ue8m0 / 32)] . B[32, 128 (e4m3, ue8m0 / 32)]^T
// with the operand layouts under test selected at run time, so that a single GPU run tells which shared-memory
// layout the tensor core expects for 4-bit A operands and where it reads the scale factors in TMEM.
//
//   a_variant 0: 16-byte chunks of 8 packed bytes (16 nibbles, low nibble = even k) + 8 bytes of padding
//                (what TMA's 16U4_ALIGN16B produces)
//   a_variant 1: one element per byte, value in bits [3:0]
//   a_variant 2: one element per byte, value in bits [5:2]
//   sf_variant 0: scale word of row r (4 ue8m0 bytes = the four 32-wide k-groups of the 128-K block) in TMEM lane r,
//                 replicated over 4 consecutive columns (covers "column base" and "column base + lane quadrant")
//   sf_variant 1: only column base is written (others hold 2^0 = 127), to tell the two hypotheses apart
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -shared -Xcompiler -fPIC -I lvllm_b200/csrc \
//              -o tools/libmx_probe.so tools/mx_probe.cu        (tools/mx_probe.py does this)
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

using namespace b200;

namespace {

__device__ __forceinline__ uint32_t idesc_mx(uint32_t n, uint32_t a_sf, uint32_t b_sf) {
  uint32_t d = 0;
  d |= (b_sf & 3u) << 4;          // b_sf_id
  d |= 5u << 7;                   // a_format  = E2M1
  d |= 0u << 10;                  // b_format  = E4M3
  d |= ((n >> 3) & 63u) << 17;    // n_dim
  d |= 1u << 23;                  // scale_format = E8M0
  d |= ((128u >> 4) & 31u) << 24; // m_dim
  d |= (a_sf & 3u) << 29;         // a_sf_id
  return d;
}

__device__ __forceinline__ void umma_mx(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate, uint32_t tmem_sfa, uint32_t tmem_sfb) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
      : "memory");
}

__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}

__global__ void __launch_bounds__(128, 1)
    mx_probe_kernel(const uint8_t* __restrict__ a_packed,   // [128][64]
                    const uint8_t* __restrict__ sa,         // [128][4]
                    const uint8_t* __restrict__ b_fp8,      // [32][128]
                    const uint8_t* __restrict__ sb,         // [32][4]
                    float* __restrict__ out,                // [128][32]
                    int a_variant, int sf_variant) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_t = smem;                 // 128 rows x 128 B, 128B swizzle
  uint8_t* b_t = smem + 16384;         // 32 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 4096);
  uint32_t* tbase = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;

  // ---- operand tiles
  {
    const int r = tid;   // one A row per thread
    for (int c = 0; c < 8; ++c) {
      uint8_t chunk[16];
      for (int i = 0; i < 16; ++i) chunk[i] = 0;
      if (a_variant == 0) {
        for (int i = 0; i < 8; ++i) chunk[i] = a_packed[r * 64 + c * 8 + i];
      } else {
        for (int i = 0; i < 16; ++i) {
          const int k = c * 16 + i;
          const uint8_t byte = a_packed[r * 64 + (k >> 1)];
          const uint8_t nib = (k & 1) ? (byte >> 4) : (byte & 15);
          chunk[i] = (a_variant == 1) ? nib : (uint8_t)(nib << 2);
        }
      }
      *reinterpret_cast<uint4*>(a_t + sw128_offset(r, c * 16)) = *reinterpret_cast<const uint4*>(chunk);
    }
    if (tid < 32) {
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(b_t + sw128_offset(tid, c * 16)) =
            *reinterpret_cast<const uint4*>(b_fp8 + tid * 128 + c * 16);
    }
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tbase, 64);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tbase;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  const uint32_t tm_d = tm, tm_sfa = tm + 32, tm_sfb = tm + 36;

  // ---- scale factors into TMEM (one 32-bit word = the four k-groups of this 128-K block)
  {
    const uint32_t wa = *reinterpret_cast<const uint32_t*>(sa + tid * 4);
    const uint32_t wb = *reinterpret_cast<const uint32_t*>(sb + (tid & 31) * 4);
    const uint32_t one = 0x7f7f7f7fu;   // 2^0 in every byte
    for (int c = 0; c < 4; ++c) {
      const bool real = (sf_variant == 0) || c == 0;
      tmem_st1(tm_sfa + lane_off + c, real ? wa : one);
      tmem_st1(tm_sfb + lane_off + c, real ? wb : one);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // ---- four K=32 MMAs (one per k-group), scale-factor byte selected by sf_id
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t aa = smem_u32(a_t), ba = smem_u32(b_t);
      for (uint32_t ks = 0; ks < 4; ++ks)
        umma_mx(tm_d, umma_desc_sw128(aa + ks * 32, 1024), umma_desc_sw128(ba + ks * 32, 1024), idesc_mx(32, ks, ks),
                ks > 0 ? 1u : 0u, tm_sfa | (ks << 30), tm_sfb | (ks << 30));
      umma_commit(bar);
    }
    __syncwarp();
  }
  {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, 0)) {
      if (++spins > (1u << 22)) __trap();   // never hang the box
    }
  }
  tc_fence_after();
  float v[16];
  for (int c16 = 0; c16 < 2; ++c16) {
    tmem_ld16(tm_d + lane_off + c16 * 16, v);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[tid * 32 + c16 * 16 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 64);
}

}  // namespace

extern "C" int mx_probe_run(const void* a_packed, const void* sa, const void* b_fp8, const void* sb, void* out,
                            int a_variant, int sf_variant, void* stream) {
  const int smem = 16384 + 4096 + 64 + 1024;
  mx_probe_kernel<<<1, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint8_t*>(a_packed), reinterpret_cast<const uint8_t*>(sa),
      reinterpret_cast<const uint8_t*>(b_fp8), reinterpret_cast<const uint8_t*>(sb), reinterpret_cast<float*>(out),
      a_variant, sf_variant);
  return (int)cudaGetLastError();
}
