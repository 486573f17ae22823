// Bring-up probe (run on the B200 box): establishes on hardware
//   (1) that a CU_TENSOR_MAP_DATA_TYPE_16U4_ALIGN16B tensor map with SWIZZLE_128B over densely packed e2m1 bytes
//       ([rows][64 B], box 128 elements x 256 rows) produces the shared-memory image tcgen05.mma kind::mxf8f6f4 wants
//       (16-byte chunks = 8 packed bytes + 8 padding bytes inside the K-major 128B-swizzled tile, the image
//       tools/mx_probe.cu a_variant 0 built by hand), and how many bytes the mbarrier transaction counts;
//   (2) the shared-memory layout tcgen05.cp.32x128b.warpx4 expects for block-scale words (smem -> TMEM without
//       going through registers).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tma_probe tools/tma_probe.cu
//   tools/tma_probe tma <tx_bytes> | tools/tma_probe cp <lbo16> <sbo16>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      printf("CUDA error %s at %s:%d (%s)\n", cudaGetErrorName(e_), __FILE__, __LINE__, #x);    \
      return 2;                                                                                 \
    }                                                                                           \
  } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void tma_kernel(const __grid_constant__ CUtensorMap tm, int row0, uint32_t tx, uint8_t* out, int* status) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0xEEEEEEEEu;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(tx) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            s32(smem)),
        "l"(&tm), "r"(0), "r"(row0), "r"(s32(&bar))
        : "memory");
    long long t0 = clock64();
    uint32_t ok = 0;
    while (!ok && clock64() - t0 < 4000000ll) {
      asm volatile(
          "{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], 0;\n\tselp.b32 %0, 1, 0, P;\n\t}\n"
          : "=r"(ok)
          : "r"(s32(&bar))
          : "memory");
    }
    status[0] = (int)ok;
    status[1] = (int)(clock64() - t0);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 32768; i += blockDim.x) out[i] = smem[i];
}

// ----------------------------------------------------------------------------------------------- tcgen05.cp
__global__ void cp_kernel(uint32_t lbo16, uint32_t sbo16, uint32_t* out) {
  __shared__ __align__(1024) uint32_t sf[128];   // 512 B: 32 rows x 16 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const int tid = threadIdx.x, warp = tid >> 5;
  sf[tid] = 0xA0000000u | (uint32_t)tid;   // word index = tid: row i = tid/4, column c = tid%4
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(s32(&tbase)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tbase;
  // clear the 8 columns first so that stale TMEM content cannot fake a match
  {
    const uint32_t z = 0x55555555u;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(tb + ((uint32_t)(warp * 32) << 16)),
                 "r"(z)
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    uint64_t d = 0;
    d |= (uint64_t)((s32(sf) >> 4) & 0x3FFF);
    d |= (uint64_t)(lbo16 & 0x3FFF) << 16;
    d |= (uint64_t)(sbo16 & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(tb), "l"(d) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
    uint32_t ok = 0;
    long long t0 = clock64();
    while (!ok && clock64() - t0 < 4000000ll) {
      asm volatile(
          "{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], 0;\n\tselp.b32 %0, 1, 0, P;\n\t}\n"
          : "=r"(ok)
          : "r"(s32(&bar))
          : "memory");
    }
    out[128 * 8] = ok;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(tb + ((uint32_t)(warp * 32) << 16))
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int c = 0; c < 8; ++c) out[tid * 8 + c] = r[c];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tb) : "memory");
}

static uint32_t sw128(uint32_t r, uint32_t b) {
  return (r >> 3) * 1024u + (r & 7u) * 128u + ((((b >> 4) ^ (r & 7u)) & 7u) << 4) + (b & 15u);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  if (argc < 2) {
    printf("usage: tma_probe tma <tx_bytes> | cp <lbo16> <sbo16>\n");
    return 1;
  }
  CK(cudaSetDevice(0));
  CK(cudaFree(0));
  if (!strcmp(argv[1], "tma")) {
    const uint32_t tx = argc > 2 ? (uint32_t)atoi(argv[2]) : 16384u;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) {
      printf("cuTensorMapEncodeTiled entry point not found (qres %d)\n", (int)qres);
      return 2;
    }
    const int R = 1024;
    std::vector<uint8_t> h((size_t)R * 64);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t *d = nullptr, *dout = nullptr;
    int* dst = nullptr;
    CK(cudaMalloc(&d, h.size()));
    CK(cudaMalloc(&dout, 32768));
    CK(cudaMalloc(&dst, 8));
    CK(cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice));
    CUtensorMap tm;
    cuuint64_t gdim[2] = {128, (cuuint64_t)R};
    cuuint64_t gstr[1] = {64};
    cuuint32_t box[2] = {128, 256};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_16U4_ALIGN16B, 2, d, gdim, gstr, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("cuTensorMapEncodeTiled(16U4_ALIGN16B, dims {128,%d}, stride 64 B, box {128,256}, SWIZZLE_128B) -> %d\n", R, (int)r);
    if (r != CUDA_SUCCESS) return 3;
    CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 34 * 1024));
    const int row0 = 256;
    tma_kernel<<<1, 128, 34 * 1024>>>(tm, row0, tx, dout, dst);
    CK(cudaDeviceSynchronize());
    int st[2];
    std::vector<uint8_t> o(32768);
    CK(cudaMemcpy(st, dst, 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(o.data(), dout, 32768, cudaMemcpyDeviceToHost));
    printf("expect_tx %u: barrier completed = %d after %d cycles\n", tx, st[0], st[1]);
    // expected image: row r (0..255) of the box, 16-byte chunk c: bytes 0..7 = packed[r][8c..8c+7], bytes 8..15 = padding
    size_t bad_data = 0, pad_zero = 0, pad_ee = 0, pad_other = 0, untouched = 0;
    for (int rr = 0; rr < 256; ++rr) {
      for (int c = 0; c < 8; ++c) {
        const uint32_t off = (uint32_t)(rr >> 7) * 16384u + sw128(rr & 127, c * 16);
        for (int b = 0; b < 8; ++b)
          if (o[off + b] != h[(size_t)(row0 + rr) * 64 + c * 8 + b]) ++bad_data;
        for (int b = 8; b < 16; ++b) {
          if (o[off + b] == 0) ++pad_zero;
          else if (o[off + b] == 0xEE) ++pad_ee;
          else ++pad_other;
        }
      }
    }
    for (size_t i = 0; i < o.size(); ++i) untouched += (o[i] == 0xEE);
    printf("data bytes wrong: %zu of 16384 | padding bytes: zero %zu, untouched(0xEE) %zu, other %zu | 0xEE bytes in image: %zu\n",
           bad_data, pad_zero, pad_ee, pad_other, untouched);
    printf("%s\n", bad_data == 0 ? "TMA 16U4_ALIGN16B IMAGE MATCH (a_variant 0 layout)" : "TMA image MISMATCH");
    if (bad_data) {
      printf("first 64 bytes of the image: ");
      for (int i = 0; i < 64; ++i) printf("%02x ", o[i]);
      printf("\nfirst 32 packed bytes of box row 0: ");
      for (int i = 0; i < 32; ++i) printf("%02x ", h[(size_t)row0 * 64 + i]);
      printf("\n");
    }
    return 0;
  }
  if (!strcmp(argv[1], "cp")) {
    const uint32_t lbo = argc > 2 ? (uint32_t)atoi(argv[2]) : 1u, sbo = argc > 3 ? (uint32_t)atoi(argv[3]) : 8u;
    uint32_t* dout = nullptr;
    CK(cudaMalloc(&dout, (128 * 8 + 1) * 4));
    CK(cudaMemset(dout, 0, (128 * 8 + 1) * 4));
    cp_kernel<<<1, 128>>>(lbo, sbo, dout);
    CK(cudaDeviceSynchronize());
    std::vector<uint32_t> o(128 * 8 + 1);
    CK(cudaMemcpy(o.data(), dout, o.size() * 4, cudaMemcpyDeviceToHost));
    // hypothesis: lane i of EVERY quadrant, column c (0..3) = word i*4+c of the 512-byte block
    int good = 0;
    for (int t = 0; t < 128; ++t)
      for (int c = 0; c < 4; ++c) good += (o[t * 8 + c] == (0xA0000000u | (uint32_t)((t & 31) * 4 + c)));
    printf("tcgen05.cp.32x128b.warpx4 lbo16=%u sbo16=%u: commit seen %u | %d of 512 words where (lane i, col c) = word[i*4+c]\n", lbo,
           sbo, o[128 * 8], good);
    printf("  lane 0: %08x %08x %08x %08x | %08x ; lane 1: %08x %08x ; lane 9: %08x %08x ; lane 33: %08x %08x ; lane 127: %08x\n", o[0],
           o[1], o[2], o[3], o[4], o[8], o[9], o[72], o[73], o[33 * 8], o[33 * 8 + 1], o[127 * 8]);
    return 0;
  }
  return 1;
}
