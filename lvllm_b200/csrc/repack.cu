// One-time weight ingest: raw checkpoint layout -> kernel-native tiled layout in HBM.
//
// raw  : w13 [E][N1][Kbytes] row-major (N1 = 2I gated: rows [0,I) gate, [I,2I) up; reference
//        vllm/model_executor/layers/fused_moe/routed_experts.py:564-570),  w2 [E][H][Kbytes]
// tiled: [E][J][KB][NA][16 KB]; each 16 KB block is a [128 rows x 128 B] K-major operand tile that already
//        carries the 128-byte swizzle (16-B chunk index XOR row%8), so the GEMM streams it with one
//        contiguous bulk copy and hands it to tcgen05.mma untouched.  For gated w13, NA = 2: the gate tile
//        of output features [128j,128j+128) is followed by the matching up tile.
#include "common.cuh"
#include "moe_internal.cuh"

namespace b200 {

// one thread per 16-byte chunk of the destination
__global__ void __launch_bounds__(256)
    tile_weights_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int E, int J, int KB, int NA,
                        int rows_per_expert, int up_row_off, int tile_rows, int64_t row_bytes) {
  const int64_t n_chunks = (int64_t)E * J * KB * NA * (TILE_BYTES / 16);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_chunks;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = i;
    const int pc = t % 8;  t /= 8;    // physical chunk within the 128-B row
    const int r = t % 128;  t /= 128;  // row within the tile
    const int na = t % NA;  t /= NA;
    const int kb = t % KB;  t /= KB;
    const int j = t % J;  t /= J;
    const int e = (int)t;
    const int lc = pc ^ (r & 7);      // logical chunk stored at this physical position
    const int64_t srow = (int64_t)e * rows_per_expert + (na ? up_row_off : 0) + (int64_t)j * tile_rows + r;
    const uint4 v = *reinterpret_cast<const uint4*>(src + srow * row_bytes + (int64_t)kb * 128 + lc * 16);
    *reinterpret_cast<uint4*>(dst + i * 16) = v;
  }
}

// block scales [E][N/gN][K/gK] -> [E][N/128][KB]   (gN, gK multiples of 128, or covering the whole dim)
__global__ void expand_scales_kernel(const float* __restrict__ src, float* __restrict__ dst, int E, int NB, int KB,
                                     int SN, int SK, int gN, int gK) {
  const int64_t n = (int64_t)E * NB * KB;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int kb = i % KB;
    const int nb = (i / KB) % NB;
    const int e = (int)(i / ((int64_t)KB * NB));
    int sn = (nb * 128) / gN;
    int sk = (kb * 128) / gK;
    if (sn >= SN) sn = SN - 1;
    if (sk >= SK) sk = SK - 1;
    dst[i] = src[((int64_t)e * SN + sn) * SK + sk];
  }
}

int repack_weights(b200moe_layer* L, const void* w13, const void* w2, const void* s13, const void* s2,
                   const void* g13, const void* g2, cudaStream_t st) {
  (void)g13;
  (void)g2;
  const int NA = L->gated ? 2 : 1;
  const int64_t w13_bytes = (int64_t)L->E * L->J1 * L->KB1 * NA * TILE_BYTES;
  const int64_t w2_bytes = (int64_t)L->E * L->J2 * L->KB2 * TILE_BYTES;
  cudaError_t e;
  if ((e = cudaMalloc(reinterpret_cast<void**>(&L->w13t), w13_bytes)) != cudaSuccess) return cuda_fail(e, "cudaMalloc(w13 tiled)");
  if ((e = cudaMalloc(reinterpret_cast<void**>(&L->w2t), w2_bytes)) != cudaSuccess) return cuda_fail(e, "cudaMalloc(w2 tiled)");
  L->weight_bytes = w13_bytes + w2_bytes;
  const int64_t rb1 = (int64_t)L->KB1 * 128, rb2 = (int64_t)L->KB2 * 128;
  tile_weights_kernel<<<2048, 256, 0, st>>>(reinterpret_cast<const uint8_t*>(w13), L->w13t, L->E, L->J1, L->KB1, NA,
                                           L->N1, L->I, 128, rb1);
  // w2: when H/128 is even, tiles are stored in PAIRS ([E][J2/2][KB2][2][16 KB]) so that a GEMM2 stage (two
  // 128-row tiles x one k-block) is one contiguous 32 KB copy, like the (gate, up) stage of w13
  if (L->w2_paired)
    tile_weights_kernel<<<2048, 256, 0, st>>>(reinterpret_cast<const uint8_t*>(w2), L->w2t, L->E, L->J2 / 2, L->KB2, 2,
                                             L->H, 128, 256, rb2);
  else
    tile_weights_kernel<<<2048, 256, 0, st>>>(reinterpret_cast<const uint8_t*>(w2), L->w2t, L->E, L->J2, L->KB2, 1,
                                             L->H, 0, 128, rb2);
  g_launches += 2;
  if (L->esz_bits == 8) {
    const int gN = L->cfg.groupN > 0 ? L->cfg.groupN : 128, gK = L->cfg.groupK > 0 ? L->cfg.groupK : 128;
    const int NB1 = L->N1 / 128, NB2 = L->H / 128;
    const int64_t n1 = (int64_t)L->E * NB1 * L->KB1, n2 = (int64_t)L->E * NB2 * L->KB2;
    if ((e = cudaMalloc(reinterpret_cast<void**>(&L->ws13), n1 * 4)) != cudaSuccess) return cuda_fail(e, "cudaMalloc(ws13)");
    if ((e = cudaMalloc(reinterpret_cast<void**>(&L->ws2), n2 * 4)) != cudaSuccess) return cuda_fail(e, "cudaMalloc(ws2)");
    L->weight_bytes += (n1 + n2) * 4;
    const int SN1 = (L->N1 + gN - 1) / gN, SK1 = (L->H + gK - 1) / gK;
    const int SN2 = (L->H + gN - 1) / gN, SK2 = (L->I + gK - 1) / gK;
    expand_scales_kernel<<<256, 256, 0, st>>>(reinterpret_cast<const float*>(s13), L->ws13, L->E, NB1, L->KB1, SN1,
                                             SK1, gN, gK);
    expand_scales_kernel<<<256, 256, 0, st>>>(reinterpret_cast<const float*>(s2), L->ws2, L->E, NB2, L->KB2, SN2, SK2,
                                             gN, gK);
    g_launches += 2;
  }
  if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(e, "repack launch");
  return 0;
}

}  // namespace b200
