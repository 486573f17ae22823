// Grouped expert GEMM for decode-sized batches on sm_100a: weight-streaming, swap-AB, tcgen05 + TMEM.
//
// Y_e^T[n, t] = sum_k W_e[n, k] * X[t, k]       (n = output feature, t = permuted token row)
//
// The 128-row weight tile is the UMMA *A* operand (M = 128), the tokens of one expert are the *B* operand
// (N = tn, a multiple of 16), so one pass over an expert's weights serves all of its tokens and the tensor
// core never limits the stream.  Weights live in HBM pre-tiled as [128 rows x 128 B] blocks that already
// carry the 128-byte swizzle, so a pipeline stage is ONE contiguous 32 KB bulk async copy (UBLKCP) — no
// tensor maps, perfect DRAM page locality.  Activations are pre-tiled the same way by the prep kernel.
//
// Warp roles (192 threads, 1 CTA / SM, persistent, dynamic unit scheduler):
//   warps 0-3  epilogue: TMEM -> registers (tcgen05.ld), block-scale promotion (FP8), activation,
//              re-quantisation and stores
//   warp  4    producer: bulk copies global -> smem ring, full/empty mbarriers; also owns the scheduler
//   warp  5    MMA issuer: one elected thread issues tcgen05.mma, tcgen05.commit frees smem / signals TMEM
//   warps 6-17 (TNMAX = 128 only, 576 threads) three more epilogue groups: the 128 token columns of a prefill-class
//              tile are split in four windows of 32, so that a thread's accumulators (2 x 32 fp32) stay in registers
//              (two groups of 64 columns spilled them under the 168-register cap: tensor pipe 9 %) and four warps per
//              scheduler hide the TMEM-load latency of the block-scale promotion
//
// Prefill-class batches (>= ~96 rows per expert; BASELINE config 4 shapes, SURVEY.md 8d "tensor cores") use
// TNMAX = 128: one expert's weights then serve 128 token rows per pass (full-rate UMMA 128 x 128 x 32), two TMEM
// accumulator buffers of 256 columns, 48 KB pipeline stages.
//
// FP8 (W8A8, DeepSeek block-128 numerics): each 128-wide k-block is a fresh TMEM accumulation; the
// epilogue warps fold  part * w_scale[n/128,kb] * x_scale[t,kb]  into fp32 registers (TMEM buffers are
// ring-buffered so the MMA never waits).  16-bit formats accumulate the whole K in TMEM.
//
// FP8 with power-of-two (ue8m0) block scales — MODE 2, the DeepGEMM-on-Blackwell numerics of the reference
// (vllm/utils/deep_gemm.py:662-681 per_block_cast_to_fp8(use_ue8m0), fp8_utils.py:112, :986-1043) — needs no promotion at
// all: the scales ride in TMEM next to the accumulator and tcgen05.mma kind::mxf8f6f4.block_scale applies them per 32-wide
// k-group, so the whole K accumulates in TMEM like a 16-bit GEMM.  The promotion path is bound by TMEM reads (128 x 128 x 2
// fp32 per k-block against ~512 MMA cycles); this one is MMA-bound.  Scale words are written by epilogue warps 0-3 (one per
// TMEM lane quadrant), which are otherwise idle during a unit's main loop.
//
// Roofline: HBM.  Algorithmic bytes per unit = weight tile bytes (activations are <1%).
#include "common.cuh"
#include "moe_internal.cuh"

namespace b200 {

constexpr int NUM_EPI_WARPS = 4;   // per epilogue group (one warp per TMEM lane quadrant)
constexpr int GEMM_THREADS = 192;
constexpr int QDEPTH = 4;       // scheduler queue depth
constexpr int KBG = 28;         // k-blocks whose FP8 scales are staged in shared memory at a time
constexpr int SMEM_BUDGET = 225 * 1024;

enum { EPI_GATED = 0, EPI_ACT1 = 1, EPI_OUT = 2 };

struct GemmArgs {
  const uint8_t* wt;    // tiled weights
  const float* wscale;  // [E][NB][KB] (fp8) or nullptr
  const uint8_t* bt;    // tiled B operand (activations)   [rows/8][KB][1024]
  const float* bscale;  // [KB][rows_stride] (fp8)
  uint8_t* it;          // tiled intermediate out (EPI_GATED / EPI_ACT1)
  float* iscale;        // [KB2][rows_stride]
  float* y;             // [rows][n_out] fp32 (EPI_OUT)
  const Chunk* chunks;
  RouteState* state;
  int which;            // 0: GEMM1, 1: GEMM2 (selects the scheduler counter)
  int KB;               // k-blocks (128 B) of this GEMM
  int J;                // output tiles per expert
  int NB;               // scale row-blocks per expert (fp8): N_total/128
  int up_block_off;     // scale row-block offset of the "up" half (= I/128) for EPI_GATED
  int KB_out;           // k-blocks of the NEXT GEMM's operand (intermediate width / elems-per-128B)
  int rows_stride;      // stride of the transposed scale arrays
  int n_out;            // leading dim of y
  int act_type;         // 0 silu, 1 swigluoai, 2 relu2
  float alpha, limit;
  int act_fp16;         // activation dtype of the 16-bit intermediate: 0 bf16, 1 fp16
  int e8m0;             // FP8 layers in ue8m0 mode: activation / intermediate group scales are rounded up to powers of two
  int sfb_variant;      // MODE 2 bring-up: TMEM placement of the token scale words (0 = lane n % 32, column n / 32)
};

enum { MODE_16 = 0, MODE_FP8 = 1, MODE_FP8_MX = 2 };
constexpr int GEMM_SF_COLS = 16;   // MODE 2: TMEM columns per pipeline stage (SFA gate / SFA up / SFB, 4 each, + 4 spare)

template <int MODE, int NA, int TNMAX, bool PAIR = false>
struct Cfg {
  static constexpr bool MXS = (MODE == MODE_FP8_MX);
  static constexpr int NCH = PAIR ? 2 : 1;                 // expert chunks that share one weight stage (see PAIR below)
  static constexpr int KBS = 2 / NA;                       // k-blocks per stage
  static constexpr int A_STAGE = 2 * TILE_BYTES;           // 32 KB
  static constexpr int B_CHUNK = KBS * TNMAX * 128;        // activation bytes of one chunk per stage
  static constexpr int B_STAGE = NCH * B_CHUNK;            // bytes
  static constexpr int STAGE = A_STAGE + B_STAGE;
  static constexpr int MISC = 21504;
  static constexpr int EW = TNMAX > 64 ? 4 : 1;            // epilogue groups of 4 warps (each owns TNMAX / EW token columns)
  static constexpr int CW = TNMAX / EW;                    // token columns per epilogue warp
  static constexpr int NTHREADS = GEMM_THREADS + (EW - 1) * 128;
  static constexpr int STAGES_RAW = (SMEM_BUDGET - MISC - 1024) / STAGE;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int BUFCOLS = NCH * NA * TNMAX;
  static constexpr int NBUF_RAW = (512 - (MXS ? STAGES * GEMM_SF_COLS : 0)) / BUFCOLS;
  static constexpr int NBUF = NBUF_RAW > 4 ? 4 : NBUF_RAW;
  static constexpr int SFCOL = NBUF * BUFCOLS;             // MODE 2: first scale-factor column
  static constexpr int TMEM_COLS_RAW = NBUF * BUFCOLS + (MXS ? STAGES * GEMM_SF_COLS : 0);
  static constexpr int TMEM_COLS = TMEM_COLS_RAW <= 32 ? 32 : TMEM_COLS_RAW <= 64 ? 64 : TMEM_COLS_RAW <= 128 ? 128
                                   : TMEM_COLS_RAW <= 256 ? 256 : 512;
  static constexpr int SMEM = STAGES * STAGE + MISC + 1024;
};

struct __align__(16) Misc {
  uint64_t full[8], empty[8];
  uint64_t sfready[8];             // MODE 2: the stage's scale words are in TMEM
  uint64_t tfull[4], tempty[4];
  uint64_t qfull[QDEPTH], qempty[QDEPTH];
  int32_t qunit[QDEPTH];
  uint32_t tmem_base;
  float red[NUM_EPI_WARPS][128];   // [lane quadrant][token column]
  alignas(16) float gsx[KBG][128];   // FP8: activation scales of the current group of k-blocks [k-block][token column]
  float gws[2][KBG];                 // FP8: weight block scales of the group [gate|up][k-block]
};

B200_DEVICE void bounded_wait(uint64_t* bar, uint32_t parity) {
  // A protocol bug must not hang the GPU box: trap after ~2^27 polls (seconds).
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 27)) __trap();
  }
}

B200_DEVICE float silu_f(float x) { return x / (1.0f + expf(-x)); }

// PAIR (16-bit mode, gated layers, 128-row chunks): a unit is an expert's chunk PAIR x output tile.  Both chunks run against
// the SAME weight stage (A 32 KB + 2 x 16 KB of activations), accumulating in the two halves of TMEM.  The prefill kernel is
// bound by the bytes a CTA can keep in flight over HBM latency (three or four stages of shared memory), not by L2 or the
// tensor core: twice the MMA work per stage is twice the throughput wherever an expert has two chunks.  The routing tables
// pad every expert to an even number of chunk entries (an empty one, nrows = 0, comes second in its pair).
template <int MODE, int NA, int EPI, int TNMAX, bool PAIR = false>
__global__ void __launch_bounds__((Cfg<MODE, NA, TNMAX, PAIR>::NTHREADS), 1) moe_gemm_kernel(const GemmArgs a) {
  using C = Cfg<MODE, NA, TNMAX, PAIR>;
  static_assert(!PAIR || (MODE == MODE_16 && NA == 2 && TNMAX == 128), "chunk pairs: 16-bit gated / paired tiles, 128-row chunks");
  constexpr int NCH = C::NCH;
  constexpr bool FP8 = (MODE != MODE_16);        // 8-bit operands, fp8 intermediate + group scales
  constexpr bool PROMO = (MODE == MODE_FP8);     // fp32 block scales: per-k-block promotion in the epilogue warps
  constexpr bool MXS = (MODE == MODE_FP8_MX);    // ue8m0 block scales applied by the tensor core
  constexpr int EW = C::EW, CW = C::CW;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Misc* ms = reinterpret_cast<Misc*>(smem + C::STAGES * C::STAGE);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&ms->full[i], 1);
      mbar_init(&ms->empty[i], 1);
      mbar_init(&ms->sfready[i], 4);
    }
    for (int i = 0; i < C::NBUF; ++i) {
      mbar_init(&ms->tfull[i], 1);
      mbar_init(&ms->tempty[i], NUM_EPI_WARPS * EW);
    }
    for (int i = 0; i < QDEPTH; ++i) {
      mbar_init(&ms->qfull[i], 1);
      mbar_init(&ms->qempty[i], 1 + NUM_EPI_WARPS * EW);
    }
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(&ms->tmem_base, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ms->tmem_base;

  const int n_units = (PAIR ? a.state->n_chunks >> 1 : a.state->n_chunks) * a.J;
  const int KB = a.KB;
  const int n_iters = (KB + C::KBS - 1) / C::KBS;  // pipeline iterations per unit

  if (warp == 4) {
    // ======================================================================= producer + scheduler
    if (elect_one()) {
      const uint64_t pol = policy_evict_first();
      uint32_t it = 0, q = 0;
      for (;;) {
        const int u = atomicAdd(&a.state->unit_ctr[a.which], 1);
        const int qs = q % QDEPTH;
        bounded_wait(&ms->qempty[qs], ((q / QDEPTH) & 1) ^ 1);
        ms->qunit[qs] = (u < n_units) ? u : -1;
        mbar_arrive(&ms->qfull[qs]);
        ++q;
        if (u >= n_units) break;
        const int ci = u / a.J, j = u % a.J;
        const Chunk ch = a.chunks[PAIR ? 2 * ci : ci];
        const int tn = (ch.nrows + 15) & ~15;
        const Chunk ch1 = PAIR ? a.chunks[2 * ci + 1] : ch;        // second chunk of the pair (nrows may be 0)
        const int tn1 = PAIR ? (ch1.nrows + 15) & ~15 : 0;
        const uint8_t* wsrc = a.wt + ((size_t)(ch.expert * a.J + j) * KB) * (size_t)(NA * TILE_BYTES);
        for (int i = 0; i < n_iters; ++i, ++it) {
          const int s = it % C::STAGES;
          bounded_wait(&ms->empty[s], ((it / C::STAGES) & 1) ^ 1);
          const int kb0 = i * C::KBS;
          const int nkb = (KB - kb0 < C::KBS) ? (KB - kb0) : C::KBS;
          const uint32_t abytes = nkb * NA * TILE_BYTES;
          const uint32_t bbytes = (tn >> 3) * nkb * 1024;
          const uint32_t bbytes1 = (tn1 >> 3) * nkb * 1024;
          uint8_t* sa = smem + s * C::STAGE;
          uint8_t* sb = sa + C::A_STAGE;
          mbar_arrive_expect_tx(&ms->full[s], abytes + bbytes + bbytes1);
          bulk_g2s_hint(sa, wsrc + (size_t)kb0 * (NA * TILE_BYTES), abytes, &ms->full[s], pol);
          // the chunk's activation tiles are chunk-contiguous ([k-block][row group][1 KB]): ONE bulk copy per stage
          // (round 1 issued tn/8 copies of 1 KB each — 16 per stage at tn = 128, which bounded the prefill-class GEMM)
          bulk_g2s(sb, a.bt + (size_t)ch.row0 * KB * 128 + (size_t)kb0 * (size_t)((tn >> 3) * 1024), bbytes, &ms->full[s]);
          if (PAIR && bbytes1)
            bulk_g2s(sb + C::B_CHUNK, a.bt + (size_t)ch1.row0 * KB * 128 + (size_t)kb0 * (size_t)((tn1 >> 3) * 1024), bbytes1,
                     &ms->full[s]);
        }
      }
    }
  } else if (warp == 5) {
    // ======================================================================= MMA issuer
    if (elect_one()) {
      uint32_t it = 0, q = 0, acc_it = 0;
      for (;;) {
        const int qs = q % QDEPTH;
        bounded_wait(&ms->qfull[qs], (q / QDEPTH) & 1);
        const int u = ms->qunit[qs];
        mbar_arrive(&ms->qempty[qs]);
        ++q;
        if (u < 0) break;
        const Chunk ch = a.chunks[PAIR ? 2 * (u / a.J) : u / a.J];
        const int tn = (ch.nrows + 15) & ~15;
        const int tn1 = PAIR ? (a.chunks[2 * (u / a.J) + 1].nrows + 15) & ~15 : 0;
        const uint32_t idesc = MXS   ? umma_idesc_mx(0, 0, tn)
                               : FP8 ? umma_idesc(0, 0, 128, tn)
                                     : umma_idesc(a.act_fp16 ? 0 : 1, a.act_fp16 ? 0 : 1, 128, tn);
        const uint32_t idesc1 = umma_idesc(a.act_fp16 ? 0 : 1, a.act_fp16 ? 0 : 1, 128, tn1 > 0 ? tn1 : 16);   // PAIR only
        uint32_t buf = 0;
        if (!PROMO) {
          buf = acc_it % C::NBUF;
          bounded_wait(&ms->tempty[buf], ((acc_it / C::NBUF) & 1) ^ 1);
          tc_fence_after();
        }
        for (int i = 0; i < n_iters; ++i, ++it) {
          const int s = it % C::STAGES;
          bounded_wait(&ms->full[s], (it / C::STAGES) & 1);
          tc_fence_after();
          const int kb0 = i * C::KBS;
          const int nkb = (KB - kb0 < C::KBS) ? (KB - kb0) : C::KBS;
          const uint32_t sa = smem_u32(smem + s * C::STAGE);
          const uint32_t sb = sa + C::A_STAGE;
          if (MXS) {
            bounded_wait(&ms->sfready[s], (it / C::STAGES) & 1);   // the stage's scale words are in TMEM
            tc_fence_after();
          }
          const uint32_t sfc = tmem_base + C::SFCOL + s * GEMM_SF_COLS;
          for (int kk = 0; kk < nkb; ++kk) {
            if (PROMO) {
              buf = acc_it % C::NBUF;
              bounded_wait(&ms->tempty[buf], ((acc_it / C::NBUF) & 1) ^ 1);
              tc_fence_after();
            }
#pragma unroll
            for (int cn = 0; cn < NCH * NA; ++cn) {
              const int cc = cn / NA, na = cn % NA;   // chunk of the pair, accumulator (gate | up, or tile of a w2 pair)
              if (PAIR && cc == 1 && tn1 == 0) break;
              // stage layout: NA==2 -> [gate tile][up tile] of one k-block; NA==1 -> [kb0 tile][kb1 tile]
              const uint32_t abase = sa + (NA == 2 ? na : kk) * TILE_BYTES;
              const uint32_t bbase = sb + cc * C::B_CHUNK + kk * (uint32_t)(((cc ? tn1 : tn) >> 3) * 1024);
              const uint32_t dcol = tmem_base + buf * C::BUFCOLS + (cc * NA + na) * TNMAX;
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint64_t ad = umma_desc_sw128(abase + ks * 32, 1024);
                const uint64_t bd = umma_desc_sw128(bbase + ks * 32, 1024);
                const uint32_t accum = PROMO ? (ks > 0) : ((kb0 + kk) > 0 || ks > 0);
                if (PAIR && cc == 1) {
                  umma_f16(dcol, ad, bd, idesc1, accum);
                  continue;
                }
                if (MXS) {
                  // scale columns of the stage: NA == 2 -> [gate 0-3][up 4-7][tokens 8-11]; NA == 1 -> per k-block
                  // [weights 8kk..+3][tokens 8kk+4..+7]; ks selects the byte (= 32-wide k-group) of the scale words
                  const uint32_t sfa = sfc + (NA == 2 ? na * 4 : kk * 8), sfb = sfc + (NA == 2 ? 8 : kk * 8 + 4);
                  umma_mx(dcol, ad, bd, idesc | ((uint32_t)ks << 4) | ((uint32_t)ks << 29), accum, sfa | ((uint32_t)ks << 30),
                          sfb | ((uint32_t)ks << 30));
                } else if (FP8)
                  umma_f8(dcol, ad, bd, idesc, accum);
                else
                  umma_f16(dcol, ad, bd, idesc, accum);
              }
            }
            if (PROMO) {
              umma_commit(&ms->tfull[buf]);
              ++acc_it;
            }
          }
          umma_commit(&ms->empty[s]);
        }
        if (!PROMO) {
          umma_commit(&ms->tfull[buf]);
          ++acc_it;
        }
      }
    }
  } else {
    // ======================================================================= epilogue warps 0..3 (and 6..9)
    uint32_t q = 0, acc_it = 0, sf_it = 0;
    const int q4 = warp & 3;                     // TMEM lane quadrant of this warp
    const int c_base = (warp < 4 ? 0 : (warp - 6) / 4 + 1) * CW;  // first token column of this warp's window
    const uint32_t lane_off = (uint32_t)(q4 * 32) << 16;
    const int row_in_tile = q4 * 32 + lane;  // output feature within the 128-row tile
    for (;;) {
      const int qs = q % QDEPTH;
      bounded_wait(&ms->qfull[qs], (q / QDEPTH) & 1);
      const int u = ms->qunit[qs];
      __syncwarp();
      if (lane == 0) mbar_arrive(&ms->qempty[qs]);
      ++q;
      if (u < 0) break;
      const int ci = u / a.J, j = u % a.J;
      // PAIR: the unit's two chunks are drained one after the other (second half of the TMEM columns); the accumulator
      // is handed back to the MMA warp after the last live chunk's TMEM loads
      const bool second_live = PAIR && a.chunks[2 * ci + 1].nrows > 0;
      for (int pc = 0; pc < NCH; ++pc) {
      if (PAIR && pc == 1 && !second_live) break;
      const Chunk ch = a.chunks[PAIR ? 2 * ci + pc : ci];
      const int tn = (ch.nrows + 15) & ~15;
      const uint32_t pcol = (uint32_t)(pc * NA * TNMAX);   // first TMEM column of this chunk's accumulators
      const bool last_pc = !PAIR || pc == 1 || !second_live;

      float acc[NA][CW];
#pragma unroll
      for (int na = 0; na < NA; ++na)
#pragma unroll
        for (int c = 0; c < CW; ++c) acc[na][c] = 0.f;

      const int n_groups = PROMO ? KB : 1;
      // FP8: the scales of KBG k-blocks at a time are staged in shared memory by ALL epilogue threads (one global
      // latency per group instead of one per k-block: a prefetch distance of one k-block cannot cover ~2000 cycles of
      // load latency inside a ~512-cycle k-block), pre-multiplied: gsx[kk][c] = x_scale[token c, k-block], gws[na][kk]
      const float* wrow[NA];
      const int etid = warp < 4 ? threadIdx.x : threadIdx.x - 64;   // 0 .. 128*EW-1 over the epilogue threads
      if (FP8) {
#pragma unroll
        for (int na = 0; na < NA; ++na) {
          const int rb = (EPI == EPI_GATED) ? (na == 0 ? j : a.up_block_off + j) : (NA == 2 ? 2 * j + na : j);
          wrow[na] = a.wscale + ((size_t)ch.expert * a.NB + rb) * KB;
        }
      }
      if (MXS && warp < 4) {
        // ---- MODE 2: warps 0-3 feed the unit's scale words to TMEM (ue8m0 byte of a power-of-two fp32 scale = its
        // exponent field, replicated over the four 32-wide k-groups of the k-block).  Scales of SB k-blocks at a time go
        // through shared memory (thread = token, coalesced); the next batch's loads are in flight while this one is fed.
        constexpr int SB = KBG / 2;
        const int t = threadIdx.x;    // 0..127 = token column of the chunk
        const int nbatch = (KB + SB - 1) / SB;
        float pre[SB], prew[NA];
        auto fetch = [&](int b) {
#pragma unroll
          for (int kk = 0; kk < SB; ++kk) {
            const int kb = b * SB + kk;
            pre[kk] = (kb < KB && t < tn) ? __ldg(a.bscale + (size_t)kb * a.rows_stride + ch.row0 + t) : 0.f;
          }
#pragma unroll
          for (int na = 0; na < NA; ++na) prew[na] = (t < SB && b * SB + t < KB) ? __ldg(wrow[na] + b * SB + t) : 0.f;
        };
        auto stash = [&](int b) {
          const int o = (b & 1) * SB;
#pragma unroll
          for (int kk = 0; kk < SB; ++kk) ms->gsx[o + kk][t] = pre[kk];
          if (t < SB) {
#pragma unroll
            for (int na = 0; na < NA; ++na) ms->gws[na][o + t] = prew[na];
          }
        };
        auto sfword = [](float sc) { return ((__float_as_uint(sc) >> 23) & 0xFFu) * 0x01010101u; };
        fetch(0);
        asm volatile("bar.sync 2, 128;" ::: "memory");   // the previous unit's last batch has been consumed
        stash(0);
        asm volatile("bar.sync 2, 128;" ::: "memory");
        for (int b = 0; b < nbatch; ++b) {
          if (b + 1 < nbatch) fetch(b + 1);
          const int o = (b & 1) * SB;
          const int kend = (KB - b * SB < SB) ? KB - b * SB : SB;
          for (int k0 = 0; k0 < kend; k0 += C::KBS, ++sf_it) {
            const int s = sf_it % C::STAGES;
            bounded_wait(&ms->empty[s], ((sf_it / C::STAGES) & 1) ^ 1);   // the MMAs that read these columns are done
            tc_fence_after();
            const uint32_t sfc = tmem_base + lane_off + C::SFCOL + s * GEMM_SF_COLS;
#pragma unroll
            for (int kk = 0; kk < C::KBS; ++kk) {
              if (k0 + kk < kend) {
                const int ks = o + k0 + kk;
                uint32_t wb[4];
#pragma unroll
                for (int c = 0; c < 4; ++c)
                  wb[c] = sfword(ms->gsx[ks][a.sfb_variant == 1 ? q4 * 32 + lane : c * 32 + lane]);
                if (NA == 2) {
                  tmem_st4(sfc, sfword(ms->gws[0][ks]));
                  tmem_st4(sfc + 4, sfword(ms->gws[1][ks]));
                  tmem_st4v(sfc + 8, wb[0], wb[1], wb[2], wb[3]);
                } else {
                  tmem_st4(sfc + kk * 8, sfword(ms->gws[0][ks]));
                  tmem_st4v(sfc + kk * 8 + 4, wb[0], wb[1], wb[2], wb[3]);
                }
              }
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ms->sfready[s]);
          }
          if (b + 1 < nbatch) {
            stash(b + 1);   // the other half of the staging buffer: batch b - 1 was consumed before batch b started
            asm volatile("bar.sync 2, 128;" ::: "memory");
          }
        }
      }
      for (int g = 0; g < n_groups; ++g) {
        const uint32_t buf = acc_it % C::NBUF;
        const int rel = g % KBG;
        if (PROMO && rel == 0) {
          asm volatile("bar.sync 2, %0;" ::"n"(128 * EW) : "memory");   // every epilogue warp is done with the old group
          const int n = (n_groups - g < KBG) ? n_groups - g : KBG;
          for (int i = etid; i < n * TNMAX; i += 128 * EW) {
            const int kk = i / TNMAX, c = i - kk * TNMAX;
            ms->gsx[kk][c] = (c < tn) ? __ldg(a.bscale + (size_t)(g + kk) * a.rows_stride + ch.row0 + c) : 0.f;
          }
          for (int i = etid; i < NA * n; i += 128 * EW) {
            const int na = i / n, kk = i - na * n;
            ms->gws[na][kk] = __ldg(wrow[na] + g + kk);
          }
          asm volatile("bar.sync 2, %0;" ::"n"(128 * EW) : "memory");
        }
        if (!PAIR || pc == 0) bounded_wait(&ms->tfull[buf], (acc_it / C::NBUF) & 1);
        tc_fence_after();
        if (PROMO) {
          // gate and up partial sums of the same 16 token columns are fetched together (two TMEM loads in flight per
          // wait) and promoted with  part * (w_scale[na] * x_scale[token])
          float ws_cur[NA];
#pragma unroll
          for (int na = 0; na < NA; ++na) ws_cur[na] = ms->gws[na][rel];
          const float* sxw = &ms->gsx[rel][c_base];
#pragma unroll
          for (int c8 = 0; c8 < CW / 8; ++c8) {
            if (c_base + c8 * 8 < tn) {
              // 8 token columns of gate and up at a time: the accumulators (NA x CW) plus 2 x 8 partials fit the
              // register budget of the 576-thread variant
              float part[NA][8];
#pragma unroll
              for (int na = 0; na < NA; ++na)
                tmem_ld8(tmem_base + lane_off + buf * C::BUFCOLS + na * TNMAX + c_base + c8 * 8, part[na]);
              tmem_ld_wait();
#pragma unroll
              for (int c4 = 0; c4 < 2; ++c4) {
                const float4 xs4 = *reinterpret_cast<const float4*>(sxw + c8 * 8 + c4 * 4);
                const float xv[4] = {xs4.x, xs4.y, xs4.z, xs4.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
#pragma unroll
                  for (int na = 0; na < NA; ++na)
                    acc[na][c8 * 8 + c4 * 4 + q] = fmaf(part[na][c4 * 4 + q], ws_cur[na] * xv[q], acc[na][c8 * 8 + c4 * 4 + q]);
                }
              }
            }
          }
        } else {
#pragma unroll
          for (int na = 0; na < NA; ++na) {
#pragma unroll
            for (int c16 = 0; c16 < CW / 16; ++c16) {
              if (c_base + c16 * 16 < tn) {
                float part[16];
                tmem_ld16(tmem_base + lane_off + buf * C::BUFCOLS + pcol + na * TNMAX + c_base + c16 * 16, part);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 16; ++c) acc[na][c16 * 16 + c] = part[c];
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (last_pc) {
          if (lane == 0) mbar_arrive(&ms->tempty[buf]);
          ++acc_it;
        }
      }

      // ------------------------------------------------------------------ final epilogue of the unit
      if (EPI == EPI_OUT) {
        // NA == 2: paired layout, accumulator na is output tile 2j+na
#pragma unroll
        for (int na = 0; na < NA; ++na) {
          const int jt = (NA == 2) ? 2 * j + na : j;
          float* yb = a.y + (size_t)ch.row0 * a.n_out + (size_t)jt * 128 + row_in_tile;
#pragma unroll
          for (int c = 0; c < CW; ++c)
            if (c_base + c < ch.nrows) yb[(size_t)(c_base + c) * a.n_out] = acc[na][c];
        }
      } else {
        // activation in fp32 on values rounded to the activation dtype (matches the reference chain:
        // GEMM output -> act dtype -> act -> act dtype -> (fp8 group quant))
        float v[CW];
#pragma unroll
        for (int c = 0; c < CW; ++c) {
          float g0, r;
          if (a.act_fp16)
            g0 = __half2float(__float2half_rn(acc[0][c]));
          else
            g0 = __bfloat162float(__float2bfloat16_rn(acc[0][c]));
          if (EPI == EPI_GATED) {
            float u0;
            if (a.act_fp16)
              u0 = __half2float(__float2half_rn(acc[NA - 1][c]));
            else
              u0 = __bfloat162float(__float2bfloat16_rn(acc[NA - 1][c]));
            if (a.act_type == 1) {
              const float gg = fminf(g0, a.limit);
              const float uu = fminf(fmaxf(u0, -a.limit), a.limit);
              r = (uu + 1.0f) * gg * (1.0f / (1.0f + expf(-a.alpha * gg)));
            } else {
              r = silu_f(g0) * u0;
            }
          } else {
            const float t = fmaxf(g0, 0.f);
            r = t * t;
          }
          if (a.act_fp16)
            v[c] = __half2float(__float2half_rn(r));
          else
            v[c] = __bfloat162float(__float2bfloat16_rn(r));
        }
        if (FP8) {
          // per-token group-128 quantisation: the 128 features of this tile are exactly one group
#pragma unroll
          for (int c = 0; c < CW; ++c) {
            if (c_base + c < tn) {
              const float m = warp_max(fabsf(v[c]));
              if (lane == 0) ms->red[q4][c_base + c] = m;
            }
          }
          asm volatile("bar.sync 1, %0;" ::"n"(128 * EW) : "memory");
          const int kb2 = j;  // 128 features == one 128-B k-block of the next GEMM
#pragma unroll
          for (int c = 0; c < CW; ++c) {
            const int cc = c_base + c;
            if (cc < ch.nrows) {
              float m = fmaxf(fmaxf(ms->red[0][cc], ms->red[1][cc]), fmaxf(ms->red[2][cc], ms->red[3][cc]));
              const float sc = fp8_group_scale(m, a.e8m0);
              const int r = ch.row0 + cc;
              const __nv_fp8_e4m3 qv(v[c] / sc);
              uint8_t* dst = a.it + (size_t)ch.row0 * a.KB_out * 128 + (size_t)kb2 * (size_t)((tn >> 3) * 1024) +
                             (size_t)(cc >> 3) * 1024 + sw128_offset(cc & 7, row_in_tile);
              *dst = *reinterpret_cast<const uint8_t*>(&qv);
              if (row_in_tile == 0) a.iscale[(size_t)kb2 * a.rows_stride + r] = sc;
            }
          }
          asm volatile("bar.sync 1, %0;" ::"n"(128 * EW) : "memory");
        } else {
          const int kb2 = j * 2 + (row_in_tile >> 6);  // 64 16-bit features per 128-B k-block
          const int boff = (row_in_tile & 63) * 2;
#pragma unroll
          for (int c = 0; c < CW; ++c) {
            if (c_base + c < ch.nrows) {
              const int r = ch.row0 + c_base + c;
              const int cc = c_base + c;
              uint8_t* dst = a.it + (size_t)ch.row0 * a.KB_out * 128 + (size_t)kb2 * (size_t)((tn >> 3) * 1024) +
                             (size_t)(cc >> 3) * 1024 + sw128_offset(cc & 7, boff);
              if (a.act_fp16)
                *reinterpret_cast<__half*>(dst) = __float2half_rn(v[c]);
              else
                *reinterpret_cast<__nv_bfloat16*>(dst) = __float2bfloat16_rn(v[c]);
            }
          }
        }
      }
      }   // chunk of the pair
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

template <int MODE, int NA, int EPI, int TNMAX, bool PAIR = false>
static int launch_one(const GemmArgs& a, cudaStream_t st, int num_sms) {
  using C = Cfg<MODE, NA, TNMAX, PAIR>;
  auto kern = moe_gemm_kernel<MODE, NA, EPI, TNMAX, PAIR>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(moe_gemm)");
    attr_set = true;
  }
  kern<<<num_sms, C::NTHREADS, C::SMEM, st>>>(a);
  ++g_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "moe_gemm launch");
  return 0;
}

template <int MODE, int TNMAX, bool PAIR = false>
static int launch_pair(const b200moe_layer* L, Workspace* ws, cudaStream_t st, int num_sms, cudaEvent_t* ev) {
  GemmArgs g1{};
  g1.wt = L->w13t;
  g1.wscale = L->ws13;
  g1.bt = ws->xt;
  g1.bscale = ws->xs;
  g1.it = ws->it;
  g1.iscale = ws->is;
  g1.y = nullptr;
  g1.chunks = ws->chunks;
  g1.state = ws->state;
  g1.which = 0;
  g1.KB = L->KB1;
  g1.J = L->J1;
  g1.NB = L->N1 / 128;
  g1.up_block_off = L->I / 128;
  g1.KB_out = L->KB2;
  g1.rows_stride = (int)ws->cap_rows;
  g1.n_out = 0;
  g1.act_type = L->cfg.activation_type;
  g1.alpha = L->cfg.swiglu_alpha;
  g1.limit = L->cfg.swiglu_limit;
  g1.act_fp16 = (L->act_dtype == B200_ACT_FP16);
  g1.e8m0 = L->fp8_e8m0;
  {
    const char* v = getenv("B200MOE_SFB_VARIANT");   // bring-up only
    g1.sfb_variant = v ? atoi(v) : 0;
  }
  int rc;
  if (ev) cudaEventRecord(ev[0], st);
  if (L->gated)
    rc = launch_one<MODE, 2, EPI_GATED, TNMAX, PAIR>(g1, st, num_sms);
  else
    rc = launch_one<MODE, 1, EPI_ACT1, TNMAX>(g1, st, num_sms);
  if (rc) return rc;
  if (ev) cudaEventRecord(ev[1], st);

  GemmArgs g2 = g1;
  g2.wt = L->w2t;
  g2.wscale = L->ws2;
  g2.bt = ws->it;
  g2.bscale = ws->is;
  g2.it = nullptr;
  g2.iscale = nullptr;
  g2.y = ws->y;
  g2.which = 1;
  g2.KB = L->KB2;
  g2.J = L->J2;
  g2.NB = L->H / 128;
  g2.up_block_off = 0;
  g2.KB_out = 0;
  g2.n_out = L->H;
  if (L->w2_paired) {
    g2.J = L->J2 / 2;
    rc = launch_one<MODE, 2, EPI_OUT, TNMAX, PAIR>(g2, st, num_sms);
  } else {
    rc = launch_one<MODE, 1, EPI_OUT, TNMAX>(g2, st, num_sms);
  }
  if (ev) cudaEventRecord(ev[2], st);
  return rc;
}

// chunk size (token rows of one expert per pass over its weights): prefill-class batches with >= ~96 rows per
// expert take 128-row chunks (full-rate UMMA N = 128, weights re-read once per 128 rows instead of once per 64)
int pick_tn_max(int M, int k, int E) {
  if (M <= 16) return 16;
  if (M <= 32) return 32;
  const long rows_per_expert = ((long)M * k) / (E > 0 ? E : 1);
  return rows_per_expert >= 96 ? 128 : 64;
}

// 16-bit layers (and the fp16 expansion of 4-bit layers), prefill-class chunks, gated experts with paired w2 tiles: the
// chunk-pair form (two chunks of an expert per weight stage).  The routing tables must then pair the chunks (launch_prep
// `pair`).  OPT-IN (B200MOE_GEMM_PAIR=1): the form was written after this round's GPU budget was spent and has not run on
// hardware yet; one chunk per unit is the measured default.
int gemm_uses_pairs(const b200moe_layer* L, int tn_max) {
  if (!(L->esz_bits == 16 && tn_max == 128 && L->gated && L->w2_paired)) return 0;
  const char* v = getenv("B200MOE_GEMM_PAIR");
  return (v && v[0] == '1') ? 1 : 0;
}

int launch_gemms(const b200moe_layer* L, Workspace* ws, cudaStream_t st, int M, int k, int tn_max, cudaEvent_t* ev) {
  static int num_sms = 0;
  if (!num_sms) {
    cudaDeviceProp p;
    cudaError_t e = cudaGetDeviceProperties(&p, L->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties");
    num_sms = p.multiProcessorCount;
  }
  const bool fp8 = (L->esz_bits == 8);
  switch (tn_max) {
    case 16: return fp8 ? launch_pair<MODE_FP8, 16>(L, ws, st, num_sms, ev) : launch_pair<MODE_16, 16>(L, ws, st, num_sms, ev);
    case 32: return fp8 ? launch_pair<MODE_FP8, 32>(L, ws, st, num_sms, ev) : launch_pair<MODE_16, 32>(L, ws, st, num_sms, ev);
    case 128:
      // ue8m0 layers: the tensor core applies the block scales (no promotion) — the prefill-class kernel
      if (fp8 && L->fp8_e8m0 && !getenv("B200MOE_E8M0_PROMO")) return launch_pair<MODE_FP8_MX, 128>(L, ws, st, num_sms, ev);
      if (gemm_uses_pairs(L, tn_max)) return launch_pair<MODE_16, 128, true>(L, ws, st, num_sms, ev);
      return fp8 ? launch_pair<MODE_FP8, 128>(L, ws, st, num_sms, ev) : launch_pair<MODE_16, 128>(L, ws, st, num_sms, ev);
    default: return fp8 ? launch_pair<MODE_FP8, 64>(L, ws, st, num_sms, ev) : launch_pair<MODE_16, 64>(L, ws, st, num_sms, ev);
  }
}

}  // namespace b200
