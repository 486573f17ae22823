/* Plain-C restatement of the routed-expert forward (TEST INFRASTRUCTURE / CPU BASELINE ONLY).
 *
 * out[t] = sum_j w[t,j] * W2_e( silu(W1_e x[t]) * (W3_e x[t]) ),  e = ids[t,j] (ids < 0 skipped)
 * Follows the same reference lines as oracle/moe_oracle.py::experts_forward (reference
 * tests/kernels/utils.py:855-994, weight-only branch; w13 rows [0,I) gate, [I,2I) up,
 * vllm/model_executor/layers/fused_moe/routed_experts.py:564-570).  fp32 accumulation, intermediate rounded
 * to bf16 like the reference chain.  It is what `bench.py --impl reference` / cpu_baseline time on the host
 * cores (kind "port": the real lk_moe wheel is closed-source and absent); weights stay in
 * their checkpoint format in DRAM and are dequantised on the fly, as a CPU-offload engine must.
 * Pinned (tests/test_c_port.py) to the oracle and to the reference's compiled CPU fused MoE (oracle/_ref); the closed
 * lk_moe wheel itself cannot be run (see oracle/moe_oracle.py header).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline float bf16_to_f32(uint16_t v) {
  uint32_t u = (uint32_t)v << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static inline uint16_t f32_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static float fp8_lut[256];
static int fp8_lut_ready = 0;
static void init_fp8_lut(void) {
  if (fp8_lut_ready) return;
  for (int i = 0; i < 256; ++i) {
    int s = i >> 7, e = (i >> 3) & 15, m = i & 7;
    float v;
    if (e == 0) v = ldexpf((float)m, -9);
    else if (e == 15 && m == 7) v = NAN;
    else v = ldexpf(1.0f + m / 8.0f, e - 7);
    fp8_lut[i] = s ? -v : v;
  }
  fp8_lut_ready = 1;
}

int moe_ref_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to its children: the CPU arm sets the thread count it reports explicitly */
void moe_ref_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* bf16 weights.  hidden bf16 [M,H], w13 bf16 [E,2I,H], w2 bf16 [E,H,I], ids i32 [M,k], w f32 [M,k], out f32 [M,H] */
void moe_ref_forward_bf16(const uint16_t* hidden, const uint16_t* w13, const uint16_t* w2, const int32_t* ids,
                          const float* tw, float* out, int M, int k, int E, int H, int I) {
  float* x = (float*)malloc(sizeof(float) * H);
  float* act = (float*)malloc(sizeof(float) * I);
  for (int t = 0; t < M; ++t) {
    for (int h = 0; h < H; ++h) {
      x[h] = bf16_to_f32(hidden[(size_t)t * H + h]);
      out[(size_t)t * H + h] = 0.f;
    }
    for (int j = 0; j < k; ++j) {
      const int e = ids[t * k + j];
      if (e < 0 || e >= E) continue;
      const uint16_t* W1 = w13 + (size_t)e * 2 * I * H;
      const uint16_t* W2 = w2 + (size_t)e * H * I;
#pragma omp parallel for schedule(static)
      for (int i = 0; i < I; ++i) {
        const uint16_t* g = W1 + (size_t)i * H;
        const uint16_t* u = W1 + (size_t)(I + i) * H;
        float sg = 0.f, su = 0.f;
        for (int h = 0; h < H; ++h) {
          sg += bf16_to_f32(g[h]) * x[h];
          su += bf16_to_f32(u[h]) * x[h];
        }
        sg = bf16_to_f32(f32_to_bf16(sg));
        su = bf16_to_f32(f32_to_bf16(su));
        const float a = sg / (1.0f + expf(-sg)) * su;
        act[i] = bf16_to_f32(f32_to_bf16(a));
      }
      const float wt = tw[t * k + j];
#pragma omp parallel for schedule(static)
      for (int h = 0; h < H; ++h) {
        const uint16_t* r = W2 + (size_t)h * I;
        float s = 0.f;
        for (int i = 0; i < I; ++i) s += bf16_to_f32(r[i]) * act[i];
        out[(size_t)t * H + h] += wt * s;
      }
    }
  }
  free(x);
  free(act);
}

/* FP8 e4m3 weights with f32 [128,128] block scales, weight-only dequant (W8A16), bf16 activations. */
void moe_ref_forward_fp8_block(const uint16_t* hidden, const uint8_t* w13, const float* s13, const uint8_t* w2,
                               const float* s2, const int32_t* ids, const float* tw, float* out, int M, int k, int E,
                               int H, int I) {
  init_fp8_lut();
  const int HB = (H + 127) / 128, IB = (I + 127) / 128, NB1 = (2 * I + 127) / 128;
  float* x = (float*)malloc(sizeof(float) * H);
  float* act = (float*)malloc(sizeof(float) * I);
  for (int t = 0; t < M; ++t) {
    for (int h = 0; h < H; ++h) {
      x[h] = bf16_to_f32(hidden[(size_t)t * H + h]);
      out[(size_t)t * H + h] = 0.f;
    }
    for (int j = 0; j < k; ++j) {
      const int e = ids[t * k + j];
      if (e < 0 || e >= E) continue;
      const uint8_t* W1 = w13 + (size_t)e * 2 * I * H;
      const uint8_t* W2 = w2 + (size_t)e * H * I;
      const float* S1 = s13 + (size_t)e * NB1 * HB;
      const float* S2 = s2 + (size_t)e * HB * IB;
#pragma omp parallel for schedule(static)
      for (int i = 0; i < I; ++i) {
        float acc[2];
        for (int half = 0; half < 2; ++half) {
          const int row = half * I + i;
          const uint8_t* r = W1 + (size_t)row * H;
          const float* sr = S1 + (size_t)(row / 128) * HB;
          float s = 0.f;
          for (int hb = 0; hb < HB; ++hb) {
            float p = 0.f;
            const int h1 = (hb + 1) * 128 < H ? (hb + 1) * 128 : H;
            for (int h = hb * 128; h < h1; ++h) p += fp8_lut[r[h]] * x[h];
            s += p * sr[hb];
          }
          acc[half] = bf16_to_f32(f32_to_bf16(s));
        }
        const float a = acc[0] / (1.0f + expf(-acc[0])) * acc[1];
        act[i] = bf16_to_f32(f32_to_bf16(a));
      }
      const float wt = tw[t * k + j];
#pragma omp parallel for schedule(static)
      for (int h = 0; h < H; ++h) {
        const uint8_t* r = W2 + (size_t)h * I;
        const float* sr = S2 + (size_t)(h / 128) * IB;
        float s = 0.f;
        for (int ib = 0; ib < IB; ++ib) {
          float p = 0.f;
          const int i1 = (ib + 1) * 128 < I ? (ib + 1) * 128 : I;
          for (int i = ib * 128; i < i1; ++i) p += fp8_lut[r[i]] * act[i];
          s += p * sr[ib];
        }
        out[(size_t)t * H + h] += wt * s;
      }
    }
  }
  free(x);
  free(act);
}

/* 4-bit weight-only formats (W4A16), checkpoint layouts of SURVEY.md 8a row W; follows
 * oracle/moe_oracle.py::dequant_int4_group / dequant_nvfp4 / dequant_mxfp4 (weights dequantised to bf16-rounded
 * values, fp32 accumulation).  fmt 1: INT4 uint4b8, bf16 group-32 scales [E,N,K/32];  fmt 2: NVFP4, e4m3 block-16
 * scales [E,N,K/16] + f32 global dequant factors g13 [E,2] (gate, up) / g2 [E];  fmt 3: MXFP4, e8m0 block-32
 * scales [E,N,K/32].  Packed weights uint8 [E,N,K/2], low nibble = even k. */
static const float e2m1_lut[16] = {0.f, 0.5f, 1.f, 1.5f, 2.f, 3.f, 4.f, 6.f, -0.f, -0.5f, -1.f, -1.5f, -2.f, -3.f, -4.f, -6.f};

static inline float w4_group_scale(int fmt, const uint8_t* sc, size_t idx, float g) {
  if (fmt == 1) return bf16_to_f32(((const uint16_t*)sc)[idx]);
  if (fmt == 2) return fp8_lut[sc[idx]] * g;
  uint32_t u = (uint32_t)sc[idx] << 23; /* e8m0 -> 2^(E-127) */
  float f;
  memcpy(&f, &u, 4);
  return f;
}

/* dot of one packed row (K values) with x; dequantised weights rounded to bf16 like the oracle */
static float w4_row_dot(int fmt, const uint8_t* row, const uint8_t* sc, size_t sc_row0, float g, const float* x, int K) {
  if (fmt >= 16) {
    /* throughput form used by the timed CPU baseline: the group scale is factored out of the inner sum
     * (same value up to fp32 rounding, no per-weight bf16 rounding) */
    const int f = fmt - 16, grp = f == 2 ? 16 : 32;
    float s = 0.f;
    for (int k0 = 0; k0 < K; k0 += grp) {
      float p = 0.f;
      if (f == 1) {
        for (int k = k0; k < k0 + grp; k += 2) {
          const uint8_t b = row[k >> 1];
          p += (float)((int)(b & 15) - 8) * x[k] + (float)((int)(b >> 4) - 8) * x[k + 1];
        }
      } else {
        for (int k = k0; k < k0 + grp; k += 2) {
          const uint8_t b = row[k >> 1];
          p += e2m1_lut[b & 15] * x[k] + e2m1_lut[b >> 4] * x[k + 1];
        }
      }
      s += p * w4_group_scale(f, sc, sc_row0 + k0 / grp, g);
    }
    return s;
  }
  const int grp = fmt == 2 ? 16 : 32;
  float s = 0.f;
  for (int k0 = 0; k0 < K; k0 += grp) {
    const float scale = w4_group_scale(fmt, sc, sc_row0 + k0 / grp, g);
    float p = 0.f;
    for (int k = k0; k < k0 + grp; k += 2) {
      const uint8_t b = row[k >> 1];
      float w0, w1;
      if (fmt == 1) {
        w0 = (float)((int)(b & 15) - 8) * scale;
        w1 = (float)((int)(b >> 4) - 8) * scale;
      } else {
        w0 = e2m1_lut[b & 15] * scale;
        w1 = e2m1_lut[b >> 4] * scale;
      }
      p += bf16_to_f32(f32_to_bf16(w0)) * x[k] + bf16_to_f32(f32_to_bf16(w1)) * x[k + 1];
    }
    s += p;
  }
  return s;
}

void moe_ref_forward_w4(const uint16_t* hidden, const uint8_t* w13, const uint8_t* s13, const uint8_t* w2,
                        const uint8_t* s2, const float* g13, const float* g2, const int32_t* ids, const float* tw,
                        float* out, int M, int k, int E, int H, int I, int fmt) {
  init_fp8_lut();
  const int f = fmt >= 16 ? fmt - 16 : fmt; /* fmt + 16 selects the throughput form of the row dot */
  const int grp = f == 2 ? 16 : 32;
  float* x = (float*)malloc(sizeof(float) * H);
  float* act = (float*)malloc(sizeof(float) * I);
  for (int t = 0; t < M; ++t) {
    for (int h = 0; h < H; ++h) {
      x[h] = bf16_to_f32(hidden[(size_t)t * H + h]);
      out[(size_t)t * H + h] = 0.f;
    }
    for (int j = 0; j < k; ++j) {
      const int e = ids[t * k + j];
      if (e < 0 || e >= E) continue;
      const uint8_t* W1 = w13 + (size_t)e * 2 * I * (H / 2);
      const uint8_t* W2 = w2 + (size_t)e * H * (I / 2);
      const size_t s1_base = (size_t)e * 2 * I * (H / grp);
      const size_t s2_base = (size_t)e * H * (I / grp);
      const float gg = (f == 2 && g13) ? g13[e * 2] : 1.f, gu = (f == 2 && g13) ? g13[e * 2 + 1] : 1.f;
      const float gd = (f == 2 && g2) ? g2[e] : 1.f;
#pragma omp parallel for schedule(static)
      for (int i = 0; i < I; ++i) {
        float sg = w4_row_dot(fmt, W1 + (size_t)i * (H / 2), s13, s1_base + (size_t)i * (H / grp), gg, x, H);
        float su = w4_row_dot(fmt, W1 + (size_t)(I + i) * (H / 2), s13, s1_base + (size_t)(I + i) * (H / grp), gu, x, H);
        sg = bf16_to_f32(f32_to_bf16(sg));
        su = bf16_to_f32(f32_to_bf16(su));
        const float a = sg / (1.0f + expf(-sg)) * su;
        act[i] = bf16_to_f32(f32_to_bf16(a));
      }
      const float wt = tw[t * k + j];
#pragma omp parallel for schedule(static)
      for (int h = 0; h < H; ++h) {
        const float s = w4_row_dot(fmt, W2 + (size_t)h * (I / 2), s2, s2_base + (size_t)h * (I / grp), gd, act, I);
        out[(size_t)t * H + h] += wt * s;
      }
    }
  }
  free(x);
  free(act);
}

/* ------------------------------------------------------------------------------------------------------------------
 * Expert-major batched form of the 4-bit path (what a tuned CPU engine does for decode BATCHES, and what bench.py's
 * CPU arm times at the real batch): the tokens routed to an expert are processed together, every weight row is
 * dequantised ONCE into a thread-local fp32 buffer (group scale applied) and then dotted with all of the expert's
 * tokens; the dot products are `omp simd` reductions (AVX-512 on the bench hosts).  Threads split the rows of one
 * expert.  Same arithmetic as moe_ref_forward_w4 up to fp32 summation order and the per-weight bf16 rounding (which
 * the throughput form drops); pinned to it in tests/test_c_port.py. */
static void w4_dequant_row(int f, const uint8_t* row, const uint8_t* sc, size_t sc_row0, float g, float* dst, int K) {
  const int grp = f == 2 ? 16 : 32;
  for (int k0 = 0; k0 < K; k0 += grp) {
    const float scale = w4_group_scale(f, sc, sc_row0 + k0 / grp, g);
    if (f == 1) {
      for (int k = k0; k < k0 + grp; k += 2) {
        const uint8_t b = row[k >> 1];
        dst[k] = (float)((int)(b & 15) - 8) * scale;
        dst[k + 1] = (float)((int)(b >> 4) - 8) * scale;
      }
    } else {
      for (int k = k0; k < k0 + grp; k += 2) {
        const uint8_t b = row[k >> 1];
        dst[k] = e2m1_lut[b & 15] * scale;
        dst[k + 1] = e2m1_lut[b >> 4] * scale;
      }
    }
  }
}

static inline float dot_f32(const float* a, const float* b, int n) {
  float s = 0.f;
#pragma omp simd reduction(+ : s)
  for (int i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

void moe_ref_forward_w4_batched(const uint16_t* hidden, const uint8_t* w13, const uint8_t* s13, const uint8_t* w2,
                                const uint8_t* s2, const float* g13, const float* g2, const int32_t* ids, const float* tw,
                                float* out, int M, int k, int E, int H, int I, int fmt) {
  init_fp8_lut();
  const int f = fmt >= 16 ? fmt - 16 : fmt;
  const int grp = f == 2 ? 16 : 32;
  float* x = (float*)malloc(sizeof(float) * (size_t)M * H);
  for (size_t i = 0; i < (size_t)M * H; ++i) {
    x[i] = bf16_to_f32(hidden[i]);
    out[i] = 0.f;
  }
  int* slot = (int*)malloc(sizeof(int) * (size_t)M * k);   /* slots of the current expert */
  float* act = (float*)malloc(sizeof(float) * (size_t)M * k * I);
  const int nthr = moe_ref_num_threads();
  float* rowbuf = (float*)malloc(sizeof(float) * (size_t)nthr * 2 * (H > I ? H : I));
  for (int e = 0; e < E; ++e) {
    int n = 0;
    for (int s = 0; s < M * k; ++s)
      if (ids[s] == e) slot[n++] = s;
    if (!n) continue;
    const uint8_t* W1 = w13 + (size_t)e * 2 * I * (H / 2);
    const uint8_t* W2 = w2 + (size_t)e * H * (I / 2);
    const size_t s1_base = (size_t)e * 2 * I * (H / grp);
    const size_t s2_base = (size_t)e * H * (I / grp);
    const float gg = (f == 2 && g13) ? g13[e * 2] : 1.f, gu = (f == 2 && g13) ? g13[e * 2 + 1] : 1.f;
    const float gd = (f == 2 && g2) ? g2[e] : 1.f;
#pragma omp parallel
    {
#ifdef _OPENMP
      float* rb = rowbuf + (size_t)omp_get_thread_num() * 2 * (H > I ? H : I);
#else
      float* rb = rowbuf;
#endif
#pragma omp for schedule(static)
      for (int i = 0; i < I; ++i) {
        w4_dequant_row(f, W1 + (size_t)i * (H / 2), s13, s1_base + (size_t)i * (H / grp), gg, rb, H);
        w4_dequant_row(f, W1 + (size_t)(I + i) * (H / 2), s13, s1_base + (size_t)(I + i) * (H / grp), gu, rb + H, H);
        for (int q = 0; q < n; ++q) {
          const float* xt = x + (size_t)(slot[q] / k) * H;
          float sg = bf16_to_f32(f32_to_bf16(dot_f32(rb, xt, H)));
          float su = bf16_to_f32(f32_to_bf16(dot_f32(rb + H, xt, H)));
          act[(size_t)q * I + i] = bf16_to_f32(f32_to_bf16(sg / (1.0f + expf(-sg)) * su));
        }
      }
#pragma omp for schedule(static)
      for (int h = 0; h < H; ++h) {
        w4_dequant_row(f, W2 + (size_t)h * (I / 2), s2, s2_base + (size_t)h * (I / grp), gd, rb, I);
        for (int q = 0; q < n; ++q)
          out[(size_t)(slot[q] / k) * H + h] += tw[slot[q]] * dot_f32(rb, act + (size_t)q * I, I);
      }
    }
  }
  free(x);
  free(slot);
  free(act);
  free(rowbuf);
}
