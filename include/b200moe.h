/*
 * b200moe.h — C ABI of libb200moe.so: the B200-native (sm_100a) replacement for the expert hot path that
 * Lvllm delegates to the closed `lk_moe` wheel, plus the routing / permutation / decode-attention
 * operators that sit next to it in a decoder layer.
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes (no torch types) and returns 0 on
 * success or a negative error code; b200moe_last_error() returns a thread-local message.
 * All device work is enqueued on the caller's `stream` (a cudaStream_t passed as void*), is CUDA-graph
 * capturable (no allocation, no synchronisation, pointer-stable workspaces) unless stated otherwise.
 *
 * Reference interfaces replaced (file:line in guqiong96/Lvllm):
 *   lk_moe.MOEConfigV2 fields ............. vllm/model_executor/layers/fused_moe/routed_experts.py:1490-1511
 *   lk_moe.MOE_* constructors ............. routed_experts.py:1514-1533, 1598-1616, 1650-1668, 1724-1742, 1795-1813
 *   lk_moe.cpu_decode ..................... routed_experts.py:1840-1855
 *   lk_moe.cpu_prefill .................... routed_experts.py:1858-1882
 *   lk_moe.gpu_prefill .................... routed_experts.py:1884-1899
 *   _moe_C.topk_softmax / topk_sigmoid .... csrc/libtorch_stable/moe/topk_softmax_kernels.cu:822-897
 *   _moe_C.grouped_topk ................... csrc/libtorch_stable/moe/grouped_topk_kernels.cu:1447-1556
 *   GateLinear.forward (router GEMM) ...... vllm/model_executor/layers/fused_moe/router/gate_linear.py:171-221
 *   _moe_C.moe_permute / moe_unpermute .... csrc/libtorch_stable/moe/moe_permute_unpermute_op.cu:59-207
 *   _C.sm100_cutlass_mla_decode ........... csrc/libtorch_stable/attention/mla/sm100_cutlass_mla_kernel.cu:225-262
 *   paged GQA decode (FlashInfer backend) . vllm/v1/attention/backends/flashinfer.py (cache layout :398-409)
 */
#ifndef B200MOE_H_
#define B200MOE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes --------------------------------------------------------------------------- */
#define B200_OK 0
#define B200_ERR_INVALID (-1)   /* bad argument / unsupported shape */
#define B200_ERR_CUDA (-2)      /* a CUDA runtime call failed */
#define B200_ERR_NO_DEVICE (-3) /* no sm_100 device */

const char* b200moe_last_error(void);
/* library build info: "b200moe <version> sm_100a" */
const char* b200moe_version(void);

/* ---- lk_moe.MOEConfigV2 (routed_experts.py:1490-1511) ------------------------------------------ */
typedef struct b200moe_config {
  int32_t num_processes;    /* TP or EP world size the layer is sharded over */
  int32_t process_id;       /* this rank */
  int32_t gpu_id;           /* CUDA device ordinal */
  int32_t has_gate_proj;    /* 1: w13 = [gate;up] rows, 0: single up projection */
  int32_t expert_num;       /* LOCAL experts held by this object */
  int32_t top_k;
  int32_t hidden_size;      /* H */
  int32_t intermediate_size;/* I per partition */
  int32_t max_batch_size;   /* max tokens per prefill call */
  int32_t max_num_seqs;     /* max tokens per decode call */
  int32_t stride;           /* lk_moe CPU tiling knobs: accepted, unused */
  int32_t group_min_len;
  int32_t group_max_len;
  int32_t groupN;           /* weight rows per scale row */
  int32_t groupK;           /* weight cols per scale col */
  int32_t activation_type;  /* 0 silu, 1 swigluoai (layout stated by B200MOE_SWIGLUOAI_LAYOUT), 2 relu^2 (non-gated) */
  float swiglu_alpha;
  float swiglu_limit;
  int32_t use_gpu_prefill;
} b200moe_config;

/* weight formats == the lk_moe.MOE_* class families */
#define B200_FMT_16BIT 0 /* MOE_BF16 / MOE_FP16: w13 [E,2I,H], w2 [E,H,I] in the activation dtype */
#define B200_FMT_FP8 1   /* MOE_FP8: e4m3 + f32 block scales [E,2I/gN,H/gK] (or per-tensor [E,2]/[E]) */
#define B200_FMT_WNA16 2 /* MOE_WNA16: uint4b8 packed u8 [E,2I,H/2], group scales act-dtype [E,2I,H/g] */
#define B200_FMT_NVFP4 3 /* MOE_NVFP4: e2m1 u8 [E,2I,H/2], e4m3 scales [E,2I,H/16], f32 global [E,2]/[E] */
#define B200_FMT_MXFP4 4 /* MOE_MXFP4: e2m1 u8 [E,2I,H/2], e8m0 scales [E,2I,H/32] */

#define B200_ACT_BF16 0
#define B200_ACT_FP16 1

typedef struct b200moe_layer* b200moe_handle;

/* Construct one MoE layer object.  Weight pointers are raw checkpoint-layout tensors (see format list);
 * `weights_on_device` = 0: pageable/pinned HOST memory (the lk_moe contract: the caller frees them right
 * after the call, so everything is copied/repacked into HBM before returning);  1: device memory (fast
 * path used by loaders that already staged the checkpoint in HBM).  Absent tensors are NULL.
 * Synchronous; not graph-capturable. */
int b200moe_create(const b200moe_config* cfg, const void* w13, const void* w2, const void* w13_scale,
                   const void* w2_scale, const void* w13_global_scale, const void* w2_global_scale,
                   int format, int act_dtype, int weights_on_device, b200moe_handle* out);
/* Per-expert ingest (reference weight_loader routed_experts.py:644-1168 builds stacked [E, ...] CPU
 * parameters first): create the layer without weights, feed expert ranges straight from wherever the checkpoint shards
 * live (host or device memory, raw checkpoint layout of those experts only), then finalize.  Each range is staged
 * through a bounded device buffer and re-tiled into its final place; the caller may free / unmap its tensors as soon as
 * b200moe_load_experts returns.  b200moe_create is this sequence over the whole stacked tensor. */
int b200moe_create_empty(const b200moe_config* cfg, int format, int act_dtype, b200moe_handle* out);
int b200moe_load_experts(b200moe_handle h, int first_expert, int num_experts, const void* w13, const void* w2,
                         const void* w13_scale, const void* w2_scale, const void* w13_global_scale,
                         const void* w2_global_scale, int weights_on_device);
int b200moe_finalize(b200moe_handle h);
int b200moe_destroy(b200moe_handle h);
/* bytes of HBM held by the layer (repacked weights + scales + workspaces) */
int64_t b200moe_device_bytes(b200moe_handle h);

/* layer introspection (tests / bring-up): what = 0: 1 when an MXFP4 layer runs the native block-scaled W4A8-MX kernel
 * (0: W4A16 dequant kernel); 1: tokens per pass; 2: 1 when w13 arrived with interleaved gate/up rows (SwiGLU-OAI,
 * B200MOE_SWIGLUOAI_LAYOUT=interleaved); 3: 4-bit flavour (0 none, 1 int4, 2 nvfp4, 3 mxfp4); 4: 1 when an FP8 layer
 * runs in ue8m0 mode (B200MOE_FP8_E8M0=1: power-of-two block scales, the reference's DeepGEMM-on-Blackwell numerics,
 * fp8_utils.py:986-1043).  < 0: unknown. */
int b200moe_query(b200moe_handle h, int what);

/* lk_moe.cpu_decode(stream, M, k, hidden, ids, weights, out_f32): all DEVICE pointers; hidden [M,H] in
 * the activation dtype, ids int32 [M,k] (<0 or >=E: skip), weights f32 [M,k], out f32 [M,H].
 * Graph-capturable. */
int b200moe_cpu_decode(b200moe_handle h, void* stream, int num_tokens, int top_k, const void* hidden,
                       const int32_t* topk_ids, const float* topk_weights, float* out_f32);
/* lk_moe.cpu_prefill(M, k, ids, weights, hidden, out_f32): all HOST pointers; synchronous (copies in,
 * runs the same device path, copies out). */
int b200moe_cpu_prefill(b200moe_handle h, int num_tokens, int top_k, const int32_t* topk_ids_host,
                        const float* topk_weights_host, const void* hidden_host, float* out_f32_host);
/* lk_moe.gpu_prefill(hidden, out, ids, weights, M, k, stream): DEVICE pointers, out in the activation
 * dtype. */
int b200moe_gpu_prefill(b200moe_handle h, const void* hidden, void* out, const int32_t* topk_ids,
                        const float* topk_weights, int num_tokens, int top_k, void* stream);

/* ---- routing operators ---------------------------------------------------------------------------- */
/* softmax (scoring=0) / sigmoid (scoring=1) top-k; logits [M,E] f32/bf16/f16 (logits_dtype 0 f32, 1 bf16,
 * 2 f16); bias f32 [E] or NULL (selection only); ids int32 [M,k], weights f32 [M,k],
 * token_expert_indices int32 [M,k] = j*M+t or NULL.  Ties -> lower expert index. */
int b200_topk_gating(void* stream, const void* logits, int logits_dtype, const float* bias, int num_tokens,
                     int num_experts, int top_k, int scoring, int renormalize, float routed_scaling_factor,
                     float* topk_weights, int32_t* topk_ids, int32_t* token_expert_indices);
/* DeepSeek group-limited routing (scoring 0 none, 1 sigmoid); bias f32 [E] or NULL. */
int b200_grouped_topk(void* stream, const void* logits, int logits_dtype, const float* bias, int num_tokens,
                      int num_experts, int n_group, int topk_group, int top_k, int scoring, int renormalize,
                      float routed_scaling_factor, float* topk_weights, int32_t* topk_ids);
/* EP id remap: local = expert_map[clamp(id)] ; id<0 -> -1  (routed_experts.py:1332-1342) */
int b200_global_to_local_ids(void* stream, const int32_t* topk_ids, const int32_t* expert_map, int num_global,
                             int64_t numel, int32_t* local_ids);

/* Fused router: logits = hidden . gate_weight^T (fp32 accumulate on tcgen05, split-K with a deterministic fixed-order
 * reduction), then top-k routing and the optional EP id remap, in ONE kernel.  Replaces GateLinear.forward
 * (router/gate_linear.py:171-221, csrc/libtorch_stable/moe/dsv3_router_gemm_entry.cu:112) + fused_topk / grouped_topk
 * (fused_topk_router.py:81-124, grouped_topk_kernels.cu:523-678) + global_to_local_expert_ids (routed_experts.py:1332-1342).
 * hidden [M,H] and gate_weight [E,H] in the activation dtype (16-byte aligned, H % 64 == 0, E <= 1024);
 * mode 0 softmax top-k, 1 sigmoid top-k (bias for selection only), 2 grouped top-k (scoring 0 none / 1 sigmoid, bias,
 * n_group, topk_group); topk_ids are GLOBAL expert ids; local_ids (optional) = expert_map[id] (-1 stays -1; with
 * expert_map NULL a copy); logits_out (optional) f32 [M,E].
 * Shared experts (reference runner/shared_experts.py) are folded into the routed launch as always-on
 * experts: with n_shared > 0 every output row has top_k + n_shared columns, the extra ones carrying weight
 * shared_weight, global id E + s and local id shared_local_base + s (-1 when shared_local_base < 0: another rank holds
 * them); the caller appends the shared expert(s) to the layer's stacked weights and calls cpu_decode with
 * top_k + n_shared — the stream-K schedule then overlaps them with the routed experts by construction.  workspace: b200_router_workspace_bytes() bytes of device
 * memory, zero-filled once by the caller (the kernel hands its counters back clean, so it is CUDA-graph replayable). */
int64_t b200_router_workspace_bytes(int num_tokens, int num_experts, int hidden_size);
int b200_router_topk(void* stream, const void* hidden, int act_dtype, const void* gate_weight, int num_tokens,
                     int num_experts, int hidden_size, const float* bias, int mode, int scoring, int top_k, int renormalize,
                     int n_group, int topk_group, float routed_scaling_factor, const int32_t* expert_map, int n_shared,
                     int shared_local_base, float shared_weight, void* workspace, int64_t workspace_bytes,
                     float* topk_weights, int32_t* topk_ids, int32_t* local_ids, float* logits_out);

/* ---- permutation operators (stable sort by expert) ------------------------------------------------- */
/* sorted_slot int32 [M*k] (source slot t*k+j of each permuted row, valid rows first in (expert, slot)
 * order), expert_first_offset int64 [E+1], inv_perm int32 [M*k] (-1 for skipped slots);
 * permuted_hidden [M*k,H] (same dtype as hidden, 2 bytes/elt) or NULL. */
int b200_moe_permute(void* stream, const void* hidden, const int32_t* topk_ids, int num_tokens, int top_k,
                     int num_local_experts, int hidden_size, int32_t* sorted_slot,
                     int64_t* expert_first_offset, int32_t* inv_perm, void* permuted_hidden);
/* out[t,:] = sum_j w[t,j] * permuted[inv_perm[t,j],:] (fp32 accumulate); act dtype in, out_dtype
 * 0 bf16 / 1 fp16 / 2 f32 out. */
int b200_moe_unpermute(void* stream, const void* permuted, int act_dtype, const float* topk_weights,
                       const int32_t* inv_perm, int num_tokens, int top_k, int hidden_size, void* out,
                       int out_dtype);

/* fp32 -> activation dtype cast of the lk_moe output (routed_experts.py:1855) fused with an RMS
 * normalisation: out[t] = cast(in[t] * gain * rsqrt(mean(in[t]^2) + eps)); out_dtype 0 bf16 / 1 fp16. */
int b200_rmsnorm_cast(void* stream, const float* in, void* out, int num_tokens, int hidden_size, float gain,
                      float eps, int out_dtype);

/* ---- decode attention ----------------------------------------------------------------------------- */
/* Paged MLA decode, absorbed form.  q_nope [B,Hq,512], q_pe [B,Hq,64] (bf16), kv cache
 * [num_pages,page_size,576] bf16, seq_lens int32 [B], page_table int32 [B,max_pages];
 * out bf16 [B,Hq,512], lse f32 [B,Hq] (may be NULL).  workspace: b200_mla_decode_workspace_bytes(). */
int64_t b200_mla_decode_workspace_bytes(int batch, int num_heads, int num_splits);
int b200_mla_decode(void* stream, const void* q_nope, const void* q_pe, const void* kv_cache,
                    const int32_t* seq_lens, const int32_t* page_table, int batch, int num_heads,
                    int page_size, int max_pages, float sm_scale, int num_splits, void* workspace, void* out,
                    float* lse);
/* Same with an e4m3 latent cache (kv_cache_dtype "fp8": [num_pages,page_size,576] bytes, half the HBM traffic) and,
 * optionally, e4m3 queries — the fp8 mode of the reference's Blackwell MLA (backends/mla/cutlass_mla.py:44-45,
 * tests/kernels/attention/test_cutlass_mla_decode.py:101-110).  q_dtype / kv_dtype: 0 bf16, 1 e4m3; descale_q /
 * descale_k are the per-tensor dequantisation scales (q_scale * k_scale folds into the softmax scale, k_scale into the
 * output).  The cache is widened to bf16 on its way into shared memory: probabilities are not re-quantised. */
int b200_mla_decode_ex(void* stream, const void* q_nope, const void* q_pe, int q_dtype, const void* kv_cache,
                       int kv_dtype, float descale_q, float descale_k, const int32_t* seq_lens,
                       const int32_t* page_table, int batch, int num_heads, int page_size, int max_pages, float sm_scale,
                       int num_splits, void* workspace, void* out, float* lse);
/* The MLA decode kernel's neighbours, bf16:
 *  b200_mla_rope_cache_write  RoPE of q_pe [T,Hq,64] and k_pe [T,64] in place (rotary_embedding/base.py:161-201; cos_sin_cache
 *      bf16 [max_pos,64] = cos | sin; is_neox 0 = GPT-J pairs (DeepSeek), 1 = NeoX halves) fused with concat_and_cache_mla
 *      (csrc/libtorch_stable/cache_kernels.cu:403-444): row slot_mapping[t] (i64, < 0 = padded token) of the paged latent
 *      cache [blocks*block_size, 576] <- [kv_c[t] (512) | rotated k_pe[t]]; kv_dtype 1 = e4m3 cache storing value / kv_scale;
 *  b200_mla_q_absorb          ql_nope [T,Hq,512] = q_nope [T,Hq,128] x W_UK_T [Hq,128,512]  (mla_attention.py:875-893);
 *  b200_mla_decode_vup        b200_mla_decode_ex whose split merge is fused with the v up-projection
 *      out_v [B,Hq,128] = o [B,Hq,512] x W_UV [Hq,512,128] (mla_attention.py:1154-1176); out_latent (optional) = o. */
int b200_mla_rope_cache_write(void* stream, void* q_pe, void* k_pe, const void* kv_c, const int64_t* positions,
                              const void* cos_sin_cache, int is_neox, const int64_t* slot_mapping, void* kv_cache,
                              int kv_dtype, float kv_scale, int num_tokens, int num_heads);
int b200_mla_q_absorb(void* stream, const void* q_nope, const void* w_uk_t, void* out, int num_tokens, int num_heads);
int b200_mla_decode_vup(void* stream, const void* q_nope, const void* q_pe, int q_dtype, const void* kv_cache, int kv_dtype,
                        float descale_q, float descale_k, const int32_t* seq_lens, const int32_t* page_table, int batch,
                        int num_heads, int page_size, int max_pages, float sm_scale, int num_splits, void* workspace,
                        const void* w_uv, void* out_v, void* out_latent, float* lse);
/* Paged GQA decode.  q [B,Hq,D] bf16, k_cache/v_cache [num_pages,page_size,Hkv,D] bf16, D=128;
 * out bf16 [B,Hq,D], lse f32 [B,Hq] or NULL. */
int64_t b200_gqa_decode_workspace_bytes(int batch, int num_q_heads, int head_dim, int num_splits);
int b200_gqa_decode(void* stream, const void* q, const void* k_cache, const void* v_cache,
                    const int32_t* seq_lens, const int32_t* page_table, int batch, int num_q_heads,
                    int num_kv_heads, int head_dim, int page_size, int max_pages, float sm_scale,
                    int num_splits, void* workspace, void* out, float* lse);

/* ---- expert-parallel exchange over NVLink peer memory ---------------------------------------------- */
/* Sum-combine of per-rank partial MoE outputs through peer-mapped buffers (the lk_moe EP/TP contract:
 * replicated tokens, local experts, all-reduce; moe_runner.py:488-494).  peer_bufs[r] is rank r's
 * staging buffer mapped into this process (CUDA IPC), peer_flags[r] its flag word array.
 * out = sum over ranks (fixed rank order, bit-identical on every rank) of each rank's local_in[0:numel];
 * fp32 in, out_dtype out (0 bf16, 1 fp16, 2 f32).  CUDA-graph replayable (epochs live in device memory). */
/* buffers: data buffer of 2*slot_elems floats and a flag buffer of b200_ep_flag_bytes() bytes per rank,
 * both created with b200_ep_buffer_create (cudaMalloc + zero + IPC export) and opened by the peers. */
int64_t b200_ep_flag_bytes(void);
int b200_ep_buffer_create(int64_t bytes, void** dev_ptr, void* ipc_handle_64B);
int b200_ep_buffer_open(const void* ipc_handle_64B, void** dev_ptr);
int b200_ep_buffer_close(void* dev_ptr, int is_owner);
int b200_ep_allreduce(void* stream, void* const* peer_bufs, int32_t* const* peer_flags, int world, int rank,
                      const float* local_in, int64_t numel, int64_t slot_elems, void* out, int out_dtype);

/* Push all-reduce fused with residual add + RMSNorm (the op pair that follows the MoE block in a decoder layer;
 * reference moe_runner.py:462-496 + fused_add_rms_norm, flashinfer_all_reduce.py): every rank stores its fp32
 * [num_tokens, hidden] partial straight into every peer's buffer, then reduces from local memory in fixed rank order:
 *   x = sum_r partial_r (+ residual);  residual <- cast(x) (in place, may be NULL);  sum_out <- x (f32, may be NULL);
 *   out = cast(x * rsqrt(mean(x^2) + eps)) * gamma   (gamma act-dtype [hidden]; NULL: out = cast(x * rsqrt(..) * gain)).
 * peer_bufs[r]: data buffers of (2 + 2 * world) * slot_elems floats (b200_ep_buffer_create; the first two slots are the
 * pull all-reduce's), flags as b200_ep_allreduce.
 * hidden <= 8192, num_tokens * hidden <= slot_elems.  CUDA-graph replayable. */
int b200_ep_allreduce_norm(void* stream, void* const* peer_bufs, int32_t* const* peer_flags, int world, int rank,
                           const float* local_in, int num_tokens, int hidden, int64_t slot_elems, void* residual,
                           const void* gamma, float gain, float eps, void* out, float* sum_out, int act_dtype);

/* Dispatch / combine all-to-all for token-sharded callers (DP attention + EP experts): rank r owns tokens
 * [r*m_local, (r+1)*m_local) of the global batch M = world*m_local and experts [r*experts_per_rank, ...).
 * Replaces the dispatch/combine pair around the experts in the reference's EP path
 * (vllm/model_executor/layers/fused_moe/runner/moe_runner.py:462-496, 600).
 * Every rank creates one data buffer of b200_ep_a2a_layout(...) bytes with b200_ep_buffer_create and opens the
 * peers'; the flag buffers are the ones of b200_ep_allreduce.  b200_ep_a2a_layout returns the buffer size and the
 * byte offsets of X [M][hidden] 16-bit, IDS [M][top_k] i32 (ids local to the rank, -1 = skip), W [M][top_k] f32,
 * Y [M][hidden] f32 inside it.
 *   b200_ep_dispatch pushes this rank's rows to the ranks owning one of their experts; when it retires, the local
 *     X / IDS / W are complete: run b200moe_cpu_decode(handle, stream, M, top_k, buf+off_x, buf+off_ids, buf+off_w,
 *     buf+off_y) on them.
 *   b200_ep_combine then pulls the partial rows of this rank's tokens from those ranks and writes
 *     out [m_local][hidden] (out_dtype 0 bf16, 1 fp16, 2 f32), summed in ascending rank order.
 * ids_global are global expert ids (< 0 = padding).  All ranks must call both with identical shapes; both are
 * CUDA-graph capturable (epochs live in device memory). */
int64_t b200_ep_a2a_layout(int global_tokens, int hidden, int top_k, int64_t* off_x, int64_t* off_ids, int64_t* off_w,
                           int64_t* off_y);
int b200_ep_dispatch(void* stream, void* const* peer_bufs, int32_t* const* peer_flags, int world, int rank,
                     const void* hidden_local, const int32_t* ids_global, const float* weights, int m_local, int top_k,
                     int hidden, int experts_per_rank);
int b200_ep_combine(void* stream, void* const* peer_bufs, int32_t* const* peer_flags, int world, int rank,
                    const int32_t* ids_global, int m_local, int top_k, int hidden, int experts_per_rank, void* out,
                    int out_dtype);

/* b200_ep_combine fused with residual add + RMSNorm (arithmetic and arguments of b200_ep_allreduce_norm): the fp32
 * combine output never goes to HBM. */
int b200_ep_combine_norm(void* stream, void* const* peer_bufs, int32_t* const* peer_flags, int world, int rank,
                         const int32_t* ids_global, int m_local, int top_k, int hidden, int experts_per_rank,
                         void* residual, const void* gamma, float gain, float eps, void* out, int act_dtype);

/* per-kernel timing of the expert GEMMs with CUDA events on the launching stream (eager calls only;
 * ignored under stream capture).  profile(1) starts a window, profile_read returns the summed GEMM1 /
 * GEMM2 milliseconds and the number of forward calls in the window, then resets it. */
int b200moe_profile(int enable);
int b200moe_profile_read(double* gemm1_ms, double* gemm2_ms, int64_t* calls);

/* bring-up aid: copy `bytes` of an internal workspace buffer of the current device to host memory
 * (what: 0 tiled activations, 1 activation scales, 2 tiled intermediate, 3 intermediate scales,
 * 4 expert outputs y, 5 route state, 6 chunk table, 7 row_of_slot, 8 slot_of_row).  Synchronous. */
int b200moe_debug_read(int what, void* dst_host, int64_t bytes);

/* kernel-launch counter (number of this library's kernels launched by this process so far) */
int64_t b200moe_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* B200MOE_H_ */
