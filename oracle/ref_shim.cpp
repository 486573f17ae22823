// C-ABI shim around the REFERENCE's own CPU kernels for this path — fused MoE (csrc/cpu/cpu_fused_moe.cpp: prepack_moe_weight
// :640-661, cpu_fused_moe :663-702), paged MLA decode (mla_decode.cpp), RMSNorm / residual-add RMSNorm (layernorm.cpp), rotary
// embedding (pos_encoding.cpp), silu_and_mul (activation.cpp) — compiled from the sources where they lie under /root/reference by
// oracle/build_ref.py into oracle/_ref/libref_moe.so.  TEST INFRASTRUCTURE / CPU BASELINE ONLY: it validates the
// restated oracle (bf16 experts) against a compiled reference and is what `bench.py --impl reference` times for
// bf16 workloads (cpu_baseline.kind = "reference").  Nothing here is imported by the product.
#include <torch/torch.h>

#include <optional>
#include <string>

// declarations of the two reference entry points (defined in the reference translation unit)
void prepack_moe_weight(const torch::Tensor& weight, torch::Tensor& packed_weight, const std::string& isa);
void cpu_fused_moe(torch::Tensor& output, const torch::Tensor& input, const torch::Tensor& w13, const torch::Tensor& w2,
                   const std::optional<torch::Tensor>& w13_bias, const std::optional<torch::Tensor>& w2_bias,
                   const torch::Tensor& topk_weights, const torch::Tensor& topk_id, const bool skip_weighted,
                   const std::string& act, const std::string& isa);

// reference csrc/cpu/mla_decode.cpp:356-383 (paged MLA decode on the latent cache, block_size 16)
void mla_decode_kvcache(torch::Tensor& out, torch::Tensor& query, torch::Tensor& kv_cache, double scale,
                        torch::Tensor& block_tables, torch::Tensor& seq_lens);

// reference csrc/cpu/layernorm.cpp:99-136 (RMSNorm and residual-add + RMSNorm, in place on input / residual)
void rms_norm(torch::Tensor& out, torch::Tensor& input, std::optional<torch::Tensor> weight, double epsilon);
void fused_add_rms_norm(torch::Tensor& input, torch::Tensor& residual, std::optional<torch::Tensor> weight, double epsilon);
// reference csrc/cpu/pos_encoding.cpp:332-366 (rotary embedding in place on query / key)
void rotary_embedding(torch::Tensor& positions, torch::Tensor& query, std::optional<torch::Tensor> key, int64_t head_size,
                      torch::Tensor& cos_sin_cache, bool is_neox, int64_t rope_dim_offset, bool inverse);
// reference csrc/cpu/activation.cpp:87-97
void silu_and_mul(torch::Tensor& out, torch::Tensor& input);

namespace {
struct RefMoe {
  torch::Tensor w13p, w2p;
  int E, H, I;
  std::string isa;
};
std::string g_err;
}  // namespace

extern "C" {

const char* ref_moe_last_error() { return g_err.c_str(); }

// w13 bf16 [E, 2I, H], w2 bf16 [E, H, I] (checkpoint layout); both are re-packed by the reference's own prepack
void* ref_moe_create(const void* w13, const void* w2, int E, int H, int I, const char* isa) {
  try {
    auto opt = torch::TensorOptions().dtype(torch::kBFloat16);
    torch::Tensor a = torch::from_blob(const_cast<void*>(w13), {E, 2 * I, H}, opt);
    torch::Tensor b = torch::from_blob(const_cast<void*>(w2), {E, H, I}, opt);
    auto* h = new RefMoe();
    h->E = E;
    h->H = H;
    h->I = I;
    h->isa = isa;
    h->w13p = torch::empty_like(a);
    h->w2p = torch::empty_like(b);
    prepack_moe_weight(a, h->w13p, h->isa);
    prepack_moe_weight(b, h->w2p, h->isa);
    return h;
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}

// hidden bf16 [M, H], ids i32 [M, k], weights f32 [M, k] -> out bf16 [M, H]
int ref_moe_forward(void* handle, const void* hidden, const int32_t* ids, const float* weights, void* out, int M, int k,
                    const char* act) {
  try {
    auto* h = static_cast<RefMoe*>(handle);
    torch::Tensor x = torch::from_blob(const_cast<void*>(hidden), {M, h->H}, torch::TensorOptions().dtype(torch::kBFloat16));
    torch::Tensor o = torch::from_blob(out, {M, h->H}, torch::TensorOptions().dtype(torch::kBFloat16));
    torch::Tensor ti = torch::from_blob(const_cast<int32_t*>(ids), {M, k}, torch::TensorOptions().dtype(torch::kInt32));
    torch::Tensor tw = torch::from_blob(const_cast<float*>(weights), {M, k}, torch::TensorOptions().dtype(torch::kFloat32));
    cpu_fused_moe(o, x, h->w13p, h->w2p, std::nullopt, std::nullopt, tw, ti, false, act, h->isa);
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

void ref_moe_destroy(void* handle) { delete static_cast<RefMoe*>(handle); }

// q bf16 [B, Hq, 576] (nope 512 | rope 64), kv_cache bf16 [num_blocks, 16, 576], block_tables i32 [B, max_blocks],
// seq_lens i32 [B] -> out bf16 [B, Hq, 512]
int ref_mla_decode(void* out, const void* q, const void* kv_cache, double scale, const int32_t* block_tables,
                   const int32_t* seq_lens, int B, int Hq, int num_blocks, int max_blocks) {
  try {
    auto bf = torch::TensorOptions().dtype(torch::kBFloat16);
    auto i32 = torch::TensorOptions().dtype(torch::kInt32);
    torch::Tensor o = torch::from_blob(out, {B, Hq, 512}, bf);
    torch::Tensor qq = torch::from_blob(const_cast<void*>(q), {B, Hq, 576}, bf);
    torch::Tensor kv = torch::from_blob(const_cast<void*>(kv_cache), {num_blocks, 16, 576}, bf);
    torch::Tensor bt = torch::from_blob(const_cast<int32_t*>(block_tables), {B, max_blocks}, i32);
    torch::Tensor sl = torch::from_blob(const_cast<int32_t*>(seq_lens), {B}, i32);
    mla_decode_kvcache(o, qq, kv, scale, bt, sl);
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// ---- the reference's CPU RMSNorm kernels (csrc/cpu/layernorm.cpp).  dtype: 0 bf16, 1 fp16, 2 fp32; weight may be null.
static torch::ScalarType ref_dtype(int dtype) {
  return dtype == 0 ? torch::kBFloat16 : dtype == 1 ? torch::kFloat16 : torch::kFloat32;
}

// out[M, H] = rms_norm(x[M, H]) (* weight[H])
int ref_rms_norm(void* out, const void* x, const void* weight, double eps, int M, int H, int dtype) {
  try {
    auto opt = torch::TensorOptions().dtype(ref_dtype(dtype));
    torch::Tensor o = torch::from_blob(out, {M, H}, opt);
    torch::Tensor xi = torch::from_blob(const_cast<void*>(x), {M, H}, opt);
    std::optional<torch::Tensor> w;
    if (weight) w = torch::from_blob(const_cast<void*>(weight), {H}, opt);
    rms_norm(o, xi, w, eps);
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// in place: residual <- x + residual ; x <- rms_norm(x + residual) (* weight)
int ref_fused_add_rms_norm(void* x, void* residual, const void* weight, double eps, int M, int H, int dtype) {
  try {
    auto opt = torch::TensorOptions().dtype(ref_dtype(dtype));
    torch::Tensor xi = torch::from_blob(x, {M, H}, opt);
    torch::Tensor r = torch::from_blob(residual, {M, H}, opt);
    std::optional<torch::Tensor> w;
    if (weight) w = torch::from_blob(const_cast<void*>(weight), {H}, opt);
    fused_add_rms_norm(xi, r, w, eps);
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// in place on query [T, Hq * head_size] and key [T, Hk * head_size] (key may be null); positions i64 [T];
// cos_sin_cache [max_pos, rot_dim] in the activation dtype (cos half | sin half)
int ref_rotary_embedding(const int64_t* positions, void* query, void* key, int T, int Hq, int Hk, int head_size,
                         const void* cos_sin_cache, int max_pos, int rot_dim, int is_neox, int dtype) {
  try {
    auto opt = torch::TensorOptions().dtype(ref_dtype(dtype));
    torch::Tensor pos = torch::from_blob(const_cast<int64_t*>(positions), {T}, torch::TensorOptions().dtype(torch::kInt64));
    torch::Tensor q = torch::from_blob(query, {T, (int64_t)Hq * head_size}, opt);
    std::optional<torch::Tensor> k;
    if (key) k = torch::from_blob(key, {T, (int64_t)Hk * head_size}, opt);
    torch::Tensor cs = torch::from_blob(const_cast<void*>(cos_sin_cache), {max_pos, rot_dim}, opt);
    rotary_embedding(pos, q, k, head_size, cs, is_neox != 0, 0, false);
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// out[M, d] = silu(x[M, :d]) * x[M, d:]
int ref_silu_and_mul(void* out, const void* x, int M, int d, int dtype) {
  try {
    auto opt = torch::TensorOptions().dtype(ref_dtype(dtype));
    torch::Tensor o = torch::from_blob(out, {M, d}, opt);
    torch::Tensor xi = torch::from_blob(const_cast<void*>(x), {M, 2 * d}, opt);
    silu_and_mul(o, xi);
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

}  // extern "C"
