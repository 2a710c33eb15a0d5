// TEST INFRASTRUCTURE (oracle/_ref), round-2 additions to the shim over the REFERENCE's own kernels: RopePreparer,
// Marlin (repack + gemm), the native AWQ GEMM, the int8-KV cache path, and a timed chain of the reference's default decode
// kernels for one layer.  Only linked into oracle/_ref/libzl_ref.so (the drop-in library has no counterpart for these).
#include <bmengine/core/core.h>
#include <bmengine/functions/gemm.h>
#include <bmengine/functions/index_select.h>
#include <bmengine/functions/init.h>

#include "nn/quant/gptq/gptq.h"
#include "nn/quant/int8/quant_kernel.h"
#include "nn/attention/attention_kernel.h"
#include "nn/layernorm/layernorm.h"
#include "nn/position/rotary_embedding.h"
#include "nn/position/rope_preparer.h"
#include "nn/block/block_kernel.h"
#include "nn/linear/activation_kernel.h"
#include "kvcache/ragged_buffer_kernel.h"
#include "nn/quant/marlin/marlin.h"
#include "nn/quant/awq/awq.h"

#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

using namespace bmengine;
using core::DataType;
using core::Tensor;

// state owned by ref_shim.cu
core::Context* zlref_ctx();
char* zlref_err_buf();

namespace {
#define g_ctx (zlref_ctx())
Tensor wrap(std::vector<size_t> shape, DataType dt, const void* p) {
    size_t n = 1;
    for (auto s : shape) n *= s;
    return Tensor::from_external(shape, dt, const_cast<void*>(p), n * core::get_elem_size(dt), 0, false);
}
DataType dt_of(int code) { return code == 1 ? DataType::kBFloat16 : (code == 2 ? DataType::kFloat : DataType::kHalf); }
void sync() { BM_CUDART_ASSERT(cudaStreamSynchronize(g_ctx->current_stream()->ptr)); }
void copy_out(const Tensor& t, void* dst) {
    BM_CUDART_ASSERT(cudaMemcpyAsync(dst, t.data(), t.nbytes(), cudaMemcpyDeviceToDevice, g_ctx->current_stream()->ptr));
}
}  // namespace

#define ZLREF_TRY(...)                                                  \
    try {                                                               \
        __VA_ARGS__;                                                    \
        sync();                                                         \
        return 0;                                                       \
    } catch (const std::exception& e) {                                 \
        snprintf(zlref_err_buf(), 1024, "%s", e.what());                \
        return -1;                                                      \
    }

extern "C" {

// ---- Marlin (QuantType::GPTQ_Marlin = 8): the load steps of GPTQMarlin::post_load (linear.cpp:1402-1428:
// gptq_marlin_repack + the 64-column scale permutation perm[i*8+j] = i + 8j) and gptq_marlin_gemm as
// GPTQMarlin::forward calls it (linear.cpp:1300-1322: no act-order, has_zp = false, fp32 reduce, zeroed workspace) ----
namespace {
struct MarlinW {
    Tensor qweight, scales, qzeros, workspace;
};
MarlinW marlin_prepare(const void* qweight_hf, const void* scales_hf, int N, int K, int G) {
    MarlinW w;
    Tensor qw = wrap({(size_t)K / 8, (size_t)N}, DataType::kInt32, qweight_hf);
    Tensor perm;
    w.qweight = gptq_marlin_repack(*g_ctx, qw, perm, K, N, 4);
    std::vector<int> idx;
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) idx.push_back(i + 8 * j);
    Tensor idx_d = g_ctx->tensor_of(idx);
    Tensor s = wrap({(size_t)G, (size_t)N}, DataType::kHalf, scales_hf);
    Tensor s1 = s.view({s.numel() / idx.size(), idx.size()});
    w.scales = functions::index_select(*g_ctx, s1, 1, idx_d).view({(size_t)G, (size_t)N});
    w.qzeros = g_ctx->tensor({(size_t)G, (size_t)N / 8}, DataType::kInt32);
    w.workspace = g_ctx->tensor({1024 * 1024}, DataType::kFloat);
    functions::zeros_(*g_ctx, w.workspace);
    return w;
}
}  // namespace
int zlref_marlin_repack(const void* qweight_hf, int N, int K, void* out /* (K/16, 2N) int32 */) {
    ZLREF_TRY(Tensor qw = wrap({(size_t)K / 8, (size_t)N}, DataType::kInt32, qweight_hf); Tensor perm;
              Tensor r = gptq_marlin_repack(*g_ctx, qw, perm, K, N, 4); copy_out(r, out))
}
int zlref_marlin_gemm(const void* a, const void* qweight_hf, const void* scales_hf, int M, int N, int K, int G, void* out) {
    ZLREF_TRY(MarlinW w = marlin_prepare(qweight_hf, scales_hf, N, K, G);
              Tensor A = wrap({(size_t)M, (size_t)K}, DataType::kHalf, a); Tensor g_idx; Tensor perm;
              Tensor r = gptq_marlin_gemm(*g_ctx, A, w.qweight, w.scales, w.qzeros, g_idx, perm, w.workspace, M, N, K, true,
                                          false, true);
              copy_out(r, out))
}
int zlref_time_marlin(const void* a, const void* const* qweight_hf_list, const void* scales_hf, int n_rot, int M, int N, int K,
                      int G, int iters, float* us) {
    try {
        std::vector<MarlinW> ws;
        for (int i = 0; i < n_rot; ++i) ws.push_back(marlin_prepare(qweight_hf_list[i], scales_hf, N, K, G));
        Tensor A = wrap({(size_t)M, (size_t)K}, DataType::kHalf, a);
        Tensor g_idx, perm;
        auto st = g_ctx->current_stream()->ptr;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        for (int i = -3; i < iters; ++i) {
            if (i == 0) cudaEventRecord(e0, st);
            MarlinW& w = ws[(i + 3) % n_rot];
            Tensor r = gptq_marlin_gemm(*g_ctx, A, w.qweight, w.scales, w.qzeros, g_idx, perm, w.workspace, M, N, K, true, false, true);
        }
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        *us = ms * 1e3f / iters;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        return 0;
    } catch (const std::exception& e) {
        snprintf(zlref_err_buf(), 1024, "%s", e.what());
        return -1;
    }
}

// ---- native AWQ (QuantType::AWQ = 6, M < 256): awq_gemm with split_k_iters = 32 (linear.cpp:1560-1565) ----
int zlref_awq_gemm(const void* a, const void* qweight /* (K, N/8) */, const void* scales /* (K/g, N) */,
                   const void* qzeros /* (K/g, N/8) */, int M, int N, int K, int G, void* out) {
    ZLREF_TRY(Tensor A = wrap({(size_t)M, (size_t)K}, DataType::kHalf, a);
              Tensor w = wrap({(size_t)K, (size_t)N / 8}, DataType::kInt32, qweight);
              Tensor s = wrap({(size_t)G, (size_t)N}, DataType::kHalf, scales);
              Tensor z = wrap({(size_t)G, (size_t)N / 8}, DataType::kInt32, qzeros);
              Tensor r = nn::awq::awq_gemm(*g_ctx, A, w, s, z, 32); copy_out(r, out))
}

// ---- the reference's default decode chain for ONE Llama-class GPTQ layer at len_q = 1 (SURVEY.md 9.8 row 2:
// no qkv / ff fusion): RMSNorm, 3 GEMVs (q, k, v written into one fused buffer), rope_qk_cache, copy_to_rag_buffer2,
// multi_query_attention_rag_buffer, o GEMV, element_add_scale, RMSNorm, 2 GEMVs (gate, up), gate_mul_inplace (SiLU), down
// GEMV, element_add_scale.  (The default build rotates q and k with two KERNEL_rope_with_cache launches instead of the
// single fused rope_qk_cache used here -- one launch in the reference's favour.)  All buffers are allocated here and
// filled with constants; `n_rot` weight sets rotate so that no layer is served from L2.  Returns microseconds per layer.
namespace {
__global__ void k_fill_u32(uint32_t* p, size_t n, uint32_t v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
void fill32(Tensor& t, uint32_t v) {
    size_t n = t.nbytes() / 4;
    k_fill_u32<<<(unsigned)((n + 255) / 256), 256, 0, g_ctx->current_stream()->ptr>>>(reinterpret_cast<uint32_t*>(t.data()), n, v);
}
struct RefLin {
    Tensor qw, qz, sc;
    int N, K;
};
RefLin make_lin(int N, int K) {
    RefLin l;
    l.N = N;
    l.K = K;
    l.qw = g_ctx->tensor({(size_t)N, (size_t)K / 8}, DataType::kInt32);
    l.qw.set_name("zlref.linear");
    l.qz = g_ctx->tensor({(size_t)N, (size_t)K / 128}, DataType::kInt8);
    l.sc = g_ctx->tensor({(size_t)N, (size_t)K / 128}, DataType::kHalf);
    fill32(l.qw, 0x9a3c5e71u);
    fill32(l.qz, 0x08080808u);
    fill32(l.sc, 0x1c001c00u);   // 0.00391 fp16
    return l;
}
void gemv(const Tensor& x, RefLin& l, Tensor& out) {
    nn::gptq::gptq_gemm_k_major(*g_ctx, x, l.qw, l.qz, l.sc, Tensor(), Tensor(), nullptr, true, false, &out);
}
}  // namespace
int zlref_time_decode_layer(int D, int hq, int hkv, int d, int ff, int ctx_len, int n_rot, int iters, int with_lm_head_vocab,
                            float* us_per_layer, float* us_lm_head) {
    try {
        const size_t B = 1;
        struct LayerW {
            RefLin q, k, v, o, gate, up, down;
        };
        std::vector<LayerW> L;
        for (int i = 0; i < n_rot; ++i)
            L.push_back(LayerW{make_lin(hq * d, D), make_lin(hkv * d, D), make_lin(hkv * d, D), make_lin(D, hq * d),
                               make_lin(ff, D), make_lin(ff, D), make_lin(D, ff)});
        nn::LayerNorm ln(*g_ctx, D, false, 1e-5f, 1.0f, DataType::kHalf);
        Tensor lnw = g_ctx->tensor({(size_t)D}, DataType::kHalf);
        fill32(lnw, 0x3c003c00u);
        std::map<std::string, const Tensor> sd;
        sd.emplace("ln.weight", lnw);
        ln.load_state_dict(*g_ctx, sd, "ln", false);
        Tensor h = g_ctx->tensor({B, (size_t)D}, DataType::kHalf);
        fill32(h, 0x2e662e66u);   // 0.1
        Tensor qkv = g_ctx->tensor({B, (size_t)(hq + 2 * hkv) * d}, DataType::kHalf);
        Tensor ao = g_ctx->tensor({B, 1, (size_t)hq, (size_t)d}, DataType::kHalf);
        Tensor attn_out = g_ctx->tensor({B, (size_t)D}, DataType::kHalf);
        Tensor g1 = g_ctx->tensor({B, (size_t)ff}, DataType::kHalf), g2 = g_ctx->tensor({B, (size_t)ff}, DataType::kHalf);
        Tensor ffo = g_ctx->tensor({B, (size_t)D}, DataType::kHalf);
        Tensor cos = g_ctx->tensor({B, (size_t)d}, DataType::kFloat), sin = g_ctx->tensor({B, (size_t)d}, DataType::kFloat);
        fill32(cos, 0x3f800000u);
        fill32(sin, 0u);
        const size_t len_buf = (size_t)((ctx_len + 2 + 63) / 64 * 64);   // batch_generator.cpp:62-64 rounding
        Tensor kbuf = g_ctx->tensor({len_buf, (size_t)hkv, (size_t)d}, DataType::kHalf);
        Tensor vbuf = g_ctx->tensor({len_buf, (size_t)hkv, (size_t)d}, DataType::kHalf);
        fill32(kbuf, 0x2e662e66u);
        fill32(vbuf, 0x2e662e66u);
        std::vector<double> ka(1), va(1);
        void* kp = kbuf.data();
        void* vp = vbuf.data();
        memcpy(&ka[0], &kp, 8);
        memcpy(&va[0], &vp, 8);
        Tensor KA = g_ctx->tensor({B}, DataType::kDouble), VA = g_ctx->tensor({B}, DataType::kDouble);
        BM_CUDART_ASSERT(cudaMemcpy(KA.data(), ka.data(), 8, cudaMemcpyHostToDevice));
        BM_CUDART_ASSERT(cudaMemcpy(VA.data(), va.data(), 8, cudaMemcpyHostToDevice));
        std::vector<int> lens = {(int)len_buf}, place = {ctx_len};
        Tensor LEN = g_ctx->tensor_of(lens), PLACE = g_ctx->tensor_of(place).view({B, 1});
        std::vector<int8_t> mask_h(len_buf, 0);
        for (int i = 0; i <= ctx_len; ++i) mask_h[i] = 1;
        Tensor MASK = g_ctx->tensor({len_buf}, DataType::kInt8);
        BM_CUDART_ASSERT(cudaMemcpy(MASK.data(), mask_h.data(), len_buf, cudaMemcpyHostToDevice));
        const float scale = 1.0f / sqrtf((float)d);
        auto st = g_ctx->current_stream()->ptr;
        auto one_layer = [&](LayerW& w) {
            Tensor xn = ln.forward(*g_ctx, h);
            Tensor oq = qkv.slice_dim0_len(0, 1).view({B, (size_t)(hq + 2 * hkv) * d});
            // q, k, v GEMVs write disjoint column ranges of the fused (1, (hq + 2 hkv) d) row
            Tensor q_out = Tensor::from_external({B, (size_t)hq * d}, DataType::kHalf, qkv.data<char>(), (size_t)hq * d * 2, 0, false);
            Tensor k_out = Tensor::from_external({B, (size_t)hkv * d}, DataType::kHalf, qkv.data<char>() + (size_t)hq * d * 2,
                                                 (size_t)hkv * d * 2, 0, false);
            Tensor v_out = Tensor::from_external({B, (size_t)hkv * d}, DataType::kHalf,
                                                 qkv.data<char>() + (size_t)(hq + hkv) * d * 2, (size_t)hkv * d * 2, 0, false);
            gemv(xn, w.q, q_out);
            gemv(xn, w.k, k_out);
            gemv(xn, w.v, v_out);
            Tensor rq, rk, rv;
            nn::rope_qk_cache(*g_ctx, cos, sin, qkv, rq, rk, rv, hq, hkv, d, DataType::kHalf, true);
            Tensor K4 = rk.view({B, 1, (size_t)hkv, (size_t)d}), V4 = rv.view({B, 1, (size_t)hkv, (size_t)d});
            nn::copy_to_rag_buffer2(*g_ctx, PLACE, LEN, K4, V4, &KA, &VA, false);
            Tensor Q4 = rq.view({B, 1, (size_t)hq, (size_t)d});
            nn::multi_query_attention_rag_buffer(*g_ctx, Q4, LEN, KA, VA, MASK, scale, (int)len_buf, ao, hq / hkv, -1);
            Tensor ao2 = ao.view({B, (size_t)hq * d});
            gemv(ao2, w.o, attn_out);
            nn::element_add_scale_out(*g_ctx, h, attn_out, h, 1.0f, false);
            Tensor xn2 = ln.forward(*g_ctx, h);
            gemv(xn2, w.gate, g1);
            gemv(xn2, w.up, g2);
            nn::gate_mul_inplace(*g_ctx, g1, g2, "silu");
            gemv(g1, w.down, ffo);
            nn::element_add_scale_out(*g_ctx, h, ffo, h, 1.0f, false);
        };
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        for (int i = -2 * n_rot; i < iters; ++i) {
            if (i == 0) cudaEventRecord(e0, st);
            one_layer(L[(i + 2 * n_rot) % n_rot]);
        }
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        *us_per_layer = ms * 1e3f / iters;
        if (us_lm_head) *us_lm_head = 0.f;
        if (with_lm_head_vocab > 0 && us_lm_head) {
            // lm_head at decode: fp16 GEMM through bmengine's cuBLASLt wrapper (embedding.cu:353-392)
            Tensor W = g_ctx->tensor({(size_t)with_lm_head_vocab, (size_t)D}, DataType::kHalf);
            fill32(W, 0x1c001c00u);
            functions::Gemm gemm(*g_ctx, DataType::kHalf, false, true);
            for (int i = -2; i < 10; ++i) {
                if (i == 0) cudaEventRecord(e0, st);
                Tensor logits = gemm.forward(*g_ctx, h, W);
            }
            cudaEventRecord(e1, st);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            *us_lm_head = ms * 1e3f / 10;
        }
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        return 0;
    } catch (const std::exception& e) {
        snprintf(zlref_err_buf(), 1024, "%s", e.what());
        return -1;
    }
}

}  // extern "C"
