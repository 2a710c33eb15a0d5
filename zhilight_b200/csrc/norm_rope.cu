// RMSNorm (+residual), element add, gate-mul, RoPE tables, qkv split + RoPE (+ fused KV append).
//
// Reference behaviour restated: src/nn/layernorm/layernorm.cu:9-42, src/nn/block/block_kernel.cu:7-17,
// src/nn/linear/activation_kernel.cu:55-106, src/nn/position/rope_preparer.cu:49-161,
// src/nn/position/rope_common.cuh:3-34, src/nn/position/rotary_embedding_fuse_cache.cu:23-63,
// src/kvcache/ragged_buffer_kernel.cu:194-222.
#include "common.cuh"

namespace zl {

constexpr int kNormThreads = 256;
constexpr int kNormMaxChunks = 8;   // D <= 256*8*8 = 16384 kept in registers

__device__ __forceinline__ float block_sum(float v, float* sbuf) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) sbuf[warp] = v;
    __syncthreads();
    float r = (lane < (int)(blockDim.x >> 5)) ? sbuf[lane] : 0.f;
    r = warp_sum(r);
    return r;
}

// one CTA per token.  mode: -1 plain norm of a; 0 = round(a+b) then norm; 1 = norm of unrounded a+b
template <typename T>
__global__ void __launch_bounds__(kNormThreads)
k_add_rmsnorm(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ weight,
              T* __restrict__ out_sum, T* __restrict__ y, int D, float eps, float scale, int mode) {
    __shared__ float sbuf[32];
    const size_t off = (size_t)blockIdx.x * D;
    const int nchunk = D / 8;
    pdl_trigger();
    // weights are constants: fetch before waiting on the producer of a/b
    uint4 wreg[kNormMaxChunks];
#pragma unroll
    for (int c = 0; c < kNormMaxChunks; ++c) {
        const int ch = threadIdx.x + c * kNormThreads;
        if (ch < nchunk) wreg[c] = *reinterpret_cast<const uint4*>(weight + ch * 8);
    }
    pdl_wait();

    float v[kNormMaxChunks][8];
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < kNormMaxChunks; ++c) {
        const int ch = threadIdx.x + c * kNormThreads;
        if (ch < nchunk) {
            uint4 ua = ld_cg_u4(a + off + ch * 8);
            unpack8<T>(ua, v[c]);
            if (b) {
                uint4 ub = ld_cg_u4(b + off + ch * 8);
                float fb[8];
                unpack8<T>(ub, fb);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[c][i] += fb[i];
                uint4 us = pack8<T>(v[c]);
                if (out_sum) *reinterpret_cast<uint4*>(out_sum + off + ch * 8) = us;
                if (mode == 0) unpack8<T>(us, v[c]);   // normalise the T-rounded sum (block.cpp:124-131)
            } else if (out_sum && out_sum != a) {
                *reinterpret_cast<uint4*>(out_sum + off + ch * 8) = ua;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) ss += v[c][i] * v[c][i];
        }
    }
    ss = block_sum(ss, sbuf);
    const float r = rsqrtf(ss / (float)D + eps);
    if (y) {
#pragma unroll
        for (int c = 0; c < kNormMaxChunks; ++c) {
            const int ch = threadIdx.x + c * kNormThreads;
            if (ch < nchunk) {
                float fw[8], o[8];
                unpack8<T>(wreg[c], fw);
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = v[c][i] * r * fw[i] / scale;
                *reinterpret_cast<uint4*>(y + off + ch * 8) = pack8<T>(o);
            }
        }
    }
}

template <typename T>
__global__ void k_element_add_scale(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ c, size_t n,
                                    float scale) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // c = a + b * T(scale), every op rounded to T (block_kernel.cu:15)
    float bs = round_to<T>(to_f32<T>(b[i]) * round_to<T>(scale));
    c[i] = from_f32<T>(to_f32<T>(a[i]) + bs);
}

template <typename T>
__global__ void k_gate_mul(const T* __restrict__ gate, int ldg, const T* __restrict__ up, int ldu,
                           T* __restrict__ out, int ldo, int F, int act) {
    const int tok = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F) return;
    float gv = to_f32<T>(gate[(size_t)tok * ldg + i]);
    const float uv = up ? to_f32<T>(up[(size_t)tok * ldu + i]) : 1.f;   // up == NULL: plain activation
    float a = act == 0 ? silu_f(gv) : gelu_f(gv);
    out[(size_t)tok * ldo + i] = from_f32<T>(a * uv);
}

// grid T, block dim_head
__global__ void k_rope_cos_sin(const int32_t* __restrict__ pos, float* __restrict__ g_cos,
                               float* __restrict__ g_sin, float base, float factor, float low_f, float high_f,
                               float old_len, int neox) {
    const int m = pos[blockIdx.x];
    const int d = blockDim.x, half_dim = d / 2, col = threadIdx.x;
    const int i = neox ? (col < half_dim ? col : col - half_dim) : col / 2;
    float inv_freq = powf(base, -float(i * 2) / d);
    if (factor > 0.f) {   // llama3 wavelength-dependent scaling (rope_preparer.cu:124-161)
        const float low_wl = old_len / low_f;
        const float high_wl = old_len / high_f;
        const float pi = 3.141592653589793f;
        const float wavelen = 2.f * pi / inv_freq;
        if (wavelen < high_wl) {
        } else if (wavelen > low_wl) {
            inv_freq = inv_freq / factor;
        } else {
            const float smooth = (old_len / wavelen - low_f) / (high_f - low_f);
            inv_freq = (1.f - smooth) * inv_freq / factor + smooth * inv_freq;
        }
    }
    const float freq = m * inv_freq;
    const size_t o = (size_t)blockIdx.x * d + col;
    g_cos[o] = cosf(freq);
    g_sin[o] = sinf(freq);
}

template <typename T>
__device__ __forceinline__ float rope_value(const T* __restrict__ in, size_t offset, float c, float s, int col,
                                            int half_dim, int neox) {
    if (neox) {
        return col < half_dim ? to_f32<T>(in[offset]) * c - to_f32<T>(in[offset + half_dim]) * s
                              : to_f32<T>(in[offset]) * c + to_f32<T>(in[offset - half_dim]) * s;
    }
    return (col & 1) == 0 ? to_f32<T>(in[offset]) * c - to_f32<T>(in[offset + 1]) * s
                          : to_f32<T>(in[offset]) * c + to_f32<T>(in[offset - 1]) * s;
}

// grid (T, Hq + 2*Hkv), block dim_head.  k_addrs == nullptr -> plain rope_qk_cache (k, v dense outputs).
template <typename T>
__global__ void k_qkv_rope(const float* __restrict__ g_cos, const float* __restrict__ g_sin,
                           const T* __restrict__ in, T* __restrict__ q, T* __restrict__ k, T* __restrict__ v,
                           const int32_t* __restrict__ token_batch, const int32_t* __restrict__ placement,
                           T* const* __restrict__ k_addrs, T* const* __restrict__ v_addrs,
                           const int32_t* __restrict__ buf_lens, int num_heads, int num_kv_heads, int neox,
                           int bshd) {
    const int tok = blockIdx.x, head = blockIdx.y;
    const int d = blockDim.x, col = threadIdx.x, half_dim = d / 2;
    const int all_heads = num_heads + 2 * num_kv_heads;
    const size_t offset = ((size_t)tok * all_heads + head) * d + col;
    pdl_trigger();
    pdl_wait();

    const bool is_v = head >= num_heads + num_kv_heads;
    const bool is_k = !is_v && head >= num_heads;
    float val;
    if (is_v) {
        val = to_f32<T>(in[offset]);
    } else {
        const float c = g_cos[(size_t)tok * d + col];
        const float s = g_sin[(size_t)tok * d + col];
        val = rope_value<T>(in, offset, c, s, col, half_dim, neox);
    }
    const T out = is_v ? in[offset] : from_f32<T>(val);
    if (!is_k && !is_v) {
        q[((size_t)tok * num_heads + head) * d + col] = out;
        return;
    }
    const int hk = is_k ? head - num_heads : head - num_heads - num_kv_heads;
    if (k_addrs) {
        const int p = placement[tok];
        if (p < 0) return;
        const int b = token_batch[tok];
        T* dst = is_k ? k_addrs[b] : v_addrs[b];
        const size_t o = bshd ? ((size_t)p * num_kv_heads + hk) * d + col
                              : ((size_t)hk * buf_lens[b] + p) * d + col;
        dst[o] = out;
    } else {
        T* dst = is_k ? k : v;
        dst[((size_t)tok * num_kv_heads + hk) * d + col] = out;
    }
}

// grid (B, len_q, Hkv), block dim_head
template <typename T>
__global__ void k_copy_to_rag(const int32_t* __restrict__ placement, const int32_t* __restrict__ buf_lens,
                              const T* __restrict__ k_src, const T* __restrict__ v_src,
                              T* const* __restrict__ k_addrs, T* const* __restrict__ v_addrs, int bshd) {
    const int b = blockIdx.x, x = blockIdx.x * gridDim.y + blockIdx.y;
    const int p = placement[x];
    if (p < 0) return;
    const int num_heads = gridDim.z, head = blockIdx.z, d = blockDim.x;
    const size_t src = ((size_t)x * num_heads + head) * d + threadIdx.x;
    const size_t dst = bshd ? ((size_t)p * num_heads + head) * d + threadIdx.x
                            : ((size_t)head * buf_lens[b] + p) * d + threadIdx.x;
    k_addrs[b][dst] = k_src[src];
    v_addrs[b][dst] = v_src[src];
}

template <typename T>
static int run_add_rmsnorm(const void* a, const void* b, const void* w, void* out_sum, void* y, int T_, int D,
                           float eps, float scale, int mode, bool pdl, cudaStream_t stream) {
    ZL_CHECK_CUDA(launch(k_add_rmsnorm<T>, dim3(T_), dim3(kNormThreads), 0, stream, pdl, (const T*)a, (const T*)b,
                         (const T*)w, (T*)out_sum, (T*)y, D, eps, scale, mode));
    return ZL_OK;
}

}  // namespace zl

using namespace zl;

#define ZL_DISPATCH_T(dtype, ...)                          \
    if ((dtype) == ZL_F16) {                               \
        using scalar_t = __half;                           \
        __VA_ARGS__                                        \
    } else if ((dtype) == ZL_BF16) {                       \
        using scalar_t = __nv_bfloat16;                    \
        __VA_ARGS__                                        \
    } else {                                               \
        ZL_CHECK_SUPPORTED((dtype) == ZL_F16 || (dtype) == ZL_BF16); \
    }

extern "C" int zl_rmsnorm(const void* x, const void* weight, void* y, int T, int D, float eps, float scale,
                          int dtype, int pdl, zl_stream_t stream) {
    ZL_CHECK_ARG(x && weight && y && T > 0 && D > 0);
    ZL_CHECK_SUPPORTED(D % 8 == 0 && D <= kNormThreads * 8 * kNormMaxChunks);
    ZL_DISPATCH_T(dtype, { return run_add_rmsnorm<scalar_t>(x, nullptr, weight, nullptr, y, T, D, eps, scale, -1, pdl != 0, stream); })
    return ZL_OK;
}

extern "C" int zl_add_rmsnorm(const void* a, const void* b, const void* weight, void* out_sum, void* y, int T,
                              int D, float eps, float scale, int mode, int dtype, int pdl, zl_stream_t stream) {
    ZL_CHECK_ARG(a && weight && T > 0 && D > 0 && (mode == 0 || mode == 1));
    ZL_CHECK_ARG(y || out_sum);
    ZL_CHECK_SUPPORTED(D % 8 == 0 && D <= kNormThreads * 8 * kNormMaxChunks);
    ZL_DISPATCH_T(dtype, { return run_add_rmsnorm<scalar_t>(a, b, weight, out_sum, y, T, D, eps, scale, mode, pdl != 0, stream); })
    return ZL_OK;
}

extern "C" int zl_element_add_scale(const void* a, const void* b, void* c, size_t n, float scale, int dtype,
                                    zl_stream_t stream) {
    ZL_CHECK_ARG(a && b && c && n > 0);
    const unsigned blocks = (unsigned)((n + 255) / 256);
    ZL_DISPATCH_T(dtype, {
        k_element_add_scale<scalar_t><<<blocks, 256, 0, stream>>>((const scalar_t*)a, (const scalar_t*)b,
                                                                  (scalar_t*)c, n, scale);
    })
    ZL_CHECK_LAUNCH();
    return ZL_OK;
}

extern "C" int zl_gate_mul(const void* gate, int ld_gate, const void* up, int ld_up, void* out, int ld_out, int T,
                           int F, int act, int dtype, zl_stream_t stream) {
    ZL_CHECK_ARG(gate && out && T > 0 && F > 0 && (act == 0 || act == 1));
    dim3 grid(cdiv(F, 256), T);
    ZL_DISPATCH_T(dtype, {
        k_gate_mul<scalar_t><<<grid, 256, 0, stream>>>((const scalar_t*)gate, ld_gate, (const scalar_t*)up, ld_up,
                                                       (scalar_t*)out, ld_out, F, act);
    })
    ZL_CHECK_LAUNCH();
    return ZL_OK;
}

extern "C" int zl_rope_cos_sin(const int32_t* pos, float* cos_out, float* sin_out, int T, int dim_head,
                               float theta, float llama3_factor, float low_freq_factor, float high_freq_factor,
                               float old_context_len, int neox, zl_stream_t stream) {
    ZL_CHECK_ARG(pos && cos_out && sin_out && T > 0 && dim_head > 0 && dim_head <= 1024 && dim_head % 2 == 0);
    k_rope_cos_sin<<<T, dim_head, 0, stream>>>(pos, cos_out, sin_out, theta, llama3_factor, low_freq_factor,
                                               high_freq_factor, old_context_len, neox);
    ZL_CHECK_LAUNCH();
    return ZL_OK;
}

extern "C" int zl_rope_qk_cache(const float* cos, const float* sin, const void* qkv, void* q, void* k, void* v,
                                int T, int num_heads, int num_kv_heads, int dim_head, int neox, int dtype,
                                zl_stream_t stream) {
    ZL_CHECK_ARG(cos && sin && qkv && q && k && v && T > 0 && num_heads > 0 && num_kv_heads > 0);
    ZL_CHECK_ARG(dim_head > 0 && dim_head <= 1024 && dim_head % 2 == 0);
    dim3 grid(T, num_heads + 2 * num_kv_heads);
    ZL_DISPATCH_T(dtype, {
        k_qkv_rope<scalar_t><<<grid, dim_head, 0, stream>>>(cos, sin, (const scalar_t*)qkv, (scalar_t*)q,
                                                            (scalar_t*)k, (scalar_t*)v, nullptr, nullptr, nullptr,
                                                            nullptr, nullptr, num_heads, num_kv_heads, neox, 1);
    })
    ZL_CHECK_LAUNCH();
    return ZL_OK;
}

extern "C" int zl_copy_to_rag_buffer2(const int32_t* placement, const int32_t* buf_lens, const void* k_src,
                                      const void* v_src, void* const* k_addrs, void* const* v_addrs, int B,
                                      int len_q, int num_kv_heads, int dim_head, int bshd, int dtype,
                                      zl_stream_t stream) {
    ZL_CHECK_ARG(placement && buf_lens && k_src && v_src && k_addrs && v_addrs);
    ZL_CHECK_ARG(B > 0 && len_q > 0 && num_kv_heads > 0 && dim_head > 0 && dim_head <= 1024);
    dim3 grid(B, len_q, num_kv_heads);
    ZL_DISPATCH_T(dtype, {
        k_copy_to_rag<scalar_t><<<grid, dim_head, 0, stream>>>(placement, buf_lens, (const scalar_t*)k_src,
                                                               (const scalar_t*)v_src, (scalar_t* const*)k_addrs,
                                                               (scalar_t* const*)v_addrs, bshd);
    })
    ZL_CHECK_LAUNCH();
    return ZL_OK;
}

extern "C" int zl_qkv_rope_append(const float* cos, const float* sin, const void* qkv, void* q_out,
                                  const int32_t* token_batch, const int32_t* placement, void* const* k_addrs,
                                  void* const* v_addrs, int T, int num_heads, int num_kv_heads, int dim_head,
                                  int neox, int bshd, const int32_t* buf_lens, int dtype, int pdl,
                                  zl_stream_t stream) {
    ZL_CHECK_ARG(cos && sin && qkv && q_out && token_batch && placement && k_addrs && v_addrs);
    ZL_CHECK_ARG(T > 0 && num_heads > 0 && num_kv_heads > 0 && dim_head > 0 && dim_head <= 1024);
    ZL_CHECK_ARG(bshd || buf_lens);
    dim3 grid(T, num_heads + 2 * num_kv_heads);
    ZL_DISPATCH_T(dtype, {
        ZL_CHECK_CUDA(launch(k_qkv_rope<scalar_t>, grid, dim3(dim_head), 0, stream, pdl != 0, cos, sin,
                             (const scalar_t*)qkv, (scalar_t*)q_out, (scalar_t*)nullptr, (scalar_t*)nullptr,
                             token_batch, placement, (scalar_t* const*)k_addrs, (scalar_t* const*)v_addrs,
                             buf_lens, num_heads, num_kv_heads, neox, bshd));
    })
    return ZL_OK;
}
