// Library plumbing (errors, launch counter) and the small kernels around the decode step:
// embedding gather (src/nn/embedding/embedding.cu:20-60 semantics), greedy argmax over logits
// (generator/beam_util.cu pick_top_k, beam 1), synthetic-weight generators for the bench.
#include "common.cuh"

#include <atomic>
#include <cstdio>
#include <cstring>

static thread_local char g_last_error[512] = "";
static std::atomic<long long> g_launches{0};

extern "C" void zl_set_last_error(const char* file, int line, const char* msg) {
    const char* base = strrchr(file, '/');
    snprintf(g_last_error, sizeof(g_last_error), "%s:%d: %s", base ? base + 1 : file, line, msg);
}
extern "C" const char* zl_last_error(void) { return g_last_error; }
extern "C" int zl_version(void) { return 100; }
extern "C" void zl_count_launch(void) { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" long long zl_launch_count(int reset) {
    long long v = g_launches.load();
    if (reset) g_launches.store(0);
    return v;
}

namespace zl {

// out[t, :] = table[ids[t], :]   (16-byte chunks)
template <typename T>
__global__ void k_embedding(const int32_t* __restrict__ ids, const T* __restrict__ table, T* __restrict__ out,
                            int D, int vocab) {
    pdl_trigger();
    pdl_wait();
    int id = ids[blockIdx.x];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    const uint4* src = reinterpret_cast<const uint4*>(table + (size_t)id * D);
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)blockIdx.x * D);
    for (int i = threadIdx.x; i < D / 8; i += blockDim.x) dst[i] = src[i];
}

// stage 1: per (token, chunk) partial argmax; stage 2: merge.  Ties -> lowest index.
__global__ void k_argmax_partial(const float* __restrict__ logits, int V, int chunks, float* __restrict__ pval,
                                 int* __restrict__ pidx) {
    pdl_trigger();
    pdl_wait();
    const int tok = blockIdx.y, ch = blockIdx.x;
    const int per = (V + chunks - 1) / chunks;
    const int lo = ch * per, hi = min(V, lo + per);
    const float* row = logits + (size_t)tok * V;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const float v = row[i];
        if (v > best || (v == best && i < bi)) {
            best = v;
            bi = i;
        }
    }
    __shared__ float sv[256];
    __shared__ int si[256];
    sv[threadIdx.x] = best;
    si[threadIdx.x] = bi;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            const float ov = sv[threadIdx.x + s];
            const int oi = si[threadIdx.x + s];
            if (ov > sv[threadIdx.x] || (ov == sv[threadIdx.x] && oi < si[threadIdx.x])) {
                sv[threadIdx.x] = ov;
                si[threadIdx.x] = oi;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        pval[tok * chunks + ch] = sv[0];
        pidx[tok * chunks + ch] = si[0];
    }
}

__global__ void k_argmax_final(const float* __restrict__ pval, const int* __restrict__ pidx, int chunks,
                               int32_t* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    const int tok = blockIdx.x;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < chunks; i += 32) {
        const float v = pval[tok * chunks + i];
        const int id = pidx[tok * chunks + i];
        if (v > best || (v == best && id < bi)) {
            best = v;
            bi = id;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) {
            best = ov;
            bi = oi;
        }
    }
    if (threadIdx.x == 0) out[tok] = bi;
}

// vocab-parallel greedy pick: per-token (value, global index) candidate of this rank's logits shard ...
__global__ void k_argmax_cand(const float* __restrict__ pval, const int* __restrict__ pidx, int chunks, int idx_offset,
                              int2* __restrict__ cand) {
    pdl_trigger();
    pdl_wait();
    const int tok = blockIdx.x;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < chunks; i += 32) {
        const float v = pval[tok * chunks + i];
        const int id = pidx[tok * chunks + i];
        if (v > best || (v == best && id < bi)) {
            best = v;
            bi = id;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) {
            best = ov;
            bi = oi;
        }
    }
    if (threadIdx.x == 0) cand[tok] = make_int2(__float_as_int(best), bi + idx_offset);
}
// ... and the merge over the gathered candidates [ranks][stride] (ties -> lowest index, like the single-rank pick)
__global__ void k_argmax_merge(const int2* __restrict__ cand_all, int ranks, int stride, int T, int32_t* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    const int tok = blockIdx.x * blockDim.x + threadIdx.x;
    if (tok >= T) return;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int r = 0; r < ranks; ++r) {
        const int2 c = cand_all[(size_t)r * stride + tok];
        const float v = __int_as_float(c.x);
        if (v > best || (v == best && c.y < bi)) {
            best = v;
            bi = c.y;
        }
    }
    out[tok] = bi;
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__global__ void k_fill_u32(uint32_t* __restrict__ p, size_t n, uint64_t seed) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)splitmix64(seed * 0x100000001B3ull + i);
}
__global__ void k_fill_const_u32(uint32_t* __restrict__ p, size_t n, uint32_t v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
template <typename T>
__global__ void k_fill_uniform(T* __restrict__ p, size_t n, float lo, float hi, uint64_t seed) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint32_t r = (uint32_t)(splitmix64(seed * 0x100000001B3ull + i) >> 40);   // 24 bits
        p[i] = from_f32<T>(lo + (hi - lo) * (r * (1.0f / 16777216.0f)));
    }
}

}  // namespace zl

using namespace zl;

extern "C" int zl_embedding(const int32_t* ids, const void* table, void* out, int T, int D, int vocab, int dtype,
                            int pdl, zl_stream_t stream) {
    ZL_CHECK_ARG(ids && table && out && T > 0 && D > 0 && D % 8 == 0 && vocab > 0);
    if (dtype == ZL_F16)
        ZL_CHECK_CUDA(launch(k_embedding<__half>, dim3(T), dim3(256), 0, stream, pdl != 0, ids, (const __half*)table,
                             (__half*)out, D, vocab));
    else if (dtype == ZL_BF16)
        ZL_CHECK_CUDA(launch(k_embedding<__nv_bfloat16>, dim3(T), dim3(256), 0, stream, pdl != 0, ids,
                             (const __nv_bfloat16*)table, (__nv_bfloat16*)out, D, vocab));
    else
        ZL_CHECK_SUPPORTED(dtype == ZL_F16 || dtype == ZL_BF16);
    return ZL_OK;
}

extern "C" size_t zl_argmax_workspace_bytes(int T) { return (size_t)T * 64 * (sizeof(float) + sizeof(int)); }

extern "C" int zl_argmax(const float* logits, int32_t* out, int T, int V, void* workspace, size_t workspace_bytes,
                         int pdl, zl_stream_t stream) {
    ZL_CHECK_ARG(logits && out && T > 0 && V > 0 && workspace);
    const int chunks = 64;
    ZL_CHECK_ARG(workspace_bytes >= zl_argmax_workspace_bytes(T));
    float* pval = static_cast<float*>(workspace);
    int* pidx = reinterpret_cast<int*>(pval + (size_t)T * chunks);
    ZL_CHECK_CUDA(launch(k_argmax_partial, dim3(chunks, T), dim3(256), 0, stream, pdl != 0, logits, V, chunks, pval,
                         pidx));
    ZL_CHECK_CUDA(launch(k_argmax_final, dim3(T), dim3(32), 0, stream, pdl != 0, (const float*)pval,
                         (const int*)pidx, chunks, out));
    return ZL_OK;
}

extern "C" int zl_argmax_candidates(const float* logits, void* cand_out, int T, int V, int idx_offset, void* workspace,
                                    size_t workspace_bytes, int pdl, zl_stream_t stream) {
    ZL_CHECK_ARG(logits && cand_out && T > 0 && V > 0 && workspace);
    const int chunks = 64;
    ZL_CHECK_ARG(workspace_bytes >= zl_argmax_workspace_bytes(T));
    float* pval = static_cast<float*>(workspace);
    int* pidx = reinterpret_cast<int*>(pval + (size_t)T * chunks);
    ZL_CHECK_CUDA(launch(k_argmax_partial, dim3(chunks, T), dim3(256), 0, stream, pdl != 0, logits, V, chunks, pval,
                         pidx));
    ZL_CHECK_CUDA(launch(k_argmax_cand, dim3(T), dim3(32), 0, stream, pdl != 0, (const float*)pval, (const int*)pidx,
                         chunks, idx_offset, (int2*)cand_out));
    return ZL_OK;
}

extern "C" int zl_argmax_merge(const void* cand_all, int32_t* out, int T, int ranks, int stride, int pdl,
                               zl_stream_t stream) {
    ZL_CHECK_ARG(cand_all && out && T > 0 && ranks > 0 && stride >= T);
    ZL_CHECK_CUDA(launch(k_argmax_merge, dim3(cdiv(T, 64)), dim3(64), 0, stream, pdl != 0, (const int2*)cand_all, ranks,
                         stride, T, out));
    return ZL_OK;
}

extern "C" int zl_fill_random_u32(uint32_t* p, size_t n, uint64_t seed, zl_stream_t stream) {
    ZL_CHECK_ARG(p && n > 0);
    k_fill_u32<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(p, n, seed);
    ZL_CHECK_LAUNCH();
    return ZL_OK;
}
extern "C" int zl_fill_const_u32(uint32_t* p, size_t n, uint32_t value, zl_stream_t stream) {
    ZL_CHECK_ARG(p && n > 0);
    k_fill_const_u32<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(p, n, value);
    ZL_CHECK_LAUNCH();
    return ZL_OK;
}
extern "C" int zl_fill_uniform(void* p, size_t n, float lo, float hi, uint64_t seed, int dtype,
                               zl_stream_t stream) {
    ZL_CHECK_ARG(p && n > 0);
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (dtype == ZL_F16)
        k_fill_uniform<__half><<<blocks, 256, 0, stream>>>((__half*)p, n, lo, hi, seed);
    else if (dtype == ZL_BF16)
        k_fill_uniform<__nv_bfloat16><<<blocks, 256, 0, stream>>>((__nv_bfloat16*)p, n, lo, hi, seed);
    else if (dtype == ZL_F32)
        k_fill_uniform<float><<<blocks, 256, 0, stream>>>((float*)p, n, lo, hi, seed);
    else
        ZL_CHECK_SUPPORTED(dtype >= 0 && dtype <= 2);
    ZL_CHECK_LAUNCH();
    return ZL_OK;
}
