// Parameter block of the persistent whole-model decode kernel (llama_mega.cu).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace zl {

// one decoder layer: ZLW4I weight blobs in phase order {qkv, attn_out, gate_up, down}
struct MegaLayer {
    const uint8_t* packed[4];
    const __half* bias[4];      // packed-row order, may be null
    const __half* ln_attn;
    const __half* ln_ff;
    __half* const* k_addrs;     // per-task KV buffers of this layer (device pointer tables)
    __half* const* v_addrs;
};

struct MegaParams {
    const MegaLayer* layers;    // device array [num_layers]
    int num_layers;
    int mc;                     // tokens in the step (<= 8)
    int gN[4], gK[4], gTall[4]; // GEMM shapes in phase order; gTall: k split over 16 warps (else 2 x 8)
    int num_heads, num_kv_heads, dim_head;
    float eps, attn_scale;
    __half* h;                  // (mc, D) residual stream, in/out
    __half* q;                  // (mc, Hq*d)
    __half* ao;                 // (mc, Hq*d)
    __half* act;                // (mc, ff)
    const float* cos;
    const float* sin;
    const int32_t* token_batch;
    const int32_t* placement;
    const int32_t* buf_lens;
    int attn_splits;            // ceil(len bucket / 256)
    float* part_o;
    float* part_m;
    float* part_l;
    unsigned* sync;             // [0] grid barrier counter (zeroed before launch), [1] abort flag
    unsigned long long* trace;  // debug: [0] = number of records, then (id, globaltimer) pairs written by CTA 0
};

size_t mega_smem_bytes(int mc, int k_max, int dim_head, int stages);
// stages: 4 or 3.  Returns cudaErrorInvalidValue when the shape does not fit.
cudaError_t launch_llama_mega(const MegaParams& p, int stages, bool pdl, cudaStream_t stream);
cudaError_t prepare_llama_mega();

}  // namespace zl
