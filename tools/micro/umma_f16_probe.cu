// Bring-up probe for the conventions k_w4a16_tc (zhilight_b200/csrc/w4a16_tc.cu) relies on, one tile at a time:
//   D(128 x 32, f32, TMEM) = A(128 x 64 fp16, K-major, SWIZZLE_128B, written by threads) . B(32 x 64 fp16)^T
// B is staged either by threads (same manual swizzle) or by the TMA engine from a (rows_valid x 64) tensor whose box
// (64 x 32) is TALLER than the tensor (out-of-bounds rows must arrive as zeros).  Variants: LBO field 0 / 1.
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o umma_f16_probe umma_f16_probe.cu && ./umma_f16_probe
// Not part of libzhilight_b200.so.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int M = 128, N = 32, K = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t lbo) {
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)lbo << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t sw128_off(int row, int k) {   // byte offset of element (row, k), k < 64
    return (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u + (uint32_t)((((k >> 3) ^ (row & 7)) << 4) + (k & 7) * 2);
}

struct alignas(64) Params {
    CUtensorMap bmap;
    const __half* a;
    const __half* b;
    float* d;
    int lbo, b_via_tma;
};

__global__ void __launch_bounds__(128) k_probe(const __grid_constant__ Params p) {
    extern __shared__ uint8_t dyn[];
    uint8_t* smem = dyn + ((1024u - (smem_u32(dyn) & 1023u)) & 1023u);
    uint8_t* sa = smem;                 // 128 x 128 B
    uint8_t* sb = smem + 16384;         // 32 x 128 B
    uint64_t* bar_tma = reinterpret_cast<uint64_t*>(smem + 16384 + 4096);
    uint64_t* bar_mma = bar_tma + 1;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_mma + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < M * K; i += 128) *reinterpret_cast<__half*>(sa + sw128_off(i / K, i % K)) = p.a[i];
    if (!p.b_via_tma)
        for (int i = tid; i < N * K; i += 128) *reinterpret_cast<__half*>(sb + sw128_off(i / K, i % K)) = p.b[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar_tma)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar_mma)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(s_tmem)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = *s_tmem;
    if (tid == 0) {
        if (p.b_via_tma) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar_tma)), "r"(N * 128)
                         : "memory");
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
                    "r"(smem_u32(sb)),
                "l"(reinterpret_cast<uint64_t>(&p.bmap)), "r"(0), "r"(0), "r"(smem_u32(bar_tma))
                : "memory");
            uint32_t done = 0;
            while (!done)
                asm volatile(
                    "{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                    : "=r"(done)
                    : "r"(smem_u32(bar_tma))
                    : "memory");
        }
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        for (int k16 = 0; k16 < 4; ++k16) {
            const uint64_t ad = desc_sw128(smem_u32(sa), p.lbo) + (uint64_t)(k16 * 2);
            const uint64_t bd = desc_sw128(smem_u32(sb), p.lbo) + (uint64_t)(k16 * 2);
            asm volatile(
                "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n\t}" ::
                    "r"(tmem),
                "l"(ad), "l"(bd), "r"(idesc), "r"(k16 > 0 ? 1u : 0u)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar_mma))
                     : "memory");
    }
    {
        uint32_t done = 0;
        while (!done)
            asm volatile(
                "{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                : "=r"(done)
                : "r"(smem_u32(bar_mma))
                : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    uint32_t r[16];
    for (int c0 = 0; c0 < N; c0 += 16) {
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int c = 0; c < 16; ++c) p.d[(warp * 32 + lane) * N + c0 + c] = __uint_as_float(r[c]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    std::vector<__half> a(M * K), b(N * K);
    srand(11);
    for (auto& v : a) v = __float2half((float)(rand() % 31 - 15) * 0.125f);
    for (auto& v : b) v = __float2half((float)(rand() % 201 - 100) * 0.01f);
    const int rows_valid = 20;   // TMA variant: the tensor has 20 rows, the box 32
    __half *da, *db;
    float* dd;
    cudaMalloc(&da, a.size() * 2);
    cudaMalloc(&db, b.size() * 2);
    cudaMalloc(&dd, M * N * 4);
    cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess || !sym) {
        printf("cuTensorMapEncodeTiled not found\n");
        return 1;
    }
    Params p;
    const cuuint64_t gdim[2] = {K, (cuuint64_t)rows_valid};
    const cuuint64_t gstride[1] = {K * 2};
    const cuuint32_t box[2] = {64, N};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = reinterpret_cast<EncodeTiledFn>(sym)(&p.bmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, db, gdim, gstride, box, estr,
                                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("cuTensorMapEncodeTiled(box rows 32 > tensor rows %d) -> %d\n", rows_valid, (int)cr);
    p.a = da;
    p.b = db;
    p.d = dd;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    std::vector<float> out(M * N);
    for (int var = 0; var < 4; ++var) {
        p.lbo = var & 1;
        p.b_via_tma = var >> 1;
        if (p.b_via_tma && cr != CUDA_SUCCESS) continue;
        cudaMemset(dd, 0xFF, M * N * 4);
        k_probe<<<1, 128, 32768>>>(p);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("variant %d: CUDA error %s\n", var, cudaGetErrorString(e));
            return 1;
        }
        cudaMemcpy(out.data(), dd, M * N * 4, cudaMemcpyDeviceToHost);
        double worst = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                double ref = 0;
                if (!(p.b_via_tma && n >= rows_valid))
                    for (int k = 0; k < K; ++k) ref += (double)__half2float(a[m * K + k]) * (double)__half2float(b[n * K + k]);
                worst = fmax(worst, fabs(ref - (double)out[m * N + n]));
            }
        printf("variant %d (LBO field %d, B by %s): max |diff| = %.3g %s\n", var, p.lbo, p.b_via_tma ? "TMA (box taller than tensor)" : "threads",
               worst, worst < 1e-3 ? "OK" : "MISMATCH");
    }
    return 0;
}
