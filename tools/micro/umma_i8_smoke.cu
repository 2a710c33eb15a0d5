// Bring-up probe for the round-2 tcgen05 path: ONE `tcgen05.mma.cta_group::1.kind::i8`
// tile D(128 x 16, s32 in TMEM) = A(128 x 32, u8, shared memory, K-major, no swizzle) . B(16 x 32, s8)^T, read back
// with `tcgen05.ld` and compared with the CPU.  It exists to pin, on real hardware, the three things the guide
// (/opt/skills/guides/blackwell_cuda_programming.md) leaves to CUTLASS: the shared-memory matrix-descriptor
// conventions (which of LBO / SBO walks K and which walks M/N for a K-major, SWIZZLE_NONE operand), the kind::i8
// instruction descriptor (u8 x s8 -> s32), and the TMEM lane/column addressing of tcgen05.ld.32x32b.
//
// NOT part of libzhilight_b200.so and NOT run in round 1 (it was written after the GPU budget was spent; it only
// passed `nvcc -arch=sm_100a` / ptxas here).  Build and run:
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o umma_i8_smoke umma_i8_smoke.cu && ./umma_i8_smoke
// Expected output: exactly one of the four descriptor conventions reports "max |diff| = 0".
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int M = 128, N = 16, K = 32;   // one UMMA: M=128 rows (TMEM lanes), N=16 columns, K=32 bytes

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 64-bit shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp, union SmemDescriptor): start address, leading
// byte offset and stride byte offset in 16-byte units, version = 1 (sm_100), layout type SWIZZLE_NONE = 0.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // version_
    return d;
}

// 32-bit instruction descriptor (union InstrDescriptor): c_format S32 = 2 (bits 4-5), a_format (bits 7-9) and b_format
// (bits 10-12): 0 = unsigned 8 bit, 1 = signed 8 bit, K-major A and B (bits 15, 16 = 0), N >> 3 (bits 17-22),
// M >> 4 (bits 24-28).
__host__ __device__ constexpr uint32_t make_idesc(int a_signed, int b_signed) {
    return (2u << 4) | ((uint32_t)a_signed << 7) | ((uint32_t)b_signed << 10) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

// Operand tiles in shared memory: 8 x 16-byte "core matrices" stored contiguously (128 B each), ordered
// [k chunk of 16 bytes][group of 8 rows][row in group][16 bytes].
__device__ __forceinline__ int core_offset(int row, int kbyte, int rows_total) {
    return (kbyte / 16) * (rows_total / 8) * 128 + (row / 8) * 128 + (row % 8) * 16 + (kbyte % 16);
}

__global__ void __launch_bounds__(128) k_umma_i8_smoke(const uint8_t* __restrict__ a, const int8_t* __restrict__ b,
                                                       int32_t* __restrict__ d, int convention) {
    __shared__ __align__(128) uint8_t sa[M * K];
    __shared__ __align__(128) uint8_t sb[N * K];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < M * K; i += 128) sa[core_offset(i / K, i % K, M)] = a[i];
    for (int i = tid; i < N * K; i += 128) sb[core_offset(i / K, i % K, N)] = (uint8_t)b[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // make the generic-proxy writes of the operand tiles visible to the tensor-core (async) proxy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    if (warp == 0) {   // one warp allocates 32 TMEM columns (>= N) and publishes the base address
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base;

    if (tid == 0) {
        // convention bit 0: which of the two offsets is "leading" (walks K) -- bit 1: A unsigned (0) or signed (1) flag order probe
        const uint32_t a_k = (M / 8) * 128, a_mn = 128, b_k = (N / 8) * 128, b_mn = 128;
        const bool lbo_is_k = (convention & 1) == 0;
        const uint64_t da = lbo_is_k ? make_desc(smem_u32(sa), a_k, a_mn) : make_desc(smem_u32(sa), a_mn, a_k);
        const uint64_t db = lbo_is_k ? make_desc(smem_u32(sb), b_k, b_mn) : make_desc(smem_u32(sb), b_mn, b_k);
        const uint32_t idesc = (convention & 2) ? make_idesc(1, 0) : make_idesc(0, 1);   // default: A u8, B s8
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(0u)   // accumulate = 0: D is overwritten
            : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar))
                     : "memory");
    }
    // everyone waits for the MMA to retire (phase 0 of the mbarrier)
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(smem_u32(&mbar))
                : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;");

    // warp w reads TMEM lanes 32w .. 32w+31 (= rows of D), 16 consecutive 32-bit columns
    uint32_t r[16];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    const int row = warp * 32 + lane;
#pragma unroll
    for (int c = 0; c < N; ++c) d[row * N + c] = (int32_t)r[c];

    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem));
}

int main() {
    std::vector<uint8_t> a(M * K);
    std::vector<int8_t> b(N * K);
    srand(7);
    for (auto& v : a) v = (uint8_t)(rand() % 16);          // what an unpacked int4 weight looks like
    for (auto& v : b) v = (int8_t)(rand() % 255 - 127);    // an activation digit
    std::vector<int32_t> ref(M * N, 0), out(M * N);
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < K; ++k) ref[m * N + n] += (int32_t)a[m * K + k] * (int32_t)b[n * K + k];
    uint8_t* da;
    int8_t* db;
    int32_t* dd;
    cudaMalloc(&da, a.size());
    cudaMalloc(&db, b.size());
    cudaMalloc(&dd, out.size() * 4);
    cudaMemcpy(da, a.data(), a.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size(), cudaMemcpyHostToDevice);
    int ok = 0;
    for (int conv = 0; conv < 4; ++conv) {
        cudaMemset(dd, 0xFF, out.size() * 4);
        k_umma_i8_smoke<<<1, 128>>>(da, db, dd, conv);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("convention %d: CUDA error %s\n", conv, cudaGetErrorString(e));
            return 1;
        }
        cudaMemcpy(out.data(), dd, out.size() * 4, cudaMemcpyDeviceToHost);
        long long worst = 0;
        for (int i = 0; i < M * N; ++i) {
            const long long df = llabs((long long)out[i] - ref[i]);
            worst = df > worst ? df : worst;
        }
        printf("convention %d (%s, %s): max |diff| = %lld\n", conv, (conv & 1) ? "LBO walks M/N" : "LBO walks K",
               (conv & 2) ? "A s8 x B u8" : "A u8 x B s8", worst);
        ok += worst == 0;
    }
    printf("%s\n", ok == 1 ? "exactly one convention matches: use it" : "unexpected: inspect the descriptors");
    return 0;
}
