// Micro-benchmark: issue rate / latency of legacy mma.sync HMMA.16816 and IMMA.16832 on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int KIND, int CHAINS>
__global__ void k(int iters, int* out, long long* cyc) {
    uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    int d[CHAINS][4];
    float f[CHAINS][4];
    for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) { d[c][i] = 0; f[c][i] = 0.f; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(f[c][0]), "+f"(f[c][1]), "+f"(f[c][2]), "+f"(f[c][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(d[c][0]), "+r"(d[c][1]), "+r"(d[c][2]), "+r"(d[c][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    long long t1 = clock64();
    int s = 0;
    for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) s += d[c][i] + (int)f[c][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int KIND, int CHAINS>
void run(const char* name, int warps) {
    int* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    k<KIND, CHAINS><<<148, warps * 32>>>(iters, out, cyc);
    cudaDeviceSynchronize();
    k<KIND, CHAINS><<<148, warps * 32>>>(iters, out, cyc);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)h / (iters * CHAINS);
    printf("%s chains=%d warps/SM=%d: %.2f cycles per MMA per warp -> %.3f MMA/cycle/SM\n", name, CHAINS, warps, per, warps / per);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0, 1>("HMMA.16816", 1); run<0, 8>("HMMA.16816", 1); run<0, 4>("HMMA.16816", 4); run<0, 4>("HMMA.16816", 16);
    run<1, 1>("IMMA.16832", 1); run<1, 8>("IMMA.16832", 1); run<1, 4>("IMMA.16832", 4); run<1, 4>("IMMA.16832", 16);
    return 0;
}
