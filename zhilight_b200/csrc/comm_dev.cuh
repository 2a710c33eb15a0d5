// Device-side view of a zl_comm exchange object (comm.cu), shared with the W4A16 kernel whose prologue / epilogue take
// part in the tensor-parallel exchange directly (w4a16_gemm_v3.cu: partial sums are stored into the peers' inboxes from
// the GEMM epilogue and reduced in the next GEMM's activation staging).
#pragma once
#include "common.cuh"

namespace zl {

constexpr int kCommMaxRanks = 8;
constexpr int kCommMaxCtas = 64;         // <= #SMs so that every CTA of the stand-alone kernels is resident

struct CommDev {                        // lives in device memory (one copy per rank)
    uint8_t* inbox[kCommMaxRanks];      // inbox[r] = base of rank r's symmetric buffer as mapped HERE
    int rank, ws;
    size_t slot_bytes;                  // bytes of one (parity, source-rank) slot
    unsigned long long* epoch;          // local: number of completed exchanges
    unsigned int* done;                 // local: CTAs finished in the current exchange
    unsigned int* ll_step;              // local: decode steps that used the tagged exchange (bumped once per step)
};

// layout of a rank's symmetric buffer:
//   [2 parities][ws sources][slot_bytes payload]  then  flags [ws sources][kCommMaxCtas] (uint64, monotonic epochs)
__host__ __device__ inline size_t comm_flags_offset(int ws, size_t slot_bytes) { return 2 * (size_t)ws * slot_bytes; }
// then the tagged ("LL") region of the exchange that rides inside the GEMMs: [2 parities][ws sources][2 * slot_bytes] of
// 8-byte words {2 x fp16 payload, 32-bit tag}: a word is valid when its tag equals the tag of the exchange being consumed,
// so neither fences nor flags are needed (every 8-byte store is atomic)
__host__ __device__ inline size_t comm_ll_offset(int ws, size_t slot_bytes) {
    return comm_flags_offset(ws, slot_bytes) + (size_t)ws * kCommMaxCtas * sizeof(unsigned long long);
}
__host__ __device__ inline size_t comm_total_bytes(int ws, size_t slot_bytes) {
    return comm_ll_offset(ws, slot_bytes) + 2 * (size_t)ws * 2 * slot_bytes;
}
// tag of exchange `index` (< 512) of the current step; never 0 (the buffer is zero-initialised)
__device__ __forceinline__ uint32_t comm_ll_tag(uint32_t step, int index) { return (step << 9) + (uint32_t)index + 1u; }
__device__ __forceinline__ void st_ll(void* p, uint32_t data, uint32_t tag) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(data), "r"(tag) : "memory");
}
__device__ __forceinline__ uint4 ld_ll2(const void* p) {   // two words: {data0, tag0, data1, tag1}
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

}  // namespace zl

// host: device pointer to the CommDev of an opened exchange object and its slot size (comm.cu)
extern "C" const void* zl_comm_device_state(zl_comm_t* c);
extern "C" size_t zl_comm_slot_bytes(zl_comm_t* c);
