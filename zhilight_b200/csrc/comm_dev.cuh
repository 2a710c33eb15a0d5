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
};

// layout of a rank's symmetric buffer:
//   [2 parities][ws sources][slot_bytes payload]  then  flags [ws sources][kCommMaxCtas] (uint64, monotonic epochs)
__host__ __device__ inline size_t comm_flags_offset(int ws, size_t slot_bytes) { return 2 * (size_t)ws * slot_bytes; }
__host__ __device__ inline size_t comm_total_bytes(int ws, size_t slot_bytes) {
    return comm_flags_offset(ws, slot_bytes) + (size_t)ws * kCommMaxCtas * sizeof(unsigned long long);
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
