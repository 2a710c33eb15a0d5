// W4A16 GEMM on the 5th-generation tensor cores: tcgen05.mma (kind::f16) with TMEM accumulators, tensor-TMA for the
// activations, bulk-TMA for the int4 weight blocks.  Serves every M the exact-integer mma.sync kernel
// (w4a16_gemm_v3.cu) cannot stage: batch 17..256 decode and chunked prefill.
//
// Replaces the reference's Marlin kernel (src/nn/quant/marlin/gptq_marlin.cu:425-1620), the M <= 40 GEMV tiles of
// KERNEL_gemm_warp_reduce (src/nn/quant/gptq/q_gemm_k_major.cu:176-237, 580-686: the weight matrix is streamed
// ceil(M/16) times) and the M > 40 route "dequantise the whole matrix to fp16 in HBM, then cuBLASLt"
// (q_gemm_k_major.cu:843-905, 1083-1100).  Same arithmetic contract as that route: w = (q - z) * s rounded to fp16
// once, fp16 x fp16 products, fp32 accumulation over all of K.
//
// One CTA per SM, persistent over work items (128 weight rows x a k-slice); 15 warps in five roles:
//   warp 0      raw producer : cp.async.bulk (UBLKCP) of 4 ZLW4I blocks (32 rows x 128 k, 2128 B each) per stage
//   warp 1      x producer   : cp.async.bulk.tensor.2d (UTMALDG), box 64 k x NTOK tokens, SWIZZLE_128B, OOB rows = 0
//   warp 2      MMA issuer   : one thread, 8 x tcgen05.mma 128 x NTOK x 16 per 128-k stage, tcgen05.commit frees stages
//   warps 3-10  dequant      : nibbles -> fp16 (PRMT 0x64xx / 0x54xx magic, exact q - z, one HMUL2 by the group scale)
//                              written as the K-major SWIZZLE_128B A operand, fence.proxy.async, mbarrier arrive
//   warps 11-14 epilogue     : tcgen05.ld 32x32b of their TMEM lane quadrant, bias / residual / SwiGLU / qkv-RoPE-KV
// The accumulator is double-buffered in TMEM (2 x NTOK columns), so the drain of item i overlaps the mainloop of
// item i+1.  GEMMs with fewer than #SM row tiles split K across CTAs; the partial sums meet in an fp32 workspace and
// the last CTA of a tile reduces them in split order (deterministic) before the epilogue.
#include "common.cuh"
#include "w4_layout.cuh"
#include "w4_params.h"
#include "tc_common.cuh"

#include <cstdlib>

namespace zl {

constexpr int kTcRawStage = 4 * kW4BlockBytes;      // 4 blocks of 32 rows
constexpr int kTcAStage = 2 * kTcRows * 128;        // [k atom (64 k)][row][128 B]
constexpr int kTcWarpRaw = 0, kTcWarpX = 1, kTcWarpMma = 2, kTcWarpDq0 = 3, kTcDqWarps = 8, kTcWarpEpi0 = 11;
constexpr int kTcThreads = 15 * 32;
constexpr int kTcDqThreads = kTcDqWarps * 32;

template <int NTOK>
struct TcCfg {
    static constexpr int AS = NTOK <= 32 ? 4 : (NTOK <= 64 ? 3 : 2);   // A / x stages
    static constexpr int RS = NTOK <= 128 ? 4 : 3;                     // raw weight stages
    static constexpr int kXStage = 2 * NTOK * 128;                     // [k atom][token][128 B]
    static constexpr int kAOff = 0;
    static constexpr int kXOff = kAOff + AS * kTcAStage;
    static constexpr int kRawOff = kXOff + AS * kXStage;
    static constexpr int kBarOff = (kRawOff + RS * kTcRawStage + 15) & ~15;
    static constexpr int kNumBars = 2 * RS + 3 * AS + 4;
    static constexpr int kMiscOff = kBarOff + kNumBars * 8;
    static constexpr int kBytes = kMiscOff + 16 + 1024;                // + slack for the manual 1024-byte alignment
    static constexpr int kTmemCols = 2 * NTOK < 32 ? 32 : 2 * NTOK;    // power of two for NTOK in {16,32,64,128,256}
};

struct alignas(64) W4TcParams {
    CUtensorMap xmap;           // x (M, K) fp16 row-major, box {64, NTOK}, 128-byte swizzle
    CUtensorMap wmap;           // TS kernel: ZLW4I nibble words as [N/32][G][32 rows][64 B], box {64 B, 32, 1, 4}, 64-byte swizzle
    CUtensorMap tmap;           // TS kernel: the 80-byte scale / zero trailers as [N/32][G][80 B], box {80 B, 1, 4}
    const uint8_t* packed;      // ZLW4I
    const __half* bias;         // indexed by PACKED row
    const __half* residual;
    __half* y;
    int M, N, K, epi, S;        // S = k splits
    int dbg;                    // ZL_TC_DBG timing ablations of the TS kernel (results are wrong): 1 no x loads, 2 no dequant, 4 no MMA, 8 no weight loads
    float* ws;                  // [tile][split][M][128] fp32 partial sums (S > 1)
    unsigned* counters;         // [tile] arrivals (S > 1), left at zero
    unsigned* err;              // watchdog code
    const float* cos;
    const float* sin;
    __half* q_out;
    const int32_t* token_batch;
    const int32_t* placement;
    __half* const* k_addrs;
    __half* const* v_addrs;
    int num_heads, num_kv_heads, dim_head;
};

// instruction descriptor, kind::f16: D f32 (bits 4-5 = 1), A/B f16 (0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t tc_idesc_f16(int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTcRows >> 4) << 24);
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}

// One 32-bit ZLW4I word = 4 consecutive k of row g (low nibbles) and of row g + 8 (high nibbles) -> 2 x 2 half2.
// 0x64xx is 1024 + x for x < 1024; 0x54xx is 64 + x / 16: the high nibble never has to be shifted.
__device__ __forceinline__ void tc_dequant_word(uint32_t w, __half2 c_lo, __half2 s_lo, __half2 c_hi, __half2 s_hi,
                                                uint32_t (&lo)[2], uint32_t (&hi)[2]) {
    const uint32_t wl = w & 0x0f0f0f0fu, wh = w & 0xf0f0f0f0u;
    uint32_t a0 = __byte_perm(wl, 0x64646464u, 0x4140), a1 = __byte_perm(wl, 0x64646464u, 0x4342);
    uint32_t b0 = __byte_perm(wh, 0x54545454u, 0x4140), b1 = __byte_perm(wh, 0x54545454u, 0x4342);
    __half2 r;
    r = __hmul2(__hsub2(*reinterpret_cast<__half2*>(&a0), c_lo), s_lo);
    lo[0] = *reinterpret_cast<uint32_t*>(&r);
    r = __hmul2(__hsub2(*reinterpret_cast<__half2*>(&a1), c_lo), s_lo);
    lo[1] = *reinterpret_cast<uint32_t*>(&r);
    r = __hmul2(__hsub2(*reinterpret_cast<__half2*>(&b0), c_hi), s_hi);
    hi[0] = *reinterpret_cast<uint32_t*>(&r);
    r = __hmul2(__hsub2(*reinterpret_cast<__half2*>(&b1), c_hi), s_hi);
    hi[1] = *reinterpret_cast<uint32_t*>(&r);
}

// ---- epilogue of 16 tokens (c0 .. c0+15) for the weight row owned by this lane -----------------------------------------
// Packed row p = tile * 128 + 32 * quadrant + lane.  Inside a 32-row block, rows (16 tt + g) and (16 tt + g + 8) are
// partners (gate / up, RoPE low / high half): lane and lane + 8.
__device__ __forceinline__ void tc_epilogue16(const W4TcParams& p, int prow, int lane, int c0, float (&v)[16]) {
    const float b = p.bias ? __half2float(p.bias[prow]) : 0.f;
    if (p.epi == ZL_EPI_SWIGLU) {
        const int n_out = p.N / 2;
        const int col = (prow >> 5) * 16 + ((lane >> 4) << 3) + (lane & 7);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float mine = __half2float(__float2half_rn(v[i] + b));
            const float up = __shfl_down_sync(0xffffffffu, mine, 8);
            const int tok = c0 + i;
            if (!(lane & 8) && tok < p.M) p.y[(size_t)tok * n_out + col] = __float2half_rn(silu_f(mine) * up);
        }
    } else if (p.epi == ZL_EPI_QKV_ROPE) {
        const int d = p.dim_head, half_dim = d / 2, tiles_per_head = d / 32;
        const int st = prow >> 5, head = st / tiles_per_head, jt = st % tiles_per_head;
        const int c = jt * 16 + ((lane >> 4) << 3) + (lane & 7);
        const bool is_v = head >= p.num_heads + p.num_kv_heads;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float lo = __half2float(__float2half_rn(v[i] + b));
            const float hi = __shfl_down_sync(0xffffffffu, lo, 8);
            const int tok = c0 + i;
            if ((lane & 8) || tok >= p.M) continue;
            __half olo, ohi;
            if (is_v) {
                olo = __float2half_rn(lo);
                ohi = __float2half_rn(hi);
            } else {
                const float* cs = p.cos + (size_t)tok * d;
                const float* sn = p.sin + (size_t)tok * d;
                olo = __float2half_rn(lo * cs[c] - hi * sn[c]);
                ohi = __float2half_rn(hi * cs[c + half_dim] + lo * sn[c + half_dim]);
            }
            if (head < p.num_heads) {
                __half* dst = p.q_out + ((size_t)tok * p.num_heads + head) * d;
                dst[c] = olo;
                dst[c + half_dim] = ohi;
            } else {
                const int pl = p.placement[tok];
                if (pl >= 0) {
                    const bool is_k = !is_v;
                    const int hk = is_k ? head - p.num_heads : head - p.num_heads - p.num_kv_heads;
                    __half* base = (is_k ? p.k_addrs : p.v_addrs)[p.token_batch[tok]];
                    __half* dst = base + ((size_t)pl * p.num_kv_heads + hk) * d;
                    dst[c] = olo;
                    dst[c + half_dim] = ohi;
                }
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int tok = c0 + i;
            if (tok >= p.M) continue;
            __half h = __float2half_rn(v[i] + b);
            if (p.epi == ZL_EPI_RESIDUAL)
                h = __float2half_rn(__half2float(h) + __half2float(p.residual[(size_t)tok * p.N + prow]));
            p.y[(size_t)tok * p.N + prow] = h;
        }
    }
}

template <int NTOK>
__global__ void __launch_bounds__(kTcThreads, 1) k_w4a16_tc(const __grid_constant__ W4TcParams p) {
    using C = TcCfg<NTOK>;
    extern __shared__ uint8_t smem_dyn[];
    uint8_t* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kBarOff);
    uint64_t* raw_full = bars;
    uint64_t* raw_empty = raw_full + C::RS;
    uint64_t* a_full = raw_empty + C::RS;
    uint64_t* x_full = a_full + C::AS;
    uint64_t* ax_empty = x_full + C::AS;
    uint64_t* acc_full = ax_empty + C::AS;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + C::kMiscOff);
    uint32_t* s_last = s_tmem + 1;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = p.K / kW4GroupK, n_tiles = p.N / kTcRows, S = p.S;
    const int n_items = n_tiles * S;

    if (threadIdx.x == 0) {
        for (int i = 0; i < C::RS; ++i) {
            mbar_init(&raw_full[i], 1);
            mbar_init(&raw_empty[i], kTcDqWarps / 2);
        }
        for (int i = 0; i < C::AS; ++i) {
            mbar_init(&a_full[i], kTcDqWarps / 2);
            mbar_init(&x_full[i], 1);
            mbar_init(&ax_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 4);
        }
        mbar_fence_init();
    }
    if (warp == kTcWarpMma) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                     "n"(C::kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp == kTcWarpRaw) {
        // ---------------- weight stream: constants, may run ahead of the predecessor kernel (PDL) ----------------
        if (lane == 0) {
            const uint64_t pol = l2_evict_first_policy();
            int rs = 0;
            uint32_t ph = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
                const int tile = it / S, split = it % S;
                const int g0 = (split * G) / S, g1 = ((split + 1) * G) / S;
                for (int gi = g0; gi < g1; ++gi) {
                    mbar_wait_wd(&raw_empty[rs], ph ^ 1u, p.err, 0x100 + rs);
                    mbar_expect_tx(&raw_full[rs], kTcRawStage);
                    uint8_t* dst = smem + C::kRawOff + rs * kTcRawStage;
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        bulk_g2s_hint(dst + b * kW4BlockBytes,
                                      p.packed + ((size_t)(tile * 4 + b) * G + gi) * kW4BlockBytes, kW4BlockBytes,
                                      &raw_full[rs], pol);
                    if (++rs == C::RS) {
                        rs = 0;
                        ph ^= 1u;
                    }
                }
            }
        }
    } else if (warp == kTcWarpX) {
        // ---------------- activations: produced by the predecessor kernel ----------------
        if (lane == 0) {
            pdl_wait();
            int as = 0;
            uint32_t ph = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
                const int split = it % S;
                const int g0 = (split * G) / S, g1 = ((split + 1) * G) / S;
                for (int gi = g0; gi < g1; ++gi) {
                    mbar_wait_wd(&ax_empty[as], ph ^ 1u, p.err, 0x200 + as);
                    mbar_expect_tx(&x_full[as], C::kXStage);
                    uint8_t* dst = smem + C::kXOff + as * C::kXStage;
                    tma_load_2d(dst, &p.xmap, gi * kW4GroupK, 0, &x_full[as]);
                    tma_load_2d(dst + NTOK * 128, &p.xmap, gi * kW4GroupK + 64, 0, &x_full[as]);
                    if (++as == C::AS) {
                        as = 0;
                        ph ^= 1u;
                    }
                }
            }
        }
    } else if (warp == kTcWarpMma) {
        // ---------------- one thread issues every MMA of the CTA ----------------
        if (lane == 0) {
            constexpr uint32_t idesc = tc_idesc_f16(NTOK);
            int as = 0, acc = 0;
            uint32_t ph = 0, acc_ph = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
                const int split = it % S;
                const int g0 = (split * G) / S, g1 = ((split + 1) * G) / S;
                mbar_wait_wd(&acc_empty[acc], acc_ph ^ 1u, p.err, 0x300 + acc);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * NTOK);
                for (int gi = g0; gi < g1; ++gi) {
                    mbar_wait_wd(&a_full[as], ph, p.err, 0x400 + as);
                    mbar_wait_wd(&x_full[as], ph, p.err, 0x500 + as);
                    tc_fence_after();
                    const uint32_t a_base = smem_u32(smem + C::kAOff + as * kTcAStage);
                    const uint32_t x_base = smem_u32(smem + C::kXOff + as * C::kXStage);
#pragma unroll
                    for (int ka = 0; ka < 2; ++ka) {
                        const uint64_t ad = tc_desc_sw128(a_base + ka * (kTcRows * 128));
                        const uint64_t xd = tc_desc_sw128(x_base + ka * (NTOK * 128));
#pragma unroll
                        for (int k16 = 0; k16 < 4; ++k16)   // 16 fp16 = 32 bytes along K inside the swizzle atom
                            tc_mma_f16(d_tmem, ad + (uint64_t)(k16 * 2), xd + (uint64_t)(k16 * 2), idesc,
                                       (gi > g0 || ka > 0 || k16 > 0) ? 1u : 0u);
                    }
                    tc_commit(&ax_empty[as]);   // frees the A and x stage once the MMAs above have read them
                    if (++as == C::AS) {
                        as = 0;
                        ph ^= 1u;
                    }
                }
                tc_commit(&acc_full[acc]);
                if (++acc == 2) {
                    acc = 0;
                    acc_ph ^= 1u;
                }
            }
        }
    } else if (warp < kTcWarpEpi0) {
        // ---------------- dequant: raw int4 blocks -> fp16 A operand (K-major, 128-byte swizzle) ----------------
        // Two groups of four warps take the stages alternately (group = stage parity): while one group sits in its proxy
        // fence / barrier round trip, the other one is converting -- measured with ncu (profiles/r02_w4a16_tc_*): with all
        // eight warps on the same stage the SM idled through every fence and MMA hand-over.  Inside a group warp w owns
        // the 32-row block w of the stage (rows 32 w .. 32 w + 31, all 128 k).
        const int dw = warp - kTcWarpDq0;
        const int grp = dw >> 2, blk_i = dw & 3;
        const int g = lane >> 2, t = lane & 3;
        const uint32_t ka_off = (uint32_t)(t >> 1) * (kTcRows * 128);
        // lanes t and t ^ 2 of a quarter-warp write the same chunk column of the two k atoms: they walk the k-steps in a
        // different order (j ^ 2 for t >= 2) so that one STS.128 wavefront never hits a bank twice
        const int hx = t >> 1;             // which uint4 (hh) this lane converts first
        const int c_base = (t & 1) * 4;
        int stage = 0;                     // running stage counter of the CTA (all items)
        for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
            const int split = it % S;
            const int g0 = (split * G) / S, g1 = ((split + 1) * G) / S;
            for (int gi = g0; gi < g1; ++gi, ++stage) {
                if ((stage & 1) != grp) continue;
                const int rs = stage % C::RS, as = stage % C::AS;
                const uint32_t rph = (uint32_t)(stage / C::RS) & 1u, aph = (uint32_t)(stage / C::AS) & 1u;
                // one lane polls, the warp follows
                if (lane == 0) mbar_wait_wd(&raw_full[rs], rph, p.err, 0x600 + rs);
                __syncwarp();
                const uint8_t* blk = smem + C::kRawOff + rs * kTcRawStage + blk_i * kW4BlockBytes;
                uint4 wv[2][2];            // [tt][first / second k-half in this lane's order]
                __half2 sc[2];
                int zz[2];
#pragma unroll
                for (int tt = 0; tt < 2; ++tt) {
                    wv[tt][0] = *reinterpret_cast<const uint4*>(blk + ((tt * 2 + hx) * 32 + lane) * 16);
                    wv[tt][1] = *reinterpret_cast<const uint4*>(blk + ((tt * 2 + (hx ^ 1)) * 32 + lane) * 16);
                    sc[tt] = *reinterpret_cast<const __half2*>(blk + kW4ScaleOff + (tt * 8 + g) * 4);
                    zz[tt] = blk[kW4ZeroOff + tt * 8 + g];
                }
                __syncwarp();                       // every lane's shared-memory reads of the stage have returned
                if (lane == 0) mbar_arrive(&raw_empty[rs]);
                if (lane == 0) mbar_wait_wd(&ax_empty[as], aph ^ 1u, p.err, 0x700 + as);
                __syncwarp();
                uint8_t* a_st = smem + C::kAOff + as * kTcAStage + ka_off;
#pragma unroll
                for (int tt = 0; tt < 2; ++tt) {
                    const __half2 c_lo = __float2half2_rn((float)(1024 + (zz[tt] & 0xF)));
                    const __half2 c_hi = __float2half2_rn((float)(64 + (zz[tt] >> 4)));
                    const __half2 s_lo = __half2half2(__low2half(sc[tt])), s_hi = __half2half2(__high2half(sc[tt]));
                    // rows 32 blk + 16 tt + g (low nibbles) and + 8 (high nibbles): 8-row groups 4 blk + 2 tt and + 1
                    const uint32_t row_lo = (uint32_t)(blk_i * 4 + tt * 2) * 1024u + (uint32_t)g * 128u;
                    const uint32_t row_hi = row_lo + 1024u;
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const uint4 w = wv[tt][hh];
                        const int j0 = ((hh ^ hx) << 1);          // k-steps 2 (hh ^ hx) and + 1 of the 128-k group
                        const uint32_t words[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                        for (int jj = 0; jj < 2; ++jj) {
                            uint32_t lo0[2], hi0[2], lo1[2], hi1[2];
                            tc_dequant_word(words[2 * jj], c_lo, s_lo, c_hi, s_hi, lo0, hi0);
                            tc_dequant_word(words[2 * jj + 1], c_lo, s_lo, c_hi, s_hi, lo1, hi1);
                            const uint32_t chunk = (uint32_t)(((c_base + j0 + jj) ^ g) << 4);   // swizzle: chunk ^ (row % 8)
                            *reinterpret_cast<uint4*>(a_st + row_lo + chunk) = make_uint4(lo0[0], lo0[1], lo1[0], lo1[1]);
                            *reinterpret_cast<uint4*>(a_st + row_hi + chunk) = make_uint4(hi0[0], hi0[1], hi1[0], hi1[1]);
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> tensor core
                __syncwarp();                       // all 32 lanes have fenced their stores; one arrival per warp
                if (lane == 0) mbar_arrive(&a_full[as]);
            }
        }
    } else {
        // ---------------- epilogue: TMEM -> registers -> global ----------------
        const int q = warp & 3;                  // TMEM lane quadrant this warp may read
        const int m = q * 32 + lane;             // row inside the tile
        const int et = (warp - kTcWarpEpi0) * 32 + lane;
        pdl_wait();
        int acc = 0;
        uint32_t acc_ph = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
            const int tile = it / S, split = it % S;
            const int prow = tile * kTcRows + m;
            if (lane == 0) mbar_wait_wd(&acc_full[acc], acc_ph, p.err, 0x800 + acc);
            __syncwarp();
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * NTOK);
            float* wsp = (S > 1) ? p.ws + ((size_t)(tile * S + split) * p.M) * kTcRows + m : nullptr;
#pragma unroll 1
            for (int c0 = 0; c0 < NTOK; c0 += 16) {
                if (c0 >= p.M) break;
                float v[16];
                tc_ld16(taddr + (uint32_t)c0, v);
                if (S > 1) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (c0 + i < p.M) __stcg(wsp + (size_t)(c0 + i) * kTcRows, v[i]);
                } else {
                    tc_epilogue16(p, prow, lane, c0, v);
                }
            }
            tc_fence_before();
            __syncwarp();                           // one arrival per epilogue warp
            if (lane == 0) mbar_arrive(&acc_empty[acc]);
            if (++acc == 2) {
                acc = 0;
                acc_ph ^= 1u;
            }
            if (S > 1) {
                // the last CTA of the tile reduces the S partial sums in split order and runs the epilogue
                __threadfence();
                epi_bar();
                if (et == 0) *s_last = (atomicAdd(&p.counters[tile], 1u) == (unsigned)(S - 1)) ? 1u : 0u;
                epi_bar();
                const bool last = *s_last != 0u;
                epi_bar();   // s_last may be rewritten by the next item
                if (last) {
                    __threadfence();
                    const float* base = p.ws + ((size_t)tile * S * p.M) * kTcRows + m;
#pragma unroll 1
                    for (int c0 = 0; c0 < p.M; c0 += 16) {
                        float v[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = 0.f;
                        for (int s = 0; s < S; ++s) {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (c0 + i < p.M) v[i] += __ldcg(base + ((size_t)s * p.M + c0 + i) * kTcRows);
                        }
                        tc_epilogue16(p, prow, lane, c0, v);
                    }
                    if (et == 0) p.counters[tile] = 0u;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kTcWarpMma) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::kTmemCols));
    }
}

// =======================================================================================================================
// TS variant: the A operand (dequantised weights) lives in TENSOR MEMORY, not in shared memory.
//
// Why (measured on the SS kernel above, profiles/r02_w4a16_tc_*): a 128 x 128 fp16 A stage is 32 KB that the dequant warps
// write to shared memory and the tensor core reads back -- 64 KB of shared-memory traffic per 8.5 KB of HBM traffic, above
// the 128 B/clk the SM has at the HBM rate -- plus a generic->async proxy fence per hand-over and ~10 address / store
// instructions per converted word.  tcgen05.mma takes A from TMEM: TMEM lane = weight row, 32-bit column = two consecutive
// k, so a dequant thread that owns ONE ROW converts its nibbles in registers and hands them over with tcgen05.st -- no
// shared-memory round trip, no swizzle arithmetic, no proxy fence.
//
//   warps 0-3          epilogue   (TMEM lane quadrant = warp)
//   warps 4 .. 4+4KQ-1 dequant    (quadrant = warp % 4 = 32-row block of the tile; k-slice = (warp - 4) / 4 of KQ)
//   then               raw producer (tensor-TMA, see below), x producer, MMA issuer (+ TMEM alloc)
//
// A ZLW4I word holds 4 k of row g (low nibbles) and of row g + 8 (high nibbles); the thread of row g masks the low
// nibbles, the thread of row g + 8 the high ones (same words, a shared-memory broadcast).  Lanes g = 0..7 of a
// quarter-warp read 16-byte chunks 64 B apart -- a 4-way bank conflict in the plain record layout -- so the nibble words
// arrive through a tensor map with 64-byte swizzle (chunk ^= (row >> 1) & 3): conflict-free, the HBM format is unchanged.
// The 80-byte scale / zero trailers of the four blocks of a stage come through a second (unswizzled) map.
template <int NTOK>
struct TsCfg {
    static constexpr int KQ = 2;                                        // k-slices per quadrant (dequant warps = 4 KQ)
    static constexpr int DQ = 4 * KQ;
    static constexpr int kWarpDq0 = 4, kWarpRaw = 4 + DQ, kWarpX = kWarpRaw + 1, kWarpMma = kWarpRaw + 2;
    static constexpr int kThreads = (kWarpMma + 1) * 32;
    static constexpr int AS = NTOK <= 128 ? 4 : 3;                      // A (TMEM) / x (smem) stages
    static constexpr int RS = NTOK <= 128 ? 4 : 3;                      // raw weight stages
    static constexpr int NACC = NTOK <= 128 ? 2 : 1;                    // accumulator buffers in TMEM
    static constexpr int kACol0 = NACC * NTOK;                          // first A column; 64 columns per stage
    static constexpr int kXStage = 2 * NTOK * 128;
    static constexpr int kNibStage = 4 * 2048, kTrStage = 384;          // 4 x 80 B trailers, padded to the TMA alignment
    static constexpr int kXOff = 0;
    static constexpr int kNibOff = kXOff + AS * kXStage;
    static constexpr int kTrOff = kNibOff + RS * kNibStage;
    static constexpr int kBarOff = kTrOff + RS * kTrStage;
    static constexpr int kNumBars = 2 * RS + 3 * AS + 4;
    static constexpr int kMiscOff = kBarOff + kNumBars * 8;
    static constexpr int kBytes = kMiscOff + 16 + 1024;
    static_assert(kACol0 + AS * 64 <= 512, "TMEM columns");
};

__device__ __forceinline__ void tc_mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}

template <int NTOK>
__global__ void __launch_bounds__(TsCfg<NTOK>::kThreads, 1) k_w4a16_ts(const __grid_constant__ W4TcParams p) {
    using C = TsCfg<NTOK>;
    extern __shared__ uint8_t smem_dyn[];
    uint8_t* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kBarOff);
    uint64_t* raw_full = bars;
    uint64_t* raw_empty = raw_full + C::RS;
    uint64_t* a_full = raw_empty + C::RS;
    uint64_t* x_full = a_full + C::AS;
    uint64_t* ax_empty = x_full + C::AS;
    uint64_t* acc_full = ax_empty + C::AS;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + C::kMiscOff);
    uint32_t* s_last = s_tmem + 1;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = p.K / kW4GroupK, n_tiles = p.N / kTcRows, S = p.S;
    const int n_items = n_tiles * S;

    if (threadIdx.x == 0) {
        for (int i = 0; i < C::RS; ++i) {
            mbar_init(&raw_full[i], 1);
            mbar_init(&raw_empty[i], C::DQ);
        }
        for (int i = 0; i < C::AS; ++i) {
            mbar_init(&a_full[i], C::DQ);
            mbar_init(&x_full[i], 1);
            mbar_init(&ax_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 4);
        }
        mbar_fence_init();
    }
    if (warp == C::kWarpMma) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp == C::kWarpRaw) {
        // ---------------- weight stream: constants, may run ahead of the predecessor kernel (PDL) ----------------
        if (lane == 0) {
            const uint64_t pol = l2_evict_first_policy();
            int rs = 0;
            uint32_t ph = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
                const int tile = it / S, split = it % S;
                const int g0 = (split * G) / S, g1 = ((split + 1) * G) / S;
                for (int gi = g0; gi < g1; ++gi) {
                    mbar_wait_wd(&raw_empty[rs], ph ^ 1u, p.err, 0x100 + rs);
                    if (p.dbg & 8) {
                        mbar_arrive(&raw_full[rs]);
                    } else {
                        mbar_expect_tx(&raw_full[rs], C::kNibStage + 4 * 80);
                        tma_load_4d_hint(smem + C::kNibOff + rs * C::kNibStage, &p.wmap, 0, 0, gi, tile * 4, &raw_full[rs], pol);
                        tma_load_3d_hint(smem + C::kTrOff + rs * C::kTrStage, &p.tmap, 0, gi, tile * 4, &raw_full[rs], pol);
                    }
                    if (++rs == C::RS) {
                        rs = 0;
                        ph ^= 1u;
                    }
                }
            }
        }
    } else if (warp == C::kWarpX) {
        // ---------------- activations: produced by the predecessor kernel ----------------
        if (lane == 0) {
            pdl_wait();
            int as = 0;
            uint32_t ph = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
                const int split = it % S;
                const int g0 = (split * G) / S, g1 = ((split + 1) * G) / S;
                for (int gi = g0; gi < g1; ++gi) {
                    mbar_wait_wd(&ax_empty[as], ph ^ 1u, p.err, 0x200 + as);
                    if (p.dbg & 1) {
                        mbar_arrive(&x_full[as]);
                    } else {
                        mbar_expect_tx(&x_full[as], C::kXStage);
                        uint8_t* dst = smem + C::kXOff + as * C::kXStage;
                        tma_load_2d(dst, &p.xmap, gi * kW4GroupK, 0, &x_full[as]);
                        tma_load_2d(dst + NTOK * 128, &p.xmap, gi * kW4GroupK + 64, 0, &x_full[as]);
                    }
                    if (++as == C::AS) {
                        as = 0;
                        ph ^= 1u;
                    }
                }
            }
        }
    } else if (warp == C::kWarpMma) {
        // ---------------- one thread issues every MMA of the CTA: A from TMEM, B (activations) from shared memory ----------------
        if (lane == 0) {
            constexpr uint32_t idesc = tc_idesc_f16(NTOK);
            int as = 0, acc = 0;
            uint32_t ph = 0, acc_ph = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
                const int split = it % S;
                const int g0 = (split * G) / S, g1 = ((split + 1) * G) / S;
                mbar_wait_wd(&acc_empty[acc], acc_ph ^ 1u, p.err, 0x300 + acc);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * NTOK);
                for (int gi = g0; gi < g1; ++gi) {
                    mbar_wait_wd(&a_full[as], ph, p.err, 0x400 + as);
                    mbar_wait_wd(&x_full[as], ph, p.err, 0x500 + as);
                    tc_fence_after();
                    const uint32_t a_col = tmem_base + (uint32_t)(C::kACol0 + as * 64);
                    const uint32_t x_base = smem_u32(smem + C::kXOff + as * C::kXStage);
                    if (!(p.dbg & 4) || gi == g0)
#pragma unroll
                    for (int ka = 0; ka < 2; ++ka) {
                        const uint64_t xd = tc_desc_sw128(x_base + ka * (NTOK * 128));
#pragma unroll
                        for (int k16 = 0; k16 < 4; ++k16)   // 16 k = 8 TMEM columns of A, 32 bytes inside the swizzle atom of x
                            tc_mma_f16_ts(d_tmem, a_col + (uint32_t)((ka * 4 + k16) * 8), xd + (uint64_t)(k16 * 2), idesc,
                                          (gi > g0 || ka > 0 || k16 > 0) ? 1u : 0u);
                    }
                    tc_commit(&ax_empty[as]);   // frees the A columns and the x stage once the MMAs above have read them
                    if (++as == C::AS) {
                        as = 0;
                        ph ^= 1u;
                    }
                }
                tc_commit(&acc_full[acc]);
                if (++acc == C::NACC) {
                    acc = 0;
                    acc_ph ^= 1u;
                }
            }
        }
    } else if (warp >= C::kWarpDq0) {
        // ---------------- dequant: the lane owns packed row 32 q + lane of the tile, k-slice ks of every stage ----------------
        const int q = warp & 3, ks = (warp - C::kWarpDq0) >> 2;
        const int tt = lane >> 4, hi = (lane >> 3) & 1, g = lane & 7;
        constexpr int KW = kW4GroupK / C::KQ;            // k per warp and stage: 64 (KQ = 2) or 32 (KQ = 4)
        constexpr int NCH = KW / 16;                     // 16-k chunks (one uint4 each) per lane and stage
        const uint32_t mask = hi ? 0xf0f0f0f0u : 0x0f0f0f0fu;
        const uint32_t magic = hi ? 0x54545454u : 0x64646464u;   // 64 + n / 16 for a high nibble n, 1024 + n for a low one
        const uint32_t sw = (uint32_t)(g >> 1) & 3u;     // the map's 64-byte swizzle: chunk ^= (row >> 1) & 3, row % 8 = g
        const uint32_t lane_taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(C::kACol0 + ks * (KW / 2));
        int rs = 0, as = 0;
        uint32_t rph = 0, aph = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
            const int split = it % S;
            const int g0 = (split * G) / S, g1 = ((split + 1) * G) / S;
            for (int gi = g0; gi < g1; ++gi) {
                if (lane == 0) mbar_wait_wd(&raw_full[rs], rph, p.err, 0x600 + rs);
                __syncwarp();
                const uint8_t* nib = smem + C::kNibOff + rs * C::kNibStage + q * 2048;
                const uint8_t* tr = smem + C::kTrOff + rs * C::kTrStage + q * 80;
                uint4 wv[NCH];
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    // 16-k chunk kc of the group: k = 32 t + 16 hh + ..., record row (2 tt + hh) * 8 + g, chunk t
                    const int kc = ks * NCH + c, t = kc >> 1, hh = kc & 1;
                    wv[c] = *reinterpret_cast<const uint4*>(nib + ((tt * 2 + hh) * 8 + g) * 64 + (((uint32_t)t ^ sw) << 4));
                }
                const __half2 sc2 = *reinterpret_cast<const __half2*>(tr + (tt * 8 + g) * 4);
                const int zz = tr[64 + tt * 8 + g];
                __syncwarp();                       // every lane's shared-memory reads of the stage have returned
                if (lane == 0) mbar_arrive(&raw_empty[rs]);
                const __half2 cz = __float2half2_rn(hi ? (float)(64 + (zz >> 4)) : (float)(1024 + (zz & 0xF)));
                const __half2 s2 = hi ? __half2half2(__high2half(sc2)) : __half2half2(__low2half(sc2));
                if (lane == 0) mbar_wait_wd(&ax_empty[as], aph ^ 1u, p.err, 0x700 + as);
                __syncwarp();
                tc_fence_after();
                if (!(p.dbg & 2))
#pragma unroll
                for (int c2 = 0; c2 < NCH; c2 += 2) {   // 32 k = 16 columns per tcgen05.st
                    uint32_t r[16];
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const uint32_t words[4] = {wv[c2 + c].x, wv[c2 + c].y, wv[c2 + c].z, wv[c2 + c].w};
#pragma unroll
                        for (int wq = 0; wq < 4; ++wq) {
                            const uint32_t wm = words[wq] & mask;
                            uint32_t a0 = __byte_perm(wm, magic, 0x4140), a1 = __byte_perm(wm, magic, 0x4342);
                            __half2 v0 = __hmul2(__hsub2(*reinterpret_cast<__half2*>(&a0), cz), s2);
                            __half2 v1 = __hmul2(__hsub2(*reinterpret_cast<__half2*>(&a1), cz), s2);
                            r[c * 8 + wq * 2] = *reinterpret_cast<uint32_t*>(&v0);
                            r[c * 8 + wq * 2 + 1] = *reinterpret_cast<uint32_t*>(&v1);
                        }
                    }
                    tc_st16(lane_taddr + (uint32_t)(as * 64 + c2 * 8), r);
                }
                tc_wait_st();
                tc_fence_before();
                __syncwarp();                       // one arrival per warp
                if (lane == 0) mbar_arrive(&a_full[as]);
                if (++rs == C::RS) {
                    rs = 0;
                    rph ^= 1u;
                }
                if (++as == C::AS) {
                    as = 0;
                    aph ^= 1u;
                }
            }
        }
    } else {
        // ---------------- epilogue: TMEM -> registers -> global ----------------
        const int q = warp;                      // TMEM lane quadrant this warp may read
        const int m = q * 32 + lane;             // row inside the tile
        const int et = warp * 32 + lane;
        pdl_wait();
        int acc = 0;
        uint32_t acc_ph = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
            const int tile = it / S, split = it % S;
            const int prow = tile * kTcRows + m;
            if (lane == 0) mbar_wait_wd(&acc_full[acc], acc_ph, p.err, 0x800 + acc);
            __syncwarp();
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * NTOK);
            float* wsp = (S > 1) ? p.ws + ((size_t)(tile * S + split) * p.M) * kTcRows + m : nullptr;
#pragma unroll 1
            for (int c0 = 0; c0 < NTOK; c0 += 16) {
                if (c0 >= p.M) break;
                float v[16];
                tc_ld16(taddr + (uint32_t)c0, v);
                if (S > 1) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (c0 + i < p.M) __stcg(wsp + (size_t)(c0 + i) * kTcRows, v[i]);
                } else {
                    tc_epilogue16(p, prow, lane, c0, v);
                }
            }
            tc_fence_before();
            __syncwarp();                           // one arrival per epilogue warp
            if (lane == 0) mbar_arrive(&acc_empty[acc]);
            if (++acc == C::NACC) {
                acc = 0;
                acc_ph ^= 1u;
            }
            if (S > 1) {
                // the last CTA of the tile reduces the S partial sums in split order and runs the epilogue
                __threadfence();
                epi_bar();
                if (et == 0) *s_last = (atomicAdd(&p.counters[tile], 1u) == (unsigned)(S - 1)) ? 1u : 0u;
                epi_bar();
                const bool last = *s_last != 0u;
                epi_bar();   // s_last may be rewritten by the next item
                if (last) {
                    __threadfence();
                    const float* base = p.ws + ((size_t)tile * S * p.M) * kTcRows + m;
#pragma unroll 1
                    for (int c0 = 0; c0 < p.M; c0 += 16) {
                        float v[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = 0.f;
                        for (int s = 0; s < S; ++s) {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (c0 + i < p.M) v[i] += __ldcg(base + ((size_t)s * p.M + c0 + i) * kTcRows);
                        }
                        tc_epilogue16(p, prow, lane, c0, v);
                    }
                    if (et == 0) p.counters[tile] = 0u;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == C::kWarpMma) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    }
}

// ---- host side --------------------------------------------------------------------------------------------------------
static TcDeviceState g_tc_state[64];

TcDeviceState* tc_state() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < 64) ? &g_tc_state[dev] : nullptr;
}

// allocates the per-device split-k workspace and sets the opt-in shared-memory sizes; must run outside stream capture
cudaError_t prepare_w4_tc() {
    TcDeviceState* st = tc_state();
    if (!st) return cudaErrorInvalidDevice;
    cudaError_t e;
    if (!st->ws) {
        if ((e = cudaMalloc((void**)&st->ws, kTcWsBytes)) != cudaSuccess) return e;
        if ((e = cudaMalloc((void**)&st->counters, (kTcMaxTiles + 16) * sizeof(unsigned))) != cudaSuccess) return e;
        if ((e = cudaMemset(st->counters, 0, (kTcMaxTiles + 16) * sizeof(unsigned))) != cudaSuccess) return e;
        st->err = st->counters + kTcMaxTiles;
        st->ws_bytes = kTcWsBytes;
    }
#define ZL_TC_SET(NT)                                                                                             \
    if ((e = cudaFuncSetAttribute(k_w4a16_tc<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,                     \
                                  TcCfg<NT>::kBytes)) != cudaSuccess)                                              \
        return e;
    ZL_TC_SET(32) ZL_TC_SET(64) ZL_TC_SET(128) ZL_TC_SET(256)
#undef ZL_TC_SET
#define ZL_TS_SET(NT)                                                                                             \
    if ((e = cudaFuncSetAttribute(k_w4a16_ts<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,                     \
                                  TsCfg<NT>::kBytes)) != cudaSuccess)                                              \
        return e;
    ZL_TS_SET(32) ZL_TS_SET(64) ZL_TS_SET(128) ZL_TS_SET(256)
#undef ZL_TS_SET
    return cudaSuccess;
}

bool w4_tc_supports(int mc, int N, int K) {
    return mc >= 1 && mc <= 256 && N % kTcRows == 0 && K % kW4GroupK == 0 && N / kTcRows <= kTcMaxTiles;
}

// k-split: fill the SMs when there are fewer row tiles than SMs (or an awkward number of waves), never below 4 groups
// per item; every extra split costs an fp32 round trip of the (M x 128) partial tile through L2
static int tc_pick_splits(int n_tiles, int G, int mc, size_t ws_bytes) {
    const int sms = device_sm_count();
    int best = 1;
    float best_score = -1.f;
    for (int s = 1; s <= 8 && s * 4 <= G; ++s) {
        const long long items = (long long)n_tiles * s;
        if (s > 1 && (size_t)items * mc * kTcRows * 4 > ws_bytes) break;
        const long long waves = (items + sms - 1) / sms;
        const float eff = (float)items / (float)(waves * sms);
        const float score = eff - 0.03f * (s - 1);
        if (score > best_score + 1e-6f) {
            best_score = score;
            best = s;
        }
    }
    return best;
}

// last watchdog code published by a tcgen05 kernel on this device (0 = none); debugging aid
extern "C" unsigned zl_w4_tc_watchdog(void) {
    TcDeviceState* st = tc_state();
    unsigned v = 0;
    if (st && st->err) cudaMemcpy(&v, st->err, 4, cudaMemcpyDeviceToHost);
    return v;
}

// k-split override: ZL_TC_SPLITS in the environment, or zl_w4_tc_set_splits (tests exercise the split-k reduction on
// small shapes); 0 = automatic
static int g_tc_splits = -2;
static int tc_force_splits() {
    if (g_tc_splits == -2) {
        const char* e = getenv("ZL_TC_SPLITS");
        g_tc_splits = e ? atoi(e) : 0;
    }
    return g_tc_splits;
}
extern "C" int zl_w4_tc_set_splits(int splits) {
    g_tc_splits = splits < 0 ? 0 : splits;
    return ZL_OK;
}

cudaError_t launch_w4_tc(const W4Params& p, bool pdl, cudaStream_t stream) {
    TcDeviceState* st = tc_state();
    if (!st || !st->ws) return cudaErrorNotSupported;
    const int ntok = p.mc <= 32 ? 32 : p.mc <= 64 ? 64 : p.mc <= 128 ? 128 : 256;
    W4TcParams q;
    if (!tc_make_map_2d(&q.xmap, p.x, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (uint64_t)p.mc, (uint64_t)p.K, (uint64_t)p.ldx * 2,
                        (uint32_t)ntok))
        return cudaErrorInvalidValue;
    q.packed = p.packed;
    q.bias = p.bias;
    q.residual = p.residual;
    q.y = p.y;
    q.M = p.mc;
    q.N = p.N;
    q.K = p.K;
    q.epi = p.epi;
    {
        static const int tc_dbg = getenv("ZL_TC_DBG") ? atoi(getenv("ZL_TC_DBG")) : 0;
        q.dbg = tc_dbg;
    }
    const int n_tiles = p.N / kTcRows, G = p.K / kW4GroupK;
    const int forced = tc_force_splits();
    q.S = (forced > 0 && forced <= G) ? forced : tc_pick_splits(n_tiles, G, p.mc, st->ws_bytes);
    if (q.S > 1 && (size_t)n_tiles * q.S * p.mc * kTcRows * 4 > st->ws_bytes) q.S = 1;
    q.ws = st->ws;
    q.counters = st->counters;
    q.err = st->err;
    q.cos = p.cos;
    q.sin = p.sin;
    q.q_out = p.q_out;
    q.token_batch = p.token_batch;
    q.placement = p.placement;
    q.k_addrs = p.k_addrs;
    q.v_addrs = p.v_addrs;
    q.num_heads = p.num_heads;
    q.num_kv_heads = p.num_kv_heads;
    q.dim_head = p.dim_head;
    const int items = n_tiles * q.S;
    const int sms = device_sm_count();
    static const bool use_ss = getenv("ZL_W4_TC_SS") != nullptr;   // A/B: the shared-memory-A variant
    if (!use_ss) {
        const uint64_t nb = (uint64_t)p.N / 32;
        const uint64_t wd[4] = {16, 32, (uint64_t)G, nb}, wstr[3] = {64, (uint64_t)kW4BlockBytes, (uint64_t)G * kW4BlockBytes};
        const uint32_t wbox[4] = {16, 32, 1, 4};
        const uint64_t td[3] = {20, (uint64_t)G, nb}, tstr[2] = {(uint64_t)kW4BlockBytes, (uint64_t)G * kW4BlockBytes};
        const uint32_t tbox[3] = {20, 1, 4};
        if (!tc_make_map_nd(&q.wmap, p.packed, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, wd, wstr, wbox, CU_TENSOR_MAP_SWIZZLE_64B) ||
            !tc_make_map_nd(&q.tmap, p.packed + kW4ScaleOff, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, td, tstr, tbox,
                            CU_TENSOR_MAP_SWIZZLE_NONE))
            return cudaErrorInvalidValue;
        const dim3 grid(items < sms ? items : sms);
        switch (ntok) {
            case 32: return launch(k_w4a16_ts<32>, grid, dim3(TsCfg<32>::kThreads), (size_t)TsCfg<32>::kBytes, stream, pdl, q);
            case 64: return launch(k_w4a16_ts<64>, grid, dim3(TsCfg<64>::kThreads), (size_t)TsCfg<64>::kBytes, stream, pdl, q);
            case 128: return launch(k_w4a16_ts<128>, grid, dim3(TsCfg<128>::kThreads), (size_t)TsCfg<128>::kBytes, stream, pdl, q);
            default: return launch(k_w4a16_ts<256>, grid, dim3(TsCfg<256>::kThreads), (size_t)TsCfg<256>::kBytes, stream, pdl, q);
        }
    }
    const dim3 grid(items < sms ? items : sms), block(kTcThreads);
    switch (ntok) {
        case 32: return launch(k_w4a16_tc<32>, grid, block, (size_t)TcCfg<32>::kBytes, stream, pdl, q);
        case 64: return launch(k_w4a16_tc<64>, grid, block, (size_t)TcCfg<64>::kBytes, stream, pdl, q);
        case 128: return launch(k_w4a16_tc<128>, grid, block, (size_t)TcCfg<128>::kBytes, stream, pdl, q);
        default: return launch(k_w4a16_tc<256>, grid, block, (size_t)TcCfg<256>::kBytes, stream, pdl, q);
    }
}

}  // namespace zl
