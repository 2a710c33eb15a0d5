// W4A16 skinny GEMM, exact-integer variant (ZLW4I layout): the default for small M.
//
// Why: on B200 a decode GEMM must consume ~23 weight bytes per cycle per SM to keep up with HBM.  The classic
// "dequantise int4 -> fp16 on the CUDA cores, then HMMA" recipe (our v1/v2 kernels, Marlin, the reference's
// GEMV) costs ~10 issue slots per packed word, which makes the kernel ISSUE-bound before it is HBM-bound
// (profiles/r01b_w4a16_v2.txt: 48 % issue utilisation at 32 % of HBM peak).  Here the int4 weights never
// become floating point:
//   * activations are decomposed once per CTA into a block-floating-point integer  x_k = 2^e_g * m_k,
//     m_k a 16-bit integer per 128-group (error <= 2^-16 of the group maximum, far below fp16's own 2^-11),
//     split into a signed high byte and an unsigned low byte;
//   * a packed weight word turns into two IMMA A registers with two LOP3s (w & 0x0f0f0f0f, w & 0xf0f0f0f0 --
//     the odd nibbles stay multiplied by 16 and the factor is folded into that row's scale);
//   * sum_k q_k * m_k is accumulated EXACTLY in int32 by mma.sync.m16n8k32 (u8 x s8 and u8 x u8), the zero
//     point is removed with the exact group sum of m, and only then one fp32 multiply by
//     scale[row,g] * 2^e_g happens per group.
// ~3 issue slots per word instead of ~10, and the result is closer to the exact product than either the fp16
// reference kernel or our fp16 path.  Same ring / persistence / fused prologue+epilogues as v2.
#include "common.cuh"
#include "w4_layout.cuh"
#include "w4_params.h"

namespace zl {

__device__ __forceinline__ void imma_u8s8(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void imma_u8u8(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// per-warp staging area: for every token row the warp's k-slice as hi / lo bytes, plus (sum m, 2^e) per group
template <int NT, int WARPS, int STAGES>
struct V3Smem {
    static constexpr int kRingBytes = WARPS * STAGES * kW4BlockBytes;
    static constexpr int kRedBufs = (NT == 1 && WARPS == 8) ? 2 : 1;
    static constexpr int kRedFloats = WARPS * NT * 8 * 32;
    static constexpr int kBarOff = kRingBytes;
    static constexpr int kRedOff = kBarOff + WARPS * STAGES * 8;
    static constexpr int kSsOff = kRedOff + kRedBufs * kRedFloats * 4;   // [warps][NT*8]
    static constexpr int kRstdOff = kSsOff + WARPS * NT * 8 * 4;         // [NT*8]
    static constexpr int kXOff = (kRstdOff + NT * 8 * 4 + 127) & ~127;
    static constexpr int kBytes = kXOff;
};
// bytes of the staging area for one warp: mc rows x (hi + lo) + group table
__host__ __device__ inline int v3_row_bytes(int ng_max) { return ng_max * 128 + 16; }   // +16: conflict-free LDS.128
__host__ __device__ inline int v3_warp_bytes(int mc, int ng_max) {
    return 2 * mc * v3_row_bytes(ng_max) + ng_max * mc * 8;
}

template <int NT, bool NORM, int WARPS, int STAGES>
__global__ void __launch_bounds__(WARPS * 32, ((NT <= 2 && WARPS == 8) ? 2 : 1)) k_w4a16_v3(const W4Params p) {
    using S = V3Smem<NT, WARPS, STAGES>;
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int G = p.K / kW4GroupK;
    const int g_begin = (warp * G) / WARPS;
    const int g_end = ((warp + 1) * G) / WARPS;
    const int ng = g_end - g_begin;
    const int ng_max = (G + WARPS - 1) / WARPS;
    const int n_tiles = p.N / 32;
    const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total = my_tiles * ng;

    uint8_t* ring = smem + warp * (STAGES * kW4BlockBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBarOff) + warp * STAGES;
    float* red = reinterpret_cast<float*>(smem + S::kRedOff);
    float* s_ss = reinterpret_cast<float*>(smem + S::kSsOff);
    float* s_rstd = reinterpret_cast<float*>(smem + S::kRstdOff);
    const int row_b = v3_row_bytes(ng_max);
    uint8_t* xw_hi = smem + S::kXOff + warp * v3_warp_bytes(p.mc, ng_max);   // [tok][row_b]
    uint8_t* xw_lo = xw_hi + p.mc * row_b;                                   // [tok][row_b]
    int2* xw_tab = reinterpret_cast<int2*>(xw_lo + p.mc * row_b);            // [group][tok] {sum m, bits of 2^e}

    auto item_src = [&](int it) -> const uint8_t* {
        const int tile = (int)blockIdx.x + (it / ng) * (int)gridDim.x;
        const int gi = g_begin + it % ng;
        return p.packed + ((size_t)tile * G + gi) * kW4BlockBytes;
    };

    pdl_trigger();
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            if (s < total) {
                mbar_expect_tx(&bars[s], kW4BlockBytes);
                bulk_g2s(ring + s * kW4BlockBytes, item_src(s), kW4BlockBytes, &bars[s]);
            }
        }
    }
    __syncwarp();
    pdl_wait();

    // ---- stage this warp's k-slice of the activations as block-floating-point integers (warp-local) ----
    for (int tok = 0; tok < p.mc; ++tok) {
        float sq = 0.f;
        for (int gl = 0; gl < ng; ++gl) {
            const int k = (g_begin + gl) * kW4GroupK + lane * 4;
            uint2 raw = ld_cg_u2(p.x + (size_t)tok * p.ldx + k);
            __half2 h01 = *reinterpret_cast<__half2*>(&raw.x), h23 = *reinterpret_cast<__half2*>(&raw.y);
            if (NORM) {
                const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                sq = fmaf(f01.x, f01.x, sq);
                sq = fmaf(f01.y, f01.y, sq);
                sq = fmaf(f23.x, f23.x, sq);
                sq = fmaf(f23.y, f23.y, sq);
                const uint2 wr = *reinterpret_cast<const uint2*>(p.ln_w + k);
                h01 = __hmul2(h01, *reinterpret_cast<const __half2*>(&wr.x));
                h23 = __hmul2(h23, *reinterpret_cast<const __half2*>(&wr.y));
            }
            const float2 a = __half22float2(h01), b = __half22float2(h23);
            float amax = fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(b.x), fabsf(b.y)));
            amax = warp_max(amax);
            // 2^e with m = x * 2^-e in (-2^15, 2^15): e = floor(log2 amax) - 14
            const uint32_t ex = (__float_as_uint(amax) >> 23) & 0xffu;
            const bool zero = !(amax > 0.f) || !(amax < INFINITY) || ex < 20u;   // all-zero / non-finite / tiny group
            const float sg = zero ? 1.0f : __uint_as_float((ex - 14u) << 23);
            const float inv = zero ? 0.0f : __uint_as_float((254u - (ex - 14u)) << 23);
            const int m0 = __float2int_rn(a.x * inv), m1 = __float2int_rn(a.y * inv);
            const int m2 = __float2int_rn(b.x * inv), m3 = __float2int_rn(b.y * inv);
            const uint32_t lo = (uint32_t)(m0 & 255) | ((uint32_t)(m1 & 255) << 8) | ((uint32_t)(m2 & 255) << 16) |
                                ((uint32_t)(m3 & 255) << 24);
            const uint32_t hi = (uint32_t)((m0 >> 8) & 255) | ((uint32_t)((m1 >> 8) & 255) << 8) |
                                ((uint32_t)((m2 >> 8) & 255) << 16) | ((uint32_t)((m3 >> 8) & 255) << 24);
            *reinterpret_cast<uint32_t*>(xw_hi + tok * row_b + gl * 128 + lane * 4) = hi;
            *reinterpret_cast<uint32_t*>(xw_lo + tok * row_b + gl * 128 + lane * 4) = lo;
            int sm = m0 + m1 + m2 + m3;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
            if (lane == 0) xw_tab[gl * p.mc + tok] = make_int2(sm, (int)__float_as_uint(sg));
        }
        if (NORM) {
            sq = warp_sum(sq);
            if (lane == 0) s_ss[warp * (NT * 8) + tok] = sq;
        }
    }
    __syncwarp();

    int it = 0;
    for (int ti = 0; ti < my_tiles; ++ti) {
        const int st = (int)blockIdx.x + ti * (int)gridDim.x;
        float acc[2][NT][4];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < NT; ++b)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;

        for (int i = 0; i < ng; ++i, ++it) {
            const int s = it % STAGES;
            const uint32_t parity = (uint32_t)(it / STAGES) & 1u;
            // B fragments: 32 contiguous bytes of each piece per lane
            uint4 bh[NT][2], bl[NT][2];
            int2 tab[NT][2];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const int tok = nt * 8 + g;
                if (tok < p.mc) {
                    const uint8_t* ph = xw_hi + tok * row_b + i * 128 + t * 32;
                    const uint8_t* pl = xw_lo + tok * row_b + i * 128 + t * 32;
                    bh[nt][0] = *reinterpret_cast<const uint4*>(ph);
                    bh[nt][1] = *reinterpret_cast<const uint4*>(ph + 16);
                    bl[nt][0] = *reinterpret_cast<const uint4*>(pl);
                    bl[nt][1] = *reinterpret_cast<const uint4*>(pl + 16);
                } else {
                    bh[nt][0] = bh[nt][1] = bl[nt][0] = bl[nt][1] = make_uint4(0, 0, 0, 0);
                }
                // group table of the two tokens this lane's accumulators belong to (C columns 2t, 2t+1)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int tc = nt * 8 + 2 * t + e;
                    tab[nt][e] = tc < p.mc ? xw_tab[i * p.mc + tc] : make_int2(0, 0);
                }
            }

            mbar_wait(&bars[s], parity);
            const uint8_t* blk = ring + s * kW4BlockBytes;
            if (p.dbg & 1) {
                acc[0][0][0] += __uint_as_float(*reinterpret_cast<const uint32_t*>(blk + lane * 4)) * 1e-30f;
            } else {
                int ah[2][NT][4], al[2][NT][4];
#pragma unroll
                for (int tt = 0; tt < 2; ++tt)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int c = 0; c < 4; ++c) ah[tt][nt][c] = al[tt][nt][c] = 0;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    uint4 wv[2];
#pragma unroll
                    for (int tt = 0; tt < 2; ++tt)
                        wv[tt] = *reinterpret_cast<const uint4*>(blk + ((tt * 2 + hh) * 32 + lane) * 16);
#pragma unroll
                    for (int jp = 0; jp < 2; ++jp) {
                        const int j = hh * 2 + jp;   // k-step 0..3
#pragma unroll
                        for (int tt = 0; tt < 2; ++tt) {   // the two 16-row tiles interleave: independent chains
                            const uint32_t w0 = jp ? wv[tt].z : wv[tt].x;
                            const uint32_t w1 = jp ? wv[tt].w : wv[tt].y;
                            const uint32_t a[4] = {w0 & 0x0f0f0f0fu, w0 & 0xf0f0f0f0u, w1 & 0x0f0f0f0fu,
                                                   w1 & 0xf0f0f0f0u};
#pragma unroll
                            for (int nt = 0; nt < NT; ++nt) {
                                const uint4 vh = bh[nt][j >> 1], vl = bl[nt][j >> 1];
                                imma_u8s8(ah[tt][nt], a, (j & 1) ? vh.z : vh.x, (j & 1) ? vh.w : vh.y);
                                imma_u8u8(al[tt][nt], a, (j & 1) ? vl.z : vl.x, (j & 1) ? vl.w : vl.y);
                            }
                        }
                    }
                }
                // exact integer group result, then one fp32 multiply-add per element
#pragma unroll
                for (int tt = 0; tt < 2; ++tt) {
                    const __half2 sc = *reinterpret_cast<const __half2*>(blk + kW4ScaleOff + (tt * 8 + g) * 4);
                    const int zz = blk[kW4ZeroOff + tt * 8 + g];
                    const int z_lo = zz & 0xF, z_hi16 = zz & 0xF0;          // zero of row g ; 16 * zero of row g+8
                    const float s_lo = __low2float(sc), s_hi = __high2float(sc) * 0.0625f;
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const int2 tb = tab[nt][c & 1];
                            const int zc = (c >> 1) ? z_hi16 : z_lo;
                            const int r = ah[tt][nt][c] * 256 + al[tt][nt][c] - zc * tb.x;
                            const float f = ((c >> 1) ? s_hi : s_lo) * __int_as_float(tb.y);
                            acc[tt][nt][c] = fmaf((float)r, f, acc[tt][nt][c]);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0 && it + STAGES < total) {
                mbar_expect_tx(&bars[s], kW4BlockBytes);
                bulk_g2s(ring + s * kW4BlockBytes, item_src(it + STAGES), kW4BlockBytes, &bars[s]);
            }
        }

        // ---- split-k reduction across the warps + epilogue (same as v2) ----
        float* myred = red + (S::kRedBufs == 2 ? (ti & 1) * S::kRedFloats : 0) + warp * (NT * 8 * 32);
#pragma unroll
        for (int tt = 0; tt < 2; ++tt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const int tok = nt * 8 + 2 * t;
                const int row = tt * 16 + g;
                myred[tok * 32 + row] = acc[tt][nt][0];
                myred[(tok + 1) * 32 + row] = acc[tt][nt][1];
                myred[tok * 32 + row + 8] = acc[tt][nt][2];
                myred[(tok + 1) * 32 + row + 8] = acc[tt][nt][3];
            }
        __syncthreads();
        if (NORM && ti == 0) {
            if (threadIdx.x < p.mc) {
                float v = 0.f;
#pragma unroll
                for (int w = 0; w < WARPS; ++w) v += s_ss[w * (NT * 8) + threadIdx.x];
                s_rstd[threadIdx.x] = rsqrtf(v / (float)p.K + p.eps);
            }
            __syncthreads();
        }

        const float* rbase = red + (S::kRedBufs == 2 ? (ti & 1) * S::kRedFloats : 0);
        auto sum_red = [&](int idx) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) v += rbase[w * (NT * 8 * 32) + idx];
            return v;
        };
        const int n0 = st * 32;
        if (p.epi == ZL_EPI_SWIGLU) {
            const int n_out = p.N / 2;
            for (int e = threadIdx.x; e < p.mc * 16; e += blockDim.x) {
                const int tok = e >> 4, oc = e & 15;
                const int rg = (oc >> 3) * 16 + (oc & 7);
                float gate = sum_red(tok * 32 + rg);
                float up = sum_red(tok * 32 + rg + 8);
                if (NORM) {
                    gate *= s_rstd[tok];
                    up *= s_rstd[tok];
                }
                if (p.bias) {
                    gate += __half2float(p.bias[n0 + rg]);
                    up += __half2float(p.bias[n0 + rg + 8]);
                }
                const float gr = __half2float(__float2half_rn(gate));
                const float ur = __half2float(__float2half_rn(up));
                p.y[(size_t)tok * n_out + st * 16 + oc] = __float2half_rn(silu_f(gr) * ur);
            }
        } else if (p.epi == ZL_EPI_QKV_ROPE) {
            const int d = p.dim_head, half_dim = d / 2;
            const int tiles_per_head = d / 32;
            const int head = st / tiles_per_head, jt = st % tiles_per_head;
            for (int e = threadIdx.x; e < p.mc * 16; e += blockDim.x) {
                const int tok = e >> 4, oc = e & 15;
                const int rlo = (oc >> 3) * 16 + (oc & 7);
                const int c = jt * 16 + oc;
                float lo = sum_red(tok * 32 + rlo), hi = sum_red(tok * 32 + rlo + 8);
                if (NORM) {
                    lo *= s_rstd[tok];
                    hi *= s_rstd[tok];
                }
                if (p.bias) {
                    lo += __half2float(p.bias[n0 + rlo]);
                    hi += __half2float(p.bias[n0 + rlo + 8]);
                }
                lo = __half2float(__float2half_rn(lo));
                hi = __half2float(__float2half_rn(hi));
                __half olo, ohi;
                const bool is_v = head >= p.num_heads + p.num_kv_heads;
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
            for (int e = threadIdx.x; e < p.mc * 32; e += blockDim.x) {
                const int tok = e >> 5, row = e & 31;
                float v = sum_red(tok * 32 + row);
                if (NORM) v *= s_rstd[tok];
                if (p.bias) v += __half2float(p.bias[n0 + row]);
                __half h = __float2half_rn(v);
                if (p.epi == ZL_EPI_RESIDUAL)
                    h = __float2half_rn(__half2float(h) + __half2float(p.residual[(size_t)tok * p.N + n0 + row]));
                p.y[(size_t)tok * p.N + n0 + row] = h;
            }
        }
        if (S::kRedBufs == 1) __syncthreads();
    }
}

static int v3_num_sms() {
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        if (n_sm <= 0) n_sm = 148;
    }
    return n_sm;
}

constexpr int kV3TwoCtaBudget = 115000;
constexpr int kV3OneCtaBudget = 231000;

template <int NT, bool NORM, int WARPS, int STAGES>
static cudaError_t launch_v3_t(const W4Params& p, int smem, bool pdl, cudaStream_t stream) {
    const int tiles = p.N / 32;
    const int max_ctas = v3_num_sms() * (WARPS == 8 ? 2 : 1);
    const int grid = tiles < max_ctas ? tiles : max_ctas;
    return launch(k_w4a16_v3<NT, NORM, WARPS, STAGES>, dim3(grid), dim3(WARPS * 32), (size_t)smem, stream, pdl, p);
}

template <int NT, int WARPS, int STAGES>
static bool v3_try(const W4Params& p, bool pdl, cudaStream_t stream, cudaError_t* err) {
    const int G = p.K / kW4GroupK;
    const int ng_max = (G + WARPS - 1) / WARPS;
    const int smem = V3Smem<NT, WARPS, STAGES>::kBytes + WARPS * v3_warp_bytes(p.mc, ng_max);
    if (smem > (WARPS == 8 ? kV3TwoCtaBudget : kV3OneCtaBudget)) return false;
    *err = p.ln_w ? launch_v3_t<NT, true, WARPS, STAGES>(p, smem, pdl, stream)
                  : launch_v3_t<NT, false, WARPS, STAGES>(p, smem, pdl, stream);
    return true;
}

// returns false when the staged activations do not fit shared memory (caller falls back to the fp16 kernels)
bool launch_w4_v3(const W4Params& p, bool pdl, cudaStream_t stream, cudaError_t* err) {
    const bool tall = p.N / 32 <= v3_num_sms() && p.K / kW4GroupK >= 32;
    if (p.mc <= 8) {
        if (tall && v3_try<1, 16, 4>(p, pdl, stream, err)) return true;
        if (v3_try<1, 8, 5>(p, pdl, stream, err)) return true;
        if (v3_try<1, 8, 4>(p, pdl, stream, err)) return true;
        return v3_try<1, 8, 3>(p, pdl, stream, err);
    }
    if (p.mc <= 16) {
        if (tall && v3_try<2, 16, 4>(p, pdl, stream, err)) return true;
        return v3_try<2, 8, 4>(p, pdl, stream, err);
    }
    return false;
}

bool w4_v3_fits(int mc, int N, int K) {
    const int G = K / kW4GroupK;
    auto fits = [&](int fixed, int warps, int budget) {
        const int ng_max = (G + warps - 1) / warps;
        return fixed + warps * v3_warp_bytes(mc, ng_max) <= budget;
    };
    const bool tall = N / 32 <= v3_num_sms() && G >= 32;
    if (mc <= 8) {
        if (tall && fits(V3Smem<1, 16, 4>::kBytes, 16, kV3OneCtaBudget)) return true;
        return fits(V3Smem<1, 8, 3>::kBytes, 8, kV3TwoCtaBudget);
    }
    if (mc <= 16) {
        if (tall && fits(V3Smem<2, 16, 4>::kBytes, 16, kV3OneCtaBudget)) return true;
        return fits(V3Smem<2, 8, 4>::kBytes, 8, kV3TwoCtaBudget);
    }
    return false;
}

cudaError_t prepare_w4_v3() {
    cudaError_t e;
#define ZL_SET(NT, NORM, W, ST)                                                                              \
    e = cudaFuncSetAttribute(k_w4a16_v3<NT, NORM, W, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                             W == 8 ? kV3TwoCtaBudget : kV3OneCtaBudget);                                     \
    if (e != cudaSuccess) return e;
    ZL_SET(1, false, 8, 5) ZL_SET(1, true, 8, 5) ZL_SET(1, false, 16, 4) ZL_SET(1, true, 16, 4)
    ZL_SET(1, false, 8, 4) ZL_SET(1, true, 8, 4) ZL_SET(1, false, 8, 3) ZL_SET(1, true, 8, 3)
    ZL_SET(2, false, 8, 4) ZL_SET(2, true, 8, 4) ZL_SET(2, false, 16, 4) ZL_SET(2, true, 16, 4)
#undef ZL_SET
    return cudaSuccess;
}

}  // namespace zl
