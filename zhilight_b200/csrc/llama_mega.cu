// Persistent whole-model decode kernel: every transformer layer of one decode step in ONE launch.
//
// Why: at batch 1 a Llama-3.1-8B GPTQ step is 160 dependent kernels of 4-15 us.  Each kernel boundary costs a drain,
// a launch/wait and -- worst -- an empty weight pipeline: ~200 KB of shared memory per CTA means a PDL-launched
// successor cannot become resident before its predecessor exits, so HBM idles through every activation-staging
// phase, every tail and the whole attention kernel (profiles/r01_inchain_timeline.txt: ~43 us per layer for 17 us
// worth of weight bytes).  The reference has the same structure with more launches (block.cpp:86-166).
//
// Here one CTA per SM stays resident for the whole step.  Each of its 16 warps owns a private ring of bulk-TMA
// stages and walks a STATIC schedule of 2128-byte ZLW4I weight blocks that spans all GEMMs of all layers
// (weights never depend on activations), so the producer side keeps ~20-25 MB of loads in flight across phase
// boundaries: while the grid barrier, the activation staging or the attention phase of layer l run, the rings
// already fill with the blocks of the next GEMM.  Phases of a layer (grid barrier between them, all through L2):
//   0  qkv  = RMSNorm(h) . Wqkv  + RoPE + KV append     (x staged as block-floating-point ints, IMMA, exact)
//   A  decode attention over the ragged KV buffers     (decode_attn_short.cuh work items, optional split + combine)
//   1  h   += ao . Wo
//   2  act  = SwiGLU(RMSNorm(h) . Wgate_up)
//   3  h   += act . Wdown
// The arithmetic of every phase is the one of k_w4a16_v3 / k_decode_attn_short (same operation order), so the
// results are bit-identical to the kernel-per-op path (tests/test_llama_gpu.py compares them for equality).
#include "common.cuh"
#include "decode_attn_short.cuh"
#include "mega_params.h"
#include "w4_layout.cuh"

namespace zl {

constexpr int kMegaWarps = 16;
constexpr int kMegaThreads = kMegaWarps * 32;
constexpr int kMegaMaxNg = 8;   // staging groups per warp kept in registers (K <= 8*16*128)

enum { kMegaEpiQkv = 0, kMegaEpiResidual = 1, kMegaEpiSwiglu = 2 };

__device__ __forceinline__ void mg_imma_u8s8(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mg_imma_u8u8(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mg_sub_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float ld_cg_half(const __half* p) {
    unsigned short v;
    asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(v) : "l"(p));
    return __half2float(__ushort_as_half(v));
}

// shared memory: [rings][mbarriers][R: GEMM view {red, ss, rstd, staged x} | attention view {logits, red, m/l}]
template <int STAGES>
struct MegaSmem {
    static constexpr int kRingBytes = kMegaWarps * STAGES * kW4BlockBytes;
    static constexpr int kBarOff = kRingBytes;
    static constexpr int kFlagOff = kBarOff + kMegaWarps * STAGES * 8;
    static constexpr int kROff = (kFlagOff + 16 + 127) & ~127;
    static constexpr int kRedFloats = kMegaWarps * 8 * 32;     // one buffer set: 16 warps x (8 tokens x 32 rows)
    static constexpr int kSsOff = kROff + 2 * kRedFloats * 4;  // [16 warps][8]
    static constexpr int kRstdOff = kSsOff + kMegaWarps * 8 * 4;
    static constexpr int kXOff = (kRstdOff + 8 * 4 + 127) & ~127;
};
__host__ __device__ inline int mega_row_bytes(int K) { return K + 16; }
__host__ __device__ inline int mega_stage_bytes(int mc, int K) {
    return 2 * mc * mega_row_bytes(K) + (K / kW4GroupK) * mc * 8;
}

// how warp `warp` of CTA `cta` walks GEMM (N, K): tiles tile0, tile0+stride, ... ; k-groups [g_begin, g_begin+ng)
struct MegaGeom {
    int G, g_begin, ng, tile0, tile_stride, my_tiles, sub, wl, warps;
};
__device__ __forceinline__ MegaGeom mega_geom(int N, int K, int tall, int warp) {
    MegaGeom m;
    m.warps = tall ? 16 : 8;
    const int subs = tall ? 1 : 2;
    m.sub = warp / m.warps;
    m.wl = warp % m.warps;
    m.G = K / kW4GroupK;
    m.g_begin = (m.wl * m.G) / m.warps;
    m.ng = ((m.wl + 1) * m.G) / m.warps - m.g_begin;
    const int n_tiles = N / 32;
    m.tile0 = (int)blockIdx.x * subs + m.sub;
    m.tile_stride = (int)gridDim.x * subs;
    m.my_tiles = m.tile0 < n_tiles ? (n_tiles - m.tile0 + m.tile_stride - 1) / m.tile_stride : 0;
    return m;
}

// Producer side of one warp's ring (used by lane 0 only): the static block schedule over all layers and phases.
// Kept small so that it lives in registers: the refill sits on the consumer warp's critical path.
template <int STAGES>
struct MegaProducer {
    const uint8_t* next;     // address of the next block to request
    long long tile_step;     // bytes from the end of this warp's k-range in one tile to its start in the next tile
    int remaining;           // blocks still to request over the whole step
    int ng;                  // blocks per tile for this warp in the current phase
    int left_in_tile;        // blocks left in the current tile
    int left_tiles;          // tiles left in the current phase after the current one
    int layer, phase, slot;

    __device__ __forceinline__ void seek(const MegaParams& p, int warp) {   // first (layer, phase) >= current with work
        while (layer < p.num_layers) {
            const MegaGeom m = mega_geom(p.gN[phase], p.gK[phase], p.gTall[phase], warp);
            if (m.my_tiles * m.ng > 0) {
                next = p.layers[layer].packed[phase] + ((size_t)m.tile0 * m.G + m.g_begin) * kW4BlockBytes;
                ng = left_in_tile = m.ng;
                left_tiles = m.my_tiles - 1;
                tile_step = ((long long)m.tile_stride * m.G - m.ng) * kW4BlockBytes;
                return;
            }
            if (++phase == 4) {
                phase = 0;
                ++layer;
            }
        }
    }
    __device__ __forceinline__ void init(const MegaParams& p, int warp) {
        layer = phase = slot = 0;
        int per_layer = 0;
        for (int ph = 0; ph < 4; ++ph) {
            const MegaGeom m = mega_geom(p.gN[ph], p.gK[ph], p.gTall[ph], warp);
            per_layer += m.my_tiles * m.ng;
        }
        remaining = per_layer * p.num_layers;
        next = nullptr;
        ng = left_in_tile = left_tiles = 0;
        tile_step = 0;
        if (remaining > 0) seek(p, warp);
    }
    __device__ __forceinline__ void issue(const MegaParams& p, int warp, uint8_t* ring, uint64_t* bars, uint64_t pol) {
        mbar_expect_tx(&bars[slot], kW4BlockBytes);
        bulk_g2s_hint(ring + slot * kW4BlockBytes, next, kW4BlockBytes, &bars[slot], pol);
        if (++slot == STAGES) slot = 0;
        --remaining;
        next += kW4BlockBytes;
        if (--left_in_tile == 0) {
            if (left_tiles > 0) {
                --left_tiles;
                next += tile_step;
                left_in_tile = ng;
            } else {
                if (++phase == 4) {
                    phase = 0;
                    ++layer;
                }
                if (remaining > 0) seek(p, warp);
            }
        }
    }
};

// grid-wide barrier (all CTAs resident: one per SM).  Returns false when the step was aborted (a CTA waited > 2 s:
// something is wrong, e.g. the grid is not co-resident) -- every CTA then leaves the kernel instead of hanging.
__device__ __forceinline__ bool mega_grid_sync(unsigned* sync, unsigned& target, volatile int* s_flag) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(&sync[0], 1u);
        bool ok = true;
        unsigned spins = 0;
        unsigned long long t0 = 0;
        for (;;) {
            unsigned v;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(sync) : "memory");
            if (v >= target) break;
            if ((++spins & 1023u) == 0) {
                unsigned ab;
                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(ab) : "l"(sync + 1) : "memory");
                if (ab) {
                    ok = false;
                    break;
                }
                const unsigned long long now = globaltimer_ns();
                if (t0 == 0) {
                    t0 = now;
                } else if (now - t0 > 2000000000ull) {
                    atomicExch(&sync[1], 1u);
                    ok = false;
                    break;
                }
            }
        }
        *s_flag = ok ? 1 : 0;
    }
    __syncthreads();
    return *s_flag != 0;
}

struct MegaGemmArgs {
    const __half* x;      // (mc, K) activations of this phase (read through L2)
    int ldx;
    int N, K, tall;
    const __half* ln_w;   // fused RMSNorm prologue when non-null
    const __half* bias;
    __half* y;            // RESIDUAL: h (in/out); SWIGLU: act
    const MegaLayer* L;   // QKV: KV pointer tables
    int epi;
    unsigned long long* tr;   // debug timeline of CTA 0 (null: off)
    int* tr_n;
    int tr_id;
};
__device__ __forceinline__ void mega_stamp(const MegaGemmArgs& a, int sub_id) {
    if (a.tr && *a.tr_n < 255) {
        a.tr[1 + 2 * *a.tr_n] = (unsigned long long)(a.tr_id + sub_id);
        a.tr[2 + 2 * *a.tr_n] = globaltimer_ns();
        a.tr[0] = ++*a.tr_n;
    }
}

// One GEMM phase for this CTA: stage x (whole K) as block-floating-point ints, then walk this warp's tiles.
// NORM / EPI are runtime (CTA-uniform) so that the kernel holds ONE copy of this code, called from one site.
template <int STAGES>
__device__ __forceinline__ void mega_gemm_phase(const MegaParams& p, const MegaGemmArgs& a, uint8_t* smem,
                                                MegaProducer<STAGES>& prod, uint64_t pol, int& c_slot,
                                                uint32_t& c_parity) {
    const bool NORM = a.ln_w != nullptr;
    const int EPI = a.epi;
    using S = MegaSmem<STAGES>;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const MegaGeom m = mega_geom(a.N, a.K, a.tall, warp);
    const int WARPS = m.warps, sub = m.sub, wl = m.wl, G = m.G;
    const int stid = threadIdx.x - sub * (WARPS * 32);
    const int mc = p.mc;

    uint8_t* ring = smem + warp * (STAGES * kW4BlockBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBarOff) + warp * STAGES;
    // red: [sub][buffer][warp in sub][8 tokens][32 rows]
    float* red = reinterpret_cast<float*>(smem + S::kROff) + sub * (2 * WARPS * 256);
    float* s_ss = reinterpret_cast<float*>(smem + S::kSsOff);
    float* s_rstd = reinterpret_cast<float*>(smem + S::kRstdOff);
    const int row_b = mega_row_bytes(a.K);
    uint8_t* xs_hi = smem + S::kXOff;
    uint8_t* xs_lo = xs_hi + mc * row_b;
    int2* xs_tab = reinterpret_cast<int2*>(xs_lo + mc * row_b);

    // ---- stage the activations: group gi -> warp gi % 16 (same arithmetic as k_w4a16_v3) ----
    uint2 lnw[kMegaMaxNg];
    if (NORM) {
#pragma unroll
        for (int gl = 0; gl < kMegaMaxNg; ++gl) {
            const int gi = warp + gl * kMegaWarps;
            if (gi < G) lnw[gl] = *reinterpret_cast<const uint2*>(a.ln_w + gi * kW4GroupK + lane * 4);
        }
    }
    for (int tok = 0; tok < mc; ++tok) {
        uint2 raw[kMegaMaxNg];
#pragma unroll
        for (int gl = 0; gl < kMegaMaxNg; ++gl) {
            const int gi = warp + gl * kMegaWarps;
            if (gi < G) raw[gl] = ld_cg_u2(a.x + (size_t)tok * a.ldx + gi * kW4GroupK + lane * 4);
        }
        float sq = 0.f;
#pragma unroll
        for (int gl = 0; gl < kMegaMaxNg; ++gl) {
            const int gi = warp + gl * kMegaWarps;
            if (gi < G) {
                __half2 h01 = *reinterpret_cast<__half2*>(&raw[gl].x), h23 = *reinterpret_cast<__half2*>(&raw[gl].y);
                if (NORM) {
                    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                    sq = fmaf(f01.x, f01.x, sq);
                    sq = fmaf(f01.y, f01.y, sq);
                    sq = fmaf(f23.x, f23.x, sq);
                    sq = fmaf(f23.y, f23.y, sq);
                    h01 = __hmul2(h01, *reinterpret_cast<const __half2*>(&lnw[gl].x));
                    h23 = __hmul2(h23, *reinterpret_cast<const __half2*>(&lnw[gl].y));
                }
                const float2 fa = __half22float2(h01), fb = __half22float2(h23);
                const float amax_l = fmaxf(fmaxf(fabsf(fa.x), fabsf(fa.y)), fmaxf(fabsf(fb.x), fabsf(fb.y)));
                const uint32_t amax_bits = __reduce_max_sync(0xffffffffu, __float_as_uint(amax_l));
                const uint32_t ex = (amax_bits >> 23) & 0xffu;
                const bool zero = ex < 20u || ex == 0xffu;
                const float sg = zero ? 1.0f : __uint_as_float((ex - 14u) << 23);
                const float inv = zero ? 0.0f : __uint_as_float((268u - ex) << 23);
                const int m0 = __float2int_rn(fa.x * inv), m1 = __float2int_rn(fa.y * inv);
                const int m2 = __float2int_rn(fb.x * inv), m3 = __float2int_rn(fb.y * inv);
                const uint32_t lo = __byte_perm(__byte_perm((uint32_t)m0, (uint32_t)m1, 0x0040),
                                                __byte_perm((uint32_t)m2, (uint32_t)m3, 0x0040), 0x5410);
                const uint32_t hi = __byte_perm(__byte_perm((uint32_t)m0, (uint32_t)m1, 0x0051),
                                                __byte_perm((uint32_t)m2, (uint32_t)m3, 0x0051), 0x5410);
                *reinterpret_cast<uint32_t*>(xs_hi + tok * row_b + gi * 128 + lane * 4) = hi;
                *reinterpret_cast<uint32_t*>(xs_lo + tok * row_b + gi * 128 + lane * 4) = lo;
                const int sm = __reduce_add_sync(0xffffffffu, m0 + m1 + m2 + m3);
                if (lane == 0) xs_tab[gi * mc + tok] = make_int2(sm, (int)__float_as_uint(sg));
            }
        }
        if (NORM) {
            sq = warp_sum(sq);
            if (lane == 0) s_ss[warp * 8 + tok] = sq;
        }
    }
    __syncthreads();
    if (NORM) {
        if (threadIdx.x < mc) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < kMegaWarps; ++w) v += s_ss[w * 8 + threadIdx.x];
            s_rstd[threadIdx.x] = rsqrtf(v / (float)a.K + p.eps);
        }
        __syncthreads();
    }

    mega_stamp(a, 100);   // staged
    // ---- tiles ----
    for (int ti = 0; ti < m.my_tiles; ++ti) {
        const int st = m.tile0 + ti * m.tile_stride;
        float acc[2][4];
#pragma unroll
        for (int x = 0; x < 2; ++x)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[x][c] = 0.f;

        for (int i = 0; i < m.ng; ++i) {
            const int s = c_slot;
            const uint32_t parity = c_parity;
            if (++c_slot == STAGES) {
                c_slot = 0;
                c_parity ^= 1u;
            }
            const int gi = m.g_begin + i;
            uint4 bh[2], bl[2];
            int2 tab[2];
            if (g < mc) {
                const uint8_t* ph = xs_hi + g * row_b + gi * 128 + t * 32;
                const uint8_t* pl = xs_lo + g * row_b + gi * 128 + t * 32;
                bh[0] = *reinterpret_cast<const uint4*>(ph);
                bh[1] = *reinterpret_cast<const uint4*>(ph + 16);
                bl[0] = *reinterpret_cast<const uint4*>(pl);
                bl[1] = *reinterpret_cast<const uint4*>(pl + 16);
            } else {
                bh[0] = bh[1] = bl[0] = bl[1] = make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int tc = 2 * t + e;
                tab[e] = tc < mc ? xs_tab[gi * mc + tc] : make_int2(0, 0);
            }
            mbar_wait(&bars[s], parity);
            const uint8_t* blk = ring + s * kW4BlockBytes;
            int ah[2][4], al[2][4];
#pragma unroll
            for (int tt = 0; tt < 2; ++tt)
#pragma unroll
                for (int c = 0; c < 4; ++c) ah[tt][c] = al[tt][c] = 0;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                uint4 wv[2];
#pragma unroll
                for (int tt = 0; tt < 2; ++tt)
                    wv[tt] = *reinterpret_cast<const uint4*>(blk + ((tt * 2 + hh) * 32 + lane) * 16);
#pragma unroll
                for (int jp = 0; jp < 2; ++jp) {
                    const int j = hh * 2 + jp;
#pragma unroll
                    for (int tt = 0; tt < 2; ++tt) {
                        const uint32_t w0 = jp ? wv[tt].z : wv[tt].x;
                        const uint32_t w1 = jp ? wv[tt].w : wv[tt].y;
                        const uint32_t af[4] = {w0 & 0x0f0f0f0fu, w0 & 0xf0f0f0f0u, w1 & 0x0f0f0f0fu,
                                                w1 & 0xf0f0f0f0u};
                        const uint4 vh = bh[j >> 1], vl = bl[j >> 1];
                        mg_imma_u8s8(ah[tt], af, (j & 1) ? vh.z : vh.x, (j & 1) ? vh.w : vh.y);
                        mg_imma_u8u8(al[tt], af, (j & 1) ? vl.z : vl.x, (j & 1) ? vl.w : vl.y);
                    }
                }
            }
#pragma unroll
            for (int tt = 0; tt < 2; ++tt) {
                const __half2 sc = *reinterpret_cast<const __half2*>(blk + kW4ScaleOff + (tt * 8 + g) * 4);
                const int zz = blk[kW4ZeroOff + tt * 8 + g];
                const int z_lo = zz & 0xF, z_hi16 = zz & 0xF0;
                const float s_lo = __low2float(sc), s_hi = __high2float(sc) * 0.0625f;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int2 tb = tab[c & 1];
                    const int zc = (c >> 1) ? z_hi16 : z_lo;
                    const int r = ah[tt][c] * 256 + al[tt][c] - zc * tb.x;
                    const float f = ((c >> 1) ? s_hi : s_lo) * __int_as_float(tb.y);
                    acc[tt][c] = fmaf((float)r, f, acc[tt][c]);
                }
            }
            __syncwarp();
            if (lane == 0 && prod.remaining > 0) prod.issue(p, warp, ring, bars, pol);   // refills the slot just drained
        }

        mega_stamp(a, 200 + ti);   // this warp's blocks of tile ti consumed
        // ---- split-k reduction across the warps of this sub-CTA + epilogue ----
        float* rbuf = red + (ti & 1) * (WARPS * 256);
        float* myred = rbuf + wl * 256;
#pragma unroll
        for (int tt = 0; tt < 2; ++tt) {
            const int tok = 2 * t;
            const int row = tt * 16 + g;
            myred[tok * 32 + row] = acc[tt][0];
            myred[(tok + 1) * 32 + row] = acc[tt][1];
            myred[tok * 32 + row + 8] = acc[tt][2];
            myred[(tok + 1) * 32 + row + 8] = acc[tt][3];
        }
        mg_sub_barrier(1 + sub, WARPS * 32);

        auto sum_red = [&](int idx) {
            float v = 0.f;
            for (int w = 0; w < WARPS; ++w) v += rbuf[w * 256 + idx];
            return v;
        };
        const int n0 = st * 32;
        const int sub_threads = WARPS * 32;
        if (EPI == kMegaEpiSwiglu) {
            const int n_out = a.N / 2;
            for (int e = stid; e < mc * 16; e += sub_threads) {
                const int tok = e >> 4, oc = e & 15;
                const int rg = (oc >> 3) * 16 + (oc & 7);
                float gate = sum_red(tok * 32 + rg);
                float up = sum_red(tok * 32 + rg + 8);
                if (NORM) {
                    gate *= s_rstd[tok];
                    up *= s_rstd[tok];
                }
                if (a.bias) {
                    gate += __half2float(a.bias[n0 + rg]);
                    up += __half2float(a.bias[n0 + rg + 8]);
                }
                const float gr = __half2float(__float2half_rn(gate));
                const float ur = __half2float(__float2half_rn(up));
                a.y[(size_t)tok * n_out + st * 16 + oc] = __float2half_rn(silu_f(gr) * ur);
            }
        } else if (EPI == kMegaEpiQkv) {
            const int d = p.dim_head, half_dim = d / 2;
            const int tiles_per_head = d / 32;
            const int head = st / tiles_per_head, jt = st % tiles_per_head;
            for (int e = stid; e < mc * 16; e += sub_threads) {
                const int tok = e >> 4, oc = e & 15;
                const int rlo = (oc >> 3) * 16 + (oc & 7);
                const int c = jt * 16 + oc;
                float lo = sum_red(tok * 32 + rlo), hi = sum_red(tok * 32 + rlo + 8);
                if (NORM) {
                    lo *= s_rstd[tok];
                    hi *= s_rstd[tok];
                }
                if (a.bias) {
                    lo += __half2float(a.bias[n0 + rlo]);
                    hi += __half2float(a.bias[n0 + rlo + 8]);
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
                    __half* dst = p.q + ((size_t)tok * p.num_heads + head) * d;
                    dst[c] = olo;
                    dst[c + half_dim] = ohi;
                } else {
                    const int pl = p.placement[tok];
                    if (pl >= 0) {
                        const bool is_k = !is_v;
                        const int hk = is_k ? head - p.num_heads : head - p.num_heads - p.num_kv_heads;
                        __half* kv = (is_k ? a.L->k_addrs : a.L->v_addrs)[p.token_batch[tok]];
                        __half* dst = kv + ((size_t)pl * p.num_kv_heads + hk) * d;
                        dst[c] = olo;
                        dst[c + half_dim] = ohi;
                    }
                }
            }
        } else {
            for (int e = stid; e < mc * 32; e += sub_threads) {
                const int tok = e >> 5, row = e & 31;
                float v = sum_red(tok * 32 + row);
                if (NORM) v *= s_rstd[tok];
                if (a.bias) v += __half2float(a.bias[n0 + row]);
                __half hv = __float2half_rn(v);
                // residual stream: read through L2 (another SM may have written it in an earlier phase)
                hv = __float2half_rn(__half2float(hv) + ld_cg_half(a.y + (size_t)tok * a.N + n0 + row));
                a.y[(size_t)tok * a.N + n0 + row] = hv;
            }
        }
    }
}

template <int D>
__device__ __forceinline__ void mega_attention(const MegaParams& p, const MegaLayer* L, float* s_dyn) {
    const int m_query = p.num_heads / p.num_kv_heads;
    const int hgroups = (m_query + 7) / 8;
    const int hgs = p.num_kv_heads * hgroups;
    const int items = p.mc * hgs * p.attn_splits;
    for (int it = (int)blockIdx.x; it < items; it += (int)gridDim.x) {
        const int split = it % p.attn_splits;
        const int hg = (it / p.attn_splits) % hgs;
        const int bq = it / (p.attn_splits * hgs);
        attn_short_item<__half, D>(p.q, p.buf_lens, L->k_addrs, L->v_addrs, nullptr, p.attn_scale, p.ao, p.part_o,
                                   p.part_m, p.part_l, 1, p.num_heads, p.num_kv_heads, m_query, p.attn_splits, 1, split,
                                   hg, bq, s_dyn);
        __syncthreads();
    }
}

// LSE merge of the split partials (same arithmetic as k_attn_combine)
__device__ __forceinline__ void mega_attn_combine(const MegaParams& p) {
    const int D = p.dim_head, per = kMegaThreads / D;
    const int vheads = p.mc * p.num_heads, ns = p.attn_splits;
    const int lane_v = threadIdx.x / D, d = threadIdx.x % D;
    for (int vh0 = (int)blockIdx.x * per; vh0 < vheads; vh0 += (int)gridDim.x * per) {
        const int vh = vh0 + lane_v;
        if (vh >= vheads) continue;
        const float* pm = p.part_m + (size_t)vh * ns;
        const float* pl = p.part_l + (size_t)vh * ns;
        float gm = -1e20f;
        for (int i = 0; i < ns; ++i) gm = fmaxf(gm, __ldcg(pm + i));
        float gs = 0.f;
        for (int i = 0; i < ns; ++i) gs += __ldcg(pl + i) * expf(__ldcg(pm + i) - gm);
        float res = 0.f;
        for (int i = 0; i < ns; ++i) {
            const float w = __ldcg(pl + i) / gs * expf(__ldcg(pm + i) - gm);
            res += __ldcg(p.part_o + ((size_t)vh * ns + i) * D + d) * w;
        }
        p.ao[(size_t)vh * D + d] = __float2half_rn(res);
    }
}

template <int STAGES>
__global__ void __launch_bounds__(kMegaThreads, 1) k_llama_mega(const MegaParams p) {
    using S = MegaSmem<STAGES>;
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* ring = smem + warp * (STAGES * kW4BlockBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBarOff) + warp * STAGES;
    volatile int* s_flag = reinterpret_cast<volatile int*>(smem + S::kFlagOff);
    float* s_attn = reinterpret_cast<float*>(smem + S::kROff);

    // debug timeline (CTA 0): id 1000*layer + {1: entry/wait, 10+ph: phase ph computed, 20+ph: barrier passed,
    // 30: attention computed, 31: barrier, 32: combine + barrier}
    unsigned long long* tr = (p.trace && blockIdx.x == 0 && threadIdx.x == 0) ? p.trace : nullptr;
    int tr_n = 0;
    auto stamp = [&](int id) {
        if (tr && tr_n < 255) {
            tr[1 + 2 * tr_n] = (unsigned long long)id;
            tr[2 + 2 * tr_n] = globaltimer_ns();
            tr[0] = ++tr_n;
        }
    };
    stamp(0);
    pdl_trigger();
    const uint64_t pol = l2_evict_first_policy();
    MegaProducer<STAGES> prod;
    prod.remaining = 0;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
        prod.init(p, warp);
#pragma unroll
        for (int s = 0; s < STAGES; ++s)
            if (prod.remaining > 0) prod.issue(p, warp, ring, bars, pol);
    }
    __syncwarp();
    pdl_wait();   // everything above touched only weights; activations / KV / tables come from predecessor kernels
    stamp(1);

    int c_slot = 0;
    uint32_t c_parity = 0;
    unsigned target = 0;
    const int hd = p.num_heads * p.dim_head;
    for (int l = 0; l < p.num_layers; ++l) {
        const MegaLayer* L = p.layers + l;
        for (int ph = 0; ph < 4; ++ph) {
            MegaGemmArgs a;
            a.L = L;
            a.N = p.gN[ph];
            a.K = p.gK[ph];
            a.tall = p.gTall[ph];
            a.bias = L->bias[ph];
            if (ph == 0) {          // qkv = RMSNorm(h) . Wqkv, RoPE, KV append
                a.x = p.h; a.ldx = a.K; a.ln_w = L->ln_attn; a.y = nullptr; a.epi = kMegaEpiQkv;
            } else if (ph == 1) {   // h += ao . Wo
                a.x = p.ao; a.ldx = hd; a.ln_w = nullptr; a.y = p.h; a.epi = kMegaEpiResidual;
            } else if (ph == 2) {   // act = SwiGLU(RMSNorm(h) . Wgu)
                a.x = p.h; a.ldx = a.K; a.ln_w = L->ln_ff; a.y = p.act; a.epi = kMegaEpiSwiglu;
            } else {                // h += act . Wdown
                a.x = p.act; a.ldx = a.K; a.ln_w = nullptr; a.y = p.h; a.epi = kMegaEpiResidual;
            }
            a.tr = (l < 2) ? tr : nullptr;
            a.tr_n = &tr_n;
            a.tr_id = 1000 * l + 10000 * (ph + 1);
            mega_gemm_phase<STAGES>(p, a, smem, prod, pol, c_slot, c_parity);
            stamp(1000 * l + 10 + ph);
            if (!mega_grid_sync(p.sync, target, s_flag)) return;
            stamp(1000 * l + 20 + ph);
            if (ph == 0) {
                if (p.dim_head == 128) mega_attention<128>(p, L, s_attn);
                else mega_attention<64>(p, L, s_attn);
                stamp(1000 * l + 30);
                if (!mega_grid_sync(p.sync, target, s_flag)) return;
                stamp(1000 * l + 31);
                if (p.attn_splits > 1) {
                    mega_attn_combine(p);
                    if (!mega_grid_sync(p.sync, target, s_flag)) return;
                    stamp(1000 * l + 32);
                }
            }
        }
    }
}

size_t mega_smem_bytes(int mc, int k_max, int dim_head, int stages) {
    const int x_off = stages == 4 ? MegaSmem<4>::kXOff : MegaSmem<3>::kXOff;
    const int r_off = stages == 4 ? MegaSmem<4>::kROff : MegaSmem<3>::kROff;
    const size_t gemm = (size_t)x_off + mega_stage_bytes(mc, k_max);
    const size_t attn = (size_t)r_off + (dim_head == 128 ? short_smem_bytes<128>() : short_smem_bytes<64>());
    return gemm > attn ? gemm : attn;
}

constexpr size_t kMegaSmemMax = 232448;   // 227 KB opt-in limit per CTA on sm_100

cudaError_t prepare_llama_mega() {
    cudaError_t e = cudaFuncSetAttribute(k_llama_mega<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMegaSmemMax);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_llama_mega<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMegaSmemMax);
}

cudaError_t launch_llama_mega(const MegaParams& p, int stages, bool pdl, cudaStream_t stream) {
    int k_max = 0;
    for (int i = 0; i < 4; ++i) {
        k_max = p.gK[i] > k_max ? p.gK[i] : k_max;
        if (p.gK[i] % kW4GroupK != 0 || p.gN[i] % 32 != 0 || p.gK[i] / kW4GroupK > kMegaMaxNg * kMegaWarps)
            return cudaErrorInvalidValue;
    }
    if (p.mc < 1 || p.mc > 8 || (p.dim_head != 64 && p.dim_head != 128) || (stages != 3 && stages != 4))
        return cudaErrorInvalidValue;
    const size_t smem = mega_smem_bytes(p.mc, k_max, p.dim_head, stages);
    if (smem > kMegaSmemMax) return cudaErrorInvalidValue;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) return cudaErrorInvalidValue;
    if (stages == 4) return launch(k_llama_mega<4>, dim3(sms), dim3(kMegaThreads), smem, stream, pdl, p);
    return launch(k_llama_mega<3>, dim3(sms), dim3(kMegaThreads), smem, stream, pdl, p);
}

}  // namespace zl
