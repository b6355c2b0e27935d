// gru_tc.cuh — tensor-core (tcgen05, 3xTF32) version of the prefix-reuse chain kernel for the FP32 GRU pRNN.
//
// Same contract as gru_chain_kernel (kind 0: TFIM single flips, 1DTFIM/TrainingRNN_1DTFIM.py:43-48,74): every
// work item (slot s, 128-row tile) restarts from the stashed base state after site s and re-runs sites s+1..N-1,
// accumulating delta = log P(sigma') - log P(sigma) as per-site differences.
//
// Mapping to the hardware
//   * one persistent CTA per SM, 128 "row" threads (thread = configuration = TMEM lane) + one MMA/alloc warp;
//   * the per-site contraction [128 x (x|h)] x [(x|h) x (r|u|c)] runs on tcgen05.mma kind::tf32 with M=128:
//       A (hidden states, split hi/lo) lives in TMEM, written by the row threads with tcgen05.st;
//       B (weights, split hi/lo, K-major no-swizzle core-matrix layout) lives in shared memory, brought in with
//       1-D bulk async copies (TMA) from a pre-packed image;
//       D (gate pre-activations, FP32) lives in TMEM and is read back with tcgen05.ld for the gate math;
//     three passes hi*hi + lo*hi + hi*lo give FP32-grade products (the 1e-5 parity gate rules out plain TF32);
//   * only ONE layer's weights fit in shared memory together with hi/lo splits (172 KB), so sites are processed in
//     blocks of T: layer 0 over the T sites, then layer 1, then layer 2 (teacher forcing makes this legal); the
//     inter-layer activations of the block go through a small per-CTA global scratch that stays in L2.
//   TMEM columns:  D: r [0,64) u [64,128) cx [128,192) ch [192,256) | X_hi [256,..) X_lo | H_hi H_lo  (4 * Kp <= 256)
// Included at the end of gru.cu.
#pragma once
#include "gru_kernels.cuh"
#include "host_util.cuh"
#include "umma.cuh"

namespace rnnwf {

constexpr int kTcRows = 128;      // rows per tile = UMMA M = TMEM lanes
constexpr int kTcBlk = 64;        // column block per gate in D / row block per gate in B
constexpr int kTcT = 16;           // sites per layer block (inter-layer scratch = T * 128 * 52 * 4 B per CTA, kept small to stay in L2)
constexpr int kTcThreads = 288;   // 8 row warps + 1 MMA warp

struct TcLayout {
    int L, H, Kp, KC, N;
    int s1, s2, sx;          // floats of Bh1 (128 x Kp), Bh2 (64 x Kp), Bx (192 x Kp)
    int img_floats;          // per layer: Bh1_hi | Bh2_hi | Bh1_lo | Bh2_lo | Bx_hi | Bx_lo
    int tab_floats;          // head: Wd[64][2] | bd[2]
};

inline TcLayout make_tc_layout(const GruLayout& g) {
    TcLayout t;
    t.L = g.L; t.H = g.H; t.N = g.N;
    t.Kp = (g.H + 8) & ~7;   // at least one spare K column: column H of the A operands is the constant 1 that carries the biases
    t.KC = t.Kp / 4;
    t.s1 = 2 * kTcBlk * t.Kp; t.s2 = kTcBlk * t.Kp; t.sx = 3 * kTcBlk * t.Kp;
    t.img_floats = 2 * (t.s1 + t.s2 + t.sx);
    t.tab_floats = 2 * kTcBlk + 4;
    return t;
}

__host__ __device__ __forceinline__ int tc_core_off(int n, int k, int KC) {   // float offset inside a K-major core-matrix image
    return (n >> 3) * (KC * 32) + (k >> 2) * 32 + (n & 7) * 4 + (k & 3);
}

// flat TF-order parameters -> per-layer B images (hi/lo split) + head table.
// The gate non-linearities are evaluated as 1/(1 + 2^a), so the images carry the pre-scaled weights
//   r, u rows: -log2(e) * W        (sigmoid(x) = 1 / (1 + 2^(-x log2 e)))
//   c   rows:  2 log2(e) * W       (tanh(x)    = 1 - 2 / (1 + 2^(2 x log2 e)))
// and K column H (multiplied by the constant-1 column of the A operands) carries the biases:
//   Bh1[:, H] = bg,  Bh2[:, H] = bch,  Bx[c rows, H] = bci.   Layer 0 (one-hot input) uses Bx rows k = 0, 1.
__global__ void pack_gru_tc_kernel(GruLayout g, TcLayout t, const float* __restrict__ flat, float* __restrict__ img,
                                   float* __restrict__ tab) {
    const int H = g.H, Kp = t.Kp, KC = t.KC;
    const float kS = -1.4426950408889634f, kC = 2.8853900817779268f;
    const int per = t.s1 + t.s2 + t.sx;   // one precision half
    const int total = g.L * per;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int l = idx / per;
        int q = idx % per;
        const int d = g.d[l];
        const float* Kg = flat + g.flat_off[l];
        const float* bg = Kg + (d + H) * 2 * H;
        const float* Kci = bg + 2 * H;
        const float* Kch = Kci + d * H;
        const float* bci = Kch + H * H;
        const float* bch = bci + H;
        float v = 0.f;
        int dst;   // offset of the hi value inside the layer image
        bool xpart = false;
        if (q < t.s1) {                       // Bh1: rows [r (64) | u (64)], K over h (+ bias column)
            const int n = q / Kp, k = q % Kp, gate = n / kTcBlk, j = n % kTcBlk;
            if (j < H) {
                if (k < H) v = kS * Kg[(d + k) * 2 * H + gate * H + j];
                else if (k == H) v = kS * bg[gate * H + j];
            }
            dst = tc_core_off(n, k, KC);
        } else if (q < t.s1 + t.s2) {         // Bh2: rows [ch (64)], K over h (+ bias column)
            q -= t.s1;
            const int n = q / Kp, k = q % Kp;
            if (n < H) {
                if (k < H) v = kC * Kch[k * H + n];
                else if (k == H) v = kC * bch[n];
            }
            dst = t.s1 + tc_core_off(n, k, KC);
        } else {                              // Bx: rows [r | u | cx], K over x (+ bias column for cx)
            q -= t.s1 + t.s2;
            const int n = q / Kp, k = q % Kp, gate = n / kTcBlk, j = n % kTcBlk;
            if (j < H) {
                if (k < d) v = gate < 2 ? kS * Kg[k * 2 * H + gate * H + j] : kC * Kci[k * H + j];
                else if (k == H && gate == 2) v = kC * bci[j];
            }
            dst = 2 * (t.s1 + t.s2) + tc_core_off(n, k, KC);
            xpart = true;
        }
        uint32_t hb;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
        const float hi = __uint_as_float(hb);
        float* L0 = img + (size_t)l * t.img_floats;
        L0[dst] = hi;
        L0[dst + (xpart ? t.sx : t.s1 + t.s2)] = v - hi;     // the tensor core reads the upper 19 bits of the remainder
    }
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < t.tab_floats; idx += gridDim.x * blockDim.x) {
        float v = 0.f;                        // head: Wd[j][2] (64 x 2) | bd[2]
        if (idx < 2 * kTcBlk) { if (idx / 2 < H) v = flat[g.flat_head + idx]; }
        else if (idx < 2 * kTcBlk + 2) v = flat[g.flat_head + 2 * H + (idx - 2 * kTcBlk)];
        tab[idx] = v;
    }
}

constexpr int kTcRowThreads = 256;   // 2 threads per row: the row's units are split between warp w and warp w + 4 (same TMEM lanes)
__device__ __forceinline__ void tc_named_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kTcThreads) : "memory"); }
__device__ __forceinline__ void tc_row_sync() { asm volatile("bar.sync 2, %0;" ::"n"(kTcRowThreads) : "memory"); }

__device__ __forceinline__ float tc_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float tc_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// unit range of a row thread: part 0 owns units [0, U0), part 1 owns [U0, H); U0 is a multiple of 8 (TMEM store windows)
template <int H, int PART> struct TcPart {
    static constexpr int S = ((H / 2) / 8) * 8;
    static constexpr int U0 = PART ? S : 0;
    static constexpr int U = PART ? H - S : S;
    static constexpr int NG = (U + 7) / 8;
    static constexpr int NV = (U + 3) / 4;
};

// hi/lo split of 8 consecutive units (group GQ of the thread's range) -> TMEM operand columns
template <int H, int PART, int GQ>
__device__ __forceinline__ void tc_stage_group(uint32_t lane_addr, uint32_t colHi, uint32_t colLo, const float* v) {
    using P = TcPart<H, PART>;
    if constexpr (GQ < P::NG) {
        constexpr int CNT = P::U - 8 * GQ >= 8 ? 8 : P::U - 8 * GQ;
        float hi[8], lo[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            hi[q] = 0.f; lo[q] = 0.f;
            if (q < CNT) umma::split_tf32_fast(v[8 * GQ + q], hi[q], lo[q]);
        }
        umma::tmem_st_n<CNT>(lane_addr + colHi + P::U0 + 8 * GQ, hi);
        umma::tmem_st_n<CNT>(lane_addr + colLo + P::U0 + 8 * GQ, lo);
    }
}
template <int H, int PART>
__device__ __forceinline__ void tc_stage_g(int gq, uint32_t lane_addr, uint32_t colHi, uint32_t colLo, const float* v) {
    if (gq == 0) tc_stage_group<H, PART, 0>(lane_addr, colHi, colLo, v);
    else if (gq == 1) tc_stage_group<H, PART, 1>(lane_addr, colHi, colLo, v);
    else if (gq == 2) tc_stage_group<H, PART, 2>(lane_addr, colHi, colLo, v);
    else tc_stage_group<H, PART, 3>(lane_addr, colHi, colLo, v);
}
template <int H, int PART>
__device__ __forceinline__ void tc_stage(uint32_t lane_addr, uint32_t colHi, uint32_t colLo, const float* v) {
    tc_stage_group<H, PART, 0>(lane_addr, colHi, colLo, v);
    tc_stage_group<H, PART, 1>(lane_addr, colHi, colLo, v);
    tc_stage_group<H, PART, 2>(lane_addr, colHi, colLo, v);
    tc_stage_group<H, PART, 3>(lane_addr, colHi, colLo, v);
}
// the thread's units of one row of a [rows][HP] scratch (float4 granularity; HP = H rounded up to 4)
template <int H, int PART> __device__ __forceinline__ void tc_load_row(const float* __restrict__ row, float* v) {
    using P = TcPart<H, PART>;
    const float4* src = reinterpret_cast<const float4*>(row + P::U0);
#pragma unroll
    for (int i = 0; i < P::NV; ++i) {
        const float4 a = src[i];
        v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = a.z; v[4 * i + 3] = a.w;
    }
}
template <int H, int PART> __device__ __forceinline__ void tc_store_row(float* __restrict__ row, const float* v) {
    using P = TcPart<H, PART>;
    float4* dst = reinterpret_cast<float4*>(row + P::U0);
#pragma unroll
    for (int i = 0; i < P::NV; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

struct TcArgs {
    GruLayout g;
    TcLayout t;
    int Mold, tiles128;
    int64_t rows_total;
    const float *img, *tabg;
    const uint8_t* sigT;
    float* hstore;          // BASE: written (every layer, every site); FLIP: restart states
    double *la_sel, *la_oth; // BASE: written; FLIP: read
    double* lp;             // BASE: sum_n la_sel
    float *xbuf, *hsave;
    double* delta;          // FLIP: [tile][slot][M]
    int* counter;
    long long* dbg;         // optional [grid][4] cycle counters: {mma wait, gate math, staging/barrier, steps} of row thread 0
};

// all row-thread work of one (block, layer): restore the state, then nb sites of {stage operands, wait for the MMAs, gate math}
template <int H, int PART, bool BASE>
__device__ __forceinline__ void tc_row_block(const TcArgs& a, const float* tab, float2* zsm, uint64_t* bars, uint32_t& par_mma, uint32_t lane_addr,
                                             int rowi, bool live, size_t rowbase, int m, int s, int b0, int nb, int l, bool first_block,
                                             float* xb, float* hs, double& acc) {
    using P = TcPart<H, PART>;
    constexpr int Kp = (H + 8) & ~7, HP = ((H + 3) / 4) * 4, UP = P::NG * 8;
    constexpr uint32_t colD = 0, colXH = 256, colXL = 256 + Kp, colHH = 256 + 2 * Kp, colHL = 256 + 3 * Kp;
    const int L = a.g.L, N = a.g.N, Mold = a.Mold;
    const bool top = l == L - 1;
    float hp[UP], xr[UP];
#pragma unroll
    for (int j = 0; j < UP; ++j) { hp[j] = 0.f; xr[j] = 0.f; }
    if (live) {
        if (first_block) {
            if (!BASE) {
                const float* src = a.hstore + ((rowbase + s) * L + l) * (size_t)H * Mold + m;
#pragma unroll
                for (int j = 0; j < P::U; ++j) hp[j] = src[(size_t)(P::U0 + j) * Mold];
            }
        } else {
            tc_load_row<H, PART>(hs + ((size_t)l * kTcRows + rowi) * HP, hp);
        }
    }
    tc_stage<H, PART>(lane_addr, colHH, colHL, hp);
    // the input of the first site of the block: layer 0 -> one-hot code of the previous spin, layers >= 1 -> h^(l-1)
    int code = 2;
    auto fetch_code = [&](int n) {
        int c = 2;
        if (live && n > 0) {
            c = a.sigT[(rowbase + n - 1) * Mold + m];
            if (!BASE && n - 1 == s) c = 1 - c;
        }
        return c;
    };
    if (l == 0) {
        if (PART == 0) code = fetch_code(b0);
        else {   // columns of the X operand that belong to h^(l-1) units must read as 0 while layer 0 runs
            const float z2[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            umma::tmem_st_n<(H & 7)>(lane_addr + colXH + (H & ~7), z2);
        }
    } else {
        tc_load_row<H, PART>(xb + ((size_t)0 * kTcRows + rowi) * HP, xr);
    }
    long long tq3 = a.dbg ? clock64() : 0;
    for (int tt = 0; tt < nb; ++tt) {
        const int n = b0 + tt;
        if (l > 0) {
            if (tt == 0) tc_stage<H, PART>(lane_addr, colXH, colXL, xr);   // later sites: staged during the previous site's gate math
        } else if (PART == 0) {
            const float oh[8] = {code == 0 ? 1.f : 0.f, code == 1 ? 1.f : 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            umma::tmem_st8(lane_addr + colXH, oh);
        }
        const long long tq4 = a.dbg ? clock64() : 0;
        umma::wait_st();
        const long long tq5 = a.dbg ? clock64() : 0;
        umma::fence_before_sync();
        tc_named_sync();
        const long long tq0 = a.dbg ? clock64() : 0;
        // ---- the MMA warp issues this site's MMAs now; fetch what the next steps need meanwhile ----
        int sg = 0;
        double lsel = 0.0;
        if (top && PART == 0 && live) {
            sg = a.sigT[(rowbase + n) * Mold + m];
            if (!BASE) lsel = a.la_sel[(rowbase + n) * Mold + m];
        }
        if (tt + 1 < nb) {
            if (l > 0) tc_load_row<H, PART>(xb + ((size_t)(tt + 1) * kTcRows + rowi) * HP, xr);
            else if (PART == 0) code = fetch_code(n + 1);
        }
        umma::mbar_wait(&bars[0], par_mma);
        par_mma ^= 1;
        umma::fence_after_sync();
        const long long tq1 = a.dbg ? clock64() : 0;
        // ---- gate math.  D holds a_r = -log2e * pre_r, a_u likewise, a_cx, a_ch = 2 log2e * (candidate parts), biases included ----
        float z0 = 0.f, z1 = 0.f;
#pragma unroll
        for (int gq = 0; gq < P::NG; ++gq) {
            const int cnt = P::U - 8 * gq >= 8 ? 8 : P::U - 8 * gq;
            const uint32_t col = lane_addr + colD + P::U0 + 8 * gq;
            float dr[8], du[8], dc[8], dq[8];
            umma::tmem_ld8(col, dr);
            umma::tmem_ld8(col + kTcBlk, du);
            umma::tmem_ld8(col + 2 * kTcBlk, dc);
            umma::tmem_ld8(col + 3 * kTcBlk, dq);
            umma::wait_ld();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (q < cnt) {
                    const int jl = 8 * gq + q, j = P::U0 + jl;
                    const float er = 1.0f + tc_ex2(fminf(dr[q], 60.f)), eu = 1.0f + tc_ex2(fminf(du[q], 60.f));
                    const float inv = tc_rcp(er * eu);                       // one reciprocal for both gates
                    const float r = inv * eu, u = inv * er;
                    const float ec = 1.0f + tc_ex2(fmaf(r, dq[q], dc[q]));
                    const float cc = fmaf(-2.0f, tc_rcp(ec), 1.0f);
                    const float h = fmaf(u, hp[jl] - cc, cc);
                    hp[jl] = h;
                    if (top) {
                        z0 = fmaf(h, tab[2 * j], z0);
                        z1 = fmaf(h, tab[2 * j + 1], z1);
                    }
                    if (BASE && live) a.hstore[(((rowbase + n) * L + l) * (size_t)H + j) * Mold + m] = h;
                }
            }
            // operands of the next MMAs for this group of units (the MMAs of this site are complete: both regions are free),
            // interleaved with the MUFU-bound gate math of the following group
            tc_stage_g<H, PART>(gq, lane_addr, colHH, colHL, hp);
            if (l > 0 && tt + 1 < nb) tc_stage_g<H, PART>(gq, lane_addr, colXH, colXL, xr);
            if (!top) {
                float4* dst = reinterpret_cast<float4*>(xb + ((size_t)tt * kTcRows + rowi) * HP + P::U0 + 8 * gq);
                dst[0] = make_float4(hp[8 * gq], hp[8 * gq + 1], hp[8 * gq + 2], hp[8 * gq + 3]);
                if (cnt > 4) dst[1] = make_float4(hp[8 * gq + 4], hp[8 * gq + 5], hp[8 * gq + 6], hp[8 * gq + 7]);
            }
        }
        const long long tq2 = a.dbg ? clock64() : 0;
        if (top) {
            if (PART == 1) zsm[rowi] = make_float2(z0, z1);
            tc_row_sync();
            if (PART == 0 && live) {
                const float2 o = zsm[rowi];
                const float f0 = z0 + o.x + tab[2 * kTcBlk], f1 = z1 + o.y + tab[2 * kTcBlk + 1];
                // log softmax of the 2-way head in FP32 (log1pf/expf, ~1e-7 relative); the site terms are summed in FP64
                const float dsel = sg ? f0 - f1 : f1 - f0;                         // z_other - z_selected
                const double ls = dsel > 30.f ? -(double)dsel : -(double)log1pf(expf(dsel));
                const double lo_ = -dsel > 30.f ? (double)dsel : -(double)log1pf(expf(-dsel));
                if (BASE) {
                    a.la_sel[(rowbase + n) * Mold + m] = ls;
                    a.la_oth[(rowbase + n) * Mold + m] = lo_;
                    acc += ls;
                } else {
                    acc += ls - lsel;
                }
            }
        }
        if (a.dbg && rowi == 0) {
            long long* d = a.dbg + 16 * blockIdx.x + 8 * PART;
            d[4] += tq4 - tq3; d[5] += tq5 - tq4; d[6] += tq0 - tq5;
            tq3 = clock64();
            d[0] += tq1 - tq0; d[1] += tq2 - tq1; d[2] += tq3 - tq2; d[3] += 1;
        }
    }
    if (b0 + nb < N) tc_store_row<H, PART>(hs + ((size_t)l * kTcRows + rowi) * HP, hp);   // park h^l for the next block
}

template <int H, bool BASE>
__global__ void __launch_bounds__(kTcThreads, 1) gru_chain_tc_kernel(const __grid_constant__ TcArgs a) {
    constexpr int Kp = (H + 8) & ~7, KC = Kp / 4, HP = ((H + 3) / 4) * 4;
    constexpr uint32_t colD = 0, colXH = 256, colXL = 256 + Kp, colHH = 256 + 2 * Kp, colHL = 256 + 3 * Kp;
    static_assert(256 + 4 * Kp <= 512 && (H & 7) != 0, "TMEM budget / spare K column");
    extern __shared__ __align__(128) unsigned char smem_tc[];
    const TcLayout& t = a.t;
    float* Bsm = reinterpret_cast<float*>(smem_tc);                           // one layer's image
    float* tab = Bsm + t.img_floats;
    float2* zsm = reinterpret_cast<float2*>(tab + ((t.tab_floats + 3) & ~3));
    uint64_t* bars = reinterpret_cast<uint64_t*>(zsm + kTcRows);             // [0] mma, [1] weights
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    int* s_work = reinterpret_cast<int*>(tmem_slot + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = a.g.L, N = a.g.N, Mold = a.Mold;
    const bool is_row = warp < 8;
    const int part = warp >> 2 & 1, rowi = tid & 127;

    if (warp == 8) umma::tmem_alloc(tmem_slot, 512);
    if (tid == 0) {
        umma::mbar_init(&bars[0], 1);
        umma::mbar_init(&bars[1], 1);
        umma::mbar_fence_init();
    }
    for (int i = tid; i < t.tab_floats; i += blockDim.x) tab[i] = a.tabg[i];
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tbase = *tmem_slot;
    const uint32_t lane_addr = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    if (warp < 4) {   // zero the four A regions once: their K padding columns must never hold NaN/Inf bit patterns
        float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (uint32_t c = 0; c < 4 * Kp; c += 8) umma::tmem_st8(lane_addr + colXH + c, z);
        const float one[1] = {1.0f};   // K column H of X_hi and H_hi is the constant 1 that multiplies the bias rows of B
        umma::tmem_st1(lane_addr + colXH + H, one);
        umma::tmem_st1(lane_addr + colHH + H, one);
        umma::wait_st();
    }
    uint32_t par_mma = 0, par_w = 0;
    int loaded_layer = -1;
    float* xb = a.xbuf + (size_t)blockIdx.x * kTcT * kTcRows * HP;
    float* hs = a.hsave + (size_t)blockIdx.x * L * kTcRows * HP;
    const int total = (BASE ? 1 : N) * a.tiles128;
    const uint32_t idX = umma::instr_desc(umma::kFmtTF32, 128, 3 * kTcBlk), id1 = umma::instr_desc(umma::kFmtTF32, 128, 2 * kTcBlk),
                   id2 = umma::instr_desc(umma::kFmtTF32, 128, kTcBlk);
    const uint32_t sB = umma::smem_u32(Bsm);
    const uint32_t half = (uint32_t)(t.s1 + t.s2) * 4;                     // bytes of the hi half of (Bh1|Bh2)
    const uint32_t oX = 2 * half;                                          // byte offset of Bx_hi
    const uint64_t d1hi = umma::smem_desc(sB, 128, KC * 128), d1lo = umma::smem_desc(sB + half, 128, KC * 128);
    const uint64_t d2hi = umma::smem_desc(sB + (uint32_t)t.s1 * 4, 128, KC * 128),
                   d2lo = umma::smem_desc(sB + half + (uint32_t)t.s1 * 4, 128, KC * 128);
    const uint64_t dXhi = umma::smem_desc(sB + oX, 128, KC * 128), dXlo = umma::smem_desc(sB + oX + (uint32_t)t.sx * 4, 128, KC * 128);

    while (true) {
        if (tid == 0) *s_work = atomicAdd(a.counter, 1);
        __syncthreads();
        const int work = *s_work;
        __syncthreads();
        if (work >= total) break;
        const int s = BASE ? -1 : work / a.tiles128, tile = work % a.tiles128;   // ascending s = longest chains first
        const int64_t R = (int64_t)tile * kTcRows + rowi;
        const bool live = is_row && R < a.rows_total;
        const int64_t t120 = live ? R / Mold : 0;
        const int m = live ? (int)(R % Mold) : 0;
        const size_t rowbase = (size_t)t120 * N;                           // index of (old tile, site 0)
        double acc = 0.0;
        if (!BASE && live && part == 0) acc = a.la_oth[(rowbase + s) * Mold + m] - a.la_sel[(rowbase + s) * Mold + m];

        for (int b0 = s + 1; b0 < N; b0 += kTcT) {
            const int nb = min(kTcT, N - b0);
            for (int l = 0; l < L; ++l) {
                // ---- weights of layer l -> shared memory (bulk async copies, overlapped with the row threads' state restore) ----
                if (loaded_layer != l && tid == 0) {
                    const float* src = a.img + (size_t)l * t.img_floats;
                    const uint32_t bytes = (uint32_t)t.img_floats * 4;
                    umma::mbar_expect_tx(&bars[1], bytes);
                    for (uint32_t o = 0; o < bytes; o += 32768)
                        umma::bulk_g2s(reinterpret_cast<unsigned char*>(Bsm) + o, reinterpret_cast<const unsigned char*>(src) + o,
                                       min(32768u, bytes - o), &bars[1]);
                }
                if (is_row) {
                    if (part == 0)
                        tc_row_block<H, 0, BASE>(a, tab, zsm, bars, par_mma, lane_addr, rowi, live, rowbase, m, s, b0, nb, l, b0 == s + 1, xb, hs, acc);
                    else
                        tc_row_block<H, 1, BASE>(a, tab, zsm, bars, par_mma, lane_addr, rowi, live, rowbase, m, s, b0, nb, l, b0 == s + 1, xb, hs, acc);
                } else {
                    for (int tt = 0; tt < nb; ++tt) {
                        umma::fence_before_sync();
                        tc_named_sync();
                        if (lane == 0) {
                            // ---- MMA issue: 3 passes (hi*hi, lo*hi, hi*lo) over x (layers >= 1) and h ----
                            if (tt == 0 && loaded_layer != l) umma::mbar_wait(&bars[1], par_w);
                            umma::fence_after_sync();
                            // descriptors differ only in the start-address field: +16 (256 bytes >> 4) per K step
                            uint32_t accD = 0, accC = 0;
#pragma unroll
                            for (int pass = 0; pass < 3; ++pass) {
                                const uint32_t aX = tbase + (pass == 1 ? colXL : colXH), aH = tbase + (pass == 1 ? colHL : colHH);
                                const uint64_t dX = (pass == 2 ? dXlo : dXhi), d1 = (pass == 2 ? d1lo : d1hi), d2 = (pass == 2 ? d2lo : d2hi);
                                if (l > 0) {
#pragma unroll
                                    for (int ks = 0; ks < Kp / 8; ++ks) {
                                        umma::mma_tf32_ts(tbase + colD, aX + ks * 8, dX + (uint64_t)(ks * 16), idX, accD);
                                        accD = 1;
                                    }
                                } else if (pass != 1) {   // layer 0: one-hot input (K columns 0, 1) and the bias column; both exact in TF32
                                    umma::mma_tf32_ts(tbase + colD, aX, dX, idX, accD);
                                    accD = 1;
                                    umma::mma_tf32_ts(tbase + colD, aX + (H & ~7), dX + (uint64_t)((H >> 3) * 16), idX, accD);
                                }
#pragma unroll
                                for (int ks = 0; ks < Kp / 8; ++ks) {
                                    umma::mma_tf32_ts(tbase + colD, aH + ks * 8, d1 + (uint64_t)(ks * 16), id1, accD);
                                    accD = 1;
                                    umma::mma_tf32_ts(tbase + colD + 3 * kTcBlk, aH + ks * 8, d2 + (uint64_t)(ks * 16), id2, accC);
                                    accC = 1;
                                }
                            }
                            umma::commit(&bars[0]);
                        }
                    }
                }
                if (loaded_layer != l) {
                    loaded_layer = l;
                    par_w ^= 1;          // every thread tracks the weight-barrier phase; only the MMA thread waits on it
                }
            }
        }
        if (live && part == 0) {
            if (BASE) a.lp[t120 * Mold + m] = acc;
            else a.delta[((size_t)t120 * N + s) * Mold + m] = acc;
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 8) umma::tmem_dealloc(tbase, 512);
    if (a.dbg && (tid == 0 || tid == 128) && blockIdx.x < 1) {
        const long long* d = a.dbg + 16 * blockIdx.x + 8 * (tid >> 7);
        const double n = (double)d[3];
        printf("[tc dbg] base=%d part %d: steps %lld | x-stage %.0f wait_st %.0f named-sync %.0f | mma-wait %.0f gate-math %.0f h-stage+head %.0f cycles/step\n",
               (int)BASE, tid >> 7, d[3], d[4] / n, d[5] / n, d[6] / n, d[0] / n, d[1] / n, d[2] / n);
    }
}

struct TcWs {
    float *img, *tab, *xbuf, *hsave;
};

inline bool tc_supported(const GruLayout& g) { return g.H == 50 && g.nheads == 1 && g.N >= 2; }

inline size_t tc_smem_bytes(const TcLayout& t) {
    return (size_t)t.img_floats * 4 + (size_t)((t.tab_floats + 3) & ~3) * 4 + kTcRows * sizeof(float2) + 64;
}

inline TcWs carve_tc(Ws& ws, const GruLayout& g, const TcLayout& t, int sms) {
    TcWs w;
    const int HP = ((g.H + 3) / 4) * 4;
    w.img = ws.take<float>((size_t)g.L * t.img_floats);
    w.tab = ws.take<float>(t.tab_floats);
    w.xbuf = ws.take<float>((size_t)sms * kTcT * kTcRows * HP);
    w.hsave = ws.take<float>((size_t)sms * g.L * kTcRows * HP + 4096);   // + room for the optional debug counters
    return w;
}

// base pass + single-flip chains on the tensor cores (replaces launch_forward<STASH> + launch_chain for the FP32 pRNN)
static int launch_eloc_tc(const GruLayout& g, const GruLaunch& c, const GruWs<float>& w, const TcWs& tw, int tiles, const float* params,
                          bool flips, cudaStream_t s) {
    const TcLayout t = make_tc_layout(g);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    TcArgs a;
    a.g = g; a.t = t; a.Mold = c.M;
    a.rows_total = (int64_t)tiles * c.M;
    a.tiles128 = (int)cdiv(a.rows_total, kTcRows);
    a.img = tw.img; a.tabg = tw.tab; a.sigT = w.sigT; a.hstore = w.hstore; a.la_sel = w.la_sel; a.la_oth = w.la_oth; a.lp = w.lp_re;
    a.xbuf = tw.xbuf; a.hsave = tw.hsave; a.delta = w.delta_re; a.counter = w.counter;
    a.dbg = getenv("RNNWF_TC_DEBUG") ? reinterpret_cast<long long*>(tw.hsave + (size_t)sms * g.L * kTcRows * (((g.H + 3) / 4) * 4)) : nullptr;
    const int smem = (int)tc_smem_bytes(t);
    RNNWF_CHECK(smem <= kSmemLimit, -3, "tensor-core chain kernel needs %d bytes of shared memory", smem);
    prof_count(); pack_gru_tc_kernel<<<grid_for(g.L * (t.s1 + t.s2 + t.sx)), 256, 0, s>>>(g, t, params, tw.img, tw.tab);
    {
        RNNWF_CUDA(cudaMemsetAsync(w.counter, 0, sizeof(int), s));
        if (a.dbg) RNNWF_CUDA(cudaMemsetAsync(a.dbg, 0, 4096 * sizeof(float), s));
        auto k = gru_chain_tc_kernel<50, true>;
        if (int e = set_smem(k, smem)) return e;
        prof_count();
        k<<<std::min(a.tiles128, sms), kTcThreads, smem, s>>>(a);
        RNNWF_CUDA(cudaGetLastError());
    }
    if (flips) {
        RNNWF_CUDA(cudaMemsetAsync(w.counter, 0, sizeof(int), s));
        if (a.dbg) RNNWF_CUDA(cudaMemsetAsync(a.dbg, 0, 4096 * sizeof(float), s));
        auto k = gru_chain_tc_kernel<50, false>;
        if (int e = set_smem(k, smem)) return e;
        const int grid = (int)std::min<int64_t>((int64_t)g.N * a.tiles128, sms);
        prof_count();
        prof_mark(0, s);
        k<<<grid, kTcThreads, smem, s>>>(a);
        prof_mark(1, s);
        RNNWF_CUDA(cudaGetLastError());
    }
    return 0;
}

}  // namespace rnnwf
