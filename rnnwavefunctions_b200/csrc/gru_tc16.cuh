// gru_tc16.cuh — tensor-core chain kernel, generation 2: tcgen05 kind::f16 with 3xFP16 operands (hi*hi + lo*hi + hi*lo,
// FP32 accumulate) and ALL layers' weights resident in shared memory.
//
// Same contract as gru_chain_kernel / gru_tc.cuh (TFIM single flips, 1DTFIM/TrainingRNN_1DTFIM.py:43-48,74).
// Why FP16 limbs instead of TF32 limbs: FP16 has the same 11-bit significand as TF32, the operands here live in
// (-1, 1) (hidden states) or are O(1) (weights), so an (hi, lo) pair of halfs carries ~22 bits like a TF32 pair, but
//   * kind::f16 runs at twice the kind::tf32 rate (K = 16 per instruction),
//   * a weight costs 4 bytes (hi + lo) instead of 8: 3 x GRU(50) fits in 210 KB of shared memory, so there is no
//     weight swapping, no site blocking and no inter-layer scratch in global memory,
//   * the A operand of a layer (its previous hidden state, hi/lo packed pairs) needs 64 TMEM columns, and the same
//     region is the x operand of the layer above.
// TMEM columns: D: cx [0,52) r [52,104) u [104,156) ch [156,208) | R_l (h^l: 32 hi + 32 lo) at 224 + 64 l | X0 (one-hot) at 224 + 64 L
// Included by gru.cu.
#pragma once
#include <cuda_fp16.h>
#include "gru_kernels.cuh"
#include "host_util.cuh"
#include "umma.cuh"

namespace rnnwf {
namespace tc16 {

// ---- helpers shared with the pipelined generation (gru_tc16p.cuh); always compiled ------------------------------------------------
__host__ __device__ __forceinline__ int core_off(int n, int k, int KC) {   // offset in halfs inside a K-major core-matrix image
    return (n >> 3) * (KC * 64) + (k >> 3) * 64 + (n & 7) * 8 + (k & 7);
}
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {   // low half = fp16(a), high half = fp16(b)
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ float2 unpack_h2(uint32_t w) {
    const __half2 h = *reinterpret_cast<const __half2*>(&w);
    return __half22float2(h);
}

#ifdef RNNWF_LEGACY   // generation 2 (MMA and gate math alternating): kept for A/B runs (RNNWF_CHAIN=tc16), not in the product build

constexpr int kRows = 128, kRowThreads = 256, kThreads = 288;
constexpr int kBW = 52;                 // D columns per gate block (multiple of 4: TMEM load alignment)
constexpr int kNN = 160;                // N of the x-part ([cx|r|u]) and of the h-part ([r|u|ch]) MMAs
constexpr int kKp = 64, kKC = 8;        // K padded to 4 MMA steps of 16; 16-byte chunks per row
constexpr int kColD = 0, kColR = 224;   // accumulators (the ch-clearing MMA spans [156, 220)) / operand regions

struct Layout {
    int L, H, N;
    int bh_bytes, bx_bytes, bx0_bytes;   // one precision half of BH (160 x 64), BX (160 x 64), BX of layer 0 (160 x 16)
    int l0_bytes, l1_bytes;              // layer 0 / layers >= 1: BH_hi | BH_lo | BX_hi | BX_lo
    int zero_off, tab_off, img_bytes;    // zero matrix (64 x 16 halfs), head table (floats), total image
    int tab_floats;
};

inline Layout make_layout(const GruLayout& g) {
    Layout t;
    t.L = g.L; t.H = g.H; t.N = g.N;
    t.bh_bytes = kNN * kKp * 2;
    t.bx_bytes = kNN * kKp * 2;
    t.bx0_bytes = kNN * 16 * 2;
    t.l0_bytes = 2 * t.bh_bytes + 2 * t.bx0_bytes;
    t.l1_bytes = 2 * t.bh_bytes + 2 * t.bx_bytes;
    t.zero_off = t.l0_bytes + (g.L - 1) * t.l1_bytes;
    t.tab_off = t.zero_off + 64 * 16 * 2;
    t.tab_floats = g.nheads * (2 * 64 + 4);   // per head: Wd[64][2] | bd[2] | pad
    t.img_bytes = t.tab_off + t.tab_floats * 4;
    return t;
}

inline bool supported(const GruLayout& g) {
    if (!(g.H == 50 && g.N >= 2 && g.L >= 1 && g.L <= 3)) return false;
    return make_layout(g).img_bytes + 2048 <= kSmemLimit;
}


// flat TF-order parameters -> shared-memory image.  Weights are pre-scaled so that the gates are 1/(1 + 2^a):
// r, u rows by -log2(e), candidate rows by 2 log2(e); K column H (times the constant-1 column of the operand regions)
// carries the biases bg (BH r,u rows), bch (BH ch rows) and bci (BX cx rows).  Layer 0: BX has K = 16 with the two one-hot
// rows of the input kernels at k = 0, 1 and bci at k = 2.
__global__ void pack_kernel(GruLayout g, Layout t, const float* __restrict__ flat, unsigned char* __restrict__ img) {
    const int H = g.H;
    const float kS = -1.4426950408889634f, kC = 2.8853900817779268f;
    const int per_l = kNN * kKp * 2;                    // BH elements + BX elements (64-wide) per layer, one precision
    const int total = g.L * per_l;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int l = idx / per_l;
        int q = idx % per_l;
        const int d = g.d[l];
        const float* Kg = flat + g.flat_off[l];
        const float* bg = Kg + (d + H) * 2 * H;
        const float* Kci = bg + 2 * H;
        const float* Kch = Kci + d * H;
        const float* bci = Kch + H * H;
        const float* bch = bci + H;
        unsigned char* Lb = img + (l == 0 ? 0 : t.l0_bytes + (l - 1) * t.l1_bytes);
        float v = 0.f;
        __half *hi_p, *lo_p;
        if (q < kNN * kKp) {                            // BH: rows [r | u | ch] (52 each), K over h (+ bias column)
            const int n = q / kKp, k = q % kKp, gate = n / kBW, j = n % kBW;
            if (gate < 3 && j < H) {
                if (gate < 2) v = k < H ? kS * Kg[(d + k) * 2 * H + gate * H + j] : (k == H ? kS * bg[gate * H + j] : 0.f);
                else v = k < H ? kC * Kch[k * H + j] : (k == H ? kC * bch[j] : 0.f);
            }
            hi_p = reinterpret_cast<__half*>(Lb) + core_off(n, k, kKC);
            lo_p = reinterpret_cast<__half*>(Lb + t.bh_bytes) + core_off(n, k, kKC);
        } else {                                        // BX: rows [cx | r | u]
            q -= kNN * kKp;
            const int n = q / kKp, k = q % kKp, gate = n / kBW, j = n % kBW;   // gate 0: cx, 1: r, 2: u
            if (l > 0) {
                if (gate < 3 && j < H) {
                    if (gate == 0) v = k < d ? kC * Kci[k * H + j] : (k == H ? kC * bci[j] : 0.f);
                    else v = k < d ? kS * Kg[k * 2 * H + (gate - 1) * H + j] : 0.f;
                }
                hi_p = reinterpret_cast<__half*>(Lb + 2 * t.bh_bytes) + core_off(n, k, kKC);
                lo_p = reinterpret_cast<__half*>(Lb + 2 * t.bh_bytes + t.bx_bytes) + core_off(n, k, kKC);
            } else {
                if (k >= 16) continue;                  // layer 0: K = 16
                if (gate < 3 && j < H) {
                    if (gate == 0) v = k < 2 ? kC * Kci[k * H + j] : (k == 2 ? kC * bci[j] : 0.f);
                    else v = k < 2 ? kS * Kg[k * 2 * H + (gate - 1) * H + j] : 0.f;
                }
                hi_p = reinterpret_cast<__half*>(Lb + 2 * t.bh_bytes) + core_off(n, k, 2);
                lo_p = reinterpret_cast<__half*>(Lb + 2 * t.bh_bytes + t.bx0_bytes) + core_off(n, k, 2);
            }
        }
        const __half hi = __float2half_rn(v);
        *hi_p = hi;
        *lo_p = __float2half_rn(v - __half2float(hi));
    }
    const int tail = (t.img_bytes - t.zero_off) / 4;    // zero matrix + head table
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < tail; idx += gridDim.x * blockDim.x) {
        float v = 0.f;
        const int q = idx - (t.tab_off - t.zero_off) / 4;
        if (q >= 0) {                                   // per head: Wd[j][2] (64 x 2) | bd[2] | pad
            const int hd = q / 132, r = q % 132;
            const float* hw = flat + g.flat_head + hd * (2 * H + 2);
            if (r < 128) { if (r / 2 < H) v = hw[r]; }
            else if (r < 130) v = hw[2 * H + (r - 128)];
        }
        reinterpret_cast<float*>(img + t.zero_off)[idx] = v;
    }
}

__device__ __forceinline__ void named_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory"); }
__device__ __forceinline__ void row_sync() { asm volatile("bar.sync 2, %0;" ::"n"(kRowThreads) : "memory"); }

// unit range of a row thread: part 0 owns units [0, S), part 1 owns [S, H); S multiple of 8
template <int H, int PART> struct Part {
    static constexpr int S = ((H / 2) / 8) * 8;
    static constexpr int U0 = PART ? S : 0;
    static constexpr int U = PART ? H - S : S;
    static constexpr int NG = (U + 7) / 8;
};

// (hi, lo) FP16 pairs of CNT (even, <= 8) consecutive units -> operand region columns (two units per column)
template <int CNT> __device__ __forceinline__ void stage_units(uint32_t reg_addr, int unit0, const float* h) {
    float hi[4], lo[4];
#pragma unroll
    for (int c = 0; c < CNT / 2; ++c) {
        const uint32_t wh = pack_h2(h[2 * c], h[2 * c + 1]);
        const float2 f = unpack_h2(wh);
        hi[c] = __uint_as_float(wh);
        lo[c] = __uint_as_float(pack_h2(h[2 * c] - f.x, h[2 * c + 1] - f.y));
    }
    const uint32_t col = reg_addr + unit0 / 2;
    if constexpr (CNT == 8) {
        umma::tmem_st4(col, hi);
        umma::tmem_st4(col + 32, lo);
    } else {
        static_assert(CNT == 2, "unit groups are 8 wide or the 2-unit tail");
        umma::tmem_st1(col, hi);
        umma::tmem_st1(col + 32, lo);
    }
}
template <int H, int PART> __device__ __forceinline__ void stage_all(uint32_t reg_addr, const float* hp) {
    using P = Part<H, PART>;
#pragma unroll
    for (int gq = 0; gq < P::NG; ++gq) {
        if (P::U - 8 * gq >= 8) stage_units<8>(reg_addr, P::U0 + 8 * gq, hp + 8 * gq);
        else stage_units<2>(reg_addr, P::U0 + 8 * gq, hp + 8 * gq);
    }
}

struct Args {
    GruLayout g;
    Layout t;
    int Mold, tiles128;
    int64_t rows_total;
    const unsigned char* img;
    const uint8_t* sigT;
    float* hstore;            // BASE: written (every layer, every site); FLIP: restart states
    double *la_sel, *la_oth;  // BASE: written; FLIP: read
    double* lp;               // BASE: sum_n la_sel
    double* delta;            // FLIP: [tile][slot][M]
    int* counter;
    // complex cRNN / J1-J2 exchanges (CPLX instantiations): phases, imaginary parts, slot plan
    double *ph_sel, *ph_oth, *lp_im, *delta_im;
    const int* order;         // slots by decreasing chain length
    const double *j1, *j2;    // couplings: slots with a zero coupling are skipped (J1J2/TrainingRNN_J1J2.py:69,84)
    int n_kind1, n_kind2, nslots;
};

// one (site, layer) step of a row thread: wait for the MMAs, gate math on the accumulators, stage the new state
template <int H, int PART, bool BASE, bool CPLX>
__device__ __forceinline__ void row_step(const Args& a, const float* tab, float4* zsm, uint64_t* bar, uint32_t& par, uint32_t lane_addr,
                                         int rowi, bool live, size_t rowbase, int m, int n, int l, float* hp, float4& zout) {
    using P = Part<H, PART>;
    const int L = a.g.L, Mold = a.Mold;
    const bool top = l == L - 1;
    umma::mbar_wait(bar, par);
    par ^= 1;
    umma::fence_after_sync();
    const uint32_t reg = lane_addr + kColR + 64 * l;
    float z0 = 0.f, z1 = 0.f, y0 = 0.f, y1 = 0.f;
#pragma unroll
    for (int gq = 0; gq < P::NG; ++gq) {
        const int cnt = P::U - 8 * gq >= 8 ? 8 : P::U - 8 * gq;
        const uint32_t col = lane_addr + kColD + P::U0 + 8 * gq;
        float dc[8], dr[8], du[8], dq[8];
        umma::tmem_ld8(col, dc);
        umma::tmem_ld8(col + kBW, dr);
        umma::tmem_ld8(col + 2 * kBW, du);
        umma::tmem_ld8(col + 3 * kBW, dq);
        umma::wait_ld();
        // gates of two units at a time: 1/(1+2^a) for r0, u0, r1, u1 share ONE reciprocal (exponents clamped to 30, so the
        // product of the four denominators stays below 2^121), and so do the two candidates' 1/(1+2^c)
#pragma unroll
        for (int q = 0; q < 8; q += 2) {
            if (q < cnt) {
                const int jl = 8 * gq + q, j = P::U0 + jl;
                const float er0 = 1.0f + ex2(fminf(dr[q], 30.f)), eu0 = 1.0f + ex2(fminf(du[q], 30.f));
                const float er1 = 1.0f + ex2(fminf(dr[q + 1], 30.f)), eu1 = 1.0f + ex2(fminf(du[q + 1], 30.f));
                const float p0 = er0 * eu0, p1 = er1 * eu1;
                const float inv = rcp(p0 * p1);
                const float i0 = inv * p1, i1 = inv * p0;                          // 1/p0, 1/p1
                const float r0 = i0 * eu0, u0 = i0 * er0, r1 = i1 * eu1, u1 = i1 * er1;
                const float ec0 = 1.0f + ex2(fminf(fmaf(r0, dq[q], dc[q]), 60.f)), ec1 = 1.0f + ex2(fminf(fmaf(r1, dq[q + 1], dc[q + 1]), 60.f));
                const float ic = rcp(ec0 * ec1);
                const float c0 = fmaf(-2.0f, ic * ec1, 1.0f), c1 = fmaf(-2.0f, ic * ec0, 1.0f);
                const float h0 = fmaf(u0, hp[jl] - c0, c0), h1 = fmaf(u1, hp[jl + 1] - c1, c1);
                hp[jl] = h0;
                hp[jl + 1] = h1;
                if (top) {
                    z0 = fmaf(h0, tab[2 * j], z0);
                    z1 = fmaf(h0, tab[2 * j + 1], z1);
                    z0 = fmaf(h1, tab[2 * j + 2], z0);
                    z1 = fmaf(h1, tab[2 * j + 3], z1);
                    if (CPLX) {
                        y0 = fmaf(h0, tab[132 + 2 * j], y0);
                        y1 = fmaf(h0, tab[132 + 2 * j + 1], y1);
                        y0 = fmaf(h1, tab[132 + 2 * j + 2], y0);
                        y1 = fmaf(h1, tab[132 + 2 * j + 3], y1);
                    }
                }
                if (BASE && live) {
                    a.hstore[(((rowbase + n) * L + l) * (size_t)H + j) * Mold + m] = h0;
                    a.hstore[(((rowbase + n) * L + l) * (size_t)H + j + 1) * Mold + m] = h1;
                }
            }
        }
        // the MMAs of this step are complete: region l can take the new state (it is the h operand of layer l at the next
        // site and the x operand of layer l + 1 at this site); interleaved with the MUFU-bound math of the next group
        if (cnt == 8) stage_units<8>(reg, P::U0 + 8 * gq, hp + 8 * gq);
        else stage_units<2>(reg, P::U0 + 8 * gq, hp + 8 * gq);
    }
    if (top) {   // partial head sums; the log-softmax itself is finished during the next step's MMA wait (row_chain)
        if (PART == 1) zsm[rowi] = make_float4(z0, z1, y0, y1);
        else zout = make_float4(z0, z1, y0, y1);
    }
}

// all row-thread work of one chain.  kind 0: sigma with site s flipped (TFIM); kind 1 / 2: sites s and t = s + kind exchanged
// (J1-J2, J1J2/TrainingRNN_J1J2.py:68-92); BASE: the unmodified configuration from site 0 (s = -1).
template <int H, int PART, bool BASE, bool CPLX>
__device__ __forceinline__ void row_chain(const Args& a, const float* tab, float4* zsm, uint64_t* bars, uint32_t& par, uint32_t lane_addr,
                                          int rowi, bool live, size_t rowbase, int m, int s, int t, double& acc, double& acc_im) {
    using P = Part<H, PART>;
    const int L = a.g.L, N = a.g.N, Mold = a.Mold;
    float hp0[P::U], hp1[P::U], hp2[P::U];
#pragma unroll
    for (int j = 0; j < P::U; ++j) { hp0[j] = 0.f; hp1[j] = 0.f; hp2[j] = 0.f; }
    if (!BASE && live) {   // restart from the base states after site s
        const float* src = a.hstore + (rowbase + s) * L * (size_t)H * Mold + m;
#pragma unroll
        for (int j = 0; j < P::U; ++j) {
            hp0[j] = src[(size_t)(P::U0 + j) * Mold];
            if (L > 1) hp1[j] = src[((size_t)H + P::U0 + j) * Mold];
            if (L > 2) hp2[j] = src[((size_t)2 * H + P::U0 + j) * Mold];
        }
    }
    stage_all<H, PART>(lane_addr + kColR, hp0);
    if (L > 1) stage_all<H, PART>(lane_addr + kColR + 64, hp1);
    if (L > 2) stage_all<H, PART>(lane_addr + kColR + 128, hp2);
    const uint32_t x0 = lane_addr + kColR + 64 * L;
    auto spin = [&](int q) {   // spin of site q in the configuration this chain evaluates
        int c = a.sigT[(rowbase + q) * Mold + m];
        if (!BASE && (q == s || q == t)) c = 1 - c;
        return c;
    };
    int code = (PART == 0 && live && s >= 0) ? spin(s) : 2;                // input of site s + 1 (the zero vector at site 0)
    int nup = 0;                                                           // up spins among the sites before the current one
    if (CPLX && PART == 0 && live && !BASE) {
        for (int q = 0; q < s; ++q) nup += a.sigT[(rowbase + q) * Mold + m];
        nup += code;
    }
    // head of the previous site, finished while the MMAs of the next step run (part 0 only)
    float4 pz = make_float4(0.f, 0.f, 0.f, 0.f);
    int psg = 0, pn = -1;
    auto finish_head = [&]() {
        if (PART == 0 && pn >= 0 && live) {
            const float4 o = zsm[rowi];
            const float f0 = pz.x + o.x + tab[128], f1 = pz.y + o.y + tab[129];
            // log softmax of the 2-way head in FP32 (log1pf/expf, ~1e-7 relative); the site terms are summed in FP64
            const float dsel = psg ? f0 - f1 : f1 - f0;                             // z_other - z_selected
            double ls = dsel > 30.f ? -(double)dsel : -(double)log1pf(expf(dsel));
            double lo = -dsel > 30.f ? (double)dsel : -(double)log1pf(expf(-dsel));
            double ps = 0.0, po = 0.0;
            if (CPLX) {
                // amplitude = sqrt(softmax) with the zero-magnetisation mask and renormalisation
                // (J1J2/ComplexRNNwavefunction.py:85-93,147-155); phase = pi * softsign (:8-9)
                ls *= 0.5;
                lo *= 0.5;
                if (2 * pn >= N) {
                    const int half = N / 2, ndn = pn - nup;
                    const bool ok_dn = (half - 1 - ndn) >= 0, ok_up = (half - 1 - nup) >= 0;
                    const bool ok_sel = psg ? ok_up : ok_dn, ok_oth = psg ? ok_dn : ok_up;
                    const double ninf = -__longlong_as_double(0x7ff0000000000000LL);
                    if (!ok_sel) ls = ninf; else if (!ok_oth) ls = 0.0;
                    if (!ok_oth) lo = ninf; else if (!ok_sel) lo = 0.0;
                }
                const float y0 = pz.z + o.z + tab[132 + 128], y1 = pz.w + o.w + tab[132 + 129];
                const double ys = psg ? (double)y1 : (double)y0, yo = psg ? (double)y0 : (double)y1;
                ps = kPi * ys / (1.0 + fabs(ys));
                po = kPi * yo / (1.0 + fabs(yo));
                nup += psg;
            }
            const size_t o_ = (rowbase + pn) * Mold + m;
            if (BASE) {
                a.la_sel[o_] = ls;
                a.la_oth[o_] = lo;
                acc += ls;
                if (CPLX) { a.ph_sel[o_] = ps; a.ph_oth[o_] = po; acc_im += ps; }
            } else {
                acc += ls - a.la_sel[o_];
                if (CPLX) acc_im += ps - a.ph_sel[o_];
            }
        }
        pn = -1;
    };
    for (int n = s + 1; n < N; ++n) {
        // ---- layer 0: one-hot input of the previous spin ----
        if (PART == 0) {
            const float oh[1] = {__uint_as_float(pack_h2(code == 0 ? 1.f : 0.f, code == 1 ? 1.f : 0.f))};
            umma::tmem_st1(x0, oh);
        }
        umma::wait_st();
        umma::fence_before_sync();
        named_sync();                      // also orders part 1's head partials (zsm) of the previous site before finish_head
        int sg = 0;
        if (PART == 0) {
            if (live) {
                sg = spin(n);
                code = sg;
            }
            finish_head();
        }
        float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        row_step<H, PART, BASE, CPLX>(a, tab, zsm, &bars[0], par, lane_addr, rowi, live, rowbase, m, n, 0, hp0, z);
        if (L > 1) {
            umma::wait_st();
            umma::fence_before_sync();
            named_sync();
            row_step<H, PART, BASE, CPLX>(a, tab, zsm, &bars[0], par, lane_addr, rowi, live, rowbase, m, n, 1, hp1, z);
        }
        if (L > 2) {
            umma::wait_st();
            umma::fence_before_sync();
            named_sync();
            row_step<H, PART, BASE, CPLX>(a, tab, zsm, &bars[0], par, lane_addr, rowi, live, rowbase, m, n, 2, hp2, z);
        }
        pz = z; psg = sg; pn = n;
    }
    row_sync();                            // part 1's partials of the last site
    finish_head();
}

template <int H, bool BASE, bool CPLX>
__global__ void __launch_bounds__(kThreads, 1) chain_kernel(const __grid_constant__ Args a) {
    extern __shared__ __align__(128) unsigned char smem_h16[];
    const Layout& t = a.t;
    const float* tab = reinterpret_cast<const float*>(smem_h16 + t.tab_off);
    float4* zsm = reinterpret_cast<float4*>(smem_h16 + ((t.img_bytes + 15) & ~15));
    uint64_t* bars = reinterpret_cast<uint64_t*>(zsm + kRows);              // [0] MMAs, [1] weight image
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
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    if (tid == 0) {   // the whole weight image stays resident: bulk async copies (TMA), one barrier
        umma::mbar_expect_tx(&bars[1], (uint32_t)t.img_bytes);
        for (uint32_t o = 0; o < (uint32_t)t.img_bytes; o += 32768)
            umma::bulk_g2s(smem_h16 + o, a.img + o, min(32768u, (uint32_t)t.img_bytes - o), &bars[1]);
    }
    const uint32_t tbase = *tmem_slot;
    const uint32_t lane_addr = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    if (warp < 4) {   // zero the operand regions, then set their constant-1 K column (k = H -> column H/2, low half)
        const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (uint32_t c = 0; c < (uint32_t)(64 * L + 8); c += 8) umma::tmem_st8(lane_addr + kColR + c, z);
        const float one[1] = {__uint_as_float(pack_h2(1.0f, 0.0f))};
        for (int l = 0; l < L; ++l) umma::tmem_st1(lane_addr + kColR + 64 * l + H / 2, one);
        umma::tmem_st1(lane_addr + kColR + 64 * L + 1, one);               // one-hot region: k = 2 is the constant 1
        umma::wait_st();
    }
    if (is_row) umma::mbar_wait(&bars[1], 0);                               // tab is read with ordinary loads
    uint32_t par = 0;
    const int total = (BASE ? 1 : a.nslots) * a.tiles128;
    const uint32_t idN = (1u << 4) | ((uint32_t)(kNN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);     // F16 x F16 -> F32, M = 128
    const uint32_t idZ = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t sB = umma::smem_u32(smem_h16);
    bool weights_ready = false;

    while (true) {
        if (tid == 0) *s_work = atomicAdd(a.counter, 1);
        __syncthreads();
        const int work = *s_work;
        __syncthreads();
        if (work >= total) break;
        const int tile = work % a.tiles128;
        int slot = 0, s = -1, tt = -1;                                        // modified sites of this chain (none for BASE)
        if (!BASE) {
            slot = a.order ? a.order[work / a.tiles128] : work / a.tiles128;  // decreasing chain length
            if (slot < a.nslots - a.n_kind1 - a.n_kind2) s = slot;
            else if (slot < a.nslots - a.n_kind2) { s = slot - (a.nslots - a.n_kind1 - a.n_kind2); tt = s + 1; }
            else { s = slot - (a.nslots - a.n_kind2); tt = s + 2; }
            if (tt >= 0 && ((tt == s + 1 && a.j1 && a.j1[s] == 0.0) || (tt == s + 2 && a.j2 && a.j2[s] == 0.0))) continue;
        }
        const int64_t R = (int64_t)tile * kRows + rowi;
        const bool live = is_row && R < a.rows_total;
        const int64_t t120 = live ? R / Mold : 0;
        const int m = live ? (int)(R % Mold) : 0;
        const size_t rowbase = (size_t)t120 * N;                            // index of (old tile, site 0)
        double acc = 0.0, acc_im = 0.0;
        if (!BASE && live && part == 0) {
            acc = a.la_oth[(rowbase + s) * Mold + m] - a.la_sel[(rowbase + s) * Mold + m];
            if (CPLX) acc_im = a.ph_oth[(rowbase + s) * Mold + m] - a.ph_sel[(rowbase + s) * Mold + m];
        }

        if (is_row) {
            if (part == 0) row_chain<H, 0, BASE, CPLX>(a, tab, zsm, bars, par, lane_addr, rowi, live, rowbase, m, s, tt, acc, acc_im);
            else row_chain<H, 1, BASE, CPLX>(a, tab, zsm, bars, par, lane_addr, rowi, live, rowbase, m, s, tt, acc, acc_im);
        } else {
            for (int n = s + 1; n < N; ++n) {
                for (int l = 0; l < L; ++l) {
                    umma::fence_before_sync();
                    named_sync();
                    if (lane == 0) {
                        if (!weights_ready) { umma::mbar_wait(&bars[1], 0); weights_ready = true; }
                        umma::fence_after_sync();
                        const uint32_t lb = sB + (l == 0 ? 0u : (uint32_t)(t.l0_bytes + (l - 1) * t.l1_bytes));
                        const uint32_t dX = tbase + kColD, dH = tbase + kColD + kBW;
                        const uint32_t rX = tbase + kColR + 64 * (l == 0 ? L : l - 1), rH = tbase + kColR + 64 * l;
                        // x part: overwrite [cx | r | u]
                        if (l == 0) {
                            const uint64_t bhi = umma::smem_desc(lb + 2 * t.bh_bytes, 128, 2 * 128);
                            const uint64_t blo = umma::smem_desc(lb + 2 * t.bh_bytes + t.bx0_bytes, 128, 2 * 128);
                            umma::mma_f16_ts(dX, rX, bhi, idN, 0);
                            umma::mma_f16_ts(dX, rX, blo, idN, 1);
                        } else {
                            const uint64_t bhi = umma::smem_desc(lb + 2 * t.bh_bytes, 128, kKC * 128);
                            const uint64_t blo = umma::smem_desc(lb + 2 * t.bh_bytes + t.bx_bytes, 128, kKC * 128);
#pragma unroll
                            for (int ks = 0; ks < kKp / 16; ++ks) umma::mma_f16_ts(dX, rX + ks * 8, bhi + (uint64_t)(ks * 16), idN, ks > 0);
#pragma unroll
                            for (int ks = 0; ks < kKp / 16; ++ks) umma::mma_f16_ts(dX, rX + 32 + ks * 8, bhi + (uint64_t)(ks * 16), idN, 1);
#pragma unroll
                            for (int ks = 0; ks < kKp / 16; ++ks) umma::mma_f16_ts(dX, rX + ks * 8, blo + (uint64_t)(ks * 16), idN, 1);
                        }
                        // clear the ch block (the h part accumulates onto r, u but starts ch): 64-column MMA with a zero B
                        umma::mma_f16_ts(tbase + kColD + 3 * kBW, tbase + kColR + 64 * L, umma::smem_desc(sB + t.zero_off, 128, 2 * 128), idZ, 0);
                        // h part: accumulate onto [r | u | ch]
                        {
                            const uint64_t bhi = umma::smem_desc(lb, 128, kKC * 128);
                            const uint64_t blo = umma::smem_desc(lb + t.bh_bytes, 128, kKC * 128);
#pragma unroll
                            for (int ks = 0; ks < kKp / 16; ++ks) umma::mma_f16_ts(dH, rH + ks * 8, bhi + (uint64_t)(ks * 16), idN, 1);
#pragma unroll
                            for (int ks = 0; ks < kKp / 16; ++ks) umma::mma_f16_ts(dH, rH + 32 + ks * 8, bhi + (uint64_t)(ks * 16), idN, 1);
#pragma unroll
                            for (int ks = 0; ks < kKp / 16; ++ks) umma::mma_f16_ts(dH, rH + ks * 8, blo + (uint64_t)(ks * 16), idN, 1);
                        }
                        umma::commit(&bars[0]);
                    }
                }
            }
        }
        if (live && part == 0) {
            if (BASE) {
                a.lp[t120 * Mold + m] = acc;
                if (CPLX) a.lp_im[t120 * Mold + m] = acc_im;
            } else {
                a.delta[((size_t)t120 * a.nslots + slot) * Mold + m] = acc;
                if (CPLX) a.delta_im[((size_t)t120 * a.nslots + slot) * Mold + m] = acc_im;
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 8) umma::tmem_dealloc(tbase, 512);
}

inline size_t smem_bytes(const Layout& t) { return (size_t)((t.img_bytes + 15) & ~15) + kRows * sizeof(float4) + 64; }

template <bool CPLX>
static int launch_chains(Args& a, int sms, bool flips, cudaStream_t s) {
    const int smem = (int)smem_bytes(a.t);
    RNNWF_CHECK(smem <= kSmemLimit, -3, "tensor-core chain kernel needs %d bytes of shared memory", smem);
    {
        RNNWF_CUDA(cudaMemsetAsync(a.counter, 0, sizeof(int), s));
        auto k = chain_kernel<50, true, CPLX>;
        RNNWF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        prof_count();
        k<<<std::min(a.tiles128, sms), kThreads, smem, s>>>(a);
        RNNWF_CUDA(cudaGetLastError());
    }
    if (flips) {
        RNNWF_CUDA(cudaMemsetAsync(a.counter, 0, sizeof(int), s));
        auto k = chain_kernel<50, false, CPLX>;
        RNNWF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        const int grid = (int)std::min<int64_t>((int64_t)a.nslots * a.tiles128, sms);
        prof_count();
        prof_mark(0, s);
        k<<<grid, kThreads, smem, s>>>(a);
        prof_mark(1, s);
        RNNWF_CUDA(cudaGetLastError());
    }
    return 0;
}

static Args make_args(const GruLayout& g, int Mold, int tiles, unsigned char* img, const uint8_t* sigT, float* hstore, double* la_sel,
                      double* la_oth, double* lp, double* delta, int* counter, int& sms) {
    int dev = 0;
    sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    Args a;
    memset(&a, 0, sizeof(a));
    a.g = g; a.t = make_layout(g); a.Mold = Mold;
    a.rows_total = (int64_t)tiles * Mold;
    a.tiles128 = (int)cdiv(a.rows_total, kRows);
    a.img = img; a.sigT = sigT; a.hstore = hstore; a.la_sel = la_sel; a.la_oth = la_oth; a.lp = lp; a.delta = delta; a.counter = counter;
    a.nslots = g.N;
    return a;
}

// base pass + single-flip chains (replaces launch_forward<STASH> + launch_chain for the FP32 pRNN with 50 units)
static int launch_eloc(const GruLayout& g, int Mold, int tiles, const float* params, unsigned char* img, const uint8_t* sigT, float* hstore,
                       double* la_sel, double* la_oth, double* lp, double* delta, int* counter, bool flips, cudaStream_t s) {
    int sms;
    Args a = make_args(g, Mold, tiles, img, sigT, hstore, la_sel, la_oth, lp, delta, counter, sms);
    prof_count(); pack_kernel<<<148, 256, 0, s>>>(g, a.t, params, img);
    return launch_chains<false>(a, sms, flips, s);
}

// base pass + NN / NNN exchange chains of the complex cRNN (J1-J2); `order` lists the 2N-3 slots by decreasing chain length
static int launch_j1j2(const GruLayout& g, int Mold, int tiles, const float* params, unsigned char* img, const uint8_t* sigT, float* hstore,
                       double* la_sel, double* la_oth, double* ph_sel, double* ph_oth, double* lp_re, double* lp_im, double* delta_re,
                       double* delta_im, const int* order, const double* j1, const double* j2, int* counter, cudaStream_t s) {
    int sms;
    Args a = make_args(g, Mold, tiles, img, sigT, hstore, la_sel, la_oth, lp_re, delta_re, counter, sms);
    a.ph_sel = ph_sel; a.ph_oth = ph_oth; a.lp_im = lp_im; a.delta_im = delta_im; a.order = order; a.j1 = j1; a.j2 = j2;
    a.n_kind1 = g.N - 1; a.n_kind2 = g.N - 2; a.nslots = a.n_kind1 + a.n_kind2;
    prof_count(); pack_kernel<<<148, 256, 0, s>>>(g, a.t, params, img);
    return launch_chains<true>(a, sms, true, s);
}

#endif  // RNNWF_LEGACY
}  // namespace tc16
}  // namespace rnnwf
