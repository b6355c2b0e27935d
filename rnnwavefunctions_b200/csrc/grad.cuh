// grad.cuh — K3: VMC gradient  sum_s [ w_re[s] d Re log psi_s + w_im[s] d Im log psi_s ]  by BPTT.
//
// Replaces optimizer.compute_gradients(cost) (1DTFIM/TrainingRNN_1DTFIM.py:156-160; complex form
// J1J2/TrainingRNN_J1J2.py:197-201).  Structure:
//   1. teacher-forced base pass with stash (gru_forward_kernel<STASH>): every layer's state per site.
//   2. for layer l = L-1 .. 0: gru_bwd_layer_kernel walks the sites backwards on a tile of M samples,
//      recomputes the gates from the stashed states, propagates d h through time (registers) and to the
//      layer below (dxbuf in HBM), and writes the gate gradients G = [da_r, da_u, da_c, dq] to HBM.
//   3. wgrad_kernel: split-K reduction  dW = sum_{sample,site} [x; h_prev; 1]^T G  with FP32 FMAs per
//      (tile, site) block and FP64 accumulation across blocks (deterministic; no atomics).
// Included at the end of gru.cu (same translation unit as the forward launchers).
#pragma once
#include "gru_kernels.cuh"
#include "host_util.cuh"
#include "umma.cuh"
#include "wgrad_f64mma.cuh"

namespace rnnwf {

struct BwdLaunch {
    int CT, RT, M, Mp, NT, w_smem, smem_bytes;
    int tc;   // backward recurrence on the tensor cores (gru_tc16b.cuh): M is then that kernel's tile, not RT * SPT
};

struct GruLayoutT {   // transposed packed weights for the backward GEMMs
    int off_h[kMaxLayers], off_x[kMaxLayers], total;
};

inline GruLayoutT make_gru_layout_T(const GruLayout& g) {
    GruLayoutT t;
    memset(&t, 0, sizeof(t));
    int o = 0;
    for (int l = 0; l < g.L; ++l) {
        t.off_h[l] = o;
        o += align4(3 * g.H * g.CT * 2);
        t.off_x[l] = o;
        if (l > 0) o += align4(3 * g.H * g.CT * 2);
    }
    t.total = o;
    return t;
}

// WT_h[(gate*H + j)][ct][u] = W_h,gate[i = 2ct+u][j]   (gate 0:r, 1:u from Kg rows d.., 2: Kch)
// WT_x[(gate*H + j)][ct][u] = W_x,gate[i = 2ct+u][j]   (gate 0:r, 1:u from Kg rows 0..d-1, 2: Kci)
template <typename T>
__global__ void pack_gru_T_kernel(GruLayout g, GruLayoutT t, const T* __restrict__ flat, T* __restrict__ pkT) {
    const int H = g.H, CT = g.CT;
    const int per = 3 * H * CT * 2;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < t.total; idx += gridDim.x * blockDim.x) {
        int l = 0;
        while (l + 1 < g.L && idx >= t.off_h[l + 1]) ++l;
        const int d = g.d[l];
        const T* Kg = flat + g.flat_off[l];
        const T* Kci = Kg + (d + H) * 2 * H + 2 * H;
        const T* Kch = Kci + d * H;
        T val = T(0);
        int loc = idx - t.off_h[l];
        const bool xpart = l > 0 && idx >= t.off_x[l];
        if (xpart) loc = idx - t.off_x[l];
        if (loc < per) {
            const int u = loc & 1, ct = (loc >> 1) % CT, kk = (loc >> 1) / CT, gate = kk / H, j = kk % H, i = 2 * ct + u;
            if (!xpart) {
                if (i < H) val = gate == 0 ? Kg[(d + i) * 2 * H + j] : gate == 1 ? Kg[(d + i) * 2 * H + H + j] : Kch[i * H + j];
            } else {
                if (i < d) val = gate == 0 ? Kg[i * 2 * H + j] : gate == 1 ? Kg[i * 2 * H + H + j] : Kci[i * H + j];
            }
        }
        pkT[idx] = val;
    }
}

// per-row weights: PROB: roww[R] = w[s] (parity: split between the two directions by their share of P_sym)
//                  COMPLEX: roww[R] = w_re[s], roww[rows + R] = w_im[s]
__global__ void row_weight_kernel(const double* __restrict__ w, const double* __restrict__ lp, int64_t ns, int M, int tiles_s,
                                  int parity, int cplx, double* __restrict__ roww) {
    const int64_t rows_dir = (int64_t)tiles_s * M;
    const int ndir = parity ? 2 : 1;
    const int64_t rows = rows_dir * ndir;
    for (int64_t R = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; R < rows; R += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = R % rows_dir;
        double a = 0.0, b = 0.0;
        if (s < ns) {
            if (cplx) { a = w[2 * s]; b = w[2 * s + 1]; }
            else if (!parity) a = w[s];
            else {
                const double l_this = lp[R], l_other = lp[R < rows_dir ? R + rows_dir : R - rows_dir];
                a = w[s] / (1.0 + exp(l_other - l_this));
            }
        }
        roww[R] = a;
        if (cplx) roww[rows + R] = b;
    }
}

template <typename T, bool WSMEM, bool CPLX>
__global__ void __launch_bounds__(384, 1)
gru_bwd_layer_kernel(GruLayout g, GruLayoutT gt, BwdLaunch c, int l, const T* __restrict__ pk, const T* __restrict__ pkT,
                     const uint8_t* __restrict__ sigT, const T* __restrict__ hstore, const double* __restrict__ la_sel,
                     const double* __restrict__ la_oth, const double* __restrict__ ph_sel, const double* __restrict__ roww,
                     int64_t rows_total, T* __restrict__ dxbuf, T* __restrict__ Gbuf, T* __restrict__ dzbuf) {
    constexpr int SPT = VT<T>::SPT;
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, H = g.H, CT = g.CT, L = g.L, N = g.N, M = c.M, d = g.d[l];
    const bool top = l == L - 1;
    const int per = align4(3 * H * CT * 2);
    size_t off = 0;
    const T* wl;
    const T *WTh, *WTx = nullptr;
    if (WSMEM) {
        T* wsm = reinterpret_cast<T*>(smem);
        const T* src = pk + g.pk_off[l];
        for (int i = tid; i < g.pk_size[l]; i += blockDim.x) wsm[i] = src[i];
        wl = wsm;
        T* th = wsm + g.pk_size[l];
        for (int i = tid; i < per; i += blockDim.x) th[i] = pkT[gt.off_h[l] + i];
        WTh = th;
        int words = g.pk_size[l] + per;
        if (l > 0) {
            T* tx = th + per;
            for (int i = tid; i < per; i += blockDim.x) tx[i] = pkT[gt.off_x[l] + i];
            WTx = tx;
            words += per;
        }
        off = ((size_t)words * sizeof(T) + 15) & ~(size_t)15;
    } else {
        wl = pk + g.pk_off[l];
        WTh = pkT + gt.off_h[l];
        if (l > 0) WTx = pkT + gt.off_x[l];
    }
    T* xs = reinterpret_cast<T*>(smem + off);
    off += (size_t)(l > 0 ? d : 0) * M * sizeof(T);
    T* hp = reinterpret_cast<T*>(smem + off);
    off += (size_t)H * M * sizeof(T);
    T* G = reinterpret_cast<T*>(smem + off);
    off += (size_t)4 * H * M * sizeof(T);
    T* dz = reinterpret_cast<T*>(smem + off);       // two buffers of 4 * M
    off += (size_t)8 * M * sizeof(T);
    uint8_t* codes = smem + off;

    const bool is_compute = tid < CT * c.RT;
    const int ct = tid % CT, rt = tid / CT, row0 = rt * SPT;
    const int st = blockIdx.x;
    constexpr int NZ = CPLX ? 4 : 2;
    // head weights of this thread's two units (top layer only)
    T wd[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    if (top && is_compute) {
        const T* hw = pk + g.pk_head;
        for (int u = 0; u < 2; ++u) {
            const int j = 2 * ct + u;
            if (j < H) {
                wd[u][0] = hw[2 * j]; wd[u][1] = hw[2 * j + 1];
                if (CPLX) { wd[u][2] = hw[2 * H + 2 + 2 * j]; wd[u][3] = hw[2 * H + 2 + 2 * j + 1]; }
            }
        }
    }
    T carry[2][SPT];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int s = 0; s < SPT; ++s) carry[u][s] = T(0);
    __syncthreads();

    // (a) staging of a site's inputs, one site ahead of the arithmetic: h_prev = h^l_{n-1}, x = h^{l-1}_n (or the one-hot code of
    // sigma_{n-1}) arrive by cp.async while phase (e) of the site before runs (xs / hp / codes are only read in phase (c)); the head
    // gradients dz of the top layer are double-buffered.  (ncu on the synchronous version: long_scoreboard 1.85 stalls per issue.)
    auto stage_inputs = [&](int n) {
        const size_t blk = (size_t)st * N + n;
        const uint32_t hp_s = (uint32_t)__cvta_generic_to_shared(hp), xs_s = (uint32_t)__cvta_generic_to_shared(xs);
        const uint32_t cd_s = (uint32_t)__cvta_generic_to_shared(codes);
        if (n > 0) {
            const char* hsrc = reinterpret_cast<const char*>(hstore + ((blk - 1) * L + l) * H * M);
            for (int i = tid; i < (int)(H * M * sizeof(T) / 16); i += blockDim.x)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(hp_s + 16 * i), "l"(hsrc + 16 * (size_t)i) : "memory");
        } else {
            for (int i = tid; i < H * M; i += blockDim.x) hp[i] = T(0);
        }
        if (l > 0) {
            const char* xsrc = reinterpret_cast<const char*>(hstore + (blk * L + (l - 1)) * H * M);
            for (int i = tid; i < (int)(d * M * sizeof(T) / 16); i += blockDim.x)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(xs_s + 16 * i), "l"(xsrc + 16 * (size_t)i) : "memory");
        } else if (n > 0) {
            const uint8_t* csrc = sigT + (blk - 1) * M;
            for (int i = tid; i < M / 4; i += blockDim.x)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(cd_s + 4 * i), "l"(csrc + 4 * (size_t)i) : "memory");
        } else {
            for (int m = tid; m < M; m += blockDim.x) codes[m] = (uint8_t)2;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto stage_dz = [&](int n, T* dzd) {
        const size_t blk = (size_t)st * N + n;
        for (int m = tid; m < M; m += blockDim.x) {
            const int sg = sigT[blk * M + m];
            const double lo = la_oth[blk * M + m];
            const double wre = roww[(size_t)st * M + m];
            T z[4] = {0, 0, 0, 0};
            if (!CPLX) {
                const double t = wre * exp(lo);        // w (1 - p_sel) = w p_oth
                z[sg] = (T)t;
                z[1 - sg] = (T)(-t);
            } else {
                const double t = 0.5 * wre * exp(2.0 * lo);   // d(1/2 log p_sel)/dz ; 0 when the other outcome is masked
                z[sg] = (T)t;
                z[1 - sg] = (T)(-t);
                const double y = fabs(ph_sel[blk * M + m]) / kPi;
                z[2 + sg] = (T)(roww[rows_total + (size_t)st * M + m] * kPi * (1.0 - y) * (1.0 - y));
            }
#pragma unroll
            for (int o = 0; o < NZ; ++o) {
                dzd[o * M + m] = z[o];
                dzbuf[(blk * NZ + o) * M + m] = z[o];
            }
        }
    };
    T* dz_cur = dz;
    T* dz_nxt = dz + 4 * M;
    stage_inputs(N - 1);
    if (top) stage_dz(N - 1, dz_cur);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();

    for (int n = N - 1; n >= 0; --n) {
        const size_t blk = (size_t)st * N + n;
        T cdir[2][SPT];
        if (is_compute) {   // (c) recompute gates, local backward, publish gate gradients
            T ar[2][SPT], au[2][SPT], ac[2][SPT], aq[2][SPT];
            T dpre[2][SPT];                     // d h^l_n from the layer above: loaded ahead of the gate GEMM that hides its latency
            if (!top) {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int j = 2 * ct + u;
                    if (j < H) ldv<SPT>(dpre[u], dxbuf + (blk * H + j) * M + row0);
                }
            }
            gru_preact<T>(g, l, wl, l == 0 ? nullptr : xs, hp, codes, M, ct, rt, ar, au, ac, aq);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int j = 2 * ct + u;
                if (j < H) {
                    T hold[SPT], dout[SPT], g0[SPT], g1[SPT], g2[SPT], g3[SPT];
                    ldv<SPT>(hold, hp + j * M + row0);
                    if (top) {
#pragma unroll
                        for (int s = 0; s < SPT; ++s) {
                            T v = dz_cur[row0 + s] * wd[u][0] + dz_cur[M + row0 + s] * wd[u][1];
                            if (CPLX) v += dz_cur[2 * M + row0 + s] * wd[u][2] + dz_cur[3 * M + row0 + s] * wd[u][3];
                            dout[s] = v;
                        }
                    } else {
#pragma unroll
                        for (int s = 0; s < SPT; ++s) dout[s] = dpre[u][s];
                    }
#pragma unroll
                    for (int s = 0; s < SPT; ++s) {
                        const T r = sigmoid_(ar[u][s]);
                        const T uu = sigmoid_(au[u][s]);
                        const T q = aq[u][s];
                        const T cc = tanh_(fma(r, q, ac[u][s]));
                        const T dh = carry[u][s] + dout[s];
                        const T dcc = dh * (T(1) - uu);
                        const T duu = dh * (hold[s] - cc);
                        cdir[u][s] = dh * uu;
                        const T dac = dcc * (T(1) - cc * cc);
                        g0[s] = dac * q * r * (T(1) - r);
                        g1[s] = duu * uu * (T(1) - uu);
                        g2[s] = dac;
                        g3[s] = dac * r;
                    }
                    stv<SPT>(G + (0 * H + j) * M + row0, g0);
                    stv<SPT>(G + (1 * H + j) * M + row0, g1);
                    stv<SPT>(G + (2 * H + j) * M + row0, g2);
                    stv<SPT>(G + (3 * H + j) * M + row0, g3);
                    T* gb = Gbuf + blk * 4 * H * M;
                    stv<SPT>(gb + (0 * H + j) * M + row0, g0);
                    stv<SPT>(gb + (1 * H + j) * M + row0, g1);
                    stv<SPT>(gb + (2 * H + j) * M + row0, g2);
                    stv<SPT>(gb + (3 * H + j) * M + row0, g3);
                }
            }
        }
        __syncthreads();
        if (n > 0) {        // inputs of site n - 1: in flight while (e) runs
            stage_inputs(n - 1);
            if (top) stage_dz(n - 1, dz_nxt);
        }
        if (is_compute) {   // (e) d h_{n-1} and d x through the transposed weights
            T accH[2][SPT], accX[2][SPT];
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int s = 0; s < SPT; ++s) { accH[u][s] = T(0); accX[u][s] = T(0); }
            for (int gate = 0; gate < 3; ++gate) {
                const T* Gh = G + (gate == 2 ? 3 : gate) * H * M + row0;
                const T* Gx = G + gate * H * M + row0;
                const T* wh = WTh + (size_t)gate * H * CT * 2 + ct * 2;
                const T* wx = l > 0 ? WTx + (size_t)gate * H * CT * 2 + ct * 2 : nullptr;
#pragma unroll 2
                for (int j = 0; j < H; ++j) {
                    T a[SPT], w[2];
                    ldv<SPT>(a, Gh + j * M);
                    ldv<2>(w, wh + j * CT * 2);
#pragma unroll
                    for (int s = 0; s < SPT; ++s) {
                        accH[0][s] = fma(a[s], w[0], accH[0][s]);
                        accH[1][s] = fma(a[s], w[1], accH[1][s]);
                    }
                    if (l > 0) {
                        T b[SPT], v[2];
                        if (gate == 2) ldv<SPT>(b, Gx + j * M);
                        ldv<2>(v, wx + j * CT * 2);
#pragma unroll
                        for (int s = 0; s < SPT; ++s) {
                            const T bb = gate == 2 ? b[s] : a[s];
                            accX[0][s] = fma(bb, v[0], accX[0][s]);
                            accX[1][s] = fma(bb, v[1], accX[1][s]);
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int j = 2 * ct + u;
                if (j < H) {
#pragma unroll
                    for (int s = 0; s < SPT; ++s) carry[u][s] = cdir[u][s] + accH[u][s];
                    if (l > 0 && j < d) stv<SPT>(dxbuf + (blk * H + j) * M + row0, accX[u]);
                }
            }
        }
        // the staging above only writes xs / hp / codes / the other dz buffer, which nobody reads in (e); this barrier makes them
        // visible and orders (e) before the next (c) rewrites G.
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
        { T* tmp = dz_cur; dz_cur = dz_nxt; dz_nxt = tmp; }
    }
}

// ---------------------------------------------------------------------------------------------
// split-K weight-gradient reduction:  C[r][c] = sum_{blk} sum_m A[blk][r][m] * B[blk][c][m]
//   A rows: [x rows (rows0) | h rows (rows1) | ones],  B = gate gradients (cols) of the same (tile, site)
// ---------------------------------------------------------------------------------------------
template <typename T> struct WgradArgs {
    const T* hstore;       // [blk][L][H][M]
    const uint8_t* sigT;   // [blk][M]
    const T* B;            // [blk][cols][M]
    int L, H, M, N;
    int xmode;             // 0: no x rows, 1: x = hstore layer lx of the same block, 2: one-hot of sigma_{n-1}
    int lx, rows0;
    int lh, hshift, rows1; // h rows: hstore layer lh of block (blk - hshift), zero at n == 0 when hshift
    int cols;
    int64_t nblk;
    int ksplit, rtiles, ctiles;
};

constexpr int kWgTile = 64;

template <typename T>
__global__ void __launch_bounds__(256) wgrad_kernel(WgradArgs<T> a, double* __restrict__ partial) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int M = a.M, MS = M + 4;
    T* As = reinterpret_cast<T*>(smem);
    T* Bs = As + kWgTile * MS;
    const int tile = blockIdx.x, ks = blockIdx.y;
    const int r0 = (tile / a.ctiles) * kWgTile, c0 = (tile % a.ctiles) * kWgTile;
    const int R = a.rows0 + a.rows1 + 1;
    const int tc = threadIdx.x % 16, tr = threadIdx.x / 16;
    const int64_t b0 = a.nblk * ks / a.ksplit, b1 = a.nblk * (ks + 1) / a.ksplit;
    double accd[4][4];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) accd[i][j] = 0.0;
    for (int64_t blk = b0; blk < b1; ++blk) {
        const int n = (int)(blk % a.N);
        for (int i = threadIdx.x; i < kWgTile * (M / 4); i += blockDim.x) {
            const int row = i / (M / 4), m4 = (i % (M / 4)) * 4;
            const int r = r0 + row, cc = c0 + row;
            T va[4] = {0, 0, 0, 0}, vb[4] = {0, 0, 0, 0};
            if (r < a.rows0) {
                if (a.xmode == 1) ldv<4>(va, a.hstore + ((blk * a.L + a.lx) * a.H + r) * M + m4);
                else if (n > 0) {
                    for (int q = 0; q < 4; ++q) va[q] = a.sigT[(blk - 1) * M + m4 + q] == r ? T(1) : T(0);
                }
            } else if (r < a.rows0 + a.rows1) {
                if (!(a.hshift && n == 0)) ldv<4>(va, a.hstore + (((blk - a.hshift) * a.L + a.lh) * a.H + (r - a.rows0)) * M + m4);
            } else if (r == R - 1) {
                va[0] = va[1] = va[2] = va[3] = T(1);
            }
            if (cc < a.cols) ldv<4>(vb, a.B + (blk * a.cols + cc) * M + m4);
            stv<4>(As + row * MS + m4, va);
            stv<4>(Bs + row * MS + m4, vb);
        }
        __syncthreads();
        T acc[4][4];
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
        for (int m4 = 0; m4 < M; m4 += 4) {
            T av[4][4], bv[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) ldv<4>(av[i], As + (tr * 4 + i) * MS + m4);
#pragma unroll
            for (int j = 0; j < 4; ++j) ldv<4>(bv[j], Bs + (tc * 4 + j) * MS + m4);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[i][j] = fma(av[i][q], bv[j][q], acc[i][j]);
        }
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) accd[i][j] += (double)acc[i][j];
        __syncthreads();
    }
    const int Rp = a.rtiles * kWgTile, Cp = a.ctiles * kWgTile;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            partial[((size_t)ks * Rp + r0 + tr * 4 + i) * Cp + c0 + tc * 4 + j] = accd[i][j];
}

// FP32 fast path of the same reduction: one block computes the WHOLE [R x cols] output for its chunk of (tile, site)
// blocks, so every A / B element is read from HBM once (the tiled kernel above re-reads them rtiles*ctiles times).
// 288 threads, thread tile 8 x 10 (80 FP32 accumulators); every kWgFlush blocks the tile is flushed into the block's own
// FP64 partial (deterministic: no atomics, fixed order).
constexpr int kWgfThreads = 288, kWgfTr = 8, kWgfTc = 10, kWgFlush = 32;

__global__ void __launch_bounds__(kWgfThreads) wgrad_fast_kernel(WgradArgs<float> a, double* __restrict__ partial, int Rp, int Cp) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int M = a.M, MS = M + 4;
    const int R = a.rows0 + a.rows1 + 1;
    const int RG = (R + kWgfTr - 1) / kWgfTr, CG = (a.cols + kWgfTc - 1) / kWgfTc;
    float* As = reinterpret_cast<float*>(smem);                     // [RG * 8][MS]
    float* Bs = As + (size_t)RG * kWgfTr * MS;                       // [CG * 10][MS]
    const int tid = threadIdx.x;
    const bool active = tid < RG * CG;
    const int tr = tid / CG, tc = tid % CG;
    const int ks = blockIdx.x;
    const int64_t b0 = a.nblk * ks / a.ksplit, b1 = a.nblk * (ks + 1) / a.ksplit;
    double* out = partial + (size_t)ks * Rp * Cp;
    for (int i = tid; i < Rp * Cp; i += blockDim.x) out[i] = 0.0;
    float acc[kWgfTr][kWgfTc];
#pragma unroll
    for (int i = 0; i < kWgfTr; ++i)
#pragma unroll
        for (int j = 0; j < kWgfTc; ++j) acc[i][j] = 0.f;
    int pending = 0;
    __syncthreads();
    for (int64_t blk = b0; blk < b1; ++blk) {
        const int n = (int)(blk % a.N);
        for (int i = tid; i < (RG * kWgfTr + CG * kWgfTc) * (M / 4); i += blockDim.x) {
            const int row = i / (M / 4), m4 = (i % (M / 4)) * 4;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (row < RG * kWgfTr) {
                const int r = row;
                if (r < a.rows0) {
                    if (a.xmode == 1) ldv<4>(v, a.hstore + ((blk * a.L + a.lx) * a.H + r) * M + m4);
                    else if (n > 0) {
                        for (int q = 0; q < 4; ++q) v[q] = a.sigT[(blk - 1) * M + m4 + q] == r ? 1.f : 0.f;
                    }
                } else if (r < a.rows0 + a.rows1) {
                    if (!(a.hshift && n == 0)) ldv<4>(v, a.hstore + (((blk - a.hshift) * a.L + a.lh) * a.H + (r - a.rows0)) * M + m4);
                } else if (r == R - 1) {
                    v[0] = v[1] = v[2] = v[3] = 1.f;
                }
                stv<4>(As + (size_t)row * MS + m4, v);
            } else {
                const int c = row - RG * kWgfTr;
                if (c < a.cols) ldv<4>(v, a.B + (blk * a.cols + c) * M + m4);
                stv<4>(Bs + (size_t)c * MS + m4, v);
            }
        }
        __syncthreads();
        if (active) {
            const float* ap = As + (size_t)tr * kWgfTr * MS;
            const float* bp = Bs + (size_t)tc * kWgfTc * MS;
            for (int m4 = 0; m4 < M; m4 += 4) {
                float av[kWgfTr][4], bv[kWgfTc][4];
#pragma unroll
                for (int i = 0; i < kWgfTr; ++i) ldv<4>(av[i], ap + (size_t)i * MS + m4);
#pragma unroll
                for (int j = 0; j < kWgfTc; ++j) ldv<4>(bv[j], bp + (size_t)j * MS + m4);
#pragma unroll
                for (int i = 0; i < kWgfTr; ++i)
#pragma unroll
                    for (int j = 0; j < kWgfTc; ++j)
#pragma unroll
                        for (int q = 0; q < 4; ++q) acc[i][j] = fmaf(av[i][q], bv[j][q], acc[i][j]);
            }
        }
        if (++pending == kWgFlush || blk + 1 == b1) {
            if (active) {
#pragma unroll
                for (int i = 0; i < kWgfTr; ++i)
#pragma unroll
                    for (int j = 0; j < kWgfTc; ++j) {
                        const int r = tr * kWgfTr + i, c = tc * kWgfTc + j;
                        if (r < Rp && c < Cp) out[(size_t)r * Cp + c] += (double)acc[i][j];
                        acc[i][j] = 0.f;
                    }
            }
            pending = 0;
        }
        __syncthreads();
    }
}

// Tensor-core path of the same reduction (FP32 models): tcgen05.mma kind::tf32 with 3xTF32 operands (hi * hi + hi * lo + lo * hi,
// FP32 accumulate in TMEM: FP32-grade products -- a single TF32 pass does not survive the cancellation in
// sum_s (E_s - mean) d log psi_s).  Both operands are K-major with K = the samples of a (tile, site) block, which is how the stash
// and the gate gradients already lie in HBM ([row][M], M contiguous): A = [x; h; 1] (R <= 128 rows), B = G (cols <= 256 rows), D =
// A B^T [128 x cols] stays in TMEM across the blocks of this CTA and is flushed into its FP64 partial every kWgFlush blocks
// (deterministic: no atomics, fixed order).  Per block: the threads split the operands they prefetched into registers during the
// previous block's MMAs into (hi, lo) core-matrix images in shared memory, one thread issues 3 MMAs per 8 samples, and the global
// loads of the next block are in flight while the tensor pipe works.  (The warp-level mma.sync m16n8k8 TF32 path was measured first:
// 94 ms against the 74 ms of the FFMA kernel at cfg2 -- on sm_100a it runs at ~4x the FFMA rate, which three passes eat.)
// Sub-blocks: the samples of a (tile, site) block are taken in nsub slices of Ms (68 rows: 36 + 32).  The images of a slice are
// half as large (110 instead of 198 KB at cfg2), so TWO CTAs share an SM: one splits / stores its operands or waits for its loads
// while the tensor pipe works for the other, and twice as many loads are in flight (the one-CTA version was bound by the latency of
// its one-block-deep register prefetch: ncu long_scoreboard 4.3 stalls per issue, tensor pipe 22 %, 1.87 TB/s of DRAM reads).
#ifndef RNNWF_WGRAD_ASYNC
#define RNNWF_WGRAD_ASYNC 1   // 1: operands go global -> shared by cp.async straight into the hi images (below); 0: register prefetch + split
#endif
namespace wgtc {
constexpr int kThreads = 256, kMaxItems = 16;

struct Geo {
    int Rp8, Cp8, Np, Kp, M4p, items, nit;      // padded A / B rows, MMA N, padded samples, float4 columns (multiple of 4)
    int Ms, nsub;                               // samples per slice (multiple of 4), slices per block
    size_t img_floats, smem;
};
inline Geo make_geo(int R, int cols, int M, int nsub = 0) {
    Geo q;
    q.Rp8 = (R + 7) / 8 * 8;
    q.Cp8 = (cols + 7) / 8 * 8;
    q.Np = (cols + 15) / 16 * 16;
    if (nsub == 0) {   // whole blocks while two CTAs still share an SM (narrow outputs: the head weights), else two slices
        const Geo whole = make_geo(R, cols, M, 1);
        nsub = (whole.smem <= (size_t)(kSmemLimit / 2 - 1024) && whole.nit <= kMaxItems && whole.img_floats / 4 < 4096) ? 1 : (M > 40 ? 2 : 1);
    }
    q.nsub = nsub;
    q.Ms = ((M + q.nsub - 1) / q.nsub + 3) / 4 * 4;
    q.Kp = (q.Ms + 7) / 8 * 8;
    q.M4p = (q.Ms / 4 + 3) / 4 * 4;
    q.items = (q.Rp8 + q.Cp8) * q.M4p;
    q.nit = (q.items + kThreads - 1) / kThreads;
    q.img_floats = (size_t)(128 + q.Np + 8) * q.Kp;                 // A image | B image (+ one row group of slack)
    q.smem = 2 * q.img_floats * sizeof(float) + 64;
    return q;
}
inline bool supported(int R, int cols, int M) {
    if (R > 128 || cols > 256 || M % 4 != 0) return false;
    const Geo q = make_geo(R, cols, M);
    return q.nit <= kMaxItems && q.smem <= (size_t)kSmemLimit && q.img_floats / 4 < 4096;   // image offsets are 12-bit fields of the item descriptors
}

// what a thread moves per block is fixed for the whole launch: item -> (source kind, element offset relative to the block's base,
// destination in the images) is decoded once, so that the per-block code is ~8 instructions per 16-byte item (the first version
// re-derived rows and addresses per block: 1 700 instructions per thread and block, 42 KB of unrolled code, issue-bound with
// no_instruction stalls)
enum { kSkip = 0, kStash = 1, kGate = 2, kOneHot = 3 };

__global__ void __launch_bounds__(kThreads, 2) wgrad_kernel(WgradArgs<float> a, Geo q, double* __restrict__ partial, int Rp, int Cp) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* img_hi = reinterpret_cast<float*>(smem);
    float* img_lo = img_hi + q.img_floats;
    uint64_t* bar = reinterpret_cast<uint64_t*>(img_lo + q.img_floats);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int R = a.rows0 + a.rows1 + 1, KC = q.Kp / 4, M = a.M, Ms = q.Ms, nsub = q.nsub;
    const int ks = blockIdx.x;
    const int64_t nvb = a.nblk * nsub;                                                  // (block, slice) pairs
    const int64_t b0 = nvb * ks / a.ksplit, b1 = nvb * (ks + 1) / a.ksplit;
    double* out = partial + (size_t)ks * Rp * Cp;
    for (int i = tid; i < Rp * Cp; i += blockDim.x) out[i] = 0.0;
    for (size_t i = tid; i < 2 * q.img_floats; i += blockDim.x) img_hi[i] = 0.f;       // padding rows / samples stay zero
    __syncthreads();
    for (int i = tid; i < Ms; i += blockDim.x) {                                       // the constant-1 row (biases) never changes
        const int irow = R - 1;
        img_hi[(size_t)(irow >> 3) * (KC * 32) + (size_t)(i >> 2) * 32 + (irow & 7) * 4 + (i & 3)] = 1.0f;
    }
    if (warp == 0) umma::tmem_alloc(tmem_slot, 256);
    if (tid == 0) { umma::mbar_init(bar, 1); umma::mbar_fence_init(); }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tbase = *tmem_slot;
    const uint32_t sA_hi = umma::smem_u32(img_hi), sB_hi = sA_hi + 128u * q.Kp * 4u;
    const uint32_t sA_lo = umma::smem_u32(img_lo), sB_lo = sA_lo + 128u * q.Kp * 4u;
    const uint32_t idesc = umma::instr_desc(umma::kFmtTF32, 128, q.Np);
    // item -> (row, float4 column): 8 consecutive lanes take the 8 rows of a core-matrix row group (one conflict-free 128-byte store),
    // the 4 lane groups of a warp take 4 consecutive float4 columns (64 contiguous bytes of every row in global memory)
    const int RG = (q.Rp8 + q.Cp8) / 8;
    // one register per item (two CTAs share an SM: 128 registers per thread; with separate offset / descriptor arrays ptxas spilled
    // them and every prefetch waited for local-memory loads): kind [0, 2) | zero-at-site-0 [2] | beyond the tile in the last slice [3] |
    // one-hot row [4] | image offset in float4 [5, 17) | signed element offset from the block's base [17, 32)
    int dsc[kMaxItems];
#pragma unroll
    for (int it = 0; it < kMaxItems; ++it) {
        const int i = tid + it * kThreads;
        dsc[it] = kSkip;
        if (it < q.nit && i < q.items) {
            const int b32 = i >> 5, row = (b32 % RG) * 8 + (i & 7), m4 = ((b32 / RG) * 4 + ((i >> 3) & 3)) * 4;
            if (m4 < Ms) {
                const int irow = row < q.Rp8 ? row : 128 + (row - q.Rp8);
                const int o4 = (int)(((size_t)(irow >> 3) * (KC * 32) + (size_t)(m4 >> 2) * 32 + (irow & 7) * 4) >> 2);
                int kind = kSkip, flags = 0, off = 0;
                if (row < q.Rp8) {
                    const int r = row;
                    if (r < a.rows0) {
                        if (a.xmode == 1) { kind = kStash; off = (a.lx * a.H + r) * M + m4; }
                        else { kind = kOneHot; flags = r << 4; off = m4 - M; }                 // sigT of block blk - 1
                    } else if (r < a.rows0 + a.rows1) {
                        kind = kStash;
                        off = ((a.lh - a.hshift * a.L) * a.H + (r - a.rows0)) * M + m4;        // hstore of block blk - hshift
                        flags = a.hshift ? 4 : 0;
                    }
                } else if (row - q.Rp8 < a.cols) {
                    kind = kGate;
                    off = (row - q.Rp8) * M + m4;
                }
                dsc[it] = kind | flags | ((nsub - 1) * Ms + m4 >= M ? 8 : 0) | (o4 << 5) | (int)((uint32_t)off << 17);
            }
        }
    }
    const int64_t hspan = (int64_t)a.L * a.H * M, gspan = (int64_t)a.cols * M;
#if RNNWF_WGRAD_ASYNC
    // Operand path without registers (round 2): kind::tf32 reads the upper 19 bits of an FP32 word and ignores the rest
    // (scripts/tf32_raw_probe.py: the product of raw operands equals the product of the TRUNCATED operands to 6e-8), so the raw data
    // are their own hi image: every thread copies its items global -> shared with cp.async (16 bytes = 4 samples of one row, which is
    // one core-matrix row of the image), then computes lo = x - trunc(x) from shared memory into the lo image.  No load occupies a
    // register while it is in flight; while this CTA waits for its copies the other CTA of the SM works.
    auto copy_in = [&](int64_t vb) {
        const int64_t blk = vb / nsub;
        const int sub = (int)(vb - blk * nsub), m0 = sub * Ms;
        const bool first_site = blk % a.N == 0, last = sub == nsub - 1;
        const float* hb = a.hstore + blk * hspan + m0;
        const float* gb = a.B + blk * gspan + m0;
        const uint8_t* sb = a.sigT + blk * M + m0;
#pragma unroll
        for (int it = 0; it < kMaxItems; ++it) {
            const int kind = dsc[it] & 3;
            if (kind == kSkip) continue;
            float* dst = img_hi + ((size_t)((dsc[it] >> 5) & 0xfff) << 2);
            const int goff = dsc[it] >> 17;
            const float* src = nullptr;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!(last && (dsc[it] & 8))) {
                if (kind == kGate) src = gb + goff;
                else if (kind == kStash) { if (!((dsc[it] & 4) && first_site)) src = hb + goff; }
                else if (!first_site) {                                   // one-hot rows: computed, exact in TF32
                    const uint32_t sg = *reinterpret_cast<const uint32_t*>(sb + goff);
                    const int r = (dsc[it] >> 4) & 1;
                    v.x = (int)(sg & 0xff) == r ? 1.f : 0.f;
                    v.y = (int)((sg >> 8) & 0xff) == r ? 1.f : 0.f;
                    v.z = (int)((sg >> 16) & 0xff) == r ? 1.f : 0.f;
                    v.w = (int)(sg >> 24) == r ? 1.f : 0.f;
                }
            }
            if (src != nullptr) cp_async16(dst, src);
            else *reinterpret_cast<float4*>(dst) = v;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto make_lo = [&]() {
        asm volatile("cp.async.wait_all;" ::: "memory");                  // this thread's own copies: it reads only what it copied
#pragma unroll
        for (int it = 0; it < kMaxItems; ++it) {
            if ((dsc[it] & 3) != kSkip) {
                const size_t o = (size_t)((dsc[it] >> 5) & 0xfff) << 2;
                const float4 x = *reinterpret_cast<const float4*>(img_hi + o);
                float4 lo;
                lo.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
                lo.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
                lo.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
                lo.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
                *reinterpret_cast<float4*>(img_lo + o) = lo;
            }
        }
    };
#else
    float4 pre[kMaxItems];
    auto prefetch = [&](int64_t vb) {
        const int64_t blk = vb / nsub;
        const int sub = (int)(vb - blk * nsub), m0 = sub * Ms;
        const bool first_site = blk % a.N == 0, last = sub == nsub - 1;
        const float* hb = a.hstore + blk * hspan + m0;
        const float* gb = a.B + blk * gspan + m0;
        const uint8_t* sb = a.sigT + blk * M + m0;
#pragma unroll
        for (int it = 0; it < kMaxItems; ++it) {
            const int kind = dsc[it] & 3;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            const int goff = dsc[it] >> 17;
            if (last && (dsc[it] & 8)) { pre[it] = v; continue; }
            if (kind == kGate) v = *reinterpret_cast<const float4*>(gb + goff);
            else if (kind == kStash) { if (!((dsc[it] & 4) && first_site)) v = *reinterpret_cast<const float4*>(hb + goff); }
            else if (kind == kOneHot && !first_site) {
                const uint32_t sg = *reinterpret_cast<const uint32_t*>(sb + goff);
                const int r = (dsc[it] >> 4) & 1;
                v.x = (int)(sg & 0xff) == r ? 1.f : 0.f;
                v.y = (int)((sg >> 8) & 0xff) == r ? 1.f : 0.f;
                v.z = (int)((sg >> 16) & 0xff) == r ? 1.f : 0.f;
                v.w = (int)(sg >> 24) == r ? 1.f : 0.f;
            }
            pre[it] = v;
        }
    };
#endif
    uint32_t commits = 0;
    int pending = 0;
#if !RNNWF_WGRAD_ASYNC
    if (b0 < b1) prefetch(b0);
#endif
    for (int64_t blk = b0; blk < b1; ++blk) {
        if (commits > 0) umma::mbar_wait(bar, (commits - 1) & 1);        // the previous block's MMAs have read the images
#if RNNWF_WGRAD_ASYNC
        copy_in(blk);
        make_lo();
#else
#pragma unroll
        for (int it = 0; it < kMaxItems; ++it) {
            if ((dsc[it] & 3) != kSkip) {
                const size_t o = (size_t)((dsc[it] >> 5) & 0xfff) << 2;
                float4 hi, lo;
                umma::split_tf32_fast(pre[it].x, hi.x, lo.x);      // integer rounding: cvt.rna.tf32 issues at a quarter of the ALU rate
                umma::split_tf32_fast(pre[it].y, hi.y, lo.y);
                umma::split_tf32_fast(pre[it].z, hi.z, lo.z);
                umma::split_tf32_fast(pre[it].w, hi.w, lo.w);
                *reinterpret_cast<float4*>(img_hi + o) = hi;
                *reinterpret_cast<float4*>(img_lo + o) = lo;
            }
        }
#endif
        umma::fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            umma::fence_after_sync();
            const uint32_t lbo = 128, sbo = (uint32_t)KC * 128;
            for (int kb = 0; kb < q.Kp / 8; ++kb) {
                const uint64_t ah = umma::smem_desc(sA_hi + kb * 256, lbo, sbo), al = umma::smem_desc(sA_lo + kb * 256, lbo, sbo);
                const uint64_t bh = umma::smem_desc(sB_hi + kb * 256, lbo, sbo), bl = umma::smem_desc(sB_lo + kb * 256, lbo, sbo);
                umma::mma_tf32_ss(tbase, ah, bh, idesc, (pending > 0 || kb > 0) ? 1u : 0u);
                umma::mma_tf32_ss(tbase, ah, bl, idesc, 1u);
                umma::mma_tf32_ss(tbase, al, bh, idesc, 1u);
            }
            umma::commit(bar);
        }
        ++commits;
#if !RNNWF_WGRAD_ASYNC
        if (blk + 1 < b1) prefetch(blk + 1);                             // in flight while the tensor pipe works
#endif
        if (++pending == kWgFlush * nsub || blk + 1 == b1) {      // the same number of samples per FP32 accumulator as with whole blocks
            umma::mbar_wait(bar, (commits - 1) & 1);
            umma::fence_after_sync();
            const int row = 32 * (warp & 3) + lane, half = q.Np / 2 / 8 * 8;     // warps 0-3: columns [0, half), warps 4-7: the rest
            const int c_lo = warp < 4 ? 0 : half, c_hi = warp < 4 ? half : q.Np;
            for (int c0 = c_lo; c0 < c_hi; c0 += 8) {
                float v[8];
                umma::tmem_ld8(tbase + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)c0, v);
                umma::wait_ld();
                if (row < R && row < Rp) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (c0 + j < a.cols && c0 + j < Cp) out[(size_t)row * Cp + c0 + j] += (double)v[j];
                }
            }
            pending = 0;
            umma::fence_before_sync();
            __syncthreads();                                            // every warp has read its accumulators
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tbase, 256);
}
}  // namespace wgtc


// operands of the GRU reduction for wgdm::wgrad_kernel (wgrad_f64mma.cuh)
struct WgdmGruSrc {
    WgradArgs<double> a;
    int M, R, C;
    int64_t nblk;
    struct Ctx { int n; };
    __device__ __forceinline__ Ctx begin(int64_t blk) const { return Ctx{(int)(blk % a.N)}; }
    __device__ __forceinline__ const double* a_src(const Ctx& cx, int64_t blk, int r, int k, bool& special, double& v0, double& v1) const {
        const int n = cx.n;
        if (r < a.rows0) {
            if (a.xmode == 1) return a.hstore + ((blk * a.L + a.lx) * a.H + r) * M + k;
            special = true;                                                // one-hot input row: sigma of the site before
            if (n > 0) {
                v0 = a.sigT[(blk - 1) * M + k] == r ? 1.0 : 0.0;
                v1 = a.sigT[(blk - 1) * M + k + 1] == r ? 1.0 : 0.0;
            }
            return nullptr;
        }
        if (r < a.rows0 + a.rows1) {
            if (a.hshift && n == 0) return nullptr;
            return a.hstore + (((blk - a.hshift) * a.L + a.lh) * a.H + (r - a.rows0)) * M + k;
        }
        special = true;                                                    // the constant-1 row (biases)
        v0 = v1 = 1.0;
        return nullptr;
    }
    __device__ __forceinline__ const double* b_src(const Ctx&, int64_t blk, int c, int k) const { return a.B + (blk * a.cols + c) * M + k; }
};

// sum the split-K partials (fixed order) and scatter into the flat gradient.
//   layer mode : rows [x (d) | h (H) | ones], cols [da_r (H) | da_u (H) | da_c (H) | dq (H)]
//   head mode  : rows [h_top (H) | ones],     cols [dz (2) | dz_phase (2)]
__global__ void wgrad_scatter_kernel(const double* __restrict__ partial, int ksplit, int Rp, int Cp, GruLayout g, int l, int head,
                                     double* __restrict__ grad) {
    const int H = g.H;
    const int d = head ? 0 : g.d[l];
    const int R = d + H + 1, Ccols = head ? 2 * g.nheads : 4 * H;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < R * Ccols; idx += gridDim.x * blockDim.x) {
        const int r = idx / Ccols, cidx = idx % Ccols;
        int dst = -1;
        if (head) {
            const int hd = cidx / 2, o = cidx % 2;
            const int base = g.flat_head + hd * (2 * H + 2);
            dst = r < H ? base + r * 2 + o : base + 2 * H + o;
        } else {
            const int gate = cidx / H, j = cidx % H;
            const int Kg = g.flat_off[l], bg = Kg + (d + H) * 2 * H, Kci = bg + 2 * H, Kch = Kci + d * H, bci = Kch + H * H,
                      bch = bci + H;
            if (r < d) {
                if (gate < 2) dst = Kg + r * 2 * H + gate * H + j;
                else if (gate == 2) dst = Kci + r * H + j;
            } else if (r < d + H) {
                const int k = r - d;
                if (gate < 2) dst = Kg + (d + k) * 2 * H + gate * H + j;
                else if (gate == 3) dst = Kch + k * H + j;
            } else {
                dst = gate < 2 ? bg + gate * H + j : gate == 2 ? bci + j : bch + j;
            }
        }
        if (dst < 0) continue;
        double s = 0.0;
        for (int k = 0; k < ksplit; ++k) s += partial[((size_t)k * Rp + r) * Cp + cidx];
        grad[dst] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <typename T> inline BwdLaunch choose_bwd_launch(const GruLayout& g, int64_t rows_hint = 0, int ndir = 1) {
    constexpr int SPT = VT<T>::SPT;
    BwdLaunch best;
    memset(&best, 0, sizeof(best));
    const int per = align4(3 * g.H * g.CT * 2);
    int maxpk = 0;
    for (int l = 0; l < g.L; ++l) maxpk = std::max(maxpk, g.pk_size[l]);
    const int dmax = g.L > 1 ? g.H : 0;
    for (int wsm = 1; wsm >= 0; --wsm) {
        double best_eff = -1.0, best_cost = 1e300;
        for (int RT = 1; RT <= 64; ++RT) {
            const int nt = g.CT * RT, M = RT * SPT;
            if (nt > 384 || M > 2 * kHeadThreads) break;
            const int Mp = (M + 15) & ~15;
            size_t smem = (wsm ? (((size_t)(maxpk + 2 * per) * sizeof(T) + 15) & ~(size_t)15) : 0) +
                          (size_t)(dmax + g.H + 4 * g.H + 8) * M * sizeof(T) + Mp + 64;
            // the forward pass of the gradient runs with the same M: all layers' weights + L*H*M state
            size_t fwd = (((size_t)g.PK * sizeof(T) + 15) & ~(size_t)15) + (size_t)g.L * g.H * M * sizeof(T) + 2 * (size_t)Mp + 64;
            if (smem > (size_t)kSmemLimit) break;
            if (fwd > (size_t)kSmemLimit && wsm) break;
            if ((size_t)2 * 64 * (M + 4) * sizeof(T) > (size_t)kSmemLimit) break;   // wgrad_kernel staging tiles (kWgTile = 64)
            bool take;
            if (rows_hint > 0) {
                const double cost = wave_cost(rows_hint, ndir, M);
                take = cost < best_cost || (cost == best_cost && best.RT > 0 && M > best.M);
                if (take) best_cost = cost;
            } else {
                const double eff = (double)nt / (128.0 * (double)((nt + 127) / 128));
                take = eff >= best_eff - 0.03;
                if (eff > best_eff) best_eff = eff;
            }
            if (take) {
                best.CT = g.CT; best.RT = RT; best.M = M; best.Mp = Mp; best.NT = (nt + 31) & ~31;
                best.w_smem = wsm; best.smem_bytes = (int)smem;
            }
        }
        if (best.RT > 0) return best;
    }
    return best;
}

template <typename T> struct GradWs {
    GruWs<T> f;
    T* pkT;
    double* roww;
    T *dxbuf, *Gbuf, *dzbuf;
    double* partial;
    unsigned char* img16;   // weight image of the tensor-core base pass (FP32 probability-head models it supports)
    float* gstore;          // backward factors [tile][site][layer]{[unit][M][4], [unit][M]} written by that pass for the tensor-core recurrence
    unsigned char* img16b;  // transposed-weight images of the tensor-core backward recurrence (gru_tc16b.cuh)
    double* wb64;           // float64 one-layer stacks: operands of the DMMA base pass
    int ksplit, Rp, Cp;
};

template <typename T>
static GradWs<T> carve_grad(Ws& ws, const GruLayout& g, const GruLayoutT& gt, const GruLaunch& cf, int64_t tiles, bool cplx,
                            int64_t ns, bool tc_bwd) {
    GradWs<T> w;
    w.f = carve_gru<T>(ws, g, cf, tiles, true, 0, cplx, ns);
    const size_t rows = (size_t)tiles * cf.M;
    w.pkT = ws.take<T>(gt.total);
    w.roww = ws.take<double>(rows * (cplx ? 2 : 1));
    w.dxbuf = ws.take<T>(g.L > 1 ? rows * g.N * g.H : 0);
    w.Gbuf = ws.take<T>(rows * g.N * 4 * g.H);
    w.dzbuf = ws.take<T>(rows * g.N * (cplx ? 4 : 2));
    const int R = g.H + g.H + 1, C = 4 * g.H;
    w.Rp = (int)cdiv(R, kWgTile) * kWgTile;
    w.Cp = (int)cdiv(C, kWgTile) * kWgTile;
    const int ntile = (w.Rp / kWgTile) * (w.Cp / kWgTile);
    w.ksplit = (int)std::max<int64_t>(1, std::min<int64_t>(std::is_same<T, float>::value ? 296 : (148 * 4 + ntile - 1) / ntile, tiles * g.N));
    w.partial = ws.take<double>((size_t)w.ksplit * w.Rp * w.Cp);
    w.img16 = nullptr;
    if (std::is_same<T, float>::value && !cplx && tc16p::supported(g)) w.img16 = ws.take<unsigned char>(tc16p::make_layout(g).img_bytes);
    w.wb64 = nullptr;        // float64 one-layer stacks: B fragments + table of the DMMA base pass (gru_f64mma.cuh)
    if (std::is_same<T, double>::value && !cplx && f64mma::supported(g))
        w.wb64 = ws.take<double>(f64mma::make_layout(g).wb_doubles + f64mma::make_layout(g).tab_doubles);
    w.gstore = nullptr;
    w.img16b = nullptr;
    if (tc_bwd) {
        w.gstore = ws.take<float>(rows * g.N * g.L * tc16b::kFactors * g.H);
        w.img16b = ws.take<unsigned char>(tc16b::make_layout(1, cf.M).img_bytes);
    }
    return w;
}

template <typename T> static GruLaunch fwd_launch_for(const GruLayout& g, const BwdLaunch& b) {
    GruLaunch c;
    c.CT = b.CT; c.RT = b.RT; c.M = b.M; c.Mp = b.Mp; c.NTc = b.NT;
    size_t with_w = (((size_t)g.PK * sizeof(T) + 15) & ~(size_t)15) + (size_t)g.L * g.H * b.M * sizeof(T) + 2 * (size_t)b.Mp + 64;
    c.w_smem = with_w <= (size_t)kSmemLimit;
    c.ring_kc = c.w_smem ? 0 : kRingKC;       // weights that do not fit are streamed through the shared-memory ring (gru_engine.cuh)
    c.smem_bytes = (int)(c.w_smem ? with_w : (size_t)g.L * g.H * b.M * sizeof(T) + 2 * (size_t)b.Mp + 64 + 2 * (size_t)kRingKC * g.CT * 6 * sizeof(T) + 16);
    return c;
}

// launch geometry of the gradient: the CUDA-core tile, or -- FP32 probability-head stacks the tensor-core kernels cover -- tiles that
// fill the SMs in whole waves for the tensor-core backward recurrence (RNNWF_BWD_FFMA=1 keeps the CUDA-core recurrence: A/B runs)
template <typename T> static BwdLaunch grad_launch(const GruLayout& g, int64_t ns, int ndir, bool cplx) {
    BwdLaunch b = choose_bwd_launch<T>(g, ns, ndir);
    b.tc = 0;
    if (b.RT > 0 && std::is_same<T, float>::value && !cplx && tc16b::supported(g) && chain_mode(g) == 3 && !getenv("RNNWF_BWD_FFMA")) {
        b.M = tc16b::choose_rows(ns * ndir);
        b.Mp = (b.M + 15) & ~15;
        b.tc = 1;
    }
    return b;
}

template <typename T> size_t gru_grad_workspace_bytes(const rnnwf_model& m, int64_t ns, int flags) {
    const GruLayout g = make_gru_layout(m);
    const GruLayoutT gt = make_gru_layout_T(g);
    const int ndir = (flags & RNNWF_PARITY_SYM) ? 2 : 1;
    const BwdLaunch b = grad_launch<T>(g, ns, ndir, m.head == RNNWF_HEAD_COMPLEX);
    if (b.RT == 0) return 0;
    const GruLaunch cf = fwd_launch_for<T>(g, b);
    Ws ws(nullptr, 0);
    carve_grad<T>(ws, g, gt, cf, ndir * cdiv(ns, b.M), m.head == RNNWF_HEAD_COMPLEX, ns, b.tc != 0);
    return ws.used + 256;
}
template size_t gru_grad_workspace_bytes<float>(const rnnwf_model&, int64_t, int);
template size_t gru_grad_workspace_bytes<double>(const rnnwf_model&, int64_t, int);

template <typename T, bool CPLX>
static int launch_bwd_layer(const GruLayout& g, const GruLayoutT& gt, const BwdLaunch& b, int l, const GradWs<T>& w, int tiles,
                            int64_t rows_total, cudaStream_t s) {
    // per-layer smem: this layer's weights + transposed copies + tiles
    if (b.w_smem) {
        auto k = gru_bwd_layer_kernel<T, true, CPLX>;
        if (int e = set_smem(k, b.smem_bytes)) return e;
        prof_count(); k<<<tiles, b.NT, b.smem_bytes, s>>>(g, gt, b, l, w.f.pk, w.pkT, w.f.sigT, w.f.hstore, w.f.la_sel, w.f.la_oth, w.f.ph_sel, w.roww,
                                            rows_total, w.dxbuf, w.Gbuf, w.dzbuf);
    } else {
        auto k = gru_bwd_layer_kernel<T, false, CPLX>;
        if (int e = set_smem(k, b.smem_bytes)) return e;
        prof_count(); k<<<tiles, b.NT, b.smem_bytes, s>>>(g, gt, b, l, w.f.pk, w.pkT, w.f.sigT, w.f.hstore, w.f.la_sel, w.f.la_oth, w.f.ph_sel, w.roww,
                                            rows_total, w.dxbuf, w.Gbuf, w.dzbuf);
    }
    RNNWF_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
static int launch_wgrad(const GruLayout& g, const GradWs<T>& w, int M, int64_t nblk, int l, bool head, double* grad, cudaStream_t s) {
    WgradArgs<T> a;
    memset(&a, 0, sizeof(a));
    a.hstore = w.f.hstore; a.sigT = w.f.sigT; a.L = g.L; a.H = g.H; a.M = M; a.N = g.N; a.nblk = nblk;
    if (head) {
        a.B = w.dzbuf; a.xmode = 0; a.rows0 = 0; a.lh = g.L - 1; a.hshift = 0; a.rows1 = g.H; a.cols = 2 * g.nheads;
    } else {
        a.B = w.Gbuf; a.xmode = l > 0 ? 1 : 2; a.lx = l - 1; a.rows0 = g.d[l]; a.lh = l; a.hshift = 1; a.rows1 = g.H; a.cols = 4 * g.H;
    }
    const int R = a.rows0 + a.rows1 + 1;
    a.rtiles = (int)cdiv(R, kWgTile);
    a.ctiles = (int)cdiv(a.cols, kWgTile);
    a.ksplit = w.ksplit;
    int ksplit_used = w.ksplit;
    bool fast = false;
    if constexpr (std::is_same<T, float>::value) {
        const bool offs_fit = (int64_t)g.L * g.H * M < 16384 && (int64_t)a.cols * M < 16384;   // 15-bit signed element offsets in the item descriptors
        if (wgtc::supported(R, a.cols, M) && offs_fit && !getenv("RNNWF_WGRAD_FFMA")) {   // layers and the head alike
            fast = true;
            const wgtc::Geo q = wgtc::make_geo(R, a.cols, M);
            auto k = wgtc::wgrad_kernel;
            if (int e = set_smem(k, (int)q.smem)) return e;
            prof_count(); k<<<a.ksplit, wgtc::kThreads, q.smem, s>>>(a, q, w.partial, a.rtiles * kWgTile, a.ctiles * kWgTile);
            RNNWF_CUDA(cudaGetLastError());
        }
    }
    if constexpr (std::is_same<T, float>::value) {
        const int RG = (R + kWgfTr - 1) / kWgfTr, CG = (a.cols + kWgfTc - 1) / kWgfTc;
        const size_t fsmem = (size_t)(RG * kWgfTr + CG * kWgfTc) * (M + 4) * sizeof(float);
        if (!fast && RG * CG <= kWgfThreads && fsmem <= (size_t)kSmemLimit && RG * kWgfTr <= a.rtiles * kWgTile + 0 && !getenv("RNNWF_WGRAD_TILED")) {
            fast = true;
            auto k = wgrad_fast_kernel;
            if (int e = set_smem(k, (int)fsmem)) return e;
            prof_count(); k<<<a.ksplit, kWgfThreads, fsmem, s>>>(a, w.partial, a.rtiles * kWgTile, a.ctiles * kWgTile);
            RNNWF_CUDA(cudaGetLastError());
        }
    }
    if constexpr (std::is_same<T, double>::value) {
        if (M % 2 == 0 && !getenv("RNNWF_WGRAD_FFMA")) {       // DMMA reduction (float64 models)
            const int Rp = a.rtiles * kWgTile, Cp = a.ctiles * kWgTile;
            const int rt = (int)cdiv(R, wgdm::kT), ct = (int)cdiv(a.cols, wgdm::kT);
            const int64_t slots = ((int64_t)w.ksplit * w.Rp * w.Cp) / ((int64_t)Rp * Cp);      // what the partial buffer holds
            ksplit_used = wgdm::choose_ksplit(R, a.cols, nblk, slots);
            WgdmGruSrc src;
            src.a = a; src.M = M; src.R = R; src.C = a.cols; src.nblk = nblk;
            auto k = wgdm::wgrad_kernel<WgdmGruSrc>;
            if (int e = set_smem(k, (int)wgdm::kSmem)) return e;
            prof_count(); k<<<dim3(rt * ct, ksplit_used), wgdm::kThreads, wgdm::kSmem, s>>>(src, ct, ksplit_used, w.partial, Rp, Cp);
            RNNWF_CUDA(cudaGetLastError());
            fast = true;
        }
    }
    if (!fast) {
        const int smem = 2 * kWgTile * (M + 4) * (int)sizeof(T);
        auto k = wgrad_kernel<T>;
        if (int e = set_smem(k, smem)) return e;
        prof_count(); k<<<dim3(a.rtiles * a.ctiles, a.ksplit), 256, smem, s>>>(a, w.partial);
        RNNWF_CUDA(cudaGetLastError());
    }
    prof_count(); wgrad_scatter_kernel<<<grid_for((int64_t)R * a.cols), 256, 0, s>>>(w.partial, ksplit_used, a.rtiles * kWgTile, a.ctiles * kWgTile, g, l,
                                                                    head ? 1 : 0, grad);
    RNNWF_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
int gru_vmc_grad_t(const rnnwf_model& m, const void* params, const uint8_t* samples, int64_t ns, const double* weights, int flags,
                   double* grad, void* wsp, size_t wsb, cudaStream_t s) {
    const GruLayout g = make_gru_layout(m);
    const GruLayoutT gt = make_gru_layout_T(g);
    const BwdLaunch b = grad_launch<T>(g, ns, (flags & RNNWF_PARITY_SYM) ? 2 : 1, m.head == RNNWF_HEAD_COMPLEX);
    RNNWF_CHECK(b.RT > 0, -3, "no backward launch configuration fits (units=%d layers=%d)", m.units, m.num_layers);
    const GruLaunch cf = fwd_launch_for<T>(g, b);
    const bool cplx = m.head == RNNWF_HEAD_COMPLEX;
    const int parity = (flags & RNNWF_PARITY_SYM) ? 1 : 0;
    RNNWF_CHECK(!(cplx && parity), -2, "parity symmetry is only defined for the probability head");
    const int tiles_s = (int)cdiv(ns, b.M), ndir = parity ? 2 : 1, tiles = tiles_s * ndir;
    const int64_t rows_total = (int64_t)tiles * b.M;
    Ws ws(wsp, wsb);
    GradWs<T> w = carve_grad<T>(ws, g, gt, cf, tiles, cplx, ns, b.tc != 0);
    RNNWF_CHECK(ws.ok(), -4, "workspace too small: need %zu have %zu", ws.used, wsb);
    prof_count(); pack_gru_kernel<T><<<grid_for(g.PK), 256, 0, s>>>(g, (const T*)params, w.f.pk);
    prof_count(); pack_gru_T_kernel<T><<<grid_for(gt.total), 256, 0, s>>>(g, gt, (const T*)params, w.pkT);
    prof_count(); sig_transpose_kernel<<<grid_for(rows_total * g.N), 256, 0, s>>>(samples, w.f.sigT, ns, g.N, b.M, tiles_s, ndir);
    // teacher-forced pass with stash: the tensor-core base pass where it applies (same stash layout, 128-row work items), else the
    // FFMA tile engine
    int e = 0;
    bool stashed = false, tc_bwd = false;
    if constexpr (std::is_same<T, float>::value) {
        if (w.img16 && chain_mode(g) == 3) {
            tc_bwd = b.tc != 0 && w.gstore != nullptr;
            e = tc16p::launch_eloc(g, b.M, tiles, (const float*)params, w.img16, w.f.sigT, w.f.hstore, w.f.la_sel, w.f.la_oth, w.f.la_self,
                                   w.f.lp_re, nullptr, w.f.counter, false, s, tc_bwd ? w.gstore : nullptr);
            stashed = true;
        }
    }
    if constexpr (std::is_same<T, double>::value) {
        const char* env = getenv("RNNWF_CHAIN");
        if (w.wb64 && !(env && strcmp(env, "ffma") == 0)) {      // float64 one-layer stacks: the DMMA base pass
            e = f64mma::launch(g, b.M, rows_total, (const double*)params, w.wb64, w.wb64 + f64mma::make_layout(g).wb_doubles, w.f.sigT, w.f.hstore,
                               w.f.la_sel, w.f.la_oth, w.f.lp_re, nullptr, w.f.counter, true, false, s);
            stashed = true;
        }
    }
    if (!stashed) e = cplx ? launch_forward<T, true, true>(g, cf, w.f, tiles, s) : launch_forward<T, true, false>(g, cf, w.f, tiles, s);
    if (e) return e;
    prof_count(); row_weight_kernel<<<grid_for(rows_total), 256, 0, s>>>(weights, w.f.lp_re, ns, b.M, tiles_s, parity, cplx, w.roww);
    for (int l = g.L - 1; l >= 0; --l) {
        if (tc_bwd) {
            if constexpr (std::is_same<T, float>::value)
                e = tc16b::launch(g, l, b.M, tiles, (const float*)params, w.img16b, w.gstore, w.f.sigT, w.f.la_oth, w.roww, w.dxbuf, w.Gbuf, w.dzbuf, s);
        } else {
            e = cplx ? launch_bwd_layer<T, true>(g, gt, b, l, w, tiles, rows_total, s) : launch_bwd_layer<T, false>(g, gt, b, l, w, tiles, rows_total, s);
        }
        if (e) return e;
        if (l == g.L - 1)
            if ((e = launch_wgrad<T>(g, w, b.M, (int64_t)tiles * g.N, l, true, grad, s))) return e;
        if ((e = launch_wgrad<T>(g, w, b.M, (int64_t)tiles * g.N, l, false, grad, s))) return e;
    }
    return 0;
}
template int gru_vmc_grad_t<float>(const rnnwf_model&, const void*, const uint8_t*, int64_t, const double*, int, double*, void*, size_t,
                                   cudaStream_t);
template int gru_vmc_grad_t<double>(const rnnwf_model&, const void*, const uint8_t*, int64_t, const double*, int, double*, void*, size_t,
                                    cudaStream_t);

}  // namespace rnnwf
