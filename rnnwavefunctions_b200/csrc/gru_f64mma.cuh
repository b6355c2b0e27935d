// gru_f64mma.cuh — float64 prefix-reuse chain kernel on the FP64 tensor instruction (mma.sync.m8n8k4.f64, DMMA) for one-layer GRU
// stacks: the local-energy stage of the 2-D TFIM with the 1-D RNN (2DTFIM_1DRNN/Training1DRNN_2DTFIM.py:13-81 calling
// 2DTFIM_1DRNN/RNNwavefunction.py:86-130, float64), BASELINE config 3 (12 x 12, GRU(100)).
//
// Why: the thread-tile engine (gru_engine.cuh) is bound by shared-memory bandwidth in float64 -- every DFMA wants two 8-byte operands
// and a thread tile of 4 rows x 2 units x 3 gates needs 5 LDS.128 per 24 DFMA (ncu, profiles/r2: FP64 pipe 25 % busy).  The DMMA
// fragment layout shares operands across the warp: per K-step of 4 a warp loads 4 A fragments (hidden states, shared memory) and 3 B
// fragments (weights, straight from L2, register double-buffered) for 12 DMMAs = 3 072 FMAs.
//
// One CTA = 64 rows (connected configurations of 64 samples for one flipped site), 14 warps.  Warp w owns rows 32 (w & 1) .. + 31
// (4 m-tiles of 8) and the unit blocks 2 (w >> 1), 2 (w >> 1) + 1 (8 units each; H = 100 -> 13 blocks).  Per unit block it accumulates
// the three gate tiles r | u | ch over K = H in registers; the one-hot input part and the biases are table lookups that initialise the
// accumulators; gates, new state and the restaging into the other hidden-state buffer happen in registers.  The two logits of the
// Dense head ride in two spare columns of the last unit block's r tile (h_{n-1} Wd comes out of the GEMM of step n).
// Same contract as gru_chain_kernel<double> (gru_kernels.cuh): restart states / per-site base terms in, delta[tile][slot][M] out;
// the BASE instantiation is the teacher-forced pass that produces them (contract of gru_forward_kernel<double, .., STASH>).
// Included by gru.cu.
#pragma once
#include "gru_kernels.cuh"

namespace rnnwf {
namespace f64mma {

constexpr int kRows = 64, kWarps = 14, kThreads = kWarps * 32, kMaxBlocks = 14;

struct Layout {
    int H, N, blocks, ksteps, ldk;     // unit blocks of 8, K-steps of 4, row stride of the hidden-state tile (doubles, = 4 mod 16: conflict-free)
    int head_block, head_col;          // spare columns (head_col, head_col + 1) of the r tile of block head_block carry the head logits
    size_t wb_doubles, tab_doubles;
};

inline Layout make_layout(const GruLayout& g) {
    Layout t;
    t.H = g.H; t.N = g.N;
    t.blocks = (g.H + 7) / 8;
    t.ksteps = (g.H + 3) / 4;
    t.ldk = 4 * t.ksteps;
    while (t.ldk % 16 != 4) t.ldk += 4;
    t.head_block = t.blocks - 1;
    t.head_col = g.H - 8 * (t.blocks - 1);                        // first spare column of the last block (even)
    t.wb_doubles = (size_t)t.blocks * 3 * t.ksteps * 32;
    t.tab_doubles = (size_t)11 * 8 * t.blocks;                    // [kind][unit]: r(s=0), r(s=1), u(0), u(1), cx(0), cx(1), ch-bias | 7: bd[2] | 8..10: r, u, cx without input (site 0)
    return t;
}

inline size_t smem_bytes(const Layout& t) { return (size_t)2 * kRows * t.ldk * sizeof(double) + t.tab_doubles * sizeof(double) + 64; }

inline bool supported(const GruLayout& g) {
    if (g.L != 1 || g.nheads != 1 || g.H < 2 || g.N < 2) return false;
    const int blocks = (g.H + 7) / 8;
    if (blocks > kMaxBlocks) return false;
    if (8 * blocks - g.H < 2 || (g.H % 2) != 0) return false;      // two spare (even-aligned) columns for the head logits
    return smem_bytes(make_layout(g)) <= (size_t)kSmemLimit;
}

// flat TF-order parameters -> B fragments wb[block][gate][kstep][lane] (lane = 4 * column + k: W_gate[4 ks + lane % 4][8 block + lane / 4])
// and the table of input-dependent constants.  Gate 0: r, 1: u, 2: candidate hidden projection.
__global__ void pack_kernel(GruLayout g, Layout t, const double* __restrict__ flat, double* __restrict__ wb, double* __restrict__ tab) {
    const int H = g.H, d = g.d[0];
    const double* Kg = flat + g.flat_off[0];
    const double* bg = Kg + (d + H) * 2 * H;
    const double* Kci = bg + 2 * H;
    const double* Kch = Kci + d * H;
    const double* bci = Kch + H * H;
    const double* bch = bci + H;
    const double* Wd = flat + g.flat_head;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < t.wb_doubles; idx += (size_t)gridDim.x * blockDim.x) {
        const int lane = (int)(idx % 32);
        const int ks = (int)((idx / 32) % t.ksteps);
        const int gate = (int)((idx / 32 / t.ksteps) % 3);
        const int b = (int)(idx / 32 / t.ksteps / 3);
        const int k = 4 * ks + (lane & 3), c = lane >> 2, j = 8 * b + c;
        double v = 0.0;
        if (k < H) {
            if (j < H) v = gate < 2 ? Kg[(d + k) * 2 * H + gate * H + j] : Kch[k * H + j];
            else if (b == t.head_block && gate == 0 && (c == t.head_col || c == t.head_col + 1)) v = Wd[2 * k + (c - t.head_col)];
        }
        wb[idx] = v;
    }
    const int U = 8 * t.blocks;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < (int)t.tab_doubles; idx += gridDim.x * blockDim.x) {
        const int kind = idx / U, j = idx % U;
        double v = 0.0;
        if (j < H) {
            if (kind < 2) v = bg[j] + Kg[kind * 2 * H + j];
            else if (kind < 4) v = bg[H + j] + Kg[(kind - 2) * 2 * H + H + j];
            else if (kind < 6) v = bci[j] + Kci[(kind - 4) * H + j];
            else if (kind == 6) v = bch[j];
            else v = 0.0;
        }
        if (kind == 7) v = j < 2 ? Wd[2 * H + j] : 0.0;            // head bias
        if (kind >= 8 && j < H) v = kind == 8 ? bg[j] : kind == 9 ? bg[H + j] : bci[j];   // zero input vector (site 0 of the base pass)
        tab[idx] = v;
    }
}

struct Args {
    GruLayout g;
    Layout t;
    int Mold, tiles64, nslots;
    int64_t rows_total;
    const double* wb;
    const double* tab;
    const uint8_t* sigT;
    double* hstore;               // FLIP: restart states (read); BASE: every site's state (written)
    double *la_sel, *la_oth;      // FLIP: read at the flipped site and per site; BASE: written
    double* lp;                   // BASE: sum_n la_sel
    double* delta;
    int* counter;
};

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// BASE: the teacher-forced pass over the unmodified configuration from the zero state (2DTFIM_1DRNN/RNNwavefunction.py:86-130):
// stashes every site's state and per-site head terms (what the flip chains and the backward pass start from) and sums log P.
template <bool BASE>
__global__ void __launch_bounds__(kThreads, 1) chain_kernel(const __grid_constant__ Args a) {
    extern __shared__ __align__(16) unsigned char smem_f64[];
    __shared__ int s_work;
    const Layout& t = a.t;
    const int H = t.H, N = t.N, ldk = t.ldk, Mold = a.Mold, U = 8 * t.blocks;
    double* hbuf = reinterpret_cast<double*>(smem_f64);                       // [2][kRows][ldk]
    double* tab = hbuf + (size_t)2 * kRows * ldk;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int mh = warp & 1, grp = warp >> 1;
    const int qrow = lane >> 2, qcol = lane & 3;                              // fragment coordinates of this lane
    for (int i = tid; i < (int)t.tab_doubles; i += kThreads) tab[i] = a.tab[i];
    for (int i = tid; i < 2 * kRows * ldk; i += kThreads) hbuf[i] = 0.0;      // K padding columns stay zero
    const int total = (BASE ? 1 : a.nslots) * a.tiles64;
    const bool head_warp = (2 * grp == t.head_block) || (2 * grp + 1 == t.head_block);
    const bool head_lane = head_warp && (2 * qcol == t.head_col);

    while (true) {
        __syncthreads();
        if (tid == 0) s_work = atomicAdd(a.counter, 1);
        __syncthreads();
        const int work = s_work;
        if (work >= total) break;
        const int s = BASE ? -1 : work / a.tiles64, tile = work % a.tiles64;  // slots in order of decreasing chain length
        // this lane's four rows (one per m-tile): global row -> (old tile, row in tile) of the base pass' layout
        size_t rbase[4];       // index of (old tile, site 0, row): + n * Mold for site n
        bool live[4];
        int otile[4], om[4];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            int64_t R = (int64_t)tile * kRows + 32 * mh + 8 * mt + qrow;
            live[mt] = R < a.rows_total;
            if (!live[mt]) R = a.rows_total - 1;
            otile[mt] = (int)(R / Mold);
            om[mt] = (int)(R % Mold);
            rbase[mt] = (size_t)otile[mt] * N * Mold + om[mt];
        }
        // restart state: h after site s of the base pass (BASE: the zero state) -> buffer 0
        for (int i = tid; i < kRows * H; i += kThreads) {
            const int row = i % kRows, j = i / kRows;
            int64_t R = (int64_t)tile * kRows + row;
            if (R >= a.rows_total) R = a.rows_total - 1;
            const size_t ot = (size_t)(R / Mold), m = (size_t)(R % Mold);
            hbuf[row * ldk + j] = BASE ? 0.0 : a.hstore[((ot * N + s) * H + j) * Mold + m];
        }
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        if (!BASE && head_lane) {
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) acc[mt] = a.la_oth[rbase[mt] + (size_t)s * Mold] - a.la_sel[rbase[mt] + (size_t)s * Mold];
        }
        __syncthreads();
        int cur = 0;
        // steps n = s+1 .. N-1 compute h_n; step n == N only evaluates the head of site N-1
        for (int n = s + 1; n <= N; ++n) {
            // input spin of this step = spin of site n-1 of the connected configuration (flipped at s)
            int sp[4], kr[4], ku[4], kc[4];       // input spin (2: none, site 0 of the base pass) and the table rows of its constants
            double la_prev[4];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
                int v = n > 0 ? a.sigT[rbase[mt] + (size_t)(n - 1) * Mold] : 2;
                if (!BASE && n - 1 == s) v = 1 - v;
                sp[mt] = v;
                kr[mt] = v < 2 ? v : 8; ku[mt] = v < 2 ? 2 + v : 9; kc[mt] = v < 2 ? 4 + v : 10;
                la_prev[mt] = (!BASE && head_lane && n - 1 > s) ? a.la_sel[rbase[mt] + (size_t)(n - 1) * Mold] : 0.0;
            }
            const double* hc = hbuf + (size_t)cur * kRows * ldk;
            double* hn = hbuf + (size_t)(cur ^ 1) * kRows * ldk;
            const bool last = n == N;
#pragma unroll 1
            for (int bi = 0; bi < 2; ++bi) {
                const int b = 2 * grp + bi;
                if (b >= t.blocks) break;
                if (last && b != t.head_block) continue;
                const int j0 = 8 * b + 2 * qcol;                                 // this lane's two units (columns of the 8 x 8 tiles)
                double cr[4][2], cu[4][2], cq[4][2];
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        cr[mt][e] = tab[kr[mt] * U + j0 + e];
                        cu[mt][e] = tab[ku[mt] * U + j0 + e];
                        cq[mt][e] = tab[6 * U + j0 + e];
                    }
                }
                const double* wr = a.wb + ((size_t)(b * 3 + 0) * t.ksteps) * 32 + lane;
                const double* wu = a.wb + ((size_t)(b * 3 + 1) * t.ksteps) * 32 + lane;
                const double* wq = a.wb + ((size_t)(b * 3 + 2) * t.ksteps) * 32 + lane;
                const double* ap = hc + (size_t)(32 * mh + qrow) * ldk + qcol;
                double br = wr[0], bu = wu[0], bq = wq[0];
#pragma unroll 5
                for (int ks = 0; ks < t.ksteps; ++ks) {
                    const int kn = ks + 1 < t.ksteps ? ks + 1 : ks;
                    const double nbr = wr[kn * 32], nbu = wu[kn * 32], nbq = wq[kn * 32];   // next K-step's B fragments (L2), in flight under the MMAs
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt) {
                        const double av = ap[(size_t)(8 * mt) * ldk + 4 * ks];
                        dmma(cr[mt][0], cr[mt][1], av, br);
                        if (!last) {
                            dmma(cu[mt][0], cu[mt][1], av, bu);
                            dmma(cq[mt][0], cq[mt][1], av, bq);
                        }
                    }
                    br = nbr; bu = nbu; bq = nbq;
                }
                // head of site n-1 (logits = h_{n-1} Wd + bd): this lane holds (z0, z1) of its four rows in the spare columns of the r tile
                if (b == t.head_block && head_lane && n - 1 > s) {
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt) {
                        const double z0 = cr[mt][0] - tab[kr[mt] * U + j0] + tab[7 * U], z1 = cr[mt][1] - tab[kr[mt] * U + j0 + 1] + tab[7 * U + 1];
                        const double ls = sp[mt] ? log_softmax2(z1, z0) : log_softmax2(z0, z1);
                        if (BASE) {
                            if (live[mt]) {
                                a.la_sel[rbase[mt] + (size_t)(n - 1) * Mold] = ls;
                                a.la_oth[rbase[mt] + (size_t)(n - 1) * Mold] = sp[mt] ? log_softmax2(z0, z1) : log_softmax2(z1, z0);
                            }
                            acc[mt] += ls;
                        } else {
                            acc[mt] += ls - la_prev[mt];
                        }
                    }
                }
                if (last) continue;
                // gates, candidate, new state for the lane's (row, unit) pairs
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) {
                    const int row = 32 * mh + 8 * mt + qrow;
                    if (j0 < H) {
                        const double2 ho = *reinterpret_cast<const double2*>(hc + (size_t)row * ldk + j0);
                        double hv[2] = {ho.x, ho.y}, out[2];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const double r = sigmoid_(cr[mt][e]);
                            const double u = sigmoid_(cu[mt][e]);
                            const double c = tanh_(fma(r, cq[mt][e], tab[kc[mt] * U + j0 + e]));
                            out[e] = fma(u, hv[e] - c, c);
                        }
                        *reinterpret_cast<double2*>(hn + (size_t)row * ldk + j0) = make_double2(out[0], out[1]);
                        if (BASE && live[mt]) {
                            double* hs = a.hstore + (((size_t)otile[mt] * N + n) * H + j0) * Mold + om[mt];
                            hs[0] = out[0];
                            hs[Mold] = out[1];
                        }
                    }
                }
            }
            __syncthreads();
            cur ^= 1;
        }
        if (head_lane) {
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
                if (live[mt]) {
                    if constexpr (BASE) a.lp[(size_t)otile[mt] * Mold + om[mt]] = acc[mt];
                    else a.delta[((size_t)otile[mt] * a.nslots + s) * Mold + om[mt]] = acc[mt];
                }
        }
    }
}

static Args make_args(const GruLayout& g, int Mold, int64_t rows_total, double* wb, double* tab, const uint8_t* sigT, double* hstore,
                      double* la_sel, double* la_oth, double* lp, double* delta, int* counter) {
    Args a;
    memset(&a, 0, sizeof(a));
    a.g = g; a.t = make_layout(g); a.Mold = Mold; a.rows_total = rows_total;
    a.tiles64 = (int)cdiv(rows_total, kRows);
    a.nslots = g.N;
    a.wb = wb; a.tab = tab; a.sigT = sigT; a.hstore = hstore; a.la_sel = la_sel; a.la_oth = la_oth; a.lp = lp; a.delta = delta; a.counter = counter;
    return a;
}

// teacher-forced base pass with stash (rows in the [tile][site][..][Mold] layout of the thread-tile engine) and, with `flips`, the
// single-flip chains from it
static int launch(const GruLayout& g, int Mold, int64_t rows_total, const double* params, double* wb, double* tab, const uint8_t* sigT,
                  double* hstore, double* la_sel, double* la_oth, double* lp, double* delta, int* counter, bool base, bool flips,
                  cudaStream_t s) {
    Args a = make_args(g, Mold, rows_total, wb, tab, sigT, hstore, la_sel, la_oth, lp, delta, counter);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int smem = (int)smem_bytes(a.t);
    RNNWF_CHECK(smem <= kSmemLimit, -3, "float64 DMMA chain kernel needs %d bytes of shared memory", smem);
    prof_count(); pack_kernel<<<148, 256, 0, s>>>(g, a.t, params, wb, tab);
    if (base) {
        RNNWF_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), s));
        RNNWF_CUDA(cudaFuncSetAttribute(chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        prof_count();
        chain_kernel<true><<<std::min(a.tiles64, sms), kThreads, smem, s>>>(a);
        RNNWF_CUDA(cudaGetLastError());
    }
    if (flips) {
        RNNWF_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), s));
        RNNWF_CUDA(cudaFuncSetAttribute(chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        const int grid = (int)std::min<int64_t>((int64_t)a.nslots * a.tiles64, sms);
        prof_count();
        prof_mark(0, s);
        chain_kernel<false><<<grid, kThreads, smem, s>>>(a);
        prof_mark(1, s);
        RNNWF_CUDA(cudaGetLastError());
    }
    return 0;
}

}  // namespace f64mma
}  // namespace rnnwf
