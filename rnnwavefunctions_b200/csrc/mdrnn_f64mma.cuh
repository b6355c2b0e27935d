// mdrnn_f64mma.cuh — float64 prefix-reuse chain kernel of the 2-D RNN (MDRNNcell on the zig-zag path) on the FP64 tensor instruction
// (mma.sync.m8n8k4.f64): the local-energy stage of 2DTFIM_2DRNN/Training2DRNN_2DTFIM.py:13-83 calling
// 2DTFIM_2DRNN/RNNwavefunction.py:120-200 with the cell of 2DTFIM_2DRNN/MDRNNcell.py:51-66, BASELINE config 4 (12 x 12, 100 units).
//
//   h[p] = elu( [h_left | h_up] [Wh ; Wv] + Uh[spin_left] + Uv[spin_up] + b )            K = 2 H, N = H
//
// Why: ncu of the thread-tile kernel (md_chain_kernel<double>, profiles/r2): shared-memory wavefronts 62 % of peak with 7 * 10^8 bank
// conflicts, FP64 pipe 32 % busy -- every DFMA fetches its own operands.  Here a warp shares them through the DMMA fragments: per K-step
// of 4 it loads 4 A fragments (states, shared memory) and 1 B fragment (weights, L2, register double-buffered) for 4 DMMAs.
//
// One CTA = 64 rows, 28 warps: warp w owns rows 32 (w & 1) .. + 31 (4 m-tiles) and unit block w >> 1 (8 units; H = 100 -> 13 blocks);
// the last warp pair evaluates the Dense head of the previous site (h[p-1] Wd) while the others compute h[p].  The left state is the
// previous step's output (kept in shared memory, ping-pong); the up state was produced Nx steps earlier: it comes from the CTA's
// private grid in global memory (or, at or before the flipped site, from the base pass' grid) and is prefetched with cp.async during
// the previous step.  Same contract as md_chain_kernel<double>.  Included by mdrnn.cu.
#pragma once

namespace rnnwf {
namespace mdmma {

constexpr int kRows = 64, kWarps = 28, kThreads = kWarps * 32;

struct Layout {
    int H, N, nx, ny, blocks, ksteps, ldk;
    size_t wb_doubles, tab_doubles, priv_doubles;   // B fragments, lookup table, private state grid per CTA
};


inline Layout make_layout(const MdLayout& g) {
    Layout t;
    t.H = g.H; t.N = g.N; t.nx = g.nx; t.ny = g.ny;
    t.blocks = (g.H + 7) / 8;
    t.ksteps = (g.H + 3) / 4;
    t.ldk = 4 * t.ksteps;
    while (t.ldk % 16 != 4) t.ldk += 4;
    t.wb_doubles = (size_t)(2 * t.blocks + 1) * t.ksteps * 32;        // Wh blocks | Wv blocks | head tile
    t.tab_doubles = (size_t)6 * 8 * t.blocks;                          // b | Uh[0] | Uh[1] | Uv[0] | Uv[1] | bd[2]
    t.priv_doubles = (size_t)g.N * kRows * t.ldk;
    return t;
}

inline size_t smem_bytes(const Layout& t) {
    return (size_t)4 * kRows * t.ldk * sizeof(double) + t.tab_doubles * sizeof(double) + (size_t)kRows * t.N + kRows * sizeof(size_t) + 64;
}

// up to 13 blocks of 8 units (26 compute warps + the head pair), and the four state tiles + the tile's spins must fit in shared memory
inline bool supported(const MdLayout& g) {
    return g.H >= 2 && (g.H + 7) / 8 <= 13 && g.N >= 2 && smem_bytes(make_layout(g)) <= (size_t)kSmemLimit;
}

// flat TF-order parameters (Wh[H,H] | Uh[2,H] | Wv[H,H] | Uv[2,H] | b[H] | Wd[H,2] | bd[2]) -> B fragments
//   wb[(seg * blocks + block) * ksteps + ks][lane] = W_seg[4 ks + lane % 4][8 block + lane / 4]   (seg 0: Wh, 1: Wv),
//   head tile at index 2 * blocks: Wd[4 ks + lane % 4][lane / 4] for columns 0, 1
__global__ void pack_kernel(MdLayout g, Layout t, const double* __restrict__ flat, double* __restrict__ wb, double* __restrict__ tab) {
    const int H = g.H;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < t.wb_doubles; idx += (size_t)gridDim.x * blockDim.x) {
        const int lane = (int)(idx % 32);
        const int ks = (int)((idx / 32) % t.ksteps);
        const int tile = (int)(idx / 32 / t.ksteps);
        const int k = 4 * ks + (lane & 3), c = lane >> 2;
        double v = 0.0;
        if (k < H) {
            if (tile < 2 * t.blocks) {
                const int seg = tile / t.blocks, j = 8 * (tile % t.blocks) + c;
                if (j < H) v = flat[(seg ? g.f_wv : 0) + k * H + j];
            } else if (c < 2) {
                v = flat[g.f_wd + 2 * k + c];
            }
        }
        wb[idx] = v;
    }
    const int U = 8 * t.blocks;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < (int)t.tab_doubles; idx += gridDim.x * blockDim.x) {
        const int kind = idx / U, j = idx % U;
        double v = 0.0;
        if (j < H) {
            if (kind == 0) v = flat[g.f_b + j];
            else if (kind < 3) v = flat[g.f_uh + (kind - 1) * H + j];
            else if (kind < 5) v = flat[g.f_uv + (kind - 3) * H + j];
        }
        if (kind == 5 && j < 2) v = flat[g.f_bd + j];
        tab[idx] = v;
    }
}

struct Args {
    MdLayout g;
    Layout t;
    int Mold, tiles64;
    int64_t ns;
    const double *wb, *tab;
    const uint8_t* samples;        // [ns][N], site = x * ny + y
    const double* hbase;           // base pass grid [old tile][p][H][Mold]
    double* hpriv;                 // [CTA][p][kRows][ldk]
    const double *la_sel, *la_oth; // [old tile][p][Mold]
    double* delta;                 // [old tile][site][Mold]
    int* counter;
};

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cpa16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cpa8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

// up-neighbour position of path position p (-1: first row) and whether p has a left neighbour on its row
__device__ __forceinline__ void decode(const Layout& t, int p, int& site, int& pl, int& pu) {
    const int y = p / t.nx, xi = p % t.nx;
    const int x = (y & 1) ? t.nx - 1 - xi : xi;
    site = x * t.ny + y;
    pl = xi > 0 ? p - 1 : -1;
    pu = y > 0 ? (y - 1) * t.nx + (((y - 1) & 1) ? t.nx - 1 - x : x) : -1;
}
__device__ __forceinline__ int site_of(const Layout& t, int p) {
    const int y = p / t.nx, xi = p % t.nx;
    return ((y & 1) ? t.nx - 1 - xi : xi) * t.ny + y;
}

__global__ void __launch_bounds__(kThreads, 1) chain_kernel(const __grid_constant__ Args a) {
    extern __shared__ __align__(16) unsigned char smem_md[];
    __shared__ int s_work;
    const Layout& t = a.t;
    const int H = t.H, N = t.N, ldk = t.ldk, Mold = a.Mold, U = 8 * t.blocks;
    double* hs = reinterpret_cast<double*>(smem_md);                   // [2][kRows][ldk]: h[p-1] / h[p] ping-pong
    double* hu = hs + (size_t)2 * kRows * ldk;                          // [2][kRows][ldk]: up states, double-buffered
    double* tab = hu + (size_t)2 * kRows * ldk;
    size_t* rowoff = reinterpret_cast<size_t*>(tab + t.tab_doubles);    // per row: (old tile * N * H) * Mold + row in old tile
    uint8_t* sig = reinterpret_cast<uint8_t*>(rowoff + kRows);          // [kRows][N] spins of the connected configuration
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int mh = warp & 1, blk = warp >> 1;
    const int qrow = lane >> 2, qcol = lane & 3;
    const bool head_warp = blk == t.blocks;                             // warps beyond the unit blocks (2 * blocks, 2 * blocks + 1) do the head
    const bool main_warp = blk < t.blocks;
    for (int i = tid; i < (int)t.tab_doubles; i += kThreads) tab[i] = a.tab[i];
    for (int i = tid; i < 4 * kRows * ldk; i += kThreads) hs[i] = 0.0;
    double* priv = a.hpriv + (size_t)blockIdx.x * t.priv_doubles;
    if (ldk > H)   // K padding columns of the private grid are read back as operands: zero them once (the state columns are written before use)
        for (int i = tid; i < N * kRows * (ldk - H); i += kThreads) priv[(size_t)(i / (ldk - H)) * ldk + H + i % (ldk - H)] = 0.0;
    const int total = N * a.tiles64;
    const int64_t rows_total = (int64_t)cdiv(a.ns, (int64_t)Mold) * Mold;

    // fetch the up state of position p (state of pu) into hu[buf]: private grid (contiguous) or base grid (gather)
    auto fetch_up = [&](int pu, int k, int buf) {
        double* dst = hu + (size_t)buf * kRows * ldk;
        if (pu > k) {
            const double* src = priv + (size_t)pu * kRows * ldk;
            for (int i = tid; i < kRows * ldk / 2; i += kThreads) cpa16(dst + 2 * i, src + 2 * i);
        } else {
            for (int i = tid; i < kRows * H; i += kThreads) {
                const int row = i % kRows, j = i / kRows;
                cpa8(dst + row * ldk + j, a.hbase + rowoff[row] + ((size_t)pu * H + j) * Mold);
            }
        }
    };

    while (true) {
        __syncthreads();
        if (tid == 0) s_work = atomicAdd(a.counter, 1);
        __syncthreads();
        const int work = s_work;
        if (work >= total) break;
        const int k = work / a.tiles64, tile = work % a.tiles64;      // ascending k: longest chains first
        const int ksite = site_of(t, k);
        for (int row = tid; row < kRows; row += kThreads) {
            int64_t R = (int64_t)tile * kRows + row;
            if (R >= rows_total) R = rows_total - 1;
            rowoff[row] = (size_t)(R / Mold) * N * H * Mold + (size_t)(R % Mold);
        }
        for (int i = tid; i < kRows * N; i += kThreads) {
            const int row = i / N, site = i % N;
            const int64_t R = (int64_t)tile * kRows + row;
            uint8_t v = R < a.ns ? a.samples[R * N + site] : 0;
            if (site == ksite) v = 1 - v;
            sig[i] = v;
        }
        __syncthreads();
        // left state of position k + 1 (if on the same row): the unchanged base state of k -> hs[0]
        for (int i = tid; i < kRows * H; i += kThreads) {
            const int row = i % kRows, j = i / kRows;
            hs[row * ldk + j] = a.hbase[rowoff[row] + ((size_t)k * H + j) * Mold];
        }
        int ub = 0;
        if (k + 1 < N) {
            int site1, pl1, pu1;
            decode(t, k + 1, site1, pl1, pu1);
            if (pu1 >= 0) fetch_up(pu1, k, 0);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        // this lane's rows (head warps: accumulators of the log-ratio)
        int64_t Rr[4];
        size_t lbase[4];      // index of (old tile, position 0, row) in la_sel / la_oth / delta
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            int64_t R = (int64_t)tile * kRows + 32 * mh + 8 * mt + qrow;
            Rr[mt] = R;
            if (R >= rows_total) R = rows_total - 1;
            lbase[mt] = (size_t)(R / Mold) * N * Mold + (size_t)(R % Mold);
            if (head_warp && qcol == 0) acc[mt] = a.la_oth[lbase[mt] + (size_t)k * Mold] - a.la_sel[lbase[mt] + (size_t)k * Mold];
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        int cur = 0;
        for (int p = k + 1; p <= N; ++p) {
            const bool last = p == N;
            int site = 0, pl = -1, pu = -1;
            if (!last) decode(t, p, site, pl, pu);
            // prefetch the up state of the next position
            if (p + 1 < N) {
                int s2, pl2, pu2;
                decode(t, p + 1, s2, pl2, pu2);
                if (pu2 >= 0 && pu2 != p) fetch_up(pu2, k, ub ^ 1);   // pu2 == p (row turn-around): the up state is this step's output, still in shared memory
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            const double* hc = hs + (size_t)cur * kRows * ldk;
            double* hn = hs + (size_t)(cur ^ 1) * kRows * ldk;
            const double* hup = (pu >= 0 && pu == p - 1 && p - 1 > k) ? hc : hu + (size_t)ub * kRows * ldk;   // first site of a row: up = previous position
            if (main_warp && !last) {
                const int j0 = 8 * blk + 2 * qcol;
                const int sl = pl >= 0 ? site_of(t, pl) : -1, su = pu >= 0 ? site_of(t, pu) : -1;
                double c[4][2];
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) {
                    const int row = 32 * mh + 8 * mt + qrow;
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        double v = tab[j0 + e];
                        if (sl >= 0) v += tab[(1 + sig[row * N + sl]) * U + j0 + e];
                        if (su >= 0) v += tab[(3 + sig[row * N + su]) * U + j0 + e];
                        c[mt][e] = v;
                    }
                }
#pragma unroll 1
                for (int seg = 0; seg < 2; ++seg) {
                    if ((seg == 0 && pl < 0) || (seg == 1 && pu < 0)) continue;
                    const double* w = a.wb + ((size_t)(seg * t.blocks + blk) * t.ksteps) * 32 + lane;
                    const double* ap = (seg == 0 ? hc : hup) + (size_t)(32 * mh + qrow) * ldk + qcol;
                    double b0 = w[0];
#pragma unroll 5
                    for (int ks = 0; ks < t.ksteps; ++ks) {
                        const double nb = w[(ks + 1 < t.ksteps ? ks + 1 : ks) * 32];
#pragma unroll
                        for (int mt = 0; mt < 4; ++mt) dmma(c[mt][0], c[mt][1], ap[(size_t)(8 * mt) * ldk + 4 * ks], b0);
                        b0 = nb;
                    }
                }
                if (j0 < H) {
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt) {
                        const int row = 32 * mh + 8 * mt + qrow;
                        const double2 o = make_double2(elu_(c[mt][0]), j0 + 1 < H ? elu_(c[mt][1]) : 0.0);
                        *reinterpret_cast<double2*>(hn + (size_t)row * ldk + j0) = o;
                        if (p + t.nx < N) *reinterpret_cast<double2*>(priv + ((size_t)p * kRows + row) * ldk + j0) = o;   // someone's up state later
                    }
                }
            } else if (head_warp && p - 1 > k) {
                // Dense head of position p - 1: logits = h[p-1] Wd + bd; lanes with qcol == 0 hold (z0, z1) of their rows
                const int sprev = site_of(t, p - 1);
                double z[4][2];
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) { z[mt][0] = tab[5 * U]; z[mt][1] = tab[5 * U + 1]; }
                const double* w = a.wb + ((size_t)(2 * t.blocks) * t.ksteps) * 32 + lane;
                const double* ap = hc + (size_t)(32 * mh + qrow) * ldk + qcol;
#pragma unroll 5
                for (int ks = 0; ks < t.ksteps; ++ks) {
                    const double b0 = w[ks * 32];
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt) dmma(z[mt][0], z[mt][1], ap[(size_t)(8 * mt) * ldk + 4 * ks], b0);
                }
                if (qcol == 0) {
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt) {
                        const int row = 32 * mh + 8 * mt + qrow;
                        const int sg = sig[row * N + sprev];
                        const double ls = sg ? log_softmax2(z[mt][1], z[mt][0]) : log_softmax2(z[mt][0], z[mt][1]);
                        acc[mt] += ls - a.la_sel[lbase[mt] + (size_t)(p - 1) * Mold];
                    }
                }
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
            cur ^= 1;
            ub ^= 1;
        }
        if (head_warp && qcol == 0) {
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
                if (Rr[mt] < a.ns) {
                    const int64_t R = Rr[mt];
                    a.delta[((size_t)(R / Mold) * N + ksite) * Mold + (size_t)(R % Mold)] = acc[mt];
                }
        }
    }
}

static int launch(const MdLayout& g, int Mold, int64_t ns, const double* params, double* wb, double* tab, double* hpriv,
                  const uint8_t* samples, const double* hbase, const double* la_sel, const double* la_oth, double* delta, int* counter,
                  int sms, cudaStream_t s) {
    Args a;
    memset(&a, 0, sizeof(a));
    a.g = g; a.t = make_layout(g); a.Mold = Mold; a.ns = ns;
    a.tiles64 = (int)cdiv(cdiv(ns, (int64_t)Mold) * Mold, (int64_t)kRows);
    a.wb = wb; a.tab = tab; a.samples = samples; a.hbase = hbase; a.hpriv = hpriv; a.la_sel = la_sel; a.la_oth = la_oth; a.delta = delta;
    a.counter = counter;
    const int smem = (int)smem_bytes(a.t);
    RNNWF_CHECK(smem <= kSmemLimit, -3, "float64 DMMA 2-D RNN chain kernel needs %d bytes of shared memory", smem);
    prof_count(); pack_kernel<<<148, 256, 0, s>>>(g, a.t, params, wb, tab);
    RNNWF_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), s));
    RNNWF_CUDA(cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int grid = (int)std::min<int64_t>((int64_t)g.N * a.tiles64, sms);
    prof_count();
    prof_mark(0, s);
    chain_kernel<<<grid, kThreads, smem, s>>>(a);
    prof_mark(1, s);
    RNNWF_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mdmma
}  // namespace rnnwf
