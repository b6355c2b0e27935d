// mdrnn.cu — K4: the 2-D RNN wave function (MDRNNcell on the zig-zag path) on sm_100a.
//
// Replaces 2DTFIM_2DRNN/MDRNNcell.py:51-66 (cell), 2DTFIM_2DRNN/RNNwavefunction.py:35-118 (sample),
// :120-200 (log_probability), Ising2D_local_energies (2DTFIM_2DRNN/Training2DRNN_2DTFIM.py:13-83) and the
// gradient of the cost (:160-169).
//
//   h[x,y] = elu( in_l Uh + h_l Wh + in_u Uv + h_u Wv + b ),   l = previous site on the row's path, u = (x, y-1)
//   path position p of (x,y) = y*Nx + (x if y even else Nx-1-x)                               
//
// One CTA owns a tile of M rows (samples / connected configurations).  Packed weights stay resident in shared
// memory; the state of every visited site goes to a per-tile grid in global memory (L2-resident at these
// sizes) because site p needs the state of the site above it, Nx path steps earlier.  Thread (rt, ct) owns
// SPT rows x 2 units, as in the GRU engine.  Flat parameter order (TF variable creation order,
// MDRNNcell.py:21-35 + Dense):  Wh[H,H] | Uh[2,H] | Wv[H,H] | Uv[2,H] | b[H] | Wd[H,2] | bd[2].
#include <stdlib.h>
#include <string.h>
#include <type_traits>
#include "gru_engine.cuh"
#include "host_util.cuh"
#include "api_internal.h"
#include "wgrad_f64mma.cuh"

namespace rnnwf {

struct MdLayout {
    int H, CT, N, nx, ny;
    int o_wv, o_uh, o_uv, o_b, o_head, PK;   // packed forward weights: wh[H][CT][2] | wv | uh[2][CT][2] | uv | b[CT][2] | Wd,bd
    int o_wvT, PKT;                          // packed transposed: whT[i][CT][2] (= Wh[j=2ct+u][i]) | wvT
    int f_uh, f_wv, f_uv, f_b, f_wd, f_bd, P;
};

static MdLayout make_md_layout(const rnnwf_model& m) {
    MdLayout g;
    memset(&g, 0, sizeof(g));
    g.H = m.units; g.CT = (m.units + 1) / 2; g.N = m.n_sites; g.nx = m.nx; g.ny = m.ny;
    const int H = g.H, CT = g.CT;
    int o = align4(H * CT * 2);
    g.o_wv = o; o += align4(H * CT * 2);
    g.o_uh = o; o += align4(2 * CT * 2);
    g.o_uv = o; o += align4(2 * CT * 2);
    g.o_b = o; o += align4(CT * 2);
    g.o_head = o; o += align4(2 * H + 2);
    g.PK = o;
    g.o_wvT = align4(H * CT * 2);
    g.PKT = 2 * align4(H * CT * 2);
    g.f_uh = H * H; g.f_wv = g.f_uh + 2 * H; g.f_uv = g.f_wv + H * H; g.f_b = g.f_uv + 2 * H; g.f_wd = g.f_b + H;
    g.f_bd = g.f_wd + 2 * H; g.P = g.f_bd + 2;
    return g;
}

struct MdLaunch {
    int RT, M, NT, w_smem, smem_bytes;
};

// smem: [weights] | hl[H][M] | hu[H][M] | sig[N][M] bytes | codes 2*M bytes
template <typename T> static MdLaunch choose_md_launch(const MdLayout& g, int weight_words, int extra_tiles) {
    constexpr int SPT = VT<T>::SPT;
    MdLaunch best;
    memset(&best, 0, sizeof(best));
    for (int wsm = 1; wsm >= 0; --wsm) {
        for (int RT = 1; RT <= 64; ++RT) {
            const int nt = g.CT * RT, M = RT * SPT;
            if (nt > 512 || M > 256) break;
            size_t smem = (wsm ? (((size_t)weight_words * sizeof(T) + 15) & ~(size_t)15) : 0) +
                          (size_t)(2 + extra_tiles) * g.H * M * sizeof(T) + (size_t)g.N * M + 2 * (size_t)M + 8 * (size_t)M * sizeof(T) + 64;
            if (smem > (size_t)kSmemLimit) break;
            best.RT = RT; best.M = M; best.NT = (nt + 31) & ~31; best.w_smem = wsm; best.smem_bytes = (int)smem;
        }
        if (best.RT > 0) return best;
    }
    return best;
}

template <typename T>
__global__ void md_pack_kernel(MdLayout g, const T* __restrict__ flat, T* __restrict__ pk, T* __restrict__ pkT) {
    const int H = g.H, CT = g.CT;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < g.PK + g.PKT; idx += gridDim.x * blockDim.x) {
        T val = T(0);
        if (idx < g.PK) {
            if (idx < g.o_uh) {
                const bool v = idx >= g.o_wv;
                const int q = idx - (v ? g.o_wv : 0);
                const int k = q / (CT * 2), ct = (q % (CT * 2)) / 2, j = 2 * ct + (q & 1);
                if (k < H && j < H) val = flat[(v ? g.f_wv : 0) + k * H + j];
            } else if (idx < g.o_b) {
                const bool v = idx >= g.o_uv;
                const int q = idx - (v ? g.o_uv : g.o_uh);
                const int k = q / (CT * 2), ct = (q % (CT * 2)) / 2, j = 2 * ct + (q & 1);
                if (k < 2 && j < H) val = flat[(v ? g.f_uv : g.f_uh) + k * H + j];
            } else if (idx < g.o_head) {
                const int q = idx - g.o_b;
                if (q < H) val = flat[g.f_b + q];
            } else {
                const int q = idx - g.o_head;
                if (q < 2 * H + 2) val = flat[g.f_wd + q];
            }
            pk[idx] = val;
        } else {
            const int t = idx - g.PK;
            const bool v = t >= g.o_wvT;
            const int q = t - (v ? g.o_wvT : 0);
            const int i = q / (CT * 2), ct = (q % (CT * 2)) / 2, j = 2 * ct + (q & 1);
            if (i < H && j < H) val = flat[(v ? g.f_wv : 0) + j * H + i];   // W^T[i][j] = W[j][i]
            pkT[t] = val;
        }
    }
}

__device__ __forceinline__ void md_decode(const MdLayout& g, int p, int& x, int& y, int& pl, int& pu, int& pd) {
    y = p / g.nx;
    const int xi = p % g.nx;
    x = (y & 1) ? g.nx - 1 - xi : xi;
    pl = xi > 0 ? p - 1 : -1;
    pu = y > 0 ? (y - 1) * g.nx + (((y - 1) & 1) ? g.nx - 1 - x : x) : -1;
    pd = y + 1 < g.ny ? (y + 1) * g.nx + (((y + 1) & 1) ? g.nx - 1 - x : x) : -1;
}
__device__ __forceinline__ int md_site(const MdLayout& g, int p) {   // [x][y] flat index of path position p
    const int y = p / g.nx, xi = p % g.nx;
    return ((y & 1) ? g.nx - 1 - xi : xi) * g.ny + y;
}

// pre-activation + elu for the thread's (SPT rows x 2 units) tile
template <typename T>
__device__ __forceinline__ void md_cell(const MdLayout& g, const T* __restrict__ w, const T* __restrict__ hl, const T* __restrict__ hu,
                                        const uint8_t* __restrict__ cl, const uint8_t* __restrict__ cu, int M, int ct, int rt,
                                        T (&hn)[2][VT<T>::SPT]) {
    constexpr int SPT = VT<T>::SPT;
    const int H = g.H, CT = g.CT, row0 = rt * SPT;
    T acc[2][SPT];
    {
        T b[2], l0[2], l1[2], u0[2], u1[2];
        ldv<2>(b, w + g.o_b + ct * 2);
        ldv<2>(l0, w + g.o_uh + ct * 2);
        ldv<2>(l1, w + g.o_uh + (CT + ct) * 2);
        ldv<2>(u0, w + g.o_uv + ct * 2);
        ldv<2>(u1, w + g.o_uv + (CT + ct) * 2);
#pragma unroll
        for (int s = 0; s < SPT; ++s) {
            const int a = cl[row0 + s], c = cu[row0 + s];
#pragma unroll
            for (int u = 0; u < 2; ++u)
                acc[u][s] = b[u] + (a == 0 ? l0[u] : a == 1 ? l1[u] : T(0)) + (c == 0 ? u0[u] : c == 1 ? u1[u] : T(0));
        }
    }
#pragma unroll 1
    for (int part = 0; part < 2; ++part) {
        const T* a_ = (part ? hu : hl) + row0;
        const T* w_ = w + (part ? g.o_wv : 0) + ct * 2;
#pragma unroll 2
        for (int k = 0; k < H; ++k) {
            T a[SPT], ww[2];
            ldv<SPT>(a, a_ + k * M);
            ldv<2>(ww, w_ + k * CT * 2);
#pragma unroll
            for (int s = 0; s < SPT; ++s) {
                acc[0][s] = fma(a[s], ww[0], acc[0][s]);
                acc[1][s] = fma(a[s], ww[1], acc[1][s]);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int s = 0; s < SPT; ++s) hn[u][s] = elu_(acc[u][s]);
}

template <typename T> struct MdSmem {
    const T* w;
    T *hl, *hu, *aux;
    uint8_t *sig, *cl, *cu;
};

template <typename T, bool WSMEM>
__device__ __forceinline__ MdSmem<T> md_setup(const MdLayout& g, const MdLaunch& c, const T* __restrict__ pk, int words, int extra,
                                              unsigned char* smem) {
    MdSmem<T> s;
    size_t off = 0;
    if (WSMEM) {
        T* wsm = reinterpret_cast<T*>(smem);
        for (int i = threadIdx.x; i < words; i += blockDim.x) wsm[i] = pk[i];
        s.w = wsm;
        off = ((size_t)words * sizeof(T) + 15) & ~(size_t)15;
    } else {
        s.w = pk;
    }
    s.hl = reinterpret_cast<T*>(smem + off); off += (size_t)g.H * c.M * sizeof(T);
    s.hu = reinterpret_cast<T*>(smem + off); off += (size_t)g.H * c.M * sizeof(T);
    s.aux = reinterpret_cast<T*>(smem + off); off += (size_t)extra * g.H * c.M * sizeof(T) + 8 * (size_t)c.M * sizeof(T);
    s.sig = smem + off; off += (size_t)g.N * c.M;
    s.cl = smem + off; off += c.M;
    s.cu = smem + off;
    return s;
}

// =============================================================================================
// forward: teacher-forced log-probability (optionally stashing per-site head terms) or sampling.
//   samples : uint8 [ns][Nx][Ny]  (in: teacher-forced; out: SAMPLE)
//   hgrid   : T [tiles][N(path)][H][M]      state of every site (scratch for logpsi, stash for E_loc / grad)
//   la_sel/la_oth : double [tiles][N(path)][M] or nullptr
// =============================================================================================
template <typename T, bool WSMEM, bool SAMPLE>
__global__ void __launch_bounds__(512, 1)
md_forward_kernel(MdLayout g, MdLaunch c, const T* __restrict__ pk, uint8_t* __restrict__ samples, int64_t ns, T* __restrict__ hgrid,
                  double* __restrict__ out_lp, double* __restrict__ la_sel, double* __restrict__ la_oth, uint64_t seed,
                  uint64_t sample_offset) {
    constexpr int SPT = VT<T>::SPT;
    extern __shared__ __align__(16) unsigned char smem[];
    MdSmem<T> S = md_setup<T, WSMEM>(g, c, pk, g.PK, 0, smem);
    const int tid = threadIdx.x, M = c.M, N = g.N, H = g.H, CT = g.CT;
    const int st = blockIdx.x;
    const bool is_compute = tid < CT * c.RT;
    const int ct = tid % CT, rt = tid / CT;
    for (int i = tid; i < N * M; i += blockDim.x) {
        const int site = i / M, m = i % M;
        const int64_t row = (int64_t)st * M + m;
        S.sig[i] = (!SAMPLE && row < ns) ? samples[row * N + site] : 0;
    }
    T* grid = hgrid + (size_t)st * N * H * M;
    double acc = 0.0;
    __syncthreads();
    for (int p = 0; p < N; ++p) {
        int x, y, pl, pu, pd;
        md_decode(g, p, x, y, pl, pu, pd);
        {   // stage the neighbour states and input codes
            if (pl < 0) for (int i = tid; i < H * M; i += blockDim.x) S.hl[i] = T(0);
            const T* src = pu >= 0 ? grid + (size_t)pu * H * M : nullptr;
            for (int i = tid; i < H * M; i += blockDim.x) S.hu[i] = src ? src[i] : T(0);
            for (int m = tid; m < M; m += blockDim.x) {
                S.cl[m] = pl >= 0 ? S.sig[md_site(g, pl) * M + m] : (uint8_t)2;
                S.cu[m] = pu >= 0 ? S.sig[md_site(g, pu) * M + m] : (uint8_t)2;
            }
        }
        __syncthreads();
        T hn[2][SPT];
        if (is_compute) md_cell<T>(g, S.w, S.hl, S.hu, S.cl, S.cu, M, ct, rt, hn);
        __syncthreads();
        if (is_compute) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int j = 2 * ct + u;
                if (j < H) {
                    stv<SPT>(S.hl + j * M + rt * SPT, hn[u]);
                    stv<SPT>(grid + ((size_t)p * H + j) * M + rt * SPT, hn[u]);
                }
            }
        }
        __syncthreads();
        if (tid < M) {   // head of site p
            const int m = tid, site = x * g.ny + y;
            T z0, z1;
            dense2<T>(S.hl, S.w + g.o_head, H, M, m, z0, z1);
            int sg;
            if (SAMPLE) {
                const float p0 = (float)(T(1) / (T(1) + (T)exp((double)(z1 - z0))));
                const float u = philox_uniform(seed, sample_offset + (uint64_t)st * M + m, (uint32_t)p);
                sg = u >= p0 ? 1 : 0;
                S.sig[site * M + m] = (uint8_t)sg;
            } else {
                sg = S.sig[site * M + m];
            }
            const double zs = sg ? (double)z1 : (double)z0, zo = sg ? (double)z0 : (double)z1;
            const double ls = log_softmax2(zs, zo);
            acc += ls;
            if (la_sel) {
                la_sel[((size_t)st * N + p) * M + m] = ls;
                la_oth[((size_t)st * N + p) * M + m] = log_softmax2(zo, zs);
            }
        }
        // the barrier after the next staging orders the head's reads of hl / writes of sig
        __syncthreads();
    }
    if (tid < M) out_lp[(size_t)st * M + tid] = acc;
    if (SAMPLE) {
        for (int i = tid; i < N * M; i += blockDim.x) {
            const int m = i / N, site = i % N;
            const int64_t row = (int64_t)st * M + m;
            if (row < ns) samples[row * N + site] = S.sig[site * M + m];
        }
    }
}

// =============================================================================================
// chain kernel (prefix reuse on the zig-zag path): work item = (flipped path position k, tile).
// Sites p > k are recomputed; a site's upper neighbour comes from the recomputed private grid if it lies
// after k on the path, from the stashed base grid otherwise.
//   delta : double [tiles][N (slot = x*Ny+y, the reference's queue slot - 1)][M]
// =============================================================================================
template <typename T, bool WSMEM>
__global__ void __launch_bounds__(512, 1)
md_chain_kernel(MdLayout g, MdLaunch c, int tiles, const T* __restrict__ pk, const uint8_t* __restrict__ samples, int64_t ns,
                const T* __restrict__ hbase, T* __restrict__ hpriv, const double* __restrict__ la_sel,
                const double* __restrict__ la_oth, double* __restrict__ delta, int* __restrict__ counter) {
    constexpr int SPT = VT<T>::SPT;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_work;
    MdSmem<T> S = md_setup<T, WSMEM>(g, c, pk, g.PK, 0, smem);
    const int tid = threadIdx.x, M = c.M, N = g.N, H = g.H, CT = g.CT;
    const bool is_compute = tid < CT * c.RT;
    const int ct = tid % CT, rt = tid / CT;
    T* priv = hpriv + (size_t)blockIdx.x * N * H * M;
    const int total = N * tiles;
    while (true) {
        if (tid == 0) s_work = atomicAdd(counter, 1);
        __syncthreads();
        const int work = s_work;
        if (work >= total) break;
        const int k = work / tiles, st = work % tiles;     // ascending k = longest chains first
        const T* base = hbase + (size_t)st * N * H * M;
        const int ksite = md_site(g, k);
        for (int i = tid; i < N * M; i += blockDim.x) {
            const int site = i / M, m = i % M;
            const int64_t row = (int64_t)st * M + m;
            uint8_t v = row < ns ? samples[row * N + site] : 0;
            if (site == ksite) v = 1 - v;
            S.sig[i] = v;
        }
        {   // h_l of position k+1 (if on the same row) is the unchanged base state of k
            const T* src = base + (size_t)k * H * M;
            for (int i = tid; i < H * M; i += blockDim.x) S.hl[i] = src[i];
        }
        double acc = 0.0;
        if (tid < M) acc = la_oth[((size_t)st * N + k) * M + tid] - la_sel[((size_t)st * N + k) * M + tid];
        __syncthreads();
        for (int p = k + 1; p < N; ++p) {
            int x, y, pl, pu, pd;
            md_decode(g, p, x, y, pl, pu, pd);
            {
                if (pl < 0) for (int i = tid; i < H * M; i += blockDim.x) S.hl[i] = T(0);
                const T* src = pu < 0 ? nullptr : (pu > k ? priv + (size_t)pu * H * M : base + (size_t)pu * H * M);
                for (int i = tid; i < H * M; i += blockDim.x) S.hu[i] = src ? src[i] : T(0);
                for (int m = tid; m < M; m += blockDim.x) {
                    S.cl[m] = pl >= 0 ? S.sig[md_site(g, pl) * M + m] : (uint8_t)2;
                    S.cu[m] = pu >= 0 ? S.sig[md_site(g, pu) * M + m] : (uint8_t)2;
                }
            }
            __syncthreads();
            T hn[2][SPT];
            if (is_compute) md_cell<T>(g, S.w, S.hl, S.hu, S.cl, S.cu, M, ct, rt, hn);
            __syncthreads();
            if (is_compute) {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int j = 2 * ct + u;
                    if (j < H) {
                        stv<SPT>(S.hl + j * M + rt * SPT, hn[u]);
                        stv<SPT>(priv + ((size_t)p * H + j) * M + rt * SPT, hn[u]);
                    }
                }
            }
            __syncthreads();
            if (tid < M) {
                const int m = tid, sg = S.sig[(x * g.ny + y) * M + m];
                T z0, z1;
                dense2<T>(S.hl, S.w + g.o_head, H, M, m, z0, z1);
                const double zs = sg ? (double)z1 : (double)z0, zo = sg ? (double)z0 : (double)z1;
                acc += log_softmax2(zs, zo) - la_sel[((size_t)st * N + p) * M + m];
            }
            __syncthreads();
        }
        if (tid < M) delta[((size_t)st * N + ksite) * M + tid] = acc;
        __syncthreads();
    }
}

// =============================================================================================
// backward: reverse path order.  d h[p] = head term + Wh d pre[p+1] (same row) + Wv d pre[below];
// d pre = d h * elu'(pre) with elu' = 1 (h > 0) or h + 1.  Writes G = d pre and d z for the weight-gradient
// reduction.
// =============================================================================================
template <typename T, bool WSMEM>
__global__ void __launch_bounds__(512, 1)
md_bwd_kernel(MdLayout g, MdLaunch c, const T* __restrict__ pk, const T* __restrict__ pkT, const uint8_t* __restrict__ samples,
              int64_t ns, const T* __restrict__ hgrid, const double* __restrict__ la_oth, const double* __restrict__ weights,
              T* __restrict__ Ggrid, T* __restrict__ dzbuf) {
    constexpr int SPT = VT<T>::SPT;
    extern __shared__ __align__(16) unsigned char smem[];
    MdSmem<T> S = md_setup<T, WSMEM>(g, c, pkT, g.PKT, 1, smem);   // hl := G[p+1], hu := G[below], aux := h[p] | dz[2][M]
    const int tid = threadIdx.x, M = c.M, N = g.N, H = g.H, CT = g.CT;
    const int st = blockIdx.x;
    const bool is_compute = tid < CT * c.RT;
    const int ct = tid % CT, rt = tid / CT, row0 = rt * SPT;
    T* hcur = S.aux;
    T* dz = S.aux + (size_t)H * M;
    const T* grid = hgrid + (size_t)st * N * H * M;
    T* gg = Ggrid + (size_t)st * N * H * M;
    T wd[2][2] = {{0, 0}, {0, 0}};
    if (is_compute)
        for (int u = 0; u < 2; ++u) {
            const int j = 2 * ct + u;
            if (j < H) { wd[u][0] = pk[g.o_head + 2 * j]; wd[u][1] = pk[g.o_head + 2 * j + 1]; }
        }
    for (int i = tid; i < N * M; i += blockDim.x) {
        const int site = i / M, m = i % M;
        const int64_t row = (int64_t)st * M + m;
        S.sig[i] = row < ns ? samples[row * N + site] : 0;
    }
    __syncthreads();
    for (int p = N - 1; p >= 0; --p) {
        int x, y, pl, pu, pd;
        md_decode(g, p, x, y, pl, pu, pd);
        const bool has_next = (p + 1 < N) && ((p + 1) % g.nx != 0);    // p is the horizontal neighbour of p+1
        {
            if (!has_next) for (int i = tid; i < H * M; i += blockDim.x) S.hl[i] = T(0);
            const T* src = pd >= 0 ? gg + (size_t)pd * H * M : nullptr;
            for (int i = tid; i < H * M; i += blockDim.x) S.hu[i] = src ? src[i] : T(0);
            const T* hs = grid + (size_t)p * H * M;
            for (int i = tid; i < H * M; i += blockDim.x) hcur[i] = hs[i];
            for (int m = tid; m < M; m += blockDim.x) {
                const int64_t row = (int64_t)st * M + m;
                const int sg = S.sig[(x * g.ny + y) * M + m];
                const double t = row < ns ? weights[row] * exp(la_oth[((size_t)st * N + p) * M + m]) : 0.0;   // w (1 - p_sel)
                T z[2];
                z[sg] = (T)t;
                z[1 - sg] = (T)(-t);
                dz[m] = z[0];
                dz[M + m] = z[1];
                dzbuf[(((size_t)st * N + p) * 2 + 0) * M + m] = z[0];
                dzbuf[(((size_t)st * N + p) * 2 + 1) * M + m] = z[1];
            }
        }
        __syncthreads();
        T gn[2][SPT];
        if (is_compute) {
            T acc[2][SPT];
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int s = 0; s < SPT; ++s) acc[u][s] = dz[row0 + s] * wd[u][0] + dz[M + row0 + s] * wd[u][1];
#pragma unroll 1
            for (int part = 0; part < 2; ++part) {
                const T* a_ = (part ? S.hu : S.hl) + row0;
                const T* w_ = S.w + (part ? g.o_wvT : 0) + ct * 2;
#pragma unroll 2
                for (int i = 0; i < H; ++i) {
                    T a[SPT], ww[2];
                    ldv<SPT>(a, a_ + i * M);
                    ldv<2>(ww, w_ + i * CT * 2);
#pragma unroll
                    for (int s = 0; s < SPT; ++s) {
                        acc[0][s] = fma(a[s], ww[0], acc[0][s]);
                        acc[1][s] = fma(a[s], ww[1], acc[1][s]);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int j = 2 * ct + u;
                if (j < H) {
                    T h[SPT];
                    ldv<SPT>(h, hcur + j * M + row0);
#pragma unroll
                    for (int s = 0; s < SPT; ++s) gn[u][s] = acc[u][s] * (h[s] > T(0) ? T(1) : h[s] + T(1));
                }
            }
        }
        __syncthreads();
        if (is_compute) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int j = 2 * ct + u;
                if (j < H) {
                    stv<SPT>(S.hl + j * M + row0, gn[u]);
                    stv<SPT>(gg + ((size_t)p * H + j) * M + row0, gn[u]);
                }
            }
        }
        __syncthreads();
    }
}

// weight-gradient reduction:  C[r][c] = sum_{tile, p, m} A[r] B[c]
//   cell mode: A = [h_l (H) | h_u (H) | onehot_l (2) | onehot_u (2) | 1], B = G[p] (H columns)
//   head mode: A = [h[p] (H) | 1],                                       B = dz[p] (2 columns)
template <typename T> struct MdWgArgs {
    const T* hgrid; const T* B; const uint8_t* samples;
    int64_t ns, nblk;
    int head, R, C, M, ksplit, rtiles, ctiles;
};
constexpr int kMdTile = 32;

template <typename T>
__global__ void __launch_bounds__(256) md_wgrad_kernel(MdLayout g, MdWgArgs<T> a, double* __restrict__ partial) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int M = a.M, MS = M + 1, H = g.H, N = g.N;
    T* As = reinterpret_cast<T*>(smem);
    T* Bs = As + kMdTile * MS;
    const int tile = blockIdx.x, ks = blockIdx.y;
    const int r0 = (tile / a.ctiles) * kMdTile, c0 = (tile % a.ctiles) * kMdTile;
    const int tc = threadIdx.x % 16, tr = threadIdx.x / 16;     // 16 x 16 threads, 2 x 2 outputs each
    const int64_t b0 = a.nblk * ks / a.ksplit, b1 = a.nblk * (ks + 1) / a.ksplit;
    double acc[2][2] = {{0, 0}, {0, 0}};
    for (int64_t blk = b0; blk < b1; ++blk) {
        const int p = (int)(blk % N);
        const int64_t st = blk / N;
        int x, y, pl, pu, pd;
        md_decode(g, p, x, y, pl, pu, pd);
        for (int i = threadIdx.x; i < kMdTile * M; i += blockDim.x) {
            const int row = i / M, m = i % M;
            const int r = r0 + row, cc = c0 + row;
            T va = T(0), vb = T(0);
            if (a.head) {
                if (r < H) va = a.hgrid[((size_t)blk * H + r) * M + m];
                else if (r == H) va = T(1);
            } else if (r < H) {
                if (pl >= 0) va = a.hgrid[(((size_t)st * N + pl) * H + r) * M + m];
            } else if (r < 2 * H) {
                if (pu >= 0) va = a.hgrid[(((size_t)st * N + pu) * H + (r - H)) * M + m];
            } else if (r < 2 * H + 4) {
                const int which = (r - 2 * H) >> 1, bit = (r - 2 * H) & 1, q = which ? pu : pl;
                const int64_t rowid = st * M + m;
                if (q >= 0 && rowid < a.ns) va = a.samples[rowid * N + md_site(g, q)] == bit ? T(1) : T(0);
            } else if (r == 2 * H + 4) {
                va = T(1);
            }
            if (cc < a.C) vb = a.B[((size_t)blk * a.C + cc) * M + m];
            As[row * MS + m] = va;
            Bs[row * MS + m] = vb;
        }
        __syncthreads();
        T s00 = 0, s01 = 0, s10 = 0, s11 = 0;
        for (int m = 0; m < M; ++m) {
            const T a0 = As[(tr * 2) * MS + m], a1 = As[(tr * 2 + 1) * MS + m];
            const T b0v = Bs[(tc * 2) * MS + m], b1v = Bs[(tc * 2 + 1) * MS + m];
            s00 = fma(a0, b0v, s00); s01 = fma(a0, b1v, s01); s10 = fma(a1, b0v, s10); s11 = fma(a1, b1v, s11);
        }
        acc[0][0] += (double)s00; acc[0][1] += (double)s01; acc[1][0] += (double)s10; acc[1][1] += (double)s11;
        __syncthreads();
    }
    const int Rp = a.rtiles * kMdTile, Cp = a.ctiles * kMdTile;
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) partial[((size_t)ks * Rp + r0 + tr * 2 + i) * Cp + c0 + tc * 2 + j] = acc[i][j];
}

// operands of the 2-D RNN reduction for wgdm::wgrad_kernel (wgrad_f64mma.cuh; float64 models)
struct WgdmMdSrc {
    MdLayout g;
    MdWgArgs<double> a;
    int M, R, C;
    int64_t nblk;
    struct Ctx { int64_t st; int pl, pu; };
    __device__ __forceinline__ Ctx begin(int64_t blk) const {          // neighbours of the position of block blk on the zig-zag path
        Ctx c;
        c.st = blk / g.N;
        int x, y, pd;
        md_decode(g, (int)(blk % g.N), x, y, c.pl, c.pu, pd);
        return c;
    }
    __device__ __forceinline__ const double* a_src(const Ctx& cx, int64_t blk, int r, int k, bool& special, double& v0, double& v1) const {
        const int H = g.H, N = g.N;
        if (a.head) {
            if (r < H) return a.hgrid + ((size_t)blk * H + r) * M + k;
            special = true;
            v0 = v1 = 1.0;
            return nullptr;
        }
        const int64_t st = cx.st;
        const int pl = cx.pl, pu = cx.pu;
        if (r < H) return pl >= 0 ? a.hgrid + (((size_t)st * N + pl) * H + r) * M + k : nullptr;
        if (r < 2 * H) return pu >= 0 ? a.hgrid + (((size_t)st * N + pu) * H + (r - H)) * M + k : nullptr;
        special = true;
        if (r < 2 * H + 4) {                                              // one-hot of the left / upper spin
            const int which = (r - 2 * H) >> 1, bit = (r - 2 * H) & 1, q = which ? pu : pl;
            const int64_t rowid = st * M + k;
            if (q >= 0) {
                const int site = md_site(g, q);
                if (rowid < a.ns) v0 = a.samples[rowid * N + site] == bit ? 1.0 : 0.0;
                if (rowid + 1 < a.ns) v1 = a.samples[(rowid + 1) * N + site] == bit ? 1.0 : 0.0;
            }
        } else {
            v0 = v1 = 1.0;                                                // bias row
        }
        return nullptr;
    }
    __device__ __forceinline__ const double* b_src(const Ctx&, int64_t blk, int c, int k) const { return a.B + ((size_t)blk * a.C + c) * M + k; }
};

__global__ void md_wgrad_scatter_kernel(MdLayout g, const double* __restrict__ partial, int ksplit, int Rp, int Cp, int head, int R,
                                        int C, double* __restrict__ grad) {
    const int H = g.H;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < R * C; idx += gridDim.x * blockDim.x) {
        const int r = idx / C, cidx = idx % C;
        int dst;
        if (head) dst = r < H ? g.f_wd + r * 2 + cidx : g.f_bd + cidx;
        else if (r < H) dst = r * H + cidx;
        else if (r < 2 * H) dst = g.f_wv + (r - H) * H + cidx;
        else if (r < 2 * H + 2) dst = g.f_uh + (r - 2 * H) * H + cidx;
        else if (r < 2 * H + 4) dst = g.f_uv + (r - 2 * H - 2) * H + cidx;
        else dst = g.f_b + cidx;
        double s = 0.0;
        for (int k = 0; k < ksplit; ++k) s += partial[((size_t)k * Rp + r) * Cp + cidx];
        grad[dst] = s;
    }
}

}  // namespace rnnwf
#include "mdrnn_f64mma.cuh"
namespace rnnwf {

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <typename T> struct MdWs {
    double *mm_wb, *mm_priv;     // float64 DMMA chain kernel (mdrnn_f64mma.cuh): B fragments + table, private state grids
    T *pk, *pkT, *hgrid, *hpriv, *Ggrid, *dzbuf;
    double *lp, *la_sel, *la_oth, *delta, *diag, *partial;
    int* counter;
    int ksplit, Rp, Cp;
};

static inline int md_grid_for(int64_t n, int block = 256) { return (int)std::min<int64_t>(cdiv(n, block), 148 * 16); }

template <typename T>
static MdWs<T> md_carve(Ws& ws, const MdLayout& g, const MdLaunch& c, int64_t tiles, int op, int64_t ns, int sms) {
    MdWs<T> w;
    memset(&w, 0, sizeof(w));
    const size_t rows = (size_t)tiles * c.M;
    w.pk = ws.take<T>(g.PK);
    w.pkT = ws.take<T>(g.PKT);
    w.hgrid = ws.take<T>(rows * g.N * g.H);
    w.lp = ws.take<double>(rows);
    w.counter = ws.take<int>(4);
    if (op == RNNWF_OP_TFIM_ELOC || op == RNNWF_OP_VMC_GRAD) {
        w.la_sel = ws.take<double>(rows * g.N);
        w.la_oth = ws.take<double>(rows * g.N);
    }
    if (op == RNNWF_OP_TFIM_ELOC) {
        w.hpriv = ws.take<T>((size_t)sms * c.M * g.N * g.H);
        w.delta = ws.take<double>(rows * g.N);
        w.diag = ws.take<double>((size_t)ns);
        if (std::is_same<T, double>::value && mdmma::supported(g)) {
            const mdmma::Layout t = mdmma::make_layout(g);
            w.mm_wb = ws.take<double>(t.wb_doubles + t.tab_doubles);
            w.mm_priv = ws.take<double>((size_t)sms * t.priv_doubles);
        }
    }
    if (op == RNNWF_OP_VMC_GRAD) {
        w.Ggrid = ws.take<T>(rows * g.N * g.H);
        w.dzbuf = ws.take<T>(rows * g.N * 2);
        const int R = 2 * g.H + 5;
        w.Rp = (int)cdiv(R, kMdTile) * kMdTile;
        w.Cp = (int)cdiv(g.H, kMdTile) * kMdTile;
        const int ntile = (w.Rp / kMdTile) * (w.Cp / kMdTile);
        w.ksplit = (int)std::max<int64_t>(1, std::min<int64_t>((148 * 8 + ntile - 1) / ntile, tiles * g.N));
        w.partial = ws.take<double>((size_t)w.ksplit * w.Rp * w.Cp);
    }
    return w;
}

static int md_sms() {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    else { cudaGetLastError(); }
    return sms > 0 ? sms : 148;
}

template <typename T> static MdLaunch md_launch_for(const MdLayout& g, int op) {
    if (op == RNNWF_OP_VMC_GRAD) {
        // forward (PK words, 2 tiles) and backward (PKT words, 3 tiles) share M: take the tighter one
        MdLaunch f = choose_md_launch<T>(g, g.PK, 0), b = choose_md_launch<T>(g, g.PKT, 1);
        return b.M <= f.M ? b : f;
    }
    return choose_md_launch<T>(g, g.PK, 0);
}
template <typename T> static MdLaunch md_relaunch(const MdLayout& g, const MdLaunch& ref, int words, int extra) {
    // same M / RT as `ref`, shared-memory footprint of a kernel with `words` resident weights
    MdLaunch c = ref;
    size_t with_w = (((size_t)words * sizeof(T) + 15) & ~(size_t)15) + (size_t)(2 + extra) * g.H * c.M * sizeof(T) + (size_t)g.N * c.M +
                    2 * (size_t)c.M + 8 * (size_t)c.M * sizeof(T) + 64;
    c.w_smem = with_w <= (size_t)kSmemLimit;
    c.smem_bytes = (int)(c.w_smem ? with_w : with_w - (((size_t)words * sizeof(T) + 15) & ~(size_t)15));
    return c;
}

template <typename T> size_t mdrnn_workspace_bytes_t(const rnnwf_model& m, int op, int64_t ns, int flags) {
    (void)flags;
    const MdLayout g = make_md_layout(m);
    const MdLaunch c = md_launch_for<T>(g, op);
    if (c.RT == 0) return 0;
    Ws ws(nullptr, 0);
    md_carve<T>(ws, g, c, cdiv(ns, c.M), op, ns, md_sms());
    return ws.used + 256;
}
template size_t mdrnn_workspace_bytes_t<float>(const rnnwf_model&, int, int64_t, int);
template size_t mdrnn_workspace_bytes_t<double>(const rnnwf_model&, int, int64_t, int);

template <typename K> static int md_set_smem(K kernel, int bytes) {
    RNNWF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    return 0;
}

template <typename T, bool SAMPLE>
static int md_launch_forward(const MdLayout& g, const MdLaunch& c, const MdWs<T>& w, uint8_t* samples, int64_t ns, int tiles, bool stash,
                             uint64_t seed, uint64_t off, cudaStream_t s) {
    const int block = std::max(c.NT, (c.M + 31) & ~31);
    double* ls = stash ? w.la_sel : nullptr;
    double* lo = stash ? w.la_oth : nullptr;
    prof_count();
    if (c.w_smem) {
        auto k = md_forward_kernel<T, true, SAMPLE>;
        if (int e = md_set_smem(k, c.smem_bytes)) return e;
        k<<<tiles, block, c.smem_bytes, s>>>(g, c, w.pk, samples, ns, w.hgrid, w.lp, ls, lo, seed, off);
    } else {
        auto k = md_forward_kernel<T, false, SAMPLE>;
        if (int e = md_set_smem(k, c.smem_bytes)) return e;
        k<<<tiles, block, c.smem_bytes, s>>>(g, c, w.pk, samples, ns, w.hgrid, w.lp, ls, lo, seed, off);
    }
    RNNWF_CUDA(cudaGetLastError());
    return 0;
}

__global__ void md_gather_lp_kernel(const double* __restrict__ lp, int64_t ns, double* __restrict__ out) {
    const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b < ns) out[b] = lp[b];
}

template <typename T> static int md_check(const rnnwf_model& m, const MdLaunch& c) {
    RNNWF_CHECK(m.nx >= 1 && m.ny >= 1, -1, "MDRNN needs a lattice shape");
    RNNWF_CHECK(c.RT > 0, -3, "no launch configuration fits (units=%d, %dx%d)", m.units, m.nx, m.ny);
    return 0;
}

template <typename T>
int mdrnn_sample_t(const rnnwf_model& m, const void* params, int64_t ns, uint64_t seed, uint64_t off, uint8_t* out, void* wsp,
                   size_t wsb, cudaStream_t s) {
    const MdLayout g = make_md_layout(m);
    const MdLaunch c = md_launch_for<T>(g, RNNWF_OP_SAMPLE);
    if (int e = md_check<T>(m, c)) return e;
    const int tiles = (int)cdiv(ns, c.M);
    Ws ws(wsp, wsb);
    MdWs<T> w = md_carve<T>(ws, g, c, tiles, RNNWF_OP_SAMPLE, ns, md_sms());
    RNNWF_CHECK(ws.ok(), -4, "workspace too small: need %zu have %zu", ws.used, wsb);
    prof_count(); md_pack_kernel<T><<<md_grid_for(g.PK + g.PKT), 256, 0, s>>>(g, (const T*)params, w.pk, w.pkT);
    return md_launch_forward<T, true>(g, c, w, out, ns, tiles, false, seed, off, s);
}
template int mdrnn_sample_t<float>(const rnnwf_model&, const void*, int64_t, uint64_t, uint64_t, uint8_t*, void*, size_t, cudaStream_t);
template int mdrnn_sample_t<double>(const rnnwf_model&, const void*, int64_t, uint64_t, uint64_t, uint8_t*, void*, size_t, cudaStream_t);

template <typename T>
int mdrnn_logpsi_t(const rnnwf_model& m, const void* params, const uint8_t* samples, int64_t ns, double* out, void* wsp, size_t wsb,
                   cudaStream_t s) {
    const MdLayout g = make_md_layout(m);
    const MdLaunch c = md_launch_for<T>(g, RNNWF_OP_LOGPSI);
    if (int e = md_check<T>(m, c)) return e;
    const int tiles = (int)cdiv(ns, c.M);
    Ws ws(wsp, wsb);
    MdWs<T> w = md_carve<T>(ws, g, c, tiles, RNNWF_OP_LOGPSI, ns, md_sms());
    RNNWF_CHECK(ws.ok(), -4, "workspace too small: need %zu have %zu", ws.used, wsb);
    prof_count(); md_pack_kernel<T><<<md_grid_for(g.PK + g.PKT), 256, 0, s>>>(g, (const T*)params, w.pk, w.pkT);
    if (int e = md_launch_forward<T, false>(g, c, w, const_cast<uint8_t*>(samples), ns, tiles, false, 0, 0, s)) return e;
    prof_count(); md_gather_lp_kernel<<<md_grid_for(ns), 256, 0, s>>>(w.lp, ns, out);
    RNNWF_CUDA(cudaGetLastError());
    return 0;
}
template int mdrnn_logpsi_t<float>(const rnnwf_model&, const void*, const uint8_t*, int64_t, double*, void*, size_t, cudaStream_t);
template int mdrnn_logpsi_t<double>(const rnnwf_model&, const void*, const uint8_t*, int64_t, double*, void*, size_t, cudaStream_t);

template <typename T>
int mdrnn_tfim_eloc_t(const rnnwf_model& m, const void* params, const uint8_t* samples, int64_t ns, const double* jz, double bx,
                      double* eloc, double* logp, double* ratios, void* wsp, size_t wsb, cudaStream_t s) {
    const MdLayout g = make_md_layout(m);
    const MdLaunch c = md_launch_for<T>(g, RNNWF_OP_TFIM_ELOC);
    if (int e = md_check<T>(m, c)) return e;
    const int tiles = (int)cdiv(ns, c.M), sms = md_sms();
    Ws ws(wsp, wsb);
    MdWs<T> w = md_carve<T>(ws, g, c, tiles, RNNWF_OP_TFIM_ELOC, ns, sms);
    RNNWF_CHECK(ws.ok(), -4, "workspace too small: need %zu have %zu", ws.used, wsb);
    prof_count(); md_pack_kernel<T><<<md_grid_for(g.PK + g.PKT), 256, 0, s>>>(g, (const T*)params, w.pk, w.pkT);
    if (int e = tfim_diag_impl(m, samples, ns, jz, w.diag, s)) return e;
    if (int e = md_launch_forward<T, false>(g, c, w, const_cast<uint8_t*>(samples), ns, tiles, true, 0, 0, s)) return e;
    const char* env = getenv("RNNWF_CHAIN");
    bool dmma_done = false;
    if constexpr (std::is_same<T, double>::value) {
        if (bx != 0.0 && w.mm_wb && !(env && strcmp(env, "ffma") == 0)) {   // DMMA chain kernel; RNNWF_CHAIN=ffma keeps the thread-tile kernel (A/B)
            if (int e = mdmma::launch(g, c.M, ns, (const double*)params, w.mm_wb, w.mm_wb + mdmma::make_layout(g).wb_doubles, w.mm_priv, samples,
                                      w.hgrid, w.la_sel, w.la_oth, w.delta, w.counter, sms, s))
                return e;
            dmma_done = true;
        }
    }
    if (bx != 0.0 && !dmma_done) {
        RNNWF_CUDA(cudaMemsetAsync(w.counter, 0, sizeof(int), s));
        const int block = std::max(c.NT, (c.M + 31) & ~31);
        const int grid = (int)std::min<int64_t>((int64_t)g.N * tiles, sms);
        prof_count();
        prof_mark(0, s);
        if (c.w_smem) {
            auto k = md_chain_kernel<T, true>;
            if (int e = md_set_smem(k, c.smem_bytes)) return e;
            k<<<grid, block, c.smem_bytes, s>>>(g, c, tiles, w.pk, samples, ns, w.hgrid, w.hpriv, w.la_sel, w.la_oth, w.delta, w.counter);
        } else {
            auto k = md_chain_kernel<T, false>;
            if (int e = md_set_smem(k, c.smem_bytes)) return e;
            k<<<grid, block, c.smem_bytes, s>>>(g, c, tiles, w.pk, samples, ns, w.hgrid, w.hpriv, w.la_sel, w.la_oth, w.delta, w.counter);
        }
        prof_mark(1, s);
        RNNWF_CUDA(cudaGetLastError());
    }
    return tfim_finalize_impl(w.diag, w.delta, w.lp, ns, g.N, c.M, tiles, bx, 0, eloc, logp, ratios, s);
}
template int mdrnn_tfim_eloc_t<float>(const rnnwf_model&, const void*, const uint8_t*, int64_t, const double*, double, double*,
                                      double*, double*, void*, size_t, cudaStream_t);
template int mdrnn_tfim_eloc_t<double>(const rnnwf_model&, const void*, const uint8_t*, int64_t, const double*, double, double*,
                                       double*, double*, void*, size_t, cudaStream_t);

template <typename T>
static int md_launch_wgrad(const MdLayout& g, const MdWs<T>& w, const uint8_t* samples, int64_t ns, int M, int64_t nblk, bool head,
                           double* grad, cudaStream_t s) {
    MdWgArgs<T> a;
    memset(&a, 0, sizeof(a));
    a.hgrid = w.hgrid; a.samples = samples; a.ns = ns; a.nblk = nblk; a.M = M; a.head = head ? 1 : 0;
    a.B = head ? w.dzbuf : w.Ggrid;
    a.R = head ? g.H + 1 : 2 * g.H + 5;
    a.C = head ? 2 : g.H;
    a.rtiles = (int)cdiv(a.R, kMdTile);
    a.ctiles = (int)cdiv(a.C, kMdTile);
    a.ksplit = w.ksplit;
    bool dmma_done = false;
    if constexpr (std::is_same<T, double>::value) {
        if (M % 2 == 0 && !getenv("RNNWF_WGRAD_FFMA")) {       // DMMA reduction (the reference's 2-D RNN computes in float64)
            const int Rp = a.rtiles * kMdTile, Cp = a.ctiles * kMdTile;
            const int64_t slots = ((int64_t)w.ksplit * w.Rp * w.Cp) / ((int64_t)Rp * Cp);
            a.ksplit = wgdm::choose_ksplit(a.R, a.C, nblk, slots);
            WgdmMdSrc src;
            src.g = g; src.a = a; src.M = M; src.R = a.R; src.C = a.C; src.nblk = nblk;
            auto k = wgdm::wgrad_kernel<WgdmMdSrc>;
            if (int e = md_set_smem(k, (int)wgdm::kSmem)) return e;
            const int ct = (int)cdiv(a.C, wgdm::kT);
            prof_count(); k<<<dim3((int)cdiv(a.R, wgdm::kT) * ct, a.ksplit), wgdm::kThreads, wgdm::kSmem, s>>>(src, ct, a.ksplit, w.partial, Rp, Cp);
            RNNWF_CUDA(cudaGetLastError());
            dmma_done = true;
        }
    }
    if (!dmma_done) {
        const int smem = 2 * kMdTile * (M + 1) * (int)sizeof(T);
        auto k = md_wgrad_kernel<T>;
        if (int e = md_set_smem(k, smem)) return e;
        prof_count(); k<<<dim3(a.rtiles * a.ctiles, a.ksplit), 256, smem, s>>>(g, a, w.partial);
        RNNWF_CUDA(cudaGetLastError());
    }
    prof_count(); md_wgrad_scatter_kernel<<<md_grid_for((int64_t)a.R * a.C), 256, 0, s>>>(g, w.partial, a.ksplit, a.rtiles * kMdTile,
                                                                                      a.ctiles * kMdTile, a.head, a.R, a.C, grad);
    RNNWF_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
int mdrnn_vmc_grad_t(const rnnwf_model& m, const void* params, const uint8_t* samples, int64_t ns, const double* weights, double* grad,
                     void* wsp, size_t wsb, cudaStream_t s) {
    const MdLayout g = make_md_layout(m);
    const MdLaunch c0 = md_launch_for<T>(g, RNNWF_OP_VMC_GRAD);
    if (int e = md_check<T>(m, c0)) return e;
    const MdLaunch cf = md_relaunch<T>(g, c0, g.PK, 0), cb = md_relaunch<T>(g, c0, g.PKT, 1);
    const int tiles = (int)cdiv(ns, c0.M);
    Ws ws(wsp, wsb);
    MdWs<T> w = md_carve<T>(ws, g, c0, tiles, RNNWF_OP_VMC_GRAD, ns, md_sms());
    RNNWF_CHECK(ws.ok(), -4, "workspace too small: need %zu have %zu", ws.used, wsb);
    prof_count(); md_pack_kernel<T><<<md_grid_for(g.PK + g.PKT), 256, 0, s>>>(g, (const T*)params, w.pk, w.pkT);
    if (int e = md_launch_forward<T, false>(g, cf, w, const_cast<uint8_t*>(samples), ns, tiles, true, 0, 0, s)) return e;
    const int block = std::max(cb.NT, (cb.M + 31) & ~31);
    prof_count();
    if (cb.w_smem) {
        auto k = md_bwd_kernel<T, true>;
        if (int e = md_set_smem(k, cb.smem_bytes)) return e;
        k<<<tiles, block, cb.smem_bytes, s>>>(g, cb, w.pk, w.pkT, samples, ns, w.hgrid, w.la_oth, weights, w.Ggrid, w.dzbuf);
    } else {
        auto k = md_bwd_kernel<T, false>;
        if (int e = md_set_smem(k, cb.smem_bytes)) return e;
        k<<<tiles, block, cb.smem_bytes, s>>>(g, cb, w.pk, w.pkT, samples, ns, w.hgrid, w.la_oth, weights, w.Ggrid, w.dzbuf);
    }
    RNNWF_CUDA(cudaGetLastError());
    if (int e = md_launch_wgrad<T>(g, w, samples, ns, c0.M, (int64_t)tiles * g.N, true, grad, s)) return e;
    return md_launch_wgrad<T>(g, w, samples, ns, c0.M, (int64_t)tiles * g.N, false, grad, s);
}
template int mdrnn_vmc_grad_t<float>(const rnnwf_model&, const void*, const uint8_t*, int64_t, const double*, double*, void*, size_t,
                                     cudaStream_t);
template int mdrnn_vmc_grad_t<double>(const rnnwf_model&, const void*, const uint8_t*, int64_t, const double*, double*, void*, size_t,
                                      cudaStream_t);

}  // namespace rnnwf
