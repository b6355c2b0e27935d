// gru_kernels.cuh — persistent GRU kernels: teacher-forced base pass, autoregressive sampler and the
// prefix-reuse chain kernel that evaluates every connected configuration (TFIM flips / J1-J2 exchanges).
#pragma once
#include "gru_engine.cuh"

namespace rnnwf {

constexpr double kPi = 3.14159265358979323846;

// ---------------------------------------------------------------------------------------------
// Head evaluation for one row at site n (dense + softmax [+ phase, U(1) mask]).
//  PROB    : l = log softmax(z)[.]                               1DTFIM/RNNwavefunction.py:109,113-116
//  COMPLEX : l = log amplitude, p = phase                        J1J2/ComplexRNNwavefunction.py:143-167
//            amplitude = sqrt(softmax) * mask / l2norm (:147-155): with both outcomes allowed the norm is 1;
//            with one outcome masked the other has amplitude exactly 1; a masked outcome has amplitude 0.
// n_up = number of up spins among sites < n of the configuration being evaluated.
// ---------------------------------------------------------------------------------------------
template <typename T, bool COMPLEX>
__device__ __forceinline__ void head_eval(const GruLayout& g, const T* __restrict__ htop, const T* __restrict__ wd,
                                          int M, int m, int n, int sigma, int n_up, double& l_sel, double& l_oth,
                                          double& p_sel, double& p_oth) {
    T z0, z1;
    dense2<T>(htop, wd, g.H, M, m, z0, z1);
    const double zs = sigma ? (double)z1 : (double)z0, zo = sigma ? (double)z0 : (double)z1;
    l_sel = log_softmax2(zs, zo);
    l_oth = log_softmax2(zo, zs);
    p_sel = 0.0;
    p_oth = 0.0;
    if (COMPLEX) {
        l_sel *= 0.5;
        l_oth *= 0.5;
        if (2 * n >= g.N) {
            const int half = g.N / 2, n_dn = n - n_up;
            const bool ok_dn = (half - 1 - n_dn) >= 0, ok_up = (half - 1 - n_up) >= 0;
            const bool ok_sel = sigma ? ok_up : ok_dn, ok_oth = sigma ? ok_dn : ok_up;
            const double ninf = -__longlong_as_double(0x7ff0000000000000LL);
            if (!ok_sel) l_sel = ninf; else if (!ok_oth) l_sel = 0.0;
            if (!ok_oth) l_oth = ninf; else if (!ok_sel) l_oth = 0.0;
        }
        T y0, y1;
        dense2<T>(htop, wd + 2 * g.H + 2, g.H, M, m, y0, y1);
        const double ys = sigma ? (double)y1 : (double)y0, yo = sigma ? (double)y0 : (double)y1;
        p_sel = kPi * ys / (1.0 + fabs(ys));
        p_oth = kPi * yo / (1.0 + fabs(yo));
    }
}

struct TileSmem {
    unsigned char* base;
};

template <typename T, bool WSMEM>
__device__ __forceinline__ void tile_setup(const GruLayout& g, const GruLaunch& c, const T* __restrict__ pk,
                                           unsigned char* smem, const T*& w, T*& hbuf, uint8_t*& sig, WRing<T>* ring = nullptr) {
    size_t off = 0;
    if (WSMEM) {
        T* wsm = reinterpret_cast<T*>(smem);
        for (int i = threadIdx.x; i < g.PK; i += blockDim.x) wsm[i] = pk[i];
        w = wsm;
        off = ((size_t)g.PK * sizeof(T) + 15) & ~(size_t)15;
    } else {
        w = pk;
    }
    hbuf = reinterpret_cast<T*>(smem + off);
    off += (size_t)g.L * g.H * c.M * sizeof(T);
    sig = smem + off;
    if (ring) {   // weight ring behind the spin codes (16-byte aligned); KC == 0: no ring
        off = (off + 2 * (size_t)c.Mp + 64 + 15) & ~(size_t)15;
        ring->buf = reinterpret_cast<T*>(smem + off);
        ring->KC = WSMEM ? 0 : c.ring_kc;
        ring->nthr = c.NTc;
        ring->tid = threadIdx.x;
    }
}

// =============================================================================================
// Base pass: teacher-forced forward over a tile of M samples.  Optionally stashes every layer's state
// after every site (hstore[tile][n][l][j][m]) and the selected/other head outputs, which is what the
// chain kernel (prefix reuse) and the BPTT kernels start from.
//   sigT   : uint8 [tiles][N][M]   transposed samples (row-contiguous per site)
//   out_re : double [tiles*M]      sum_n l_sel   (log-probability / log-amplitude)
//   out_im : double [tiles*M]      sum_n p_sel   (COMPLEX only)
// Head warps evaluate the head of site n-1 while the compute warps run layer 0 of site n.
// =============================================================================================
template <typename T, bool WSMEM, bool STASH, bool COMPLEX>
__global__ void __launch_bounds__(512, 1)
gru_forward_kernel(GruLayout g, GruLaunch c, const T* __restrict__ pk, const uint8_t* __restrict__ sigT,
                   double* __restrict__ out_re, double* __restrict__ out_im, T* __restrict__ hstore,
                   double* __restrict__ la_sel, double* __restrict__ la_oth, double* __restrict__ ph_sel,
                   double* __restrict__ ph_oth) {
    extern __shared__ __align__(16) unsigned char smem[];
    const T* w; T* hbuf; uint8_t* sig;
    WRing<T> ring;
    tile_setup<T, WSMEM>(g, c, pk, smem, w, hbuf, sig, &ring);
    const WRing<T>* ringp = (!WSMEM && ring.KC > 0) ? &ring : nullptr;
    const int tid = threadIdx.x, M = c.M, Mp = c.Mp, N = g.N, H = g.H, L = g.L;
    const int st = blockIdx.x;
    for (int i = tid; i < L * H * M; i += blockDim.x) hbuf[i] = T(0);
    for (int i = tid; i < 2 * Mp; i += blockDim.x) sig[i] = 2;
    __syncthreads();
    const bool is_compute = tid < c.CT * c.RT;
    const int ct = tid % c.CT, rt = tid / c.CT;
    const bool is_head = tid >= c.NTc;
    const int ht = tid - c.NTc;
    const uint8_t* sigtile = sigT + (size_t)st * N * M;
    const T* htop = hbuf + (L - 1) * H * M;
    const T* wd = w + g.pk_head;
    double acc_re[2] = {0.0, 0.0}, acc_im[2] = {0.0, 0.0};
    int sprev[2] = {0, 0}, nup[2] = {0, 0};

    for (int n = 0; n <= N; ++n) {
        if (is_head) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int m = ht + r * kHeadThreads;
                if (m < M) {
                    int s_n = 0;
                    if (n < N) {
                        s_n = sigtile[(size_t)n * M + m];
                        sig[((n + 1) & 1) * Mp + m] = (uint8_t)s_n;
                    }
                    if (n > 0) {
                        double ls, lo, ps, po;
                        head_eval<T, COMPLEX>(g, htop, wd, M, m, n - 1, sprev[r], nup[r], ls, lo, ps, po);
                        acc_re[r] += ls;
                        acc_im[r] += ps;
                        nup[r] += sprev[r];
                        if (STASH) {
                            const size_t o = ((size_t)st * N + (n - 1)) * M + m;
                            la_sel[o] = ls;
                            la_oth[o] = lo;
                            if (COMPLEX) { ph_sel[o] = ps; ph_oth[o] = po; }
                        }
                    }
                    sprev[r] = s_n;
                }
            }
        }
        if (n == N) break;
        T* stash = STASH ? hstore + ((size_t)st * N + n) * L * H * M : nullptr;
        gru_site<T, STASH>(g, w, hbuf, sig + (n & 1) * Mp, M, ct, rt, is_compute, stash, ringp);
    }
    if (is_head) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int m = ht + r * kHeadThreads;
            if (m < M) {
                out_re[(size_t)st * M + m] = acc_re[r];
                if (COMPLEX) out_im[(size_t)st * M + m] = acc_im[r];
            }
        }
    }
}

// =============================================================================================
// K1 sampler: N sequential sites in one launch; the draw of site n feeds site n+1
// (1DTFIM/RNNwavefunction.py:65-70; U(1) mask J1J2/ComplexRNNwavefunction.py:85-95).
//   sampT : uint8 [tiles][N][M] (transposed; a second kernel lays it out as [ns][N])
// =============================================================================================
template <typename T, bool WSMEM, bool COMPLEX>
__global__ void __launch_bounds__(512, 1)
gru_sample_kernel(GruLayout g, GruLaunch c, const T* __restrict__ pk, uint8_t* __restrict__ sampT, uint64_t seed,
                  uint64_t sample_offset) {
    extern __shared__ __align__(16) unsigned char smem[];
    const T* w; T* hbuf; uint8_t* sig;
    WRing<T> ring;
    tile_setup<T, WSMEM>(g, c, pk, smem, w, hbuf, sig, &ring);
    const WRing<T>* ringp = (!WSMEM && ring.KC > 0) ? &ring : nullptr;
    const int tid = threadIdx.x, M = c.M, Mp = c.Mp, N = g.N, H = g.H, L = g.L;
    const int st = blockIdx.x;
    for (int i = tid; i < L * H * M; i += blockDim.x) hbuf[i] = T(0);
    for (int i = tid; i < 2 * Mp; i += blockDim.x) sig[i] = 2;
    __syncthreads();
    const bool is_compute = tid < c.CT * c.RT;
    const int ct = tid % c.CT, rt = tid / c.CT;
    const bool is_head = tid >= c.NTc;
    const int ht = tid - c.NTc;
    const T* htop = hbuf + (L - 1) * H * M;
    const T* wd = w + g.pk_head;
    int nup[2] = {0, 0};
    for (int n = 0; n < N; ++n) {
        gru_site<T, false>(g, w, hbuf, sig + (n & 1) * Mp, M, ct, rt, is_compute, nullptr, ringp);
        if (is_head) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int m = ht + r * kHeadThreads;
                if (m < M) {
                    T z0, z1;
                    dense2<T>(htop, wd, H, M, m, z0, z1);
                    // P(sigma = 0) of the softmax in T precision, compared in float (oracle convention)
                    const float p0 = (float)(T(1) / (T(1) + (T)exp((double)(z1 - z0))));
                    const uint64_t id = sample_offset + (uint64_t)st * M + m;
                    const float u = philox_uniform(seed, id, (uint32_t)n);
                    int s = u >= p0 ? 1 : 0;
                    if (COMPLEX && 2 * n >= N) {
                        const int half = N / 2, n_dn = n - nup[r];
                        const bool ok_dn = (half - 1 - n_dn) >= 0, ok_up = (half - 1 - nup[r]) >= 0;
                        if (!ok_up) s = 0;
                        else if (!ok_dn) s = 1;
                    }
                    nup[r] += s;
                    sig[((n + 1) & 1) * Mp + m] = (uint8_t)s;
                    sampT[((size_t)st * N + n) * M + m] = (uint8_t)s;
                }
            }
        }
        __syncthreads();
    }
}

// =============================================================================================
// K2 chain kernel (prefix reuse).  A tile = (slot, sample tile): all M rows restart
// from the stashed base state after site `s` and re-run sites s+1..N-1 of the connected configuration
//   kind 0: sigma with site s flipped                        (TFIM, 1DTFIM/TrainingRNN_1DTFIM.py:43-48)
//   kind 1: sigma with sites s, s+1 exchanged                (J1-J2 NN,  J1J2/TrainingRNN_J1J2.py:68-81)
//   kind 2: sigma with sites s, s+2 exchanged                (J1-J2 NNN, :83-92)
// and accumulate  delta = log psi(sigma') - log psi(sigma)  as a sum of per-site DIFFERENCES
// (better conditioned than the reference's difference of two O(N) sums, :74).
// Tiles are handed out longest-first through an atomic counter to one persistent CTA per SM.
//   delta_re/im : double [tiles][nslots][M]
// =============================================================================================
struct ChainPlan {
    int nslots;       // connected configurations per sample
    int n_kind0;      // slots [0, n_kind0): kind 0 with s = slot
    int n_kind1;      // next n_kind1 slots: kind 1 with s = slot - n_kind0
    int n_kind2;      // next n_kind2 slots: kind 2
    int tiles;        // sample tiles (all directions)
    const double* j1; // optional couplings: a kind-1 (kind-2) slot with j1[s] == 0 (j2[s] == 0) is skipped,
    const double* j2; //   as the reference's `J[site] != 0.0` guards do (J1J2/TrainingRNN_J1J2.py:69,84)
};

__device__ __forceinline__ void chain_decode(const ChainPlan& p, int slot, int& kind, int& s) {
    if (slot < p.n_kind0) { kind = 0; s = slot; }
    else if (slot < p.n_kind0 + p.n_kind1) { kind = 1; s = slot - p.n_kind0; }
    else { kind = 2; s = slot - p.n_kind0 - p.n_kind1; }
}

template <typename T, bool WSMEM, bool COMPLEX>
__global__ void __launch_bounds__(512, 1)
gru_chain_kernel(GruLayout g, GruLaunch c, ChainPlan plan, const T* __restrict__ pk, const uint8_t* __restrict__ sigT,
                 const T* __restrict__ hstore, const double* __restrict__ la_sel, const double* __restrict__ la_oth,
                 const double* __restrict__ ph_sel, const double* __restrict__ ph_oth, const int* __restrict__ order,
                 double* __restrict__ delta_re, double* __restrict__ delta_im, int* __restrict__ counter) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_work;
    const T* w; T* hbuf; uint8_t* sig;
    WRing<T> ring;
    tile_setup<T, WSMEM>(g, c, pk, smem, w, hbuf, sig, &ring);
    const WRing<T>* ringp = (!WSMEM && ring.KC > 0) ? &ring : nullptr;
    const int tid = threadIdx.x, M = c.M, Mp = c.Mp, N = g.N, H = g.H, L = g.L;
    const bool is_compute = tid < c.CT * c.RT;
    const int ct = tid % c.CT, rt = tid / c.CT;
    const bool is_head = tid >= c.NTc;
    const int ht = tid - c.NTc;
    const T* htop = hbuf + (L - 1) * H * M;
    const T* wd = w + g.pk_head;
    const int total = plan.nslots * plan.tiles;

    while (true) {
        if (tid == 0) s_work = atomicAdd(counter, 1);
        __syncthreads();
        const int work = s_work;
        if (work >= total) break;
        // `order` lists slots by decreasing chain length; consecutive work items share the slot
        const int slot = order[work / plan.tiles], st = work % plan.tiles;
        int kind, s;
        chain_decode(plan, slot, kind, s);
        const int t = kind == 0 ? -1 : s + kind;           // second modified site (exchange partner)
        if ((kind == 1 && plan.j1 && plan.j1[s] == 0.0) || (kind == 2 && plan.j2 && plan.j2[s] == 0.0)) {
            __syncthreads();   // s_work is rewritten at the top of the loop
            continue;
        }
        const uint8_t* sigtile = sigT + (size_t)st * N * M;
        {   // restart state: every layer's h after site s of the base pass
            const T* src = hstore + ((size_t)st * N + s) * L * H * M;
            for (int i = tid; i < L * H * M; i += blockDim.x) hbuf[i] = src[i];
        }
        double acc_re[2] = {0.0, 0.0}, acc_im[2] = {0.0, 0.0};
        int sprev[2] = {0, 0}, nup[2] = {0, 0};
        if (is_head) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int m = ht + r * kHeadThreads;
                if (m < M) {
                    const size_t o = ((size_t)st * N + s) * M + m;
                    acc_re[r] = la_oth[o] - la_sel[o];
                    if (COMPLEX) {
                        acc_im[r] = ph_oth[o] - ph_sel[o];
                        int cnt = 0;
                        for (int q = 0; q < s; ++q) cnt += sigtile[(size_t)q * M + m];
                        nup[r] = cnt;
                    }
                    const int sp = 1 - (int)sigtile[(size_t)s * M + m];
                    sprev[r] = sp;
                    nup[r] += sp;
                    sig[((s + 1) & 1) * Mp + m] = (uint8_t)sp;
                }
            }
        }
        __syncthreads();
        for (int n = s + 1; n <= N; ++n) {
            if (is_head) {
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int m = ht + r * kHeadThreads;
                    if (m < M) {
                        int s_n = 0;
                        if (n < N) {
                            s_n = sigtile[(size_t)n * M + m];
                            if (n == t) s_n = 1 - s_n;
                            sig[((n + 1) & 1) * Mp + m] = (uint8_t)s_n;
                        }
                        if (n > s + 1) {
                            double ls, lo, ps, po;
                            head_eval<T, COMPLEX>(g, htop, wd, M, m, n - 1, sprev[r], nup[r] - sprev[r], ls, lo, ps, po);
                            const size_t o = ((size_t)st * N + (n - 1)) * M + m;
                            // per-site difference against the base configuration's own term
                            acc_re[r] += ls - la_sel[o];
                            if (COMPLEX) acc_im[r] += ps - ph_sel[o];
                        }
                        sprev[r] = s_n;
                        nup[r] += s_n;
                    }
                }
            }
            if (n == N) break;
            gru_site<T, false>(g, w, hbuf, sig + (n & 1) * Mp, M, ct, rt, is_compute, nullptr, ringp);
        }
        if (is_head) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int m = ht + r * kHeadThreads;
                if (m < M) {
                    const size_t o = ((size_t)st * plan.nslots + slot) * M + m;
                    delta_re[o] = acc_re[r];
                    if (COMPLEX) delta_im[o] = acc_im[r];
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace rnnwf
