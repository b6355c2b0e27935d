// gru.cu — host launchers + utility kernels for the GRU wave functions (1-D pRNN, parity-symmetric pRNN,
// 1-D RNN over a flattened 2-D lattice, complex cRNN) and the TFIM / J1-J2 local energies built on them.
#include <stdlib.h>
#include <type_traits>
#include "gru_kernels.cuh"
#include "host_util.cuh"
#include "api_internal.h"

namespace rnnwf {

// ---------------------------------------------------------------------------------------------
// utility kernels
// ---------------------------------------------------------------------------------------------
// samples [ns][N] -> sigT [ndir*tiles_s][N][M]; direction 1 holds the site-reversed configurations
// (samples[:, ::-1], 1DTFIM/RNNwavefunction_paritysym.py:125).  Rows beyond ns are zero.
__global__ void sig_transpose_kernel(const uint8_t* __restrict__ samples, uint8_t* __restrict__ sigT, int64_t ns, int N,
                                     int M, int tiles_s, int ndir) {
    const int64_t total = (int64_t)ndir * tiles_s * N * M;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(i % M);
        const int n = (int)((i / M) % N);
        const int st = (int)(i / ((int64_t)M * N));
        const int dir = st / tiles_s;
        const int64_t row = (int64_t)(st % tiles_s) * M + m;
        uint8_t v = 0;
        if (row < ns) v = samples[row * N + (dir ? N - 1 - n : n)];
        sigT[i] = v;
    }
}

__global__ void samp_untranspose_kernel(const uint8_t* __restrict__ sampT, uint8_t* __restrict__ samples, int64_t ns, int N,
                                        int M) {
    const int64_t total = ns * N;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / N;
        const int n = (int)(i % N);
        samples[i] = sampT[((row / M) * N + n) * M + (row % M)];
    }
}

// Diagonal TFIM energy, bit-exact with the reference's f64 accumulation order:
//  chain   : e += v_i * (-Jz[i]) bond by bond                  1DTFIM/TrainingRNN_1DTFIM.py:31-38
//  lattice : e += np.sum(v * (-Jz[i,:]), axis=1) for i<Nx-1, then e += np.sum(v * (-Jz[:,i]), axis=1) for i<Ny-1
//            with NumPy's pairwise order inside np.sum       2DTFIM_2DRNN/Training2DRNN_2DTFIM.py:33-49
__global__ void tfim_diag_kernel(const uint8_t* __restrict__ samples, int64_t ns, int N, int nx, int ny,
                                 const double* __restrict__ jz, double* __restrict__ diag) {
    const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= ns) return;
    const uint8_t* s = samples + b * N;
    double e = 0.0;
    if (nx <= 0) {
        for (int i = 0; i < N - 1; ++i) {
            const double v = s[i] == s[i + 1] ? 1.0 : -1.0;
            e += v * (-jz[i]);
        }
    } else {
        for (int i = 0; i < nx - 1; ++i) {
            auto f = [&](int y) { return (s[i * ny + y] == s[(i + 1) * ny + y] ? 1.0 : -1.0) * (-jz[i * ny + y]); };
            e += np_pairwise_sum(f, 0, ny);
        }
        for (int i = 0; i < ny - 1; ++i) {
            auto f = [&](int x) { return (s[x * ny + i] == s[x * ny + i + 1] ? 1.0 : -1.0) * (-jz[x * ny + i]); };
            e += np_pairwise_sum(f, 0, nx);
        }
    }
    diag[b] = e;
}

// queue[slot][b][i]: slot 0 = sample, slot k+1 = site k flipped (1DTFIM/TrainingRNN_1DTFIM.py:40-48)
__global__ void tfim_enumerate_kernel(const uint8_t* __restrict__ samples, int64_t ns, int N, int32_t* __restrict__ queue) {
    const int64_t total = (int64_t)(N + 1) * ns * N;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int site = (int)(i % N);
        const int64_t b = (i / N) % ns;
        const int slot = (int)(i / ((int64_t)N * ns));
        int v = samples[b * N + site];
        if (slot > 0 && site == slot - 1) v = 1 - v;
        queue[i] = v;
    }
}

__device__ __forceinline__ double logaddexp_(double a, double b) {
    const double mx = fmax(a, b), mn = fmin(a, b);
    if (isinf(mx)) return mx;
    return mx + log1p(exp(mn - mx));
}

// E_loc = diag - Bx * sum_k exp(0.5 * delta_k)      (1DTFIM/TrainingRNN_1DTFIM.py:74), summed in site order.
// Parity model: P_sym = (P(s) + P(rev s))/2, the flip at site k appears at N-1-k in the reversed chain.
__global__ void tfim_finalize_kernel(const double* __restrict__ diag, const double* __restrict__ delta,
                                     const double* __restrict__ lp, int64_t ns, int N, int M, int tiles_s, double bx,
                                     int parity, double* __restrict__ eloc, double* __restrict__ logp, double* __restrict__ ratios) {
    const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= ns) return;
    const int64_t st = b / M;
    const int m = (int)(b % M);
    const double ln2 = 0.69314718055994530942;
    double sum = 0.0, lpb;
    if (!parity) {
        lpb = lp[st * M + m];
        if (bx != 0.0)
            for (int k = 0; k < N; ++k) {
                const double r = exp(0.5 * delta[(st * N + k) * M + m]);
                sum += r;
                if (ratios) ratios[b * N + k] = r;
            }
    } else {
        const int64_t st2 = st + tiles_s;
        const double lp1 = lp[st * M + m], lp2 = lp[st2 * M + m];
        lpb = logaddexp_(lp1, lp2) - ln2;
        if (bx != 0.0)
            for (int k = 0; k < N; ++k) {
                const double a = lp1 + delta[(st * N + k) * M + m];
                const double c = lp2 + delta[(st2 * N + (N - 1 - k)) * M + m];
                const double r = exp(0.5 * (logaddexp_(a, c) - ln2 - lpb));
                sum += r;
                if (ratios) ratios[b * N + k] = r;
            }
    }
    eloc[b] = diag[b] - bx * sum;
    if (logp) logp[b] = lpb;
}

__global__ void gather_logpsi_kernel(const double* __restrict__ lp_re, const double* __restrict__ lp_im, int64_t ns, int M,
                                     int tiles_s, int parity, int complex_, double* __restrict__ out) {
    const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= ns) return;
    const int64_t st = b / M;
    const int m = (int)(b % M);
    if (complex_) {
        out[2 * b] = lp_re[st * M + m];
        out[2 * b + 1] = lp_im[st * M + m];
    } else if (parity) {
        out[b] = logaddexp_(lp_re[st * M + m], lp_re[(st + tiles_s) * M + m]) - 0.69314718055994530942;
    } else {
        out[b] = lp_re[st * M + m];
    }
}

// order of chain slots by decreasing length (sites re-run = N-1-s): merge the kinds by ascending s
__global__ void chain_order_kernel(ChainPlan p, int* __restrict__ order) {
    if (blockIdx.x || threadIdx.x) return;
    int o = 0;
    const int smax = max(p.n_kind0, max(p.n_kind1, p.n_kind2));
    for (int s = 0; s < smax; ++s) {
        if (s < p.n_kind0) order[o++] = s;
        if (s < p.n_kind1) order[o++] = p.n_kind0 + s;
        if (s < p.n_kind2) order[o++] = p.n_kind0 + p.n_kind1 + s;
    }
}

// ---------------------------------------------------------------------------------------------
// workspace
// ---------------------------------------------------------------------------------------------
template <typename T> struct GruWs {
    T* pk;
    uint8_t* sigT;
    T* hstore;
    double *la_sel, *la_oth, *ph_sel, *ph_oth, *delta_re, *delta_im, *lp_re, *lp_im, *diag;
    float* la_self;   // FP32 copy of la_sel for the pipelined tensor-core chains (keeps FP64 out of their site loop)
    int *counter, *order;
};

template <typename T>
static GruWs<T> carve_gru(Ws& ws, const GruLayout& g, const GruLaunch& c, int64_t tiles, bool stash, int nslots, bool cplx,
                          int64_t ns) {
    GruWs<T> w;
    memset(&w, 0, sizeof(w));
    const size_t rows = (size_t)tiles * c.M;
    w.pk = ws.take<T>(g.PK);
    w.sigT = ws.take<uint8_t>(rows * g.N);
    w.lp_re = ws.take<double>(rows);
    w.lp_im = ws.take<double>(cplx ? rows : 0);
    w.counter = ws.take<int>(4);
    if (stash) {
        w.hstore = ws.take<T>(rows * g.N * g.L * g.H);
        w.la_sel = ws.take<double>(rows * g.N);
        w.la_oth = ws.take<double>(rows * g.N);
        w.la_self = ws.take<float>(rows * g.N);
        if (cplx) {
            w.ph_sel = ws.take<double>(rows * g.N);
            w.ph_oth = ws.take<double>(rows * g.N);
        }
    }
    if (nslots > 0) {
        w.delta_re = ws.take<double>(rows * nslots);
        if (cplx) w.delta_im = ws.take<double>(rows * nslots);
        w.order = ws.take<int>(nslots);
        w.diag = ws.take<double>((size_t)ns);
    }
    return w;
}

static inline int grid_for(int64_t n, int block = 256) { return (int)std::min<int64_t>(cdiv(n, block), 148 * 16); }

template <typename K> static int set_smem(K kernel, int bytes) {
    RNNWF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    return 0;
}

// ---------------------------------------------------------------------------------------------
// launch helpers (dispatch the bool template parameters)
// ---------------------------------------------------------------------------------------------
template <typename T, bool STASH, bool CPLX>
static int launch_forward(const GruLayout& g, const GruLaunch& c, const GruWs<T>& w, int tiles, cudaStream_t s) {
    const int block = c.NTc + kHeadThreads;
    if (c.w_smem) {
        auto k = gru_forward_kernel<T, true, STASH, CPLX>;
        if (int e = set_smem(k, c.smem_bytes)) return e;
        prof_count(); k<<<tiles, block, c.smem_bytes, s>>>(g, c, w.pk, w.sigT, w.lp_re, w.lp_im, w.hstore, w.la_sel, w.la_oth, w.ph_sel, w.ph_oth);
    } else {
        auto k = gru_forward_kernel<T, false, STASH, CPLX>;
        if (int e = set_smem(k, c.smem_bytes)) return e;
        prof_count(); k<<<tiles, block, c.smem_bytes, s>>>(g, c, w.pk, w.sigT, w.lp_re, w.lp_im, w.hstore, w.la_sel, w.la_oth, w.ph_sel, w.ph_oth);
    }
    RNNWF_CUDA(cudaGetLastError());
    return 0;
}

template <typename T, bool CPLX>
static int launch_chain(const GruLayout& g, const GruLaunch& c, const ChainPlan& plan, const GruWs<T>& w, cudaStream_t s) {
    const int block = c.NTc + kHeadThreads;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = (int)std::min<int64_t>((int64_t)plan.nslots * plan.tiles, sms);
    RNNWF_CUDA(cudaMemsetAsync(w.counter, 0, sizeof(int), s));
    prof_count(); chain_order_kernel<<<1, 1, 0, s>>>(plan, w.order);
    prof_mark(0, s);
    if (c.w_smem) {
        auto k = gru_chain_kernel<T, true, CPLX>;
        if (int e = set_smem(k, c.smem_bytes)) return e;
        prof_count(); k<<<grid, block, c.smem_bytes, s>>>(g, c, plan, w.pk, w.sigT, w.hstore, w.la_sel, w.la_oth, w.ph_sel, w.ph_oth, w.order,
                                            w.delta_re, w.delta_im, w.counter);
    } else {
        auto k = gru_chain_kernel<T, false, CPLX>;
        if (int e = set_smem(k, c.smem_bytes)) return e;
        prof_count(); k<<<grid, block, c.smem_bytes, s>>>(g, c, plan, w.pk, w.sigT, w.hstore, w.la_sel, w.la_oth, w.ph_sel, w.ph_oth, w.order,
                                            w.delta_re, w.delta_im, w.counter);
    }
    prof_mark(1, s);
    RNNWF_CUDA(cudaGetLastError());
    return 0;
}

template <typename T, bool CPLX>
static int launch_sample(const GruLayout& g, const GruLaunch& c, const T* pk, uint8_t* sampT, int tiles, uint64_t seed,
                         uint64_t off, cudaStream_t s) {
    const int block = c.NTc + kHeadThreads;
    if (c.w_smem) {
        auto k = gru_sample_kernel<T, true, CPLX>;
        if (int e = set_smem(k, c.smem_bytes)) return e;
        prof_count(); k<<<tiles, block, c.smem_bytes, s>>>(g, c, pk, sampT, seed, off);
    } else {
        auto k = gru_sample_kernel<T, false, CPLX>;
        if (int e = set_smem(k, c.smem_bytes)) return e;
        prof_count(); k<<<tiles, block, c.smem_bytes, s>>>(g, c, pk, sampT, seed, off);
    }
    RNNWF_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rnnwf
#ifdef RNNWF_LEGACY
#include "gru_tc.cuh"       // generation 1 (3xTF32, site-blocked weight swaps): A/B builds only
#endif
#include "gru_tc16.cuh"     // shared 3xFP16 helpers (+ generation 2 under RNNWF_LEGACY)
#include "gru_tc16p.cuh"
#include "gru_f64mma.cuh"
#include "gru_tc16b.cuh"
namespace rnnwf {

// Chain-kernel selection for the FP32 probability-head pRNN (A/B measurements through RNNWF_CHAIN):
//   default / "tc16p": tcgen05 kind::f16, 3xFP16 operands, all weights resident, MMA / gate math software-pipelined over
//                      anti-diagonals of the (site, layer) grid (gru_tc16p.cuh)
//   "tc16"          : the same arithmetic with MMA and gate math alternating (gru_tc16.cuh)
//   "tc32"          : tcgen05 kind::tf32, 3xTF32 operands, site-blocked weight swapping (gru_tc.cuh)
//   "ffma"          : CUDA-core tile engine (gru_chain_kernel; the path every other shape / dtype takes)
static bool tc16_any_supported(const GruLayout& g) {
#ifdef RNNWF_LEGACY
    if (tc16::supported(g)) return true;
#endif
    return tc16p::supported_padded(g);
}
// layout the shared E_loc buffers are carved for: the zero-padded 50-unit layout when the tensor-core kernel runs a narrower stack
template <typename T> static GruLayout carve_layout(const GruLayout& g) {
    if (std::is_same<T, float>::value && tc16p::supported_padded(g)) return tc16p::padded_layout(g);
    return g;
}
static size_t tc16_img_bytes(const GruLayout& g) {   // one image buffer serves either 3xFP16 kernel generation
#ifdef RNNWF_LEGACY
    return std::max<size_t>(tc16::supported(g) ? tc16::make_layout(g).img_bytes : 0, tc16p::make_layout(tc16p::padded_layout(g)).img_bytes);
#else
    return tc16p::make_layout(tc16p::padded_layout(g)).img_bytes;
#endif
}
static int chain_mode(const GruLayout& g) {
    const char* e = getenv("RNNWF_CHAIN");
    if (e && strcmp(e, "ffma") == 0) return 0;
#ifdef RNNWF_LEGACY
    if (e && strcmp(e, "tc32") == 0) return tc_supported(g) ? 1 : 0;
    if (e && strcmp(e, "tc16") == 0) return tc16::supported(g) ? 2 : 0;
#endif
    return tc16p::supported_padded(g) ? 3 : 0;
}

// Sampler selection (FP32 stacks the tensor-core kernel covers): default tcgen05 (tc16p::chain_kernel<.., SAMPLE>), RNNWF_SAMPLER=ffma
// keeps the CUDA-core tile engine (A/B measurements, cross-checks)
static bool sampler_tc(const GruLayout& g) {
    const char* e = getenv("RNNWF_SAMPLER");
    if (e && strcmp(e, "ffma") == 0) return false;
    return tc16p::supported_padded(g);
}
static bool logpsi_tc(const GruLayout& g) {      // RNNWF_LOGPSI=ffma keeps the CUDA-core forward kernel (A/B measurements, cross-checks)
    const char* e = getenv("RNNWF_LOGPSI");
    if (e && strcmp(e, "ffma") == 0) return false;
    return tc16p::supported_padded(g);
}
struct LogpsiTcWs { uint8_t* sigT; double *lp_re, *lp_im; unsigned char* img; int* counter; int tiles_s, tiles; };
static LogpsiTcWs carve_logpsi_tc(Ws& ws, const GruLayout& g, int64_t ns, int ndir, bool cplx) {
    LogpsiTcWs w;
    w.tiles_s = (int)cdiv(ns, tc16p::kRows);
    w.tiles = w.tiles_s * ndir;
    const size_t rows = (size_t)w.tiles * tc16p::kRows;
    w.sigT = ws.take<uint8_t>(rows * g.N);
    w.lp_re = ws.take<double>(rows);
    w.lp_im = ws.take<double>(cplx ? rows : 0);
    w.img = ws.take<unsigned char>(tc16_img_bytes(g));
    w.counter = ws.take<int>(4);
    return w;
}
struct SampleTcWs { uint8_t* sampT; unsigned char* img; int* counter; int tiles128; };
static SampleTcWs carve_sample_tc(Ws& ws, const GruLayout& g, int64_t ns) {
    SampleTcWs w;
    w.tiles128 = (int)cdiv(ns, tc16p::kRows);
    w.sampT = ws.take<uint8_t>((size_t)w.tiles128 * tc16p::kRows * g.N);
    w.img = ws.take<unsigned char>(tc16_img_bytes(g));
    w.counter = ws.take<int>(4);
    return w;
}

// ---------------------------------------------------------------------------------------------
// typed implementations behind the C ABI
// ---------------------------------------------------------------------------------------------
template <typename T> size_t gru_workspace_bytes_t(const rnnwf_model& m, int op, int64_t ns, int flags) {
    const GruLayout g = make_gru_layout(m);
    const int ndir = (flags & RNNWF_PARITY_SYM) ? 2 : 1;
    const bool one_cta_per_tile = op == RNNWF_OP_SAMPLE || op == RNNWF_OP_LOGPSI;
    const GruLaunch c = one_cta_per_tile ? choose_gru_launch<T>(g, ns, ndir) : choose_gru_launch<T>(g);
    if (c.RT == 0) return 0;
    const bool cplx = m.head == RNNWF_HEAD_COMPLEX;
    const int64_t tiles = ndir * cdiv(ns, c.M);
    Ws ws(nullptr, 0);
    switch (op) {
        case RNNWF_OP_SAMPLE:
            if (std::is_same<T, float>::value && tc16p::supported_padded(g)) {   // whichever sampler RNNWF_SAMPLER picks at run time fits
                Ws w2(nullptr, 0);
                carve_sample_tc(w2, g, ns);
                carve_gru<T>(ws, g, c, tiles, false, 0, cplx, ns);
                ws.used = std::max(ws.used, w2.used);
                break;
            }
        case RNNWF_OP_LOGPSI:
            carve_gru<T>(ws, g, c, tiles, false, 0, cplx, ns);
            if (std::is_same<T, float>::value && tc16p::supported_padded(g)) {   // whichever kernel RNNWF_LOGPSI picks at run time fits
                Ws w2(nullptr, 0);
                carve_logpsi_tc(w2, g, ns, ndir, cplx);
                ws.used = std::max(ws.used, w2.used);
            }
            break;
        case RNNWF_OP_TFIM_ELOC:
            carve_gru<T>(ws, carve_layout<T>(g), c, tiles, true, g.N, cplx, ns);
#ifdef RNNWF_LEGACY
            if (std::is_same<T, float>::value && tc_supported(g)) carve_tc(ws, g, make_tc_layout(g), 160);
#endif
            if (std::is_same<T, float>::value && tc16_any_supported(g)) ws.take<unsigned char>(tc16_img_bytes(g));
            if (std::is_same<T, double>::value && f64mma::supported(g)) ws.take<double>(f64mma::make_layout(g).wb_doubles + f64mma::make_layout(g).tab_doubles);
            break;
        case RNNWF_OP_J1J2_ELOC:
            carve_gru<T>(ws, carve_layout<T>(g), c, tiles, true, 2 * g.N, cplx, ns);
            ws.take<float>((size_t)ns * (2 * g.N + 1));
            if (std::is_same<T, float>::value && tc16_any_supported(g)) ws.take<unsigned char>(tc16_img_bytes(g));
            break;
        case RNNWF_OP_VMC_GRAD: return gru_grad_workspace_bytes<T>(m, ns, flags);
        default: return 0;
    }
    return ws.used + 256;
}
template size_t gru_workspace_bytes_t<float>(const rnnwf_model&, int, int64_t, int);
template size_t gru_workspace_bytes_t<double>(const rnnwf_model&, int, int64_t, int);

template <typename T>
int gru_sample_t(const rnnwf_model& m, const void* params, int64_t ns, uint64_t seed, uint64_t off, uint8_t* out, void* wsp,
                 size_t wsb, cudaStream_t s) {
    const GruLayout g = make_gru_layout(m);
    if constexpr (std::is_same<T, float>::value) {
        if (sampler_tc(g)) {
            Ws ws(wsp, wsb);
            SampleTcWs w = carve_sample_tc(ws, g, ns);
            RNNWF_CHECK(ws.ok(), -4, "workspace too small: need %zu have %zu", ws.used, wsb);
            if (int e = tc16p::launch_sample(g, w.tiles128, (const float*)params, w.img, w.sampT, w.counter, seed, off, s)) return e;
            prof_count(); samp_untranspose_kernel<<<grid_for(ns * g.N), 256, 0, s>>>(w.sampT, out, ns, g.N, tc16p::kRows);
            RNNWF_CUDA(cudaGetLastError());
            return 0;
        }
    }
    const GruLaunch c = choose_gru_launch<T>(g, ns, 1);
    RNNWF_CHECK(c.RT > 0, -3, "no launch configuration fits (units=%d layers=%d)", m.units, m.num_layers);
    const bool cplx = m.head == RNNWF_HEAD_COMPLEX;
    const int tiles = (int)cdiv(ns, c.M);
    Ws ws(wsp, wsb);
    GruWs<T> w = carve_gru<T>(ws, g, c, tiles, false, 0, cplx, ns);
    RNNWF_CHECK(ws.ok(), -4, "workspace too small: need %zu have %zu", ws.used, wsb);
    prof_count(); pack_gru_kernel<T><<<grid_for(g.PK), 256, 0, s>>>(g, (const T*)params, w.pk);
    int e = cplx ? launch_sample<T, true>(g, c, w.pk, w.sigT, tiles, seed, off, s)
                 : launch_sample<T, false>(g, c, w.pk, w.sigT, tiles, seed, off, s);
    if (e) return e;
    prof_count(); samp_untranspose_kernel<<<grid_for(ns * g.N), 256, 0, s>>>(w.sigT, out, ns, g.N, c.M);
    RNNWF_CUDA(cudaGetLastError());
    return 0;
}
template int gru_sample_t<float>(const rnnwf_model&, const void*, int64_t, uint64_t, uint64_t, uint8_t*, void*, size_t, cudaStream_t);
template int gru_sample_t<double>(const rnnwf_model&, const void*, int64_t, uint64_t, uint64_t, uint8_t*, void*, size_t, cudaStream_t);

template <typename T>
int gru_logpsi_t(const rnnwf_model& m, const void* params, const uint8_t* samples, int64_t ns, int flags, double* out, void* wsp,
                 size_t wsb, cudaStream_t s) {
    const GruLayout g = make_gru_layout(m);
    const bool cplx = m.head == RNNWF_HEAD_COMPLEX;
    const int parity = (flags & RNNWF_PARITY_SYM) ? 1 : 0;
    RNNWF_CHECK(!(cplx && parity), -2, "parity symmetry is only defined for the probability head");
    if constexpr (std::is_same<T, float>::value) {
        if (logpsi_tc(g)) {      // the tensor-core base pass without its stash
            Ws ws(wsp, wsb);
            LogpsiTcWs w = carve_logpsi_tc(ws, g, ns, parity ? 2 : 1, cplx);
            RNNWF_CHECK(ws.ok(), -4, "workspace too small: need %zu have %zu", ws.used, wsb);
            prof_count(); sig_transpose_kernel<<<grid_for((int64_t)w.tiles * g.N * tc16p::kRows), 256, 0, s>>>(samples, w.sigT, ns, g.N, tc16p::kRows, w.tiles_s, parity ? 2 : 1);
            if (int e = tc16p::launch_logpsi(g, w.tiles, (const float*)params, w.img, w.sigT, w.lp_re, w.lp_im, w.counter, s)) return e;
            prof_count(); gather_logpsi_kernel<<<grid_for(ns), 256, 0, s>>>(w.lp_re, w.lp_im, ns, tc16p::kRows, w.tiles_s, parity, cplx, out);
            RNNWF_CUDA(cudaGetLastError());
            return 0;
        }
    }
    const GruLaunch c = choose_gru_launch<T>(g, ns, parity ? 2 : 1);
    RNNWF_CHECK(c.RT > 0, -3, "no launch configuration fits (units=%d layers=%d)", m.units, m.num_layers);
    const int tiles_s = (int)cdiv(ns, c.M), ndir = parity ? 2 : 1, tiles = tiles_s * ndir;
    Ws ws(wsp, wsb);
    GruWs<T> w = carve_gru<T>(ws, g, c, tiles, false, 0, cplx, ns);
    RNNWF_CHECK(ws.ok(), -4, "workspace too small: need %zu have %zu", ws.used, wsb);
    prof_count(); pack_gru_kernel<T><<<grid_for(g.PK), 256, 0, s>>>(g, (const T*)params, w.pk);
    prof_count(); sig_transpose_kernel<<<grid_for((int64_t)tiles * g.N * c.M), 256, 0, s>>>(samples, w.sigT, ns, g.N, c.M, tiles_s, ndir);
    int e = cplx ? launch_forward<T, false, true>(g, c, w, tiles, s) : launch_forward<T, false, false>(g, c, w, tiles, s);
    if (e) return e;
    prof_count(); gather_logpsi_kernel<<<grid_for(ns), 256, 0, s>>>(w.lp_re, w.lp_im, ns, c.M, tiles_s, parity, cplx, out);
    RNNWF_CUDA(cudaGetLastError());
    return 0;
}
template int gru_logpsi_t<float>(const rnnwf_model&, const void*, const uint8_t*, int64_t, int, double*, void*, size_t, cudaStream_t);
template int gru_logpsi_t<double>(const rnnwf_model&, const void*, const uint8_t*, int64_t, int, double*, void*, size_t, cudaStream_t);

template <typename T>
int gru_tfim_eloc_t(const rnnwf_model& m, const void* params, const uint8_t* samples, int64_t ns, const double* jz, double bx,
                    int flags, double* eloc, double* logp, double* ratios, void* wsp, size_t wsb, cudaStream_t s) {
    const GruLayout g = make_gru_layout(m);
    const GruLaunch c = choose_gru_launch<T>(g);
    RNNWF_CHECK(c.RT > 0, -3, "no launch configuration fits (units=%d layers=%d)", m.units, m.num_layers);
    RNNWF_CHECK(m.head == RNNWF_HEAD_PROB, -2, "TFIM local energies need the probability head");
    const int parity = (flags & RNNWF_PARITY_SYM) ? 1 : 0;
    const int tiles_s = (int)cdiv(ns, c.M), ndir = parity ? 2 : 1, tiles = tiles_s * ndir;
    Ws ws(wsp, wsb);
    GruWs<T> w = carve_gru<T>(ws, carve_layout<T>(g), c, tiles, true, g.N, false, ns);
#ifdef RNNWF_LEGACY
    TcWs tw{};
    const bool tc = std::is_same<T, float>::value && tc_supported(g);
    if (tc) tw = carve_tc(ws, g, make_tc_layout(g), 160);
#endif
    unsigned char* img16 = nullptr;
    if (std::is_same<T, float>::value && tc16_any_supported(g)) img16 = ws.take<unsigned char>(tc16_img_bytes(g));
    const int mode = std::is_same<T, float>::value ? chain_mode(g) : 0;
    double* wb64 = nullptr;       // float64 one-layer stacks: B fragments + constant table of the DMMA chain kernel (gru_f64mma.cuh)
    if (std::is_same<T, double>::value && f64mma::supported(g)) wb64 = ws.take<double>(f64mma::make_layout(g).wb_doubles + f64mma::make_layout(g).tab_doubles);
    RNNWF_CHECK(ws.ok(), -4, "workspace too small: need %zu have %zu", ws.used, wsb);
    prof_count(); pack_gru_kernel<T><<<grid_for(g.PK), 256, 0, s>>>(g, (const T*)params, w.pk);
    prof_count(); sig_transpose_kernel<<<grid_for((int64_t)tiles * g.N * c.M), 256, 0, s>>>(samples, w.sigT, ns, g.N, c.M, tiles_s, ndir);
    prof_count(); tfim_diag_kernel<<<(int)cdiv(ns, 128), 128, 0, s>>>(samples, ns, g.N, m.nx, m.ny, jz, w.diag);
    if (mode == 3) {
        if constexpr (std::is_same<T, float>::value) {
            if (int e = tc16p::launch_eloc(g, c.M, tiles, (const float*)params, img16, w.sigT, w.hstore, w.la_sel, w.la_oth, w.la_self, w.lp_re, w.delta_re,
                                           w.counter, bx != 0.0, s))
                return e;
        }
#ifdef RNNWF_LEGACY
    } else if (mode == 2) {
        if constexpr (std::is_same<T, float>::value) {
            if (int e = tc16::launch_eloc(g, c.M, tiles, (const float*)params, img16, w.sigT, w.hstore, w.la_sel, w.la_oth, w.lp_re, w.delta_re,
                                          w.counter, bx != 0.0, s))
                return e;
        }
    } else if (mode == 1) {
        if constexpr (std::is_same<T, float>::value) {
            if (int e = launch_eloc_tc(g, c, w, tw, tiles, (const float*)params, bx != 0.0, s)) return e;
        }
#endif
    } else if (bx != 0.0) {   // reference skips the off-diagonal work when Bx == 0 (1DTFIM/TrainingRNN_1DTFIM.py:42)
        const char* env = getenv("RNNWF_CHAIN");
        bool dmma_done = false;
        if constexpr (std::is_same<T, double>::value) {
            if (wb64 && !(env && strcmp(env, "ffma") == 0)) {   // DMMA base pass + chain kernel; RNNWF_CHAIN=ffma keeps the thread-tile engine (A/B)
                if (int e = f64mma::launch(g, c.M, (int64_t)tiles * c.M, (const double*)params, wb64, wb64 + f64mma::make_layout(g).wb_doubles, w.sigT,
                                           w.hstore, w.la_sel, w.la_oth, w.lp_re, w.delta_re, w.counter, true, true, s))
                    return e;
                dmma_done = true;
            }
        }
        if (!dmma_done) {
            if (int e = launch_forward<T, true, false>(g, c, w, tiles, s)) return e;
            ChainPlan plan{g.N, g.N, 0, 0, tiles, nullptr, nullptr};
            if (int e = launch_chain<T, false>(g, c, plan, w, s)) return e;
        }
    } else {
        if (int e = launch_forward<T, false, false>(g, c, w, tiles, s)) return e;
    }
    prof_count(); tfim_finalize_kernel<<<(int)cdiv(ns, 128), 128, 0, s>>>(w.diag, w.delta_re, w.lp_re, ns, g.N, c.M, tiles_s, bx, parity, eloc, logp, ratios);
    RNNWF_CUDA(cudaGetLastError());
    return 0;
}
template int gru_tfim_eloc_t<float>(const rnnwf_model&, const void*, const uint8_t*, int64_t, const double*, double, int, double*,
                                    double*, double*, void*, size_t, cudaStream_t);
template int gru_tfim_eloc_t<double>(const rnnwf_model&, const void*, const uint8_t*, int64_t, const double*, double, int, double*,
                                     double*, double*, void*, size_t, cudaStream_t);

int tfim_chain_mode_impl(const rnnwf_model& m) {
    if (m.cell != RNNWF_CELL_GRU || m.dtype != RNNWF_F32) return 0;   // complex head: the J1-J2 exchange chains make the same choice (j1j2.cuh)
    return chain_mode(make_gru_layout(m));
}

int tfim_diag_impl(const rnnwf_model& m, const uint8_t* samples, int64_t ns, const double* jz, double* diag, cudaStream_t s) {
    prof_count(); tfim_diag_kernel<<<(int)cdiv(ns, 128), 128, 0, s>>>(samples, ns, m.n_sites, m.nx, m.ny, jz, diag);
    RNNWF_CUDA(cudaGetLastError());
    return 0;
}

int tfim_finalize_impl(const double* diag, const double* delta, const double* lp, int64_t ns, int N, int M, int tiles_s, double bx,
                       int parity, double* eloc, double* logp, double* ratios, cudaStream_t s) {
    prof_count(); tfim_finalize_kernel<<<(int)cdiv(ns, 128), 128, 0, s>>>(diag, delta, lp, ns, N, M, tiles_s, bx, parity, eloc, logp, ratios);
    RNNWF_CUDA(cudaGetLastError());
    return 0;
}

int tfim_enumerate_impl(const uint8_t* samples, int64_t ns, int N, int32_t* queue, cudaStream_t s) {
    prof_count(); tfim_enumerate_kernel<<<grid_for((int64_t)(N + 1) * ns * N), 256, 0, s>>>(samples, ns, N, queue);
    RNNWF_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rnnwf

#include "grad.cuh"
#include "j1j2.cuh"
