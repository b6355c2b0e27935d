// j1j2.cuh — J1-J2 chain: connected-configuration enumeration and fused local energies for the complex cRNN.
//
// Replaces J1J2MatrixElements / J1J2Slices (J1J2/TrainingRNN_J1J2.py:12-93, :95-127) and the local-energy
// combine (:255-279).  The fused path never materialises the exchanged configurations: every NN / NNN
// exchange is a prefix-reuse chain of gru_chain_kernel (kinds 1, 2).
// Included at the end of gru.cu (same translation unit as the launchers).
#pragma once
#include "gru_kernels.cuh"
#include "host_util.cuh"

namespace rnnwf {

// diagonal element in the reference's accumulation order (:30-57), f64, then rounded to f32 (:59 / buffer :239)
__device__ __forceinline__ float j1j2_diag(const uint8_t* __restrict__ s, int N, const double* __restrict__ j1,
                                           const double* __restrict__ j2, const double* __restrict__ bz, int periodic) {
    const int lim1 = periodic ? N : N - 1, lim2 = periodic ? N : N - 2;
    double diag = 0.0;
    for (int i = 0; i < N; ++i) diag += ((double)s[i] - 0.5) * bz[i];
    for (int i = 0; i < lim1; ++i) diag += (s[i] != s[(i + 1) % N] ? -0.25 : 0.25) * j1[i];
    for (int i = 0; i < lim2; ++i)
        if (j2[i] != 0.0) diag += (s[i] != s[(i + 2) % N] ? -0.25 : 0.25) * j2[i];
    return (float)diag;
}

// One block per sample.  Row order: diagonal, NN exchanges (ascending site), NNN exchanges (ascending site).
// Fixed-slot output: row r of sample b at b*(2N+1)+r; rows >= counts[b] are left untouched.
__global__ void j1j2_enumerate_kernel(const uint8_t* __restrict__ samples, int N, const double* __restrict__ j1,
                                      const double* __restrict__ j2, const double* __restrict__ bz, int periodic, int marshall,
                                      int32_t* __restrict__ sigmas, float* __restrict__ elements, int32_t* __restrict__ counts) {
    extern __shared__ int rowmap[];   // rowmap[r] = dist * N + site for exchange rows r >= 1
    __shared__ int s_num;
    const int64_t b = blockIdx.x;
    const uint8_t* s = samples + b * N;
    const int rows = 2 * N + 1;
    if (threadIdx.x == 0) {
        const int lim1 = periodic ? N : N - 1, lim2 = periodic ? N : N - 2;
        float* el = elements + b * rows;
        el[0] = j1j2_diag(s, N, j1, j2, bz, periodic);
        int num = 1;
        for (int i = 0; i < lim1; ++i)
            if (j1[i] != 0.0 && s[i] != s[(i + 1) % N]) {
                el[num] = (float)(marshall ? -j1[i] / 2 : j1[i] / 2);
                rowmap[num++] = 1 * N + i;
            }
        for (int i = 0; i < lim2; ++i)
            if (j2[i] != 0.0 && s[i] != s[(i + 2) % N]) {
                el[num] = (float)(j2[i] / 2);
                rowmap[num++] = 2 * N + i;
            }
        counts[b] = num;
        s_num = num;
    }
    __syncthreads();
    if (sigmas == nullptr) return;
    const int num = s_num;
    for (int idx = threadIdx.x; idx < num * N; idx += blockDim.x) {
        const int r = idx / N, site = idx % N;
        int v = s[site];
        if (r > 0) {
            const int dist = rowmap[r] / N, a = rowmap[r] % N, t = (a + dist) % N;
            if (site == a) v = s[t];
            else if (site == t) v = s[a];
        }
        sigmas[(b * rows + r) * N + site] = v;
    }
}

int j1j2_enumerate_impl(const uint8_t* samples, int64_t ns, int N, const double* j1, const double* j2, const double* bz,
                        int periodic, int marshall, int32_t* sigmas, float* elements, int32_t* counts, cudaStream_t s) {
    prof_count(); j1j2_enumerate_kernel<<<(unsigned)ns, 128, (2 * N + 1) * sizeof(int), s>>>(samples, N, j1, j2, bz, periodic, marshall, sigmas,
                                                                               elements, counts);
    RNNWF_CUDA(cudaGetLastError());
    return 0;
}

// E_b = diag + sum_{antiparallel NN bonds} (+-J1/2) exp(delta) + sum_{antiparallel NNN bonds} (J2/2) exp(delta)
// with delta = log psi(sigma') - log psi(sigma) complex (J1J2/TrainingRNN_J1J2.py:277-279); elements rounded to
// f32 as the reference stores them (:239).  Open chain.
__global__ void j1j2_finalize_kernel(const uint8_t* __restrict__ samples, int64_t ns, int N, int M, const double* __restrict__ j1,
                                     const double* __restrict__ j2, const double* __restrict__ bz, int marshall,
                                     const double* __restrict__ dre, const double* __restrict__ dim, const double* __restrict__ lre,
                                     const double* __restrict__ lim, int nslots, double* __restrict__ eloc,
                                     double* __restrict__ logpsi) {
    const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= ns) return;
    const uint8_t* s = samples + b * N;
    const int64_t st = b / M;
    const int m = (int)(b % M);
    double er = (double)j1j2_diag(s, N, j1, j2, bz, 0), ei = 0.0;
    for (int i = 0; i < N - 1; ++i)
        if (j1[i] != 0.0 && s[i] != s[i + 1]) {
            const double el = (double)(float)(marshall ? -j1[i] / 2 : j1[i] / 2);
            const size_t o = ((size_t)st * nslots + i) * M + m;
            const double a = el * exp(dre[o]);
            double sn, cs;
            sincos(dim[o], &sn, &cs);
            er += a * cs;
            ei += a * sn;
        }
    for (int i = 0; i < N - 2; ++i)
        if (j2[i] != 0.0 && s[i] != s[i + 2]) {
            const double el = (double)(float)(j2[i] / 2);
            const size_t o = ((size_t)st * nslots + (N - 1) + i) * M + m;
            const double a = el * exp(dre[o]);
            double sn, cs;
            sincos(dim[o], &sn, &cs);
            er += a * cs;
            ei += a * sn;
        }
    eloc[2 * b] = er;
    eloc[2 * b + 1] = ei;
    if (logpsi) {
        logpsi[2 * b] = lre[st * M + m];
        logpsi[2 * b + 1] = lim[st * M + m];
    }
}

template <typename T>
int gru_j1j2_eloc_t(const rnnwf_model& m, const void* params, const uint8_t* samples, int64_t ns, const double* j1,
                    const double* j2, const double* bz, int marshall, double* eloc, double* logpsi, void* wsp, size_t wsb,
                    cudaStream_t s) {
    const GruLayout g = make_gru_layout(m);
    const GruLaunch c = choose_gru_launch<T>(g);
    RNNWF_CHECK(c.RT > 0, -3, "no launch configuration fits (units=%d layers=%d)", m.units, m.num_layers);
    RNNWF_CHECK(g.N >= 4, -1, "J1-J2 needs at least 4 sites");
    const int tiles = (int)cdiv(ns, c.M);
    const int nslots = 2 * g.N;
    Ws ws(wsp, wsb);
    GruWs<T> w = carve_gru<T>(ws, carve_layout<T>(g), c, tiles, true, nslots, true, ns);
    ws.take<float>((size_t)ns * (2 * g.N + 1));                         // (kept for layout compatibility with rnnwf_workspace_bytes)
    unsigned char* img16 = nullptr;
    if (std::is_same<T, float>::value && tc16_any_supported(g)) img16 = ws.take<unsigned char>(tc16_img_bytes(g));
    RNNWF_CHECK(ws.ok(), -4, "workspace too small: need %zu have %zu", ws.used, wsb);
    prof_count(); pack_gru_kernel<T><<<grid_for(g.PK), 256, 0, s>>>(g, (const T*)params, w.pk);
    prof_count(); sig_transpose_kernel<<<grid_for((int64_t)tiles * g.N * c.M), 256, 0, s>>>(samples, w.sigT, ns, g.N, c.M, tiles, 1);
    ChainPlan plan{g.N - 1 + g.N - 2, 0, g.N - 1, g.N - 2, tiles, j1, j2};
    const char* env = getenv("RNNWF_CHAIN");
    if (img16 && !(env && strcmp(env, "ffma") == 0)) {   // tcgen05 3xFP16 kernels (gru_tc16p.cuh; RNNWF_CHAIN=tc16: gru_tc16.cuh)
        if constexpr (std::is_same<T, float>::value) {
            prof_count(); chain_order_kernel<<<1, 1, 0, s>>>(plan, w.order);
#ifdef RNNWF_LEGACY
            const bool gen2 = env && strcmp(env, "tc16") == 0;
            auto launch = gen2 ? tc16::launch_j1j2 : tc16p::launch_j1j2;
#else
            auto launch = tc16p::launch_j1j2;
#endif
            if (int e = launch(g, c.M, tiles, (const float*)params, img16, w.sigT, w.hstore, w.la_sel, w.la_oth, w.ph_sel, w.ph_oth,
                               w.lp_re, w.lp_im, w.delta_re, w.delta_im, w.order, j1, j2, w.counter, s))
                return e;
        }
    } else {
        if (int e = launch_forward<T, true, true>(g, c, w, tiles, s)) return e;
        if (int e = launch_chain<T, true>(g, c, plan, w, s)) return e;
    }
    prof_count(); j1j2_finalize_kernel<<<(int)cdiv(ns, 128), 128, 0, s>>>(samples, ns, g.N, c.M, j1, j2, bz, marshall, w.delta_re, w.delta_im,
                                                                           w.lp_re, w.lp_im, plan.nslots, eloc, logpsi);
    RNNWF_CUDA(cudaGetLastError());
    return 0;
}
template int gru_j1j2_eloc_t<float>(const rnnwf_model&, const void*, const uint8_t*, int64_t, const double*, const double*,
                                    const double*, int, double*, double*, void*, size_t, cudaStream_t);
template int gru_j1j2_eloc_t<double>(const rnnwf_model&, const void*, const uint8_t*, int64_t, const double*, const double*,
                                     const double*, int, double*, double*, void*, size_t, cudaStream_t);

}  // namespace rnnwf
