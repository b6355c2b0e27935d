// capi.cu — extern "C" surface of librnnwf_b200.so (see include/rnnwf.h for the contract).
#include <stdarg.h>
#include "api_internal.h"

namespace rnnwf {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
// measurement state of rnnwf_profile_begin / end: host only and per calling thread, like the error string -- the library keeps no
// state that two host threads (or two wave functions) share
static constexpr int kProfPairs = 256;
static thread_local struct {
    bool on = false;
    long launches = 0;
    int pairs = 0;
    bool open = false;
    cudaEvent_t ev[kProfPairs][2] = {};
} g_prof;
void prof_count(int n) { if (g_prof.on) g_prof.launches += n; }
void prof_mark(int which, cudaStream_t s) {
    if (!g_prof.on) return;
    if (which == 0) {
        if (g_prof.pairs >= kProfPairs) { g_prof.open = false; return; }
        for (int k = 0; k < 2; ++k)
            if (!g_prof.ev[g_prof.pairs][k]) cudaEventCreate(&g_prof.ev[g_prof.pairs][k]);
        cudaEventRecord(g_prof.ev[g_prof.pairs][0], s);
        g_prof.open = true;
    } else if (g_prof.open) {
        cudaEventRecord(g_prof.ev[g_prof.pairs][1], s);
        g_prof.pairs++;
        g_prof.open = false;
    }
}
static int check_model(const rnnwf_model* m) {
    RNNWF_CHECK(m != nullptr, -1, "model is NULL");
    RNNWF_CHECK(m->cell == RNNWF_CELL_GRU || m->cell == RNNWF_CELL_MDRNN, -1, "unknown cell %d", m->cell);
    RNNWF_CHECK(m->head == RNNWF_HEAD_PROB || m->head == RNNWF_HEAD_COMPLEX, -1, "unknown head %d", m->head);
    RNNWF_CHECK(m->dtype == RNNWF_F32 || m->dtype == RNNWF_F64, -1, "unknown dtype %d", m->dtype);
    RNNWF_CHECK(m->units >= 1 && m->units <= 1024, -1, "units=%d out of range", m->units);
    RNNWF_CHECK(m->n_sites >= 1, -1, "n_sites=%d", m->n_sites);
    if (m->cell == RNNWF_CELL_GRU) RNNWF_CHECK(m->num_layers >= 1 && m->num_layers <= kMaxLayers, -1, "num_layers=%d (1..%d)", m->num_layers, kMaxLayers);
    if (m->cell == RNNWF_CELL_MDRNN) {
        RNNWF_CHECK(m->nx >= 1 && m->ny >= 1 && m->nx * m->ny == m->n_sites, -1, "MDRNN needs nx*ny == n_sites");
        RNNWF_CHECK(m->head == RNNWF_HEAD_PROB, -1, "MDRNN supports the probability head only");
    }
    if (m->nx > 0 || m->ny > 0) RNNWF_CHECK(m->nx * m->ny == m->n_sites, -1, "nx*ny != n_sites");
    if (m->head == RNNWF_HEAD_COMPLEX) RNNWF_CHECK(m->n_sites % 2 == 0, -1, "zero-magnetisation sector needs even N");
    return 0;
}
}  // namespace rnnwf

using namespace rnnwf;

#define DISPATCH(m, fn, ...) ((m)->dtype == RNNWF_F32 ? fn<float>(__VA_ARGS__) : fn<double>(__VA_ARGS__))

#define RNNWF_API __attribute__((visibility("default")))
extern "C" {

RNNWF_API const char* rnnwf_last_error(void) { return g_err; }
RNNWF_API int rnnwf_abi_version(void) { return RNNWF_ABI_VERSION; }

RNNWF_API int rnnwf_profile_begin(void) {
    g_prof.on = true;
    g_prof.launches = 0;
    g_prof.pairs = 0;
    g_prof.open = false;
    return 0;
}

RNNWF_API int rnnwf_profile_end(int64_t* launches_out, int64_t* dominant_launches_out, double* dominant_ms_out) {
    g_prof.on = false;
    double ms = 0.0;
    for (int i = 0; i < g_prof.pairs; ++i) {
        RNNWF_CUDA(cudaEventSynchronize(g_prof.ev[i][1]));
        float t = 0.f;
        RNNWF_CUDA(cudaEventElapsedTime(&t, g_prof.ev[i][0], g_prof.ev[i][1]));
        ms += t;
    }
    if (launches_out) *launches_out = g_prof.launches;
    if (dominant_launches_out) *dominant_launches_out = g_prof.pairs;
    if (dominant_ms_out) *dominant_ms_out = ms;
    return 0;
}

RNNWF_API int64_t rnnwf_param_count(const rnnwf_model* m) {
    if (check_model(m)) return -1;
    const int64_t H = m->units;
    if (m->cell == RNNWF_CELL_MDRNN) return 2 * H * H + 2 * 2 * H + H + 2 * H + 2;
    int64_t p = 0, d = 2;
    for (int l = 0; l < m->num_layers; ++l) {
        p += (d + H) * 2 * H + 2 * H + d * H + H * H + 2 * H;
        d = H;
    }
    return p + (m->head == RNNWF_HEAD_COMPLEX ? 2 : 1) * (2 * H + 2);
}

RNNWF_API size_t rnnwf_workspace_bytes(const rnnwf_model* m, int op, int64_t ns, int flags) {
    if (check_model(m) || ns <= 0) return 0;
    if (m->cell == RNNWF_CELL_MDRNN) return DISPATCH(m, mdrnn_workspace_bytes_t, *m, op, ns, flags);
    return DISPATCH(m, gru_workspace_bytes_t, *m, op, ns, flags);
}

RNNWF_API int rnnwf_sample(const rnnwf_model* m, const void* params, int64_t ns, uint64_t seed, uint64_t sample_offset,
                 uint8_t* samples_out, void* ws, size_t ws_bytes, void* stream) {
    if (int e = check_model(m)) return e;
    RNNWF_CHECK(params && samples_out && ns > 0, -1, "bad arguments to rnnwf_sample");
    cudaStream_t s = (cudaStream_t)stream;
    if (m->cell == RNNWF_CELL_MDRNN) return DISPATCH(m, mdrnn_sample_t, *m, params, ns, seed, sample_offset, samples_out, ws, ws_bytes, s);
    return DISPATCH(m, gru_sample_t, *m, params, ns, seed, sample_offset, samples_out, ws, ws_bytes, s);
}

RNNWF_API int rnnwf_logpsi(const rnnwf_model* m, const void* params, const uint8_t* samples, int64_t ns, int flags, double* out,
                 void* ws, size_t ws_bytes, void* stream) {
    if (int e = check_model(m)) return e;
    RNNWF_CHECK(params && samples && out && ns > 0, -1, "bad arguments to rnnwf_logpsi");
    cudaStream_t s = (cudaStream_t)stream;
    if (m->cell == RNNWF_CELL_MDRNN) return DISPATCH(m, mdrnn_logpsi_t, *m, params, samples, ns, out, ws, ws_bytes, s);
    return DISPATCH(m, gru_logpsi_t, *m, params, samples, ns, flags, out, ws, ws_bytes, s);
}

RNNWF_API int rnnwf_tfim_eloc(const rnnwf_model* m, const void* params, const uint8_t* samples, int64_t ns, const double* jz,
                    double bx, int flags, double* eloc_out, double* logp_out, void* ws, size_t ws_bytes, void* stream) {
    if (int e = check_model(m)) return e;
    RNNWF_CHECK(params && samples && jz && eloc_out && ns > 0, -1, "bad arguments to rnnwf_tfim_eloc");
    cudaStream_t s = (cudaStream_t)stream;
    if (m->cell == RNNWF_CELL_MDRNN)
        return DISPATCH(m, mdrnn_tfim_eloc_t, *m, params, samples, ns, jz, bx, eloc_out, logp_out, nullptr, ws, ws_bytes, s);
    return DISPATCH(m, gru_tfim_eloc_t, *m, params, samples, ns, jz, bx, flags, eloc_out, logp_out, nullptr, ws, ws_bytes, s);
}

RNNWF_API int rnnwf_tfim_flip_ratios(const rnnwf_model* m, const void* params, const uint8_t* samples, int64_t ns, const double* jz,
                           double bx, int flags, double* eloc_out, double* logp_out, double* ratios_out, void* ws, size_t ws_bytes,
                           void* stream) {
    if (int e = check_model(m)) return e;
    RNNWF_CHECK(params && samples && jz && eloc_out && ratios_out && ns > 0, -1, "bad arguments to rnnwf_tfim_flip_ratios");
    RNNWF_CHECK(bx != 0.0, -1, "rnnwf_tfim_flip_ratios: the single-flip chains are skipped when bx == 0 (as in the reference); pass bx != 0");
    cudaStream_t s = (cudaStream_t)stream;
    if (m->cell == RNNWF_CELL_MDRNN)
        return DISPATCH(m, mdrnn_tfim_eloc_t, *m, params, samples, ns, jz, bx, eloc_out, logp_out, ratios_out, ws, ws_bytes, s);
    return DISPATCH(m, gru_tfim_eloc_t, *m, params, samples, ns, jz, bx, flags, eloc_out, logp_out, ratios_out, ws, ws_bytes, s);
}

RNNWF_API int rnnwf_tfim_chain_mode(const rnnwf_model* m) {
    if (check_model(m)) return -1;
    return tfim_chain_mode_impl(*m);
}

RNNWF_API int rnnwf_tfim_diag(const rnnwf_model* m, const uint8_t* samples, int64_t ns, const double* jz, double* diag_out, void* stream) {
    if (int e = check_model(m)) return e;
    RNNWF_CHECK(samples && jz && diag_out && ns > 0, -1, "bad arguments to rnnwf_tfim_diag");
    return tfim_diag_impl(*m, samples, ns, jz, diag_out, (cudaStream_t)stream);
}

RNNWF_API int rnnwf_tfim_enumerate(const uint8_t* samples, int64_t ns, int32_t n_sites, int32_t* queue_out, void* stream) {
    RNNWF_CHECK(samples && queue_out && ns > 0 && n_sites > 0, -1, "bad arguments to rnnwf_tfim_enumerate");
    return tfim_enumerate_impl(samples, ns, n_sites, queue_out, (cudaStream_t)stream);
}

RNNWF_API int rnnwf_j1j2_enumerate(const uint8_t* samples, int64_t ns, int32_t n_sites, const double* j1, const double* j2,
                         const double* bz, int periodic, int marshall_sign, int32_t* sigmas_out, float* elements_out,
                         int32_t* counts_out, void* stream) {
    RNNWF_CHECK(samples && j1 && j2 && bz && elements_out && counts_out && ns > 0 && n_sites > 2, -1, "bad arguments to rnnwf_j1j2_enumerate");
    return j1j2_enumerate_impl(samples, ns, n_sites, j1, j2, bz, periodic, marshall_sign, sigmas_out, elements_out, counts_out,
                               (cudaStream_t)stream);
}

RNNWF_API int rnnwf_j1j2_eloc(const rnnwf_model* m, const void* params, const uint8_t* samples, int64_t ns, const double* j1,
                    const double* j2, const double* bz, int marshall_sign, double* eloc_out, double* logpsi_out, void* ws,
                    size_t ws_bytes, void* stream) {
    if (int e = check_model(m)) return e;
    RNNWF_CHECK(m->cell == RNNWF_CELL_GRU && m->head == RNNWF_HEAD_COMPLEX, -2, "J1-J2 local energies need the complex GRU wave function");
    RNNWF_CHECK(params && samples && j1 && j2 && bz && eloc_out && ns > 0, -1, "bad arguments to rnnwf_j1j2_eloc");
    return DISPATCH(m, gru_j1j2_eloc_t, *m, params, samples, ns, j1, j2, bz, marshall_sign, eloc_out, logpsi_out, ws, ws_bytes,
                    (cudaStream_t)stream);
}

RNNWF_API int rnnwf_vmc_grad(const rnnwf_model* m, const void* params, const uint8_t* samples, int64_t ns, const double* weights,
                   int flags, double* grad_out, void* ws, size_t ws_bytes, void* stream) {
    if (int e = check_model(m)) return e;
    RNNWF_CHECK(params && samples && weights && grad_out && ns > 0, -1, "bad arguments to rnnwf_vmc_grad");
    cudaStream_t s = (cudaStream_t)stream;
    if (m->cell == RNNWF_CELL_MDRNN) return DISPATCH(m, mdrnn_vmc_grad_t, *m, params, samples, ns, weights, grad_out, ws, ws_bytes, s);
    return DISPATCH(m, gru_vmc_grad_t, *m, params, samples, ns, weights, flags, grad_out, ws, ws_bytes, s);
}

RNNWF_API int rnnwf_adam_step(int dtype, int64_t n, void* theta, void* mom, void* vel, const double* grad, double grad_scale, double lr,
                    double beta1, double beta2, double eps, int64_t t, void* stream) {
    RNNWF_CHECK(theta && mom && vel && grad && n > 0 && t >= 1, -1, "bad arguments to rnnwf_adam_step");
    return adam_step_impl(dtype, n, theta, mom, vel, grad, grad_scale, lr, beta1, beta2, eps, t, (cudaStream_t)stream);
}

RNNWF_API int rnnwf_energy_moments(const double* eloc, int64_t ns, int stride, double* stats_out, void* stream) {
    RNNWF_CHECK(eloc && stats_out && ns > 0 && stride >= 1, -1, "bad arguments to rnnwf_energy_moments");
    return energy_moments_impl(eloc, ns, stride, stats_out, (cudaStream_t)stream);
}

RNNWF_API int rnnwf_ffma_peak(int iters, double* tflops_out, void* stream) {
    RNNWF_CHECK(iters > 0 && tflops_out, -1, "bad arguments to rnnwf_ffma_peak");
    return ffma_peak_impl(iters, tflops_out, (cudaStream_t)stream);
}

RNNWF_API int rnnwf_fp64_peak(int mode, int iters, double* tflops_out, void* stream) {
    RNNWF_CHECK((mode == 0 || mode == 1) && iters > 0 && tflops_out, -1, "bad arguments to rnnwf_fp64_peak");
    return fp64_peak_impl(mode, iters, tflops_out, (cudaStream_t)stream);
}

RNNWF_API int rnnwf_umma_selftest(int n, int k, const float* a, const float* b, float* d, int passes, void* stream) {
    RNNWF_CHECK(a && b && d, -1, "bad arguments to rnnwf_umma_selftest");
    if (passes < 0) return umma_selftest_f16_impl(n, k, a, b, d, (-passes) & 3, (-passes) >> 2, (cudaStream_t)stream);
    return umma_selftest_impl(n, k, a, b, d, passes, (cudaStream_t)stream);
}

}  // extern "C"
