// api_internal.h — typed entry points behind the C ABI (one per translation unit family).
#pragma once
#include <algorithm>
#include "common.cuh"

namespace rnnwf {

// gru.cu
template <typename T> size_t gru_workspace_bytes_t(const rnnwf_model& m, int op, int64_t ns, int flags);
template <typename T> int gru_sample_t(const rnnwf_model& m, const void* params, int64_t ns, uint64_t seed, uint64_t off,
                                       uint8_t* out, void* ws, size_t wsb, cudaStream_t s);
template <typename T> int gru_logpsi_t(const rnnwf_model& m, const void* params, const uint8_t* samples, int64_t ns, int flags,
                                       double* out, void* ws, size_t wsb, cudaStream_t s);
template <typename T> int gru_tfim_eloc_t(const rnnwf_model& m, const void* params, const uint8_t* samples, int64_t ns,
                                          const double* jz, double bx, int flags, double* eloc, double* logp, double* ratios,
                                          void* ws, size_t wsb, cudaStream_t s);
int tfim_chain_mode_impl(const rnnwf_model& m);
int tfim_diag_impl(const rnnwf_model& m, const uint8_t* samples, int64_t ns, const double* jz, double* diag, cudaStream_t s);
int tfim_finalize_impl(const double* diag, const double* delta, const double* lp, int64_t ns, int N, int M, int tiles_s, double bx,
                       int parity, double* eloc, double* logp, double* ratios, cudaStream_t s);
int tfim_enumerate_impl(const uint8_t* samples, int64_t ns, int N, int32_t* queue, cudaStream_t s);

// grad.cu
template <typename T> size_t gru_grad_workspace_bytes(const rnnwf_model& m, int64_t ns, int flags);
template <typename T> int gru_vmc_grad_t(const rnnwf_model& m, const void* params, const uint8_t* samples, int64_t ns,
                                         const double* weights, int flags, double* grad, void* ws, size_t wsb, cudaStream_t s);

// j1j2.cu
int j1j2_enumerate_impl(const uint8_t* samples, int64_t ns, int N, const double* j1, const double* j2, const double* bz,
                        int periodic, int marshall, int32_t* sigmas, float* elements, int32_t* counts, cudaStream_t s);
template <typename T> int gru_j1j2_eloc_t(const rnnwf_model& m, const void* params, const uint8_t* samples, int64_t ns,
                                          const double* j1, const double* j2, const double* bz, int marshall, double* eloc,
                                          double* logpsi, void* ws, size_t wsb, cudaStream_t s);

// mdrnn.cu
template <typename T> size_t mdrnn_workspace_bytes_t(const rnnwf_model& m, int op, int64_t ns, int flags);
template <typename T> int mdrnn_sample_t(const rnnwf_model& m, const void* params, int64_t ns, uint64_t seed, uint64_t off,
                                         uint8_t* out, void* ws, size_t wsb, cudaStream_t s);
template <typename T> int mdrnn_logpsi_t(const rnnwf_model& m, const void* params, const uint8_t* samples, int64_t ns,
                                         double* out, void* ws, size_t wsb, cudaStream_t s);
template <typename T> int mdrnn_tfim_eloc_t(const rnnwf_model& m, const void* params, const uint8_t* samples, int64_t ns,
                                            const double* jz, double bx, double* eloc, double* logp, double* ratios, void* ws,
                                            size_t wsb, cudaStream_t s);
template <typename T> int mdrnn_vmc_grad_t(const rnnwf_model& m, const void* params, const uint8_t* samples, int64_t ns,
                                           const double* weights, double* grad, void* ws, size_t wsb, cudaStream_t s);

// umma_selftest.cu
int umma_selftest_f16_impl(int N, int K, const float* A, const float* B, float* D, int passes, int dcol, cudaStream_t s);
int umma_selftest_impl(int N, int K, const float* A, const float* B, float* D, int passes, cudaStream_t s);

// misc.cu
int adam_step_impl(int dtype, int64_t n, void* theta, void* mom, void* vel, const double* grad, double grad_scale, double lr,
                   double b1, double b2, double eps, int64_t t, cudaStream_t s);
int ffma_peak_impl(int iters, double* tflops, cudaStream_t s);
int fp64_peak_impl(int mode, int iters, double* tflops, cudaStream_t s);
int energy_moments_impl(const double* eloc, int64_t ns, int stride, double* stats, cudaStream_t s);

}  // namespace rnnwf
