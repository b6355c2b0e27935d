// common.cuh — shared device/host helpers for librnnwf_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/rnnwf.h"

namespace rnnwf {

// ------------------------------------------------------------------------------------------------
// error plumbing (thread-local message, no exceptions across the ABI)
// ------------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define RNNWF_CHECK(cond, code, ...)          \
    do {                                      \
        if (!(cond)) {                        \
            rnnwf::set_error(__VA_ARGS__);    \
            return (code);                    \
        }                                     \
    } while (0)

#define RNNWF_CUDA(expr)                                                                   \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            rnnwf::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return -100 - (int)_e;                                                         \
        }                                                                                  \
    } while (0)

// ------------------------------------------------------------------------------------------------
// measurement hooks (capi.cu): launch counter + CUDA-event brackets around the dominant kernel.
// Inactive unless rnnwf_profile_begin() was called; never changes what is computed.
// ------------------------------------------------------------------------------------------------
void prof_count(int n = 1);
void prof_mark(int which, cudaStream_t s);   // which: 0 = before, 1 = after the dominant kernel of the call

constexpr int kMaxLayers = 8;
constexpr int kSmemLimit = 232448;  // 227 KB opt-in dynamic shared memory per CTA on sm_100
constexpr int kHeadThreads = 128;   // 4 "head" warps: dense + softmax + draw, overlapped with layer 0

// ------------------------------------------------------------------------------------------------
// per-dtype traits: samples-per-thread of the register tile, vector loads
// ------------------------------------------------------------------------------------------------
template <typename T> struct VT;
template <> struct VT<float> { static constexpr int SPT = 8; };
template <> struct VT<double> { static constexpr int SPT = 4; };

template <int N> __device__ __forceinline__ void ldv(float (&d)[N], const float* p) {
    if constexpr (N == 8) {
        float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
    } else if constexpr (N == 4) {
        float4 a = *reinterpret_cast<const float4*>(p);
        d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w;
    } else if constexpr (N == 2) {
        float2 a = *reinterpret_cast<const float2*>(p);
        d[0] = a.x; d[1] = a.y;
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) d[i] = p[i];
    }
}
template <int N> __device__ __forceinline__ void ldv(double (&d)[N], const double* p) {
    static_assert(N % 2 == 0, "even");
#pragma unroll
    for (int i = 0; i < N; i += 2) {
        double2 a = *reinterpret_cast<const double2*>(p + i);
        d[i] = a.x; d[i + 1] = a.y;
    }
}
template <int N> __device__ __forceinline__ void stv(float* p, const float (&d)[N]) {
    if constexpr (N == 8) {
        *reinterpret_cast<float4*>(p) = make_float4(d[0], d[1], d[2], d[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(d[4], d[5], d[6], d[7]);
    } else if constexpr (N == 4) {
        *reinterpret_cast<float4*>(p) = make_float4(d[0], d[1], d[2], d[3]);
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) p[i] = d[i];
    }
}
template <int N> __device__ __forceinline__ void stv(double* p, const double (&d)[N]) {
#pragma unroll
    for (int i = 0; i < N; i += 2) *reinterpret_cast<double2*>(p + i) = make_double2(d[i], d[i + 1]);
}

// ------------------------------------------------------------------------------------------------
// activations.  float: ex2.approx + rcp.approx based (abs. error ~1e-7);
// double: libdevice.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoid_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }
__device__ __forceinline__ double sigmoid_(double x) { return 1.0 / (1.0 + exp(-x)); }
__device__ __forceinline__ double tanh_(double x) { return tanh(x); }
__device__ __forceinline__ float elu_(float x) { return x > 0.f ? x : expm1f(x); }
__device__ __forceinline__ double elu_(double x) { return x > 0.0 ? x : expm1(x); }

// log softmax(z)[sel] for a 2-way head, evaluated in double from T logits:
//   log p_sel = -log(1 + exp(z_other - z_sel))
__device__ __forceinline__ double log_softmax2(double z_sel, double z_other) {
    double d = z_other - z_sel;
    return d > 36.0 ? -d : -log1p(exp(d));
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011).  counter = (sample id lo, hi, site, stream), key = seed.
// Mirrors oracle/rnnwf_oracle.py::philox4x32 (checked against the Random123 known-answer vectors).
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3,
                                                       uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
__host__ __device__ __forceinline__ float philox_uniform(uint64_t seed, uint64_t sample_id, uint32_t site, uint32_t stream = 0) {
    uint32_t c0 = (uint32_t)sample_id, c1 = (uint32_t)(sample_id >> 32), c2 = site, c3 = stream;
    philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
    return (float)(c0 >> 8) * (1.0f / 16777216.0f);
}

// ------------------------------------------------------------------------------------------------
// NumPy's pairwise summation order for a contiguous run of n doubles (numpy/core/src/umath
// loops_utils.h::DOUBLE_pairwise_sum, as used by np.sum(..., axis=last)).  `f(i)` yields element i.
// Needed for bit-exact 2-D diagonal energies (2DTFIM_*/Training*.py:33-49 use np.sum(axis=1)).
// ------------------------------------------------------------------------------------------------
template <typename F> __device__ double np_pairwise_sum(F f, int lo, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += f(lo + i);
        return res;
    } else if (n <= 128) {
        double r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = f(lo + j);
        int i;
        for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] += f(lo + i + j);
        }
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += f(lo + i);
        return res;
    } else {
        int n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_sum(f, lo, n2) + np_pairwise_sum(f, lo + n2, n - n2);
    }
}

__host__ __device__ __forceinline__ int align4(int x) { return (x + 3) & ~3; }
__host__ __device__ __forceinline__ int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace rnnwf
