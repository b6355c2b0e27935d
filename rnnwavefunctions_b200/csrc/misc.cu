// misc.cu — TF1 Adam update and energy moments.
#include "api_internal.h"

namespace rnnwf {

// tf.train.AdamOptimizer.apply_gradients (1DTFIM/TrainingRNN_1DTFIM.py:113,164):
//   lr_t = lr sqrt(1-b2^t)/(1-b1^t);  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  theta -= lr_t m/(sqrt(v)+eps)
template <typename T>
__global__ void adam_kernel(int64_t n, T* __restrict__ theta, T* __restrict__ mom, T* __restrict__ vel,
                            const double* __restrict__ grad, double gs, double lr_t, double b1, double b2, double eps) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const T g = (T)(grad[i] * gs);
        const T m = (T)b1 * mom[i] + (T)(1.0 - b1) * g;
        const T v = (T)b2 * vel[i] + (T)(1.0 - b2) * g * g;
        mom[i] = m;
        vel[i] = v;
        theta[i] = theta[i] - (T)lr_t * m / (sqrt(v) + (T)eps);
    }
}

int adam_step_impl(int dtype, int64_t n, void* theta, void* mom, void* vel, const double* grad, double gs, double lr, double b1,
                   double b2, double eps, int64_t t, cudaStream_t s) {
    const double lr_t = lr * sqrt(1.0 - pow(b2, (double)t)) / (1.0 - pow(b1, (double)t));
    const int grid = (int)std::min<int64_t>(cdiv(n, 256), 1184);
    prof_count();
    if (dtype == RNNWF_F32)
        adam_kernel<float><<<grid, 256, 0, s>>>(n, (float*)theta, (float*)mom, (float*)vel, grad, gs, lr_t, b1, b2, eps);
    else
        adam_kernel<double><<<grid, 256, 0, s>>>(n, (double*)theta, (double*)mom, (double*)vel, grad, gs, lr_t, b1, b2, eps);
    RNNWF_CUDA(cudaGetLastError());
    return 0;
}

// stats = {sum E, sum E^2, n}; single block, fixed summation order (deterministic).
__global__ void moments_kernel(const double* __restrict__ e, int64_t ns, int stride, double* __restrict__ stats) {
    __shared__ double s1[256], s2[256];
    double a = 0.0, b = 0.0;
    for (int64_t i = threadIdx.x; i < ns; i += blockDim.x) {
        const double v = e[i * stride];
        a += v;
        b += v * v;
    }
    s1[threadIdx.x] = a;
    s2[threadIdx.x] = b;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            s1[threadIdx.x] += s1[threadIdx.x + o];
            s2[threadIdx.x] += s2[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        stats[0] = s1[0];
        stats[1] = s2[0];
        stats[2] = (double)ns;
    }
}

int energy_moments_impl(const double* eloc, int64_t ns, int stride, double* stats, cudaStream_t s) {
    prof_count();
    moments_kernel<<<1, 256, 0, s>>>(eloc, ns, stride, stats);
    RNNWF_CUDA(cudaGetLastError());
    return 0;
}

// FFMA peak probe: 16 independent accumulator chains per thread, 2 FMAs per chain per iteration.
__global__ void __launch_bounds__(256) ffma_probe_kernel(int iters, float a, float b, float* __restrict__ out) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = (float)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 123.456f) out[0] = s;
}

int ffma_peak_impl(int iters, double* tflops, cudaStream_t s) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    float* out = nullptr;
    RNNWF_CUDA(cudaMalloc(&out, 4));
    cudaEvent_t e0, e1;
    RNNWF_CUDA(cudaEventCreate(&e0));
    RNNWF_CUDA(cudaEventCreate(&e1));
    const int grid = sms * 8;
    ffma_probe_kernel<<<grid, 256, 0, s>>>(iters / 8 + 1, 0.999f, 0.001f, out);
    RNNWF_CUDA(cudaEventRecord(e0, s));
    ffma_probe_kernel<<<grid, 256, 0, s>>>(iters, 0.999f, 0.001f, out);
    RNNWF_CUDA(cudaEventRecord(e1, s));
    RNNWF_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    RNNWF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *tflops = 2.0 * 64.0 * (double)iters * 256.0 * grid / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return 0;
}

// FP64 peak probes: the denominators of the roofline of the float64 models (both 2-D apps of the reference compute in float64).
// mode 0: DFMA, 16 independent accumulator chains per thread; mode 1: mma.sync.m8n8k4.f64 (DMMA), 8 independent accumulator pairs per warp.
__global__ void __launch_bounds__(256) dfma_probe_kernel(int iters, double a, double b, double* __restrict__ out) {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = (double)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}
__global__ void __launch_bounds__(256) dmma_probe_kernel(int iters, double a, double b, double* __restrict__ out) {
    double c0[8], c1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c0[i] = (double)(threadIdx.x + i); c1[i] = 0.5 * i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c0[i] + c1[i];
    if (s == 123.456) out[0] = s;
}

int fp64_peak_impl(int mode, int iters, double* tflops, cudaStream_t s) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double* out = nullptr;
    RNNWF_CUDA(cudaMalloc(&out, 8));
    cudaEvent_t e0, e1;
    RNNWF_CUDA(cudaEventCreate(&e0));
    RNNWF_CUDA(cudaEventCreate(&e1));
    const int grid = sms * 8;
    auto launch = [&](int n) {
        if (mode == 0) dfma_probe_kernel<<<grid, 256, 0, s>>>(n, 0.999, 0.001, out);
        else dmma_probe_kernel<<<grid, 256, 0, s>>>(n, 0.999, 0.001, out);
    };
    launch(iters / 8 + 1);
    RNNWF_CUDA(cudaEventRecord(e0, s));
    launch(iters);
    RNNWF_CUDA(cudaEventRecord(e1, s));
    RNNWF_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    RNNWF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    // mode 0: 64 FMAs per thread and iteration; mode 1: 16 warp-wide m8n8k4 (256 FMAs each) per warp and iteration
    const double flops = mode == 0 ? 2.0 * 64.0 * (double)iters * 256.0 * grid : 2.0 * 256.0 * 16.0 * (double)iters * 8.0 * grid;
    *tflops = flops / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return 0;
}

}  // namespace rnnwf
