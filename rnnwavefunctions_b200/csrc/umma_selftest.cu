// umma_selftest.cu — known-answer check of the tcgen05 plumbing in umma.cuh (descriptor encodings, operand layouts,
// TMEM addressing): D[128 x N] = A[128 x K] * B[N x K]^T with A staged in TMEM and B in shared memory, 1xTF32 or 3xTF32.
// Test infrastructure for the tensor-core recurrence kernels; exposed through rnnwf_umma_selftest.
#include "api_internal.h"
#include "umma.cuh"

namespace rnnwf {

constexpr int kStRows = 128;

// smem: B_hi | B_lo in the canonical K-major no-swizzle layout (LBO = 128 B between K chunks, SBO = KC * 128 B between
// 8-row groups), then the mbarrier and the TMEM base address.
__global__ void __launch_bounds__(160, 1)
umma_selftest_kernel(int N, int K, const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int passes) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int Kp = (K + 7) & ~7, KC = Kp / 4;
    const uint32_t b_bytes = (uint32_t)N * Kp * 4;
    float* Bhi = reinterpret_cast<float*>(smem);
    float* Blo = reinterpret_cast<float*>(smem + b_bytes);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * b_bytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 2 * b_bytes + 8);
    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;

    if (warp == 4) umma::tmem_alloc(tmem_slot, 512);
    if (tid == 0) {
        umma::mbar_init(bar, 1);
        umma::mbar_fence_init();
    }
    for (int i = tid; i < N * Kp; i += blockDim.x) {
        const int n = i / Kp, k = i % Kp;
        float hi = 0.f, lo = 0.f;
        if (k < K) umma::split_tf32(B[n * K + k], hi, lo);
        const int off = (n / 8) * (KC * 32) + (k / 4) * 32 + (n % 8) * 4 + (k % 4);   // in floats
        Bhi[off] = hi;
        Blo[off] = lo;
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tbase = *tmem_slot;
    const uint32_t colD = 0, colAhi = 256, colAlo = 256 + 64;

    if (warp < 4) {   // thread = row: stage A (hi / lo) into TMEM, 8 columns at a time
        const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
        for (int k0 = 0; k0 < Kp; k0 += 8) {
            float hi[8], lo[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int k = k0 + q;
                hi[q] = 0.f; lo[q] = 0.f;
                if (k < K) umma::split_tf32(A[tid * K + k], hi[q], lo[q]);
            }
            umma::tmem_st8(lane_addr + colAhi + k0, hi);
            umma::tmem_st8(lane_addr + colAlo + k0, lo);
        }
        umma::wait_st();
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 4 && lane == 0) {
        umma::fence_after_sync();
        const uint32_t idesc = umma::instr_desc(umma::kFmtTF32, 128, N);
        const uint32_t bhi = umma::smem_u32(Bhi), blo = umma::smem_u32(Blo);
        uint32_t acc = 0;
        for (int pass = 0; pass < passes; ++pass) {
            const uint32_t acol = pass == 1 ? colAlo : colAhi;       // hi*hi, lo*hi, hi*lo
            const uint32_t bsm = pass == 2 ? blo : bhi;
            for (int ks = 0; ks < Kp / 8; ++ks) {
                const uint64_t bd = umma::smem_desc(bsm + ks * 256, 128, KC * 128);
                umma::mma_tf32_ts(tbase + colD, tbase + acol + ks * 8, bd, idesc, acc);
                acc = 1;
            }
        }
        umma::commit(bar);
    }
    if (warp < 4) {
        umma::mbar_wait(bar, 0);
        umma::fence_after_sync();
        const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
        for (int n0 = 0; n0 < N; n0 += 16) {
            float v[16];
            umma::tmem_ld16(lane_addr + colD + n0, v);
            umma::wait_ld();
#pragma unroll
            for (int q = 0; q < 16; ++q)
                if (n0 + q < N) D[tid * N + n0 + q] = v[q];
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 4) umma::tmem_dealloc(tbase, 512);
}

int umma_selftest_impl(int N, int K, const float* A, const float* B, float* D, int passes, cudaStream_t s) {
    RNNWF_CHECK(N >= 16 && N <= 256 && N % 16 == 0 && K >= 1 && K <= 64 && (passes == 1 || passes == 3), -1,
                "umma selftest: N in [16,256] multiple of 16, K <= 64, passes 1 or 3");
    const int Kp = (K + 7) & ~7;
    const int smem = 2 * N * Kp * 4 + 64;
    RNNWF_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    prof_count();
    umma_selftest_kernel<<<1, 160, smem, s>>>(N, K, A, B, D, passes);
    RNNWF_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rnnwf
