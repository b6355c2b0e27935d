// umma_selftest.cu — known-answer check of the tcgen05 plumbing in umma.cuh (descriptor encodings, operand layouts,
// TMEM addressing): D[128 x N] = A[128 x K] * B[N x K]^T with A staged in TMEM and B in shared memory, 1xTF32 or 3xTF32.
// Test infrastructure for the tensor-core recurrence kernels; exposed through rnnwf_umma_selftest.
#include <stdio.h>
#include <stdlib.h>
#include "api_internal.h"
#include <cuda_fp16.h>
#include "umma.cuh"

namespace rnnwf {


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
        if (k < K) {
            if (passes == 4) hi = B[n * K + k];          // probe: raw FP32 bits as kind::tf32 operands (what does the tensor core do with the low 13?)
            else umma::split_tf32(B[n * K + k], hi, lo);
        }
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
                if (k < K) {
                    if (passes == 4) hi[q] = A[tid * K + k];
                    else umma::split_tf32(A[tid * K + k], hi[q], lo[q]);
                }
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
        for (int pass = 0; pass < (passes == 4 ? 1 : passes); ++pass) {
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

// ---- same check for the FP16 path: kind::f16, A = fp16 pairs packed in TMEM columns (k even in the low half),
// B = fp16 K-major core matrices (8 rows x 8 halfs), 3 passes hi*hi + lo*hi + hi*lo; D is read back at column offset
// `dcol` (any value: checks that tcgen05.ld/st need no column alignment).
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));   // low half = a, high half = b
    return r;
}
__device__ __forceinline__ float h_lo_of(float x) {   // x - fp16(x)
    const __half h = __float2half_rn(x);
    return x - __half2float(h);
}

__global__ void __launch_bounds__(160, 1)
umma_selftest_f16_kernel(int N, int K, const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int passes, int dcol) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int Kp = (K + 15) & ~15, KC = Kp / 8;
    const uint32_t b_bytes = (uint32_t)N * Kp * 2;
    __half* Bhi = reinterpret_cast<__half*>(smem);
    __half* Blo = reinterpret_cast<__half*>(smem + b_bytes);
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
        const float v = k < K ? B[n * K + k] : 0.f;
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn(v - __half2float(hi));
        const int off = (n / 8) * (KC * 64) + (k / 8) * 64 + (n % 8) * 8 + (k % 8);   // in halfs
        Bhi[off] = hi;
        Blo[off] = lo;
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tbase = *tmem_slot;
    const uint32_t colD = (uint32_t)dcol, colAhi = 300, colAlo = 300 + 36;     // deliberately unaligned bases
    if (warp < 4) {
        const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
        for (int c = 0; c < Kp / 2; ++c) {
            const int k = 2 * c;
            const float a0 = k < K ? A[tid * K + k] : 0.f, a1 = k + 1 < K ? A[tid * K + k + 1] : 0.f;
            const float h0 = __half2float(__float2half_rn(a0)), h1 = __half2float(__float2half_rn(a1));
            float hi[1] = {__uint_as_float(pack_h2(a0, a1))}, lo[1] = {__uint_as_float(pack_h2(a0 - h0, a1 - h1))};
            umma::tmem_st1(lane_addr + colAhi + c, hi);
            umma::tmem_st1(lane_addr + colAlo + c, lo);
        }
        umma::wait_st();
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 4 && lane == 0) {
        umma::fence_after_sync();
        const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // F16 x F16 -> F32
        const uint32_t bhi = umma::smem_u32(Bhi), blo = umma::smem_u32(Blo);
        uint32_t acc = 0;
        for (int pass = 0; pass < passes; ++pass) {
            const uint32_t acol = pass == 1 ? colAlo : colAhi;
            const uint32_t bsm = pass == 2 ? blo : bhi;
            for (int ks = 0; ks < Kp / 16; ++ks) {
                const uint64_t bd = umma::smem_desc(bsm + ks * 256, 128, KC * 128);
                umma::mma_f16_ts(tbase + colD, tbase + acol + ks * 8, bd, idesc, acc);
                acc = 1;
            }
        }
        umma::commit(bar);
    }
    if (warp < 4) {
        umma::mbar_wait(bar, 0);
        umma::fence_after_sync();
        const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
        for (int n0 = 0; n0 < N; n0 += 8) {
            float v[8];
            umma::tmem_ld8(lane_addr + colD + n0, v);
            umma::wait_ld();
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (n0 + q < N) D[tid * N + n0 + q] = v[q];
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 4) umma::tmem_dealloc(tbase, 512);
}

// TMEM read-bandwidth probe (development aid): `warps` warps (multiple of 4) issue `iters` x (4 x tcgen05.ld x8) each.
__global__ void __launch_bounds__(544, 1) tmem_ldbw_kernel(int warps, int iters, float* out) {
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid / 32;
    if (warp == 16) umma::tmem_alloc(&slot, 512);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tbase = slot;
    float acc = 0.f;
    long long t0 = 0, t1 = 0;
    if (warp < warps) {
        const uint32_t lane_addr = tbase + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 32);
        const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int c = 0; c < 32; c += 8) umma::tmem_st8(lane_addr + c, z);
        umma::wait_st();
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            float a[8], b[8], c[8], d[8];
            umma::tmem_ld8(lane_addr, a);
            umma::tmem_ld8(lane_addr + 8, b);
            umma::tmem_ld8(lane_addr + 16, c);
            umma::tmem_ld8(lane_addr + 24, d);
            umma::wait_ld();
            acc += a[0] + b[1] + c[2] + d[3];
        }
        t1 = clock64();
    }
    __syncthreads();
    if (tid == 0) printf("[tmem ld probe] %d warps x %d x 4 ld8: %lld cycles -> %.1f B/cycle/SM\n", warps, iters, t1 - t0,
                         (double)warps * iters * 4 * 1024.0 / (double)(t1 - t0));
    if (acc == 123.f) out[0] = acc;
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 16) umma::tmem_dealloc(tbase, 512);
}

// MMA cost probe (development aid): `count` back-to-back kind::f16 MMAs of M=128, N, K=16 on zeroed operands.
__global__ void __launch_bounds__(160, 1) mma_cost_kernel(int N, int count, float* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    if (warp == 4) umma::tmem_alloc(&slot, 512);
    if (tid == 0) { umma::mbar_init(&bar, 1); umma::mbar_fence_init(); }
    for (int i = tid; i < 256 * 16 * 2 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tbase = slot;
    if (warp < 4) {
        const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
        const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        umma::tmem_st8(lane_addr + 300, z);
        umma::wait_st();
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 4 && lane == 0) {
        umma::fence_after_sync();
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t bd = umma::smem_desc(umma::smem_u32(smem), 128, 256);
        const long long t0 = clock64();
        for (int i = 0; i < count; ++i) umma::mma_f16_ts(tbase, tbase + 300, bd, idesc, i > 0);
        const long long t1 = clock64();
        umma::commit(&bar);
        umma::mbar_wait(&bar, 0);
        const long long t2 = clock64();
        printf("[mma cost] N=%3d: %d MMAs issue %lld cycles, complete %lld cycles -> %.1f cycles/MMA\n", N, count, t1 - t0, t2 - t0,
               (double)(t2 - t0) / count);
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 4) umma::tmem_dealloc(tbase, 512);
    if (out == nullptr) return;
}

// MMA cost probe 2 (development aid): like mma_cost_kernel, but B walks through `bspan` bytes of shared memory in steps of `bstride`
// and `ldwarps` other warps stream tcgen05.ld x8 (+ optional tcgen05.st) over the accumulator columns while the MMAs run.
__global__ void __launch_bounds__(544, 1) mma_cost2_kernel(int N, int count, int bstride, int bspan, int ldwarps, int with_st, float* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bar;
    __shared__ volatile int done;
    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    if (warp == 16) umma::tmem_alloc(&slot, 512);
    if (tid == 0) { umma::mbar_init(&bar, 1); umma::mbar_fence_init(); done = 0; }
    for (int i = tid; i < (bspan + 20480) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tbase = slot;
    const uint32_t lane_addr = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    if (warp < 4) {
        const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int c = 0; c < 64; c += 8) umma::tmem_st8(lane_addr + 300 + c, z);
        umma::wait_st();
    }
    umma::fence_before_sync();
    __syncthreads();
    float acc = 0.f;
    if (warp == 16 && lane == 0) {
        umma::fence_after_sync();
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t sb = umma::smem_u32(smem);
        uint32_t off = 0;
        const long long t0 = clock64();
        for (int i = 0; i < count; ++i) {
            umma::mma_f16_ts(tbase, tbase + 300 + (i & 7) * 8, umma::smem_desc(sb + off, 128, 1024), idesc, i > 0);
            off += bstride;
            if (off >= (uint32_t)bspan) off = 0;
        }
        const long long t1 = clock64();
        umma::commit(&bar);
        umma::mbar_wait(&bar, 0);
        const long long t2 = clock64();
        done = 1;
        printf("[mma cost2] N=%3d bstride %d span %d, %d ld warps (st %d): %d MMAs issue %lld, complete %lld -> %.1f cycles/MMA\n", N, bstride, bspan,
               ldwarps, with_st, count, t1 - t0, t2 - t0, (double)(t2 - t0) / count);
    } else if (warp < ldwarps) {
        const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        long long n = 0;
        while (!done) {
            float a[8], b[8], c[8], d[8];
            umma::tmem_ld8(lane_addr + 256, a);
            umma::tmem_ld8(lane_addr + 264, b);
            umma::tmem_ld8(lane_addr + 272, c);
            umma::tmem_ld8(lane_addr + 280, d);
            umma::wait_ld();
            acc += a[0] + b[1] + c[2] + d[3];
            if (with_st) { umma::tmem_st8(lane_addr + 400 + (warp >> 2) * 8, z); umma::wait_st(); }
            ++n;
        }
        if (lane == 0 && warp == 0) printf("[mma cost2]   ld warp 0 did %lld x 4 ld8\n", n);
    }
    if (acc == 123.f) out[0] = acc;
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 16) umma::tmem_dealloc(tbase, 512);
}

int umma_selftest_f16_impl(int N, int K, const float* A, const float* B, float* D, int passes, int dcol, cudaStream_t s) {
    if (const char* pe = getenv("RNNWF_PROBE")) {   // development aid: "N,count,bstride,bspan,ldwarps,with_st"
        int pn = 112, pc = 400, pbs = 0, psp = 8192, plw = 0, pst = 0;
        sscanf(pe, "%d,%d,%d,%d,%d,%d", &pn, &pc, &pbs, &psp, &plw, &pst);
        const int smem = psp + 20480;
        RNNWF_CUDA(cudaFuncSetAttribute(mma_cost2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        mma_cost2_kernel<<<1, 544, smem, s>>>(pn, pc, pbs, psp, plw, pst, D);
        RNNWF_CUDA(cudaGetLastError());
        return 0;
    }
    if (passes == 2 && dcol >= 100) {   // development aid: MMA cost probe, N = dcol - 100
        RNNWF_CUDA(cudaFuncSetAttribute(mma_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
        mma_cost_kernel<<<1, 160, 16384, s>>>(dcol - 100, 400, D);
        RNNWF_CUDA(cudaGetLastError());
        return 0;
    }
    if (passes == 2) {   // development aid: TMEM read-bandwidth probe with dcol warps
        tmem_ldbw_kernel<<<1, 544, 0, s>>>(dcol, 2000, D);
        RNNWF_CUDA(cudaGetLastError());
        return 0;
    }
    RNNWF_CHECK(N >= 16 && N <= 256 && N % 16 == 0 && K >= 1 && K <= 64 && (passes == 1 || passes == 3) && dcol >= 0 && dcol + N <= 300, -1,
                "umma f16 selftest: bad shape");
    const int Kp = (K + 15) & ~15;
    const int smem = 2 * N * Kp * 2 + 64;
    RNNWF_CUDA(cudaFuncSetAttribute(umma_selftest_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    prof_count();
    umma_selftest_f16_kernel<<<1, 160, smem, s>>>(N, K, A, B, D, passes, dcol);
    RNNWF_CUDA(cudaGetLastError());
    return 0;
}

int umma_selftest_impl(int N, int K, const float* A, const float* B, float* D, int passes, cudaStream_t s) {
    RNNWF_CHECK(N >= 16 && N <= 256 && N % 16 == 0 && K >= 1 && K <= 64 && (passes == 1 || passes == 3 || passes == 4), -1,
                "umma selftest: N in [16,256] multiple of 16, K <= 64, passes 1 or 3 (4: raw FP32 operands, one pass)");
    const int Kp = (K + 7) & ~7;
    const int smem = 2 * N * Kp * 4 + 64;
    RNNWF_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    prof_count();
    umma_selftest_kernel<<<1, 160, smem, s>>>(N, K, A, B, D, passes);
    RNNWF_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rnnwf
