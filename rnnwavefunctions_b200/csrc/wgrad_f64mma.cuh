// wgrad_f64mma.cuh — float64 weight-gradient reduction on the FP64 tensor instruction (mma.sync.m8n8k4.f64, DMMA), shared by the GRU
// (grad.cuh) and the 2-D RNN (mdrnn.cu) gradients.  Both 2-D apps of the reference compute in float64
// (2DTFIM_1DRNN/RNNwavefunction.py:26-38, 2DTFIM_2DRNN/RNNwavefunction.py:28-33); this is the device work of
// optimizer.compute_gradients(cost) (2DTFIM_1DRNN/Training1DRNN_2DTFIM.py:193-197) that contracts over (sample, site).
//
// C[r][c] = sum_(blk, m) A[r][m] B[c][m] is a GEMM with K = the samples of the (tile, site) blocks, and both operands lie
// K-contiguous in HBM ([row][M]), which is what the DMMA fragments want: lane (g = lane / 4, q = lane % 4) holds A[8 mt + g][k0 + q]
// and B[8 nt + g][k0 + q], and C[8 mt + g][8 nt + 2 q + {0, 1}].  One CTA owns a [104 x 104] output tile (13 x 13 fragments over a
// 4 x 4 grid of warps, at most 4 x 4 fragments = 32 accumulators per thread, FP64, kept in registers for the whole launch: fixed
// summation order, no atomics) and a share of the blocks; the operands go through a 4-stage cp.async ring of K slices of 16 samples
// (row stride 20 doubles: conflict-free LDS.64 for the fragment pattern).  The thread-tile kernels it replaces spent 85 ms of cfg3's
// 446 ms step here (1.4 TFLOP/s: 8 LDS.128 per 16 DFMA, every operand re-read rtiles * ctiles times): 12 ms now.
//
// Src describes the operands of one model family:
//   int M, R, C; int64_t nblk;
//   Ctx begin(int64_t blk) const      what is common to every element of block blk (site index, neighbour positions, ...)
//   const double* a_src(const Ctx&, int64_t blk, int r, int k, bool& special, double& v0, double& v1) const
//        address of A[r][k], A[r][k + 1] of block blk (nullptr: zeros), or special = true with the two values computed (one-hot rows,
//        the constant-1 row);  k is even and k < M
//   const double* b_src(const Ctx&, int64_t blk, int c, int k) const
#pragma once
#include "gru_engine.cuh"

namespace rnnwf {
namespace wgdm {
constexpr int kThreads = 512, kT = 104, kKS = 16, kMS = 20, kStages = 4;
constexpr int kStageDoubles = 2 * kT * kMS;
constexpr size_t kSmem = (size_t)kStages * kStageDoubles * sizeof(double);     // 133 KB

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <class Src>
__global__ void __launch_bounds__(kThreads, 1) wgrad_kernel(Src src, int ct, int ksplit, double* __restrict__ partial, int Rp, int Cp) {
    extern __shared__ __align__(16) unsigned char smem[];
    double* buf = reinterpret_cast<double*>(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g8 = lane >> 2, q4 = lane & 3;
    const int M = src.M;
    const int tile = blockIdx.x, ks = blockIdx.y;
    const int r0 = (tile / ct) * kT, c0 = (tile % ct) * kT;
    const int64_t b0 = src.nblk * ks / ksplit, b1 = src.nblk * (ks + 1) / ksplit;
    const int nsl = (M + kKS - 1) / kKS;
    const int64_t T = (b1 - b0) * nsl;
    // fragments of this warp: 13 = 4 + 3 + 3 + 3 in both directions.  Warp w issues on SM sub-partition w % 4; the column group is rotated
    // by the row group so that every sub-partition gets 43 or 42 of the 169 fragments (unrotated: 52 / 39 / 39 / 39, and ncu showed the
    // DMMA pipe 69 % active with math_pipe_throttle as the only stall)
    const int wm = warp >> 2, wn = (warp + wm) & 3;
    const int mt0 = wm == 0 ? 0 : 1 + 3 * wm, nm = wm == 0 ? 4 : 3;
    const int nt0 = wn == 0 ? 0 : 1 + 3 * wn, nn = wn == 0 ? 4 : 3;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto issue = [&](int64_t t) {      // K slice t of this CTA -> stage t % kStages
        const int64_t blk = b0 + t / nsl;
        const int k0 = (int)(t % nsl) * kKS;
        const auto ctx = src.begin(blk);
        double* A = buf + (size_t)(t % kStages) * kStageDoubles;
        double* B = A + kT * kMS;
        for (int i = tid; i < 2 * kT * (kKS / 2); i += kThreads) {
            const int row = i / (kKS / 2), k = k0 + 2 * (i % (kKS / 2));
            double* dst = (row < kT ? A + row * kMS : B + (row - kT) * kMS) + (k - k0);
            const double* p = nullptr;
            if (k < M) {                                                  // M is even: a 16-byte piece never straddles the tile
                if (row < kT) {
                    bool special = false;
                    double v0 = 0.0, v1 = 0.0;
                    if (r0 + row < src.R) p = src.a_src(ctx, blk, r0 + row, k, special, v0, v1);
                    if (special) {
                        *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
                        continue;
                    }
                } else if (c0 + row - kT < src.C) {
                    p = src.b_src(ctx, blk, c0 + row - kT, k);
                }
            }
            if (p != nullptr) cp_async16(dst, p);
            else *reinterpret_cast<double2*>(dst) = make_double2(0.0, 0.0);
        }
    };

    for (int t = 0; t < kStages - 1; ++t) {
        if (t < T) issue(t);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int64_t t = 0; t < T; ++t) {
        asm volatile("cp.async.wait_group %0;" ::"n"(kStages - 2) : "memory");
        __syncthreads();                                                  // slice t has landed; everybody is done with slice t - 1
        if (t + kStages - 1 < T) issue(t + kStages - 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const double* A = buf + (size_t)(t % kStages) * kStageDoubles + (size_t)(8 * mt0 + g8) * kMS + q4;
        const double* B = buf + (size_t)(t % kStages) * kStageDoubles + (size_t)(kT + 8 * nt0 + g8) * kMS + q4;
#pragma unroll
        for (int kk = 0; kk < kKS; kk += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = i < nm ? A[i * 8 * kMS + kk] : 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = j < nn ? B[j * 8 * kMS + kk] : 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (i < nm && j < nn) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (i < nm && j < nn) {
                const int r = r0 + 8 * (mt0 + i) + g8, c = c0 + 8 * (nt0 + j) + 2 * q4;
                if (r < Rp) {
                    if (c < Cp) partial[((size_t)ks * Rp + r) * Cp + c] = acc[i][j][0];
                    if (c + 1 < Cp) partial[((size_t)ks * Rp + r) * Cp + c + 1] = acc[i][j][1];
                }
            }
        }
}

// grid.y for a launch: about one CTA per SM, never more blocks' shares than the partial buffer has slots
inline int choose_ksplit(int R, int C, int64_t nblk, int64_t slots, int sms = 148) {
    const int tiles = (int)(((R + kT - 1) / kT) * ((C + kT - 1) / kT));
    return (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(slots, nblk), std::max(1, sms / tiles)));
}
}  // namespace wgdm
}  // namespace rnnwf
