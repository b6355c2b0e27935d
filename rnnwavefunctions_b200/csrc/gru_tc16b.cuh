// gru_tc16b.cuh — backward recurrence of the VMC gradient (BPTT through one GRU layer) on tcgen05, for the FP32 probability-head
// stacks the pipelined chain kernel covers (50 units, <= 3 layers).  Replaces gru_bwd_layer_kernel<float> (grad.cuh), i.e. the
// device work of optimizer.compute_gradients(cost) of 1DTFIM/TrainingRNN_1DTFIM.py:156-160 that is sequential in the site index.
//
// The teacher-forced stash pass (tc16p::chain_kernel<BASE>) leaves, per (sample, site, layer, unit), the five factors that turn d h
// into everything the step needs (store_bwd_factors in gru_tc16p.cuh): with d h_n = carry + d out,
//     carry' = d h * u,  d a_c = d h * alpha,  d a_u = d h * beta,  d a_r = d h * gamma,  d aq = d h * rho,
// so a step is five multiplications per unit, and the only coupling between units is
//     [carry_{n-1} | d x_n] = [d a_r | d a_u | d a_c | d aq] (K = 4 x 50) x [W_h^T | W_x^T] (N = 50 + 50)
// which runs on the tensor cores: M = 128 rows, kind::f16 with the gate gradients split into (hi, lo) halfs in TMEM and the
// transposed weights as (hi, lo) K-major core-matrix images in shared memory (hi*hi + hi*lo + lo*hi, FP32 accumulation), exactly
// the operand scheme of the forward kernels.  The recurrence is linear in d h, so it runs with UNIT sample weight (values O(1),
// safely inside the FP16 range) and the per-sample weight (E_loc - mean) / ns multiplies the gate gradients only where they are
// written out for the weight-gradient reduction.
//
// One CTA per tile of M <= 68 rows (the gradient's tiles are sized so that they fill the SMs in whole waves: the walk over the N
// sites is sequential, tiles are the only parallelism), 8 row warps (two threads per row, 26 units each, as in gru_tc16p.cuh), the
// MMA warp and a producer thread.  The stash keeps a (tile, site, layer) block of factors contiguous ([unit][row][4 factors] + [unit][row], 68 KB at
// M = 68): the producer brings the block of the next site in with one bulk async copy (TMA) into a two-deep shared-memory ring, two
// sites ahead of its use (first version: 130 four-byte cp.async per thread and site -- 2 300 instructions per thread and site and
// the HBM latency exposed every site: 8.3 us per site).
// TMEM: A = gate gradients, one column per (unit pair, gate): column 4 (j / 2) + gate holds the halfs of units j, j + 1, i.e.
// K index 8 (j / 2) + 2 gate + j % 2; hi at [0, 104), lo at [104, 208) (a row thread stages a unit pair's four gates with one
// tcgen05.st.x4 per precision); D at [208, 320): carry in columns [0, 50), d x in [52, 102).  Included by gru.cu.
#pragma once
#include "gru_tc16p.cuh"

namespace rnnwf {
namespace tc16b {

using tc16::core_off;
using tc16::pack_h2;
using tc16::unpack_h2;

constexpr int kRows = 128, kRowThreads = 256, kThreads = 384, kMmaWarp = 8;
constexpr int kH = 50, kUP = 26, kPU = 24;
constexpr int kALo = 104, kColD = 208;         // 25 unit pairs x 4 gates = 100 columns per precision, padded to 13 MMA K-steps
constexpr int kK = 2 * kALo, kKC = kK / 8;     // 208 halfs per precision, 26 chunks of 8 halfs per image row
constexpr int kChunks = kK / 16;               // 13 MMA K-steps per precision
constexpr int kNFull = 112, kNCarry = 64;      // N of the MMAs: carry | d x (layers above 0), carry only (layer 0)
constexpr int kFactors = 5;

constexpr int kMaxM = 68;                      // rows per tile: two factor buffers of 5 x 50 x M floats beside the weight images

struct Layout {
    int n_out, im_bytes, img_bytes, fac_off, fac_bytes, tab_off, smem_bytes;
};
inline Layout make_layout(int l, int M) {
    Layout t;
    t.n_out = l > 0 ? kNFull : kNCarry;
    t.im_bytes = t.n_out * kK * 2;
    t.img_bytes = 2 * t.im_bytes;
    t.fac_off = (t.img_bytes + 127) & ~127;
    t.fac_bytes = kFactors * kH * M * 4;                          // one (tile, site, layer) block of the stash: contiguous in HBM
    t.tab_off = t.fac_off + 2 * ((t.fac_bytes + 127) & ~127);
    t.smem_bytes = t.tab_off + 64 * 4 + 128;
    return t;
}
inline bool supported(const GruLayout& g) { return tc16p::supported(g) && g.nheads == 1; }
// rows per tile: one CTA per tile walks the N sites alone, so the tiles should fill the SMs in whole waves
inline int choose_rows(int64_t rows, int sms = 148) {
    int64_t M = (rows + sms - 1) / sms;
    M = (M + 3) & ~(int64_t)3;
    return (int)std::min<int64_t>(kMaxM, std::max<int64_t>(16, M));
}

// transposed weights of layer l -> (hi, lo) images: row n < 50: carry unit i = n, rows 52..101: d x unit i = n - 52;
// K index k = 8 (j / 2) + 2 gate + j % 2 for gate unit j: gates d a_r, d a_u, d a_c, d aq.
__global__ void pack_kernel(GruLayout g, int l, Layout t, const float* __restrict__ flat, unsigned char* __restrict__ img) {
    const int H = g.H, d = g.d[l];
    const float* Kg = flat + g.flat_off[l];
    const float* Kci = Kg + (d + H) * 2 * H + 2 * H;
    const float* Kch = Kci + d * H;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < t.n_out * kK; idx += gridDim.x * blockDim.x) {
        const int n = idx / kK, k = idx % kK, blk = (k % 8) / 2, j = 2 * (k / 8) + (k % 2);
        float v = 0.f;
        if (j < H) {
            if (n < H) {                       // carry_i = sum_j d a_r[j] Kg[d+i][j] + d a_u[j] Kg[d+i][H+j] + d aq[j] Kch[i][j]
                const int i = n;
                if (blk == 0) v = Kg[(d + i) * 2 * H + j];
                else if (blk == 1) v = Kg[(d + i) * 2 * H + H + j];
                else if (blk == 3) v = Kch[i * H + j];
            } else if (l > 0 && n >= 52 && n < 52 + H) {   // d x_i = sum_j d a_r[j] Kg[i][j] + d a_u[j] Kg[i][H+j] + d a_c[j] Kci[i][j]
                const int i = n - 52;
                if (blk == 0) v = Kg[i * 2 * H + j];
                else if (blk == 1) v = Kg[i * 2 * H + H + j];
                else if (blk == 2) v = Kci[i * H + j];
            }
        }
        const __half hi = __float2half_rn(v);
        reinterpret_cast<__half*>(img)[core_off(n, k, kKC)] = hi;
        reinterpret_cast<__half*>(img + t.im_bytes)[core_off(n, k, kKC)] = __float2half_rn(v - __half2float(hi));
    }
}

struct Args {
    GruLayout g;
    Layout t;
    int l, M, top;
    const unsigned char* img;
    const float* flat;            // head weights (top layer)
    const float* gstore;          // [tile][site][layer]{[unit][M][u, alpha, beta, gamma], [unit][M] rho}
    const uint8_t* sigT;          // [tile][site][M]
    const double* la_oth;         // [tile][site][M]
    const double* roww;           // [tile][M]
    float* dxbuf;                 // [tile][site][unit][M]: read (d out of this layer, written by the layer above), then written (d x for the layer below)
    float* Gbuf;                  // [tile][site][4][unit][M]  (weighted gate gradients for the weight-gradient reduction)
    float* dzbuf;                 // [tile][site][2][M]        (weighted head gradients, top layer)
};

// (d a_r, d a_u, d a_c, d aq) of a unit pair -> its 4 columns of the hi and of the lo region
__device__ __forceinline__ void stage4(uint32_t col, const float* v) {
    float hi[4], lo[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint32_t wh = pack_h2(v[2 * c], v[2 * c + 1]);
        const float2 f = unpack_h2(wh);
        hi[c] = __uint_as_float(wh);
        lo[c] = __uint_as_float(pack_h2(v[2 * c] - f.x, v[2 * c + 1] - f.y));
    }
    umma::tmem_st4(col, hi);
    umma::tmem_st4(col + kALo, lo);
}

enum { kBStaged = 0, kBFull = 1, kBWImg = 2, kBFac0 = 3, kBFac1 = 4, kBFree0 = 5, kBFree1 = 6, kBars = 7 };

__global__ void __launch_bounds__(kThreads, 1) bwd_kernel(const __grid_constant__ Args a) {
    extern __shared__ __align__(128) unsigned char smem_b16[];
    const Layout& t = a.t;
    const int fac_stride = (t.fac_bytes + 127) & ~127;
    float* wdiff = reinterpret_cast<float*>(smem_b16 + t.tab_off);        // Wd[j][0] - Wd[j][1]
    uint64_t* bars = reinterpret_cast<uint64_t*>(wdiff + 64);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kBars);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int H = kH, N = a.g.N, L = a.g.L, l = a.l, M = a.M;
    const bool is_row = warp < kMmaWarp;
    const size_t st = blockIdx.x;
    const float* gsrc = a.gstore + ((st * N * L + l) * kFactors) * (size_t)H * M;       // + n * L * 5 * H * M
    const size_t gstep = (size_t)L * kFactors * H * M;

    if (warp == kMmaWarp) umma::tmem_alloc(tmem_slot, 512);
    if (tid == 0) {
        umma::mbar_init(&bars[kBStaged], kRowThreads);
        umma::mbar_init(&bars[kBFull], 1);
        umma::mbar_init(&bars[kBWImg], 1);
        umma::mbar_init(&bars[kBFac0], 1);
        umma::mbar_init(&bars[kBFac1], 1);
        umma::mbar_init(&bars[kBFree0], kRowThreads);
        umma::mbar_init(&bars[kBFree1], kRowThreads);
        umma::mbar_fence_init();
    }
    if (tid < 64) wdiff[tid] = tid < H ? a.flat[a.g.flat_head + 2 * tid] - a.flat[a.g.flat_head + 2 * tid + 1] : 0.f;
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    if (tid == 0) {
        umma::mbar_expect_tx(&bars[kBWImg], (uint32_t)t.img_bytes);
        for (uint32_t o = 0; o < (uint32_t)t.img_bytes; o += 32768)
            umma::bulk_g2s(smem_b16 + o, a.img + o, min(32768u, (uint32_t)t.img_bytes - o), &bars[kBWImg]);
    }
    const uint32_t tbase = *tmem_slot;
    const uint32_t lane_addr = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    if (warp < 4) {   // zero the gate-gradient regions: the padding columns (and the rows beyond the tile) stay zero
        const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (uint32_t c = 0; c < (uint32_t)kColD; c += 8) umma::tmem_st8(lane_addr + c, z);
        umma::wait_st();
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();

    if (is_row) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        const int part = warp >> 2, rowi = tid & 127;
        const bool live = rowi < M;
        const int m = live ? rowi : 0;
        const float wrow = live ? (float)a.roww[st * M + m] : 0.f;
        float cdir[kUP], dnext[kUP];
#pragma unroll
        for (int j = 0; j < kUP; ++j) { cdir[j] = 0.f; dnext[j] = 0.f; }
        const size_t rb = st * N;                                           // (tile, site 0)
        auto load_dout = [&](int n, float* dst) {                           // d out of site n: from the layer above, or the head (top layer)
            if (!live) return;
            if (a.top) {
                const size_t o = (rb + n) * M + m;
                const float tt = (float)exp(a.la_oth[o]);                   // p_other: d log p_sel / d z_sel = p_other, d / d z_other = -p_other
                const int sg = a.sigT[o];
                const float ts = sg == 0 ? tt : -tt;                        // d out_j = z0 Wd[j][0] + z1 Wd[j][1] with (z0, z1) = (ts, -ts)
#pragma unroll
                for (int j = 0; j < kUP; ++j) dst[j] = ts * wdiff[kPU * part + j];
                if (part == 0) {
                    a.dzbuf[((rb + n) * 2 + 0) * M + m] = wrow * ts;
                    a.dzbuf[((rb + n) * 2 + 1) * M + m] = -wrow * ts;
                }
            } else {
                const float* src = a.dxbuf + ((rb + n) * H + kPU * part) * M + m;
#pragma unroll
                for (int j = 0; j < kUP; ++j) dst[j] = src[(size_t)j * M];
            }
        };
        load_dout(N - 1, dnext);
        const uint32_t acol = lane_addr + 4 * (kPU / 2) * part;             // this thread's 13 unit pairs x 4 gate columns
        const uint32_t dcol = lane_addr + kColD + kPU * part;
        uint32_t phase = 0;
        for (int n = N - 1; n >= -1; --n) {
            float dh[kUP];
            if (n < N - 1) {   // result of the MMAs of site n + 1: carry into site n, d x of site n + 1
                umma::mbar_wait(&bars[kBFull], phase & 1);
                ++phase;
                umma::fence_after_sync();
                float cr[kUP];
#pragma unroll
                for (int gq = 0; gq < 3; ++gq) umma::tmem_ld8p(dcol + 8 * gq, cr + 8 * gq);
                umma::tmem_ld2p(dcol + 24, cr + 24);
                if (l > 0) {
                    float dx[kUP];
#pragma unroll
                    for (int gq = 0; gq < 3; ++gq) umma::tmem_ld8p(dcol + 52 + 8 * gq, dx + 8 * gq);
                    umma::tmem_ld2p(dcol + 52 + 24, dx + 24);
                    umma::wait_ld();
                    if (live) {
                        float* dst = a.dxbuf + ((rb + n + 1) * H + kPU * part) * M + m;        // the layer below reads it as its d out
#pragma unroll
                        for (int j = 0; j < kUP; ++j) dst[(size_t)j * M] = dx[j];
                    }
                } else {
                    umma::wait_ld();
                }
#pragma unroll
                for (int j = 0; j < kUP; ++j) dh[j] = cdir[j] + cr[j];
            } else {
#pragma unroll
                for (int j = 0; j < kUP; ++j) dh[j] = 0.f;
            }
            if (n < 0) break;
#pragma unroll
            for (int j = 0; j < kUP; ++j) dh[j] += dnext[j];
            const int it = N - 1 - n, buf = it & 1;
            umma::mbar_wait(&bars[kBFac0 + buf], (uint32_t)(it >> 1) & 1);                    // the factors of site n have landed
            {   // every lane runs this (tcgen05.st is warp-collective); lanes beyond the tile repeat row 0 and store nothing
                const float* fb0 = reinterpret_cast<const float*>(smem_b16 + t.fac_off + buf * fac_stride);
                const float* f4 = fb0 + 4 * ((size_t)(kPU * part) * M + m);                   // [unit][row][u, alpha, beta, gamma]
                const float* frho = fb0 + 4 * (size_t)H * M + (size_t)(kPU * part) * M + m;  // [unit][row] rho
                float* gout = a.Gbuf + ((rb + n) * 4 * (size_t)H + kPU * part) * M + m;
                const size_t AS = (size_t)H * M;
#pragma unroll
                for (int jl = 0; jl < kUP; jl += 2) {
                    float v[8];                                             // (d a_r, d a_u, d a_c, d aq) of units jl, jl + 1
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float4 f = *reinterpret_cast<const float4*>(f4 + 4 * (size_t)(jl + e) * M);
                        const float d = dh[jl + e];
                        cdir[jl + e] = d * f.x;
                        v[4 + e] = d * f.y;
                        v[2 + e] = d * f.z;
                        v[0 + e] = d * f.w;
                        v[6 + e] = d * frho[(size_t)(jl + e) * M];
                        if (live) {
                            float* go = gout + (size_t)(jl + e) * M;
                            go[0] = wrow * v[0 + e];
                            go[AS] = wrow * v[2 + e];
                            go[2 * AS] = wrow * v[4 + e];
                            go[3 * AS] = wrow * v[6 + e];
                        }
                    }
                    stage4(acol + 4 * (jl / 2), v);
                }
            }
            umma::mbar_arrive(&bars[kBFree0 + buf]);                                          // this buffer may be refilled (site n - 2)
            umma::wait_st();
            umma::fence_before_sync();
            umma::mbar_arrive(&bars[kBStaged]);
            if (n > 0) load_dout(n - 1, dnext);                                               // in flight under the MMAs
        }
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (warp == kMmaWarp) {
            umma::mbar_wait(&bars[kBWImg], 0);
            const uint32_t sB = umma::smem_u32(smem_b16);
            const uint32_t idesc = (1u << 4) | ((uint32_t)(t.n_out >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // F16 x F16 -> F32, M = 128
            const uint64_t bhi = umma::smem_desc(sB, 128, kKC * 128), blo = umma::smem_desc(sB + t.im_bytes, 128, kKC * 128);
            const uint32_t dD = tbase + kColD;
            for (int n = N - 1; n >= 0; --n) {
                umma::mbar_wait(&bars[kBStaged], (uint32_t)(N - 1 - n) & 1);
                umma::fence_after_sync();
#pragma unroll
                for (int q = 0; q < kChunks; ++q) {
                    umma::mma_f16_ts_elect(dD, tbase + q * 8, bhi + (uint64_t)(q * 16), idesc, q > 0);
                    umma::mma_f16_ts_elect(dD, tbase + q * 8, blo + (uint64_t)(q * 16), idesc, 1);
                }
#pragma unroll
                for (int q = 0; q < kChunks; ++q) umma::mma_f16_ts_elect(dD, tbase + kALo + q * 8, bhi + (uint64_t)(q * 16), idesc, 1);
                umma::commit_elect(&bars[kBFull]);
            }
        } else if (warp == kMmaWarp + 1 && (tid & 31) == 0) {
            // producer: one bulk async copy (TMA) per site brings the (tile, site, layer) block of factors into the buffer the row
            // threads released two sites ago
            for (int it = 0; it < N; ++it) {
                const int n = N - 1 - it, buf = it & 1;
                if (it >= 2) umma::mbar_wait(&bars[kBFree0 + buf], (uint32_t)((it >> 1) - 1) & 1);
                unsigned char* dst = smem_b16 + t.fac_off + buf * fac_stride;
                const unsigned char* src = reinterpret_cast<const unsigned char*>(gsrc + (size_t)n * gstep);
                umma::mbar_expect_tx(&bars[kBFac0 + buf], (uint32_t)t.fac_bytes);
                for (uint32_t o = 0; o < (uint32_t)t.fac_bytes; o += 32768)
                    umma::bulk_g2s(dst + o, src + o, min(32768u, (uint32_t)t.fac_bytes - o), &bars[kBFac0 + buf]);
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == kMmaWarp) umma::tmem_dealloc(tbase, 512);
}

// one layer of the backward recurrence: one CTA per tile of M rows
static int launch(const GruLayout& g, int l, int M, int tiles, const float* params, unsigned char* img, const float* gstore,
                  const uint8_t* sigT, const double* la_oth, const double* roww, float* dxbuf, float* Gbuf, float* dzbuf, cudaStream_t s) {
    Args a;
    memset(&a, 0, sizeof(a));
    a.g = g; a.t = make_layout(l, M); a.l = l; a.M = M; a.top = l == g.L - 1;
    a.img = img; a.flat = params; a.gstore = gstore; a.sigT = sigT; a.la_oth = la_oth; a.roww = roww; a.dxbuf = dxbuf; a.Gbuf = Gbuf; a.dzbuf = dzbuf;
    RNNWF_CHECK(M % 4 == 0 && M >= 4 && M <= kMaxM && a.t.smem_bytes <= kSmemLimit, -3, "tensor-core backward kernel: bad tile (%d rows, %d bytes of shared memory)", M, a.t.smem_bytes);
    prof_count(); pack_kernel<<<148, 256, 0, s>>>(g, l, a.t, params, img);
    RNNWF_CUDA(cudaFuncSetAttribute(bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, a.t.smem_bytes));
    prof_count();
    bwd_kernel<<<tiles, kThreads, a.t.smem_bytes, s>>>(a);
    RNNWF_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace tc16b
}  // namespace rnnwf
