// gru_tc16p.cuh — tensor-core chain kernel, generation 3: the tcgen05 3xFP16 recurrence of gru_tc16.cuh, software-pipelined so
// that the tensor pipe and the gate math (MUFU / FMA pipes) run concurrently instead of alternating, with a hot loop small
// enough for the SM's instruction cache.
//
// Same contract as tc16::chain_kernel (TFIM single flips 1DTFIM/TrainingRNN_1DTFIM.py:43-48,74; J1-J2 exchanges
// J1J2/TrainingRNN_J1J2.py:68-92; base pass = teacher-forced log-probability 1DTFIM/RNNwavefunction.py:76-118).
//
// What changed against generation 2 (which serialises MMA -> gate math -> operand restaging per (site, layer) because TMEM has
// no room for a second accumulator set):
//   * the accumulators of one (site, layer) step are split by gate into D_ru = [r | u] and D_c = [cx] [ch], written by two
//     MMA groups M_ru / M_c with their own commit barriers.  The row warps run G_ru (reset / update gates from D_ru) while the
//     tensor pipe is still busy with M_c, then G_c (candidate, new state, restaging).
//   * the (site, layer) steps of a chain are visited along anti-diagonals of the (site, layer) grid, top layer first:
//     ..., (n-2, 2), (n-1, 1), (n, 0), (n-1, 2), (n, 1), (n+1, 0), ...  With three layers no step consumes the output of the
//     step right before it, so M_ru of step k+1 is issued as soon as G_ru of step k has drained D_ru and runs under G_c of
//     step k; M_c of step k+1 runs under G_ru of step k+1.  (With fewer layers the dependent steps wait for the restaged
//     state; M_c still overlaps G_ru.)
//   * all row-warp <-> MMA-warp hand-offs are mbarriers (no CTA-wide bar.sync in the site loop).
//   * ONE copy of the step code serves both unit halves of a row and every layer (measured: with the fully specialised
//     150 KB loop body of the first version the SMs of a GPC ran at the speed of their shared instruction fetch path — the
//     first TPC of each GPC at 4 500 cycles per step, the others at 6 500 - 10 000).  The two row threads of a sample own 26
//     units each (0..25 and 24..49: two units are computed twice), which puts both halves at 4-aligned TMEM columns (tcgen05.ld/st
//     fault on others) with gate blocks only 52 wide; the same instructions work for either half with a runtime offset.
//     (3c: three-layer stacks run one statically specialised copy of the step per layer instead -- see row_chain.)
// Generation 3b: the x operand feeds ONE instruction group [cx | r | u] (N = 160) and the h operand ONE group [r | u | ch]
// (N = 160): a tcgen05.mma with a new A chunk costs ~81 cycles of TMEM operand fetch whatever its N (scripts/mma_probe2.py), so
// 24 wide instructions (tensor-throughput bound) replace 48 narrow ones (~3 300 cycles per step, operand-fetch bound).
// One commit per step; the row warps pull the whole accumulator set into registers first, release it (acc_free) and do all the
// gate math while the tensor pipe already runs the next step.
// TMEM columns: D_cx [0,52) | D_r [52,104) D_u [104,156) | D_ch [156,208) | junk [208,220) |
//               R_l (h^l: 32 hi + 32 lo packed half pairs) at 224 + 64 l | X0 (one-hot input of layer 0) at 224 + 64 L.
//   gate block (52 columns): unit j at column j, columns 50, 51 unused
//   operand region (K = 64):  unit j at k = j, the constant 1 (bias column) at k = 50 (set once per CTA)
// Shared memory: per layer the K-major core-matrix images H_hi | H_lo | X_hi | X_lo, 160 rows each: H = [r(52) | u(52) | ch(52) | 4
// zero rows], X = [cx(52) | r(52) | u(52) | 4 zero rows].  The x group writes D columns [0, 160) (its zero rows add nothing to the
// first ch columns), the h group [52, 212); the instruction that opens the h group is split in two because it accumulates onto
// r, u but must overwrite ch: [r | u] N = 112 (8 stray columns into ch) and then ch N = 64 from B row 104 (a row-group boundary; it
// runs 8 rows into whatever follows the image: junk columns).
// Included by gru.cu.
#pragma once
#include "gru_tc16.cuh"

#ifndef RNNWF_GATES
#define RNNWF_GATES 0
#endif
#ifndef RNNWF_CAND
#define RNNWF_CAND 2
#endif
#ifndef RNNWF_FUSED
#define RNNWF_FUSED 0
#endif
#ifndef RNNWF_UNROLL3
#define RNNWF_UNROLL3 1
#endif
#ifndef RNNWF_LATE_HEAD
#define RNNWF_LATE_HEAD 1
#endif
#ifndef RNNWF_EX2POLY
#define RNNWF_EX2POLY 0   // number of unit pairs (of 13 per row thread) whose candidate 2^a runs on the FMA pipe (ex2_poly2) instead of MUFU
#endif
#ifndef RNNWF_PARTS
#define RNNWF_PARTS 2     // row threads per sample: 2 (26 units each, 8 row warps at 224 registers) or 3 (18 units each, 12 row warps at 152)
#endif
#ifndef RNNWF_SKEW
#define RNNWF_SKEW 0
#endif
#ifndef RNNWF_KPACK
#define RNNWF_KPACK 1     // 3e: the three split passes packed densely along K (10 MMAs of K = 16 per operand group instead of 12)
#endif

namespace rnnwf {
namespace tc16p {

using tc16::ex2;
using tc16::rcp;
using tc16::pack_h2;
using tc16::unpack_h2;
using tc16::core_off;

// packed FP32 pairs (FFMA2 / FADD2 on sm_100a): one issue slot for two lanes of the gate arithmetic
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t f2_make(float a, float b) { f2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void f2_split(f2_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f2_t f2_add(f2_t a, f2_t b) { f2_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2_t f2_mul(f2_t a, f2_t b) { f2_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2_t f2_sub(f2_t a, f2_t b) { f2_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
// (-1/a, -1/b) on the FMA pipe: integer first guess (5 % off) + three Newton steps y <- y + y (1 + x y) in packed FP32 (1 ulp, as
// MUFU.RCP).  8 issue slots instead of two trips through the 4-lane XU pipe; callers absorb the sign.  a, b in [1, 2^126).
__device__ __forceinline__ f2_t f2_fma(f2_t a, f2_t b, f2_t c);
__device__ __forceinline__ f2_t f2_make(float a, float b);
__device__ __forceinline__ f2_t neg_rcp2(float a, float b) {
    const f2_t x = f2_make(a, b), one2 = f2_make(1.0f, 1.0f);
    f2_t y = f2_make(__uint_as_float(0xFEF311C7u - __float_as_uint(a)), __uint_as_float(0xFEF311C7u - __float_as_uint(b)));
#pragma unroll
    for (int it = 0; it < 3; ++it) y = f2_fma(y, f2_fma(x, y, one2), y);
    return y;
}
__device__ __forceinline__ f2_t f2_fma(f2_t a, f2_t b, f2_t c) { f2_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// (2^a0, 2^a1) on the FMA pipe: round to the nearest integer with the 1.5 * 2^23 trick, degree-5 minimax polynomial of 2^f on
// [-0.5, 0.5] in packed FP32 (relative error 2.3e-7, the same as ex2.approx), exponent added to the result's bits.  The XU pipe
// (16 lanes per SM) is the busiest pipe of the chain kernel (ncu: 72 %); each pair moved here frees 16 XU cycles per warp for
// about 12 issue slots.  Inputs are clamped to [-125, 126] (2^126 + 1 and 2^-125 + 1 give the saturated gate values).
__device__ __forceinline__ f2_t ex2_poly2(float a0, float a1) {
    a0 = fminf(fmaxf(a0, -125.f), 126.f);
    a1 = fminf(fmaxf(a1, -125.f), 126.f);
    const f2_t a = f2_make(a0, a1), magic = f2_make(12582912.f, 12582912.f);
    const f2_t t = f2_add(a, magic);
    const f2_t f = f2_sub(a, f2_sub(t, magic));
    f2_t p = f2_fma(f2_make(0x1.5c08ccp-10f, 0x1.5c08ccp-10f), f, f2_make(0x1.3d0c52p-7f, 0x1.3d0c52p-7f));
    p = f2_fma(p, f, f2_make(0x1.c6b6e6p-5f, 0x1.c6b6e6p-5f));
    p = f2_fma(p, f, f2_make(0x1.ebf918p-3f, 0x1.ebf918p-3f));
    p = f2_fma(p, f, f2_make(0x1.62e428p-1f, 0x1.62e428p-1f));
    p = f2_fma(p, f, f2_make(0x1.000002p+0f, 0x1.000002p+0f));
    float p0, p1, t0, t1;
    f2_split(p, p0, p1);
    f2_split(t, t0, t1);
    return f2_make(__uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23)),
                   __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23)));
}

constexpr int kParts = RNNWF_PARTS;                      // row threads per sample (one warp group of 4 warps per part: TMEM lane quadrants)
constexpr int kRows = 128, kRowThreads = 128 * kParts, kThreads = kRowThreads + 128, kMmaWarp = 4 * kParts;   // the 3 warps after the MMA warp only complete its warpgroup (setmaxnreg)
constexpr int kRowRegs = kParts == 2 ? 224 : 152, kMmaRegs = 56;   // kRowThreads * kRowRegs + 128 * kMmaRegs <= 65 536
constexpr int kUP = kParts == 2 ? 26 : 18, kPU = kParts == 2 ? 24 : 16;   // units per row thread; part p owns units [kPU p, kPU p + kUP),
constexpr int kG8 = kUP / 8;                             // full groups of 8 units (4 operand columns); kUP = 8 kG8 + 2
static_assert(kUP == 8 * kG8 + 2 && kPU * (kParts - 1) + kUP == 50 && (kParts == 2 || kParts == 3), "unit partition");
                                                         // units 24 and 25 are computed (identically) by both threads of a sample, so that
                                                         // both halves start at a 4-aligned column and the gate blocks are 52 wide, not 56
constexpr int kBW = 52;                                  // accumulator columns per gate block: unit j at column j, 2 junk columns
constexpr int kNAll = 160, kNRU = 112, kNC = 64;         // N of the merged instructions; of the two that open the h group
constexpr int kRowsImg = 160;                            // stored B rows of an image: three gate blocks + 4 zero rows
constexpr int kKp = 64, kKC = 8;                         // K padded to 4 MMA steps of 16; 16-byte chunks per row
constexpr int kColX = 0, kColCX = 0, kColRU = 52, kColCH = 156, kColR = 224;   // kColX: D base of the x group (cx | r | u)
#if RNNWF_KPACK
// Generation 3e, dense K packing.  Operand region (56 columns = K 112): columns 0..24 hi halves of units 0..49, column 25 a second
// copy of hi(48, 49), column 26 the constant (1, 1), column 27 zero, columns 28..52 lo halves of units 0..49, 53..55 zero.
// Image B1 (K = 64, 8 core-matrix columns c0..c7): c0..c5 = W_hi of units 0..47; c6 = W_hi[48], W_hi[49], W_lo[48], W_lo[49], b_hi,
// b_lo, 0, 0; c7 = a copy of c0.  Image B2 (K = 48): W_lo of units 0..47.  The MMA of operand chunk q (columns 8q..8q+7) takes the
// B1 core columns (0,1) (2,3) (4,5) (6,7) (1,2) (3,4) (5,6) for q = 0..6 -- a descriptor start may sit on any core column -- which
// pairs hi(0..49) with W_hi, the copy of hi(48,49) with W_lo, the constant with the two halves of the bias and lo(0..49) with W_hi;
// chunks 0..2 run a second time against B2: hi(0..47) x W_lo.  7 + 3 = 10 instructions of K = 16 carry the 152 products
// per output that 3 x 64 = 12 carried before.
constexpr int kLoCol = 28, kOneCol = 26, kK2 = 48, kKC2 = 6;
#else
constexpr int kLoCol = 32;
#endif
constexpr int kKOne = 50;                                // K index of the constant-1 (bias) column (unpacked layout)
enum { kFull = 0, kAccFree = 1, kCDone = 2, kWImg = 3, kZDone = 4, kNumBars = 5 };
// kCDone counts ARRIVALS, not steps: in a three-layer stack a fast warp may finish step g + 1 (whose MMAs wait for kAccFree only)
// before a slow warp has finished step g, so its phase g can complete with the slow warp's arrival still missing.  That is harmless
// for its one waiter in the site loop -- the MMA warp waits for the phase of the step right before the one it issues, which no thread
// can over-arrive on -- but a row thread must not read another warp's shared-memory partials behind it: the head partial sums have
// their own barrier kZDone (arrivals: the part >= 1 threads of a top-layer step, after their zsm store; top-layer steps are at least
// two steps apart and the skew between warps is at most one step, so its phases are exact).

__host__ __device__ __forceinline__ int unit_of_col(int c) { return c < 50 ? c : -1; }   // gate-block column -> unit
__host__ __device__ __forceinline__ int unit_of_k(int k) { return k < 50 ? k : -1; }     // operand K index -> unit

struct Layout {
    int L, H, N;
    int im_bytes, im0_bytes;                      // one precision half of an image with K = 64 / K = 16 (layer 0 input)
    int im2_bytes;                                // second image of an operand (KPACK: B2 with K = 48; else the lo half, K = 64)
    int l0_bytes, l1_bytes;
    int tab_off, tab_floats, img_bytes;
};

inline Layout make_layout(const GruLayout& g) {
    Layout t;
    t.L = g.L; t.H = g.H; t.N = g.N;
    t.im_bytes = kRowsImg * kKp * 2;
    t.im0_bytes = kRowsImg * 16 * 2;
#if RNNWF_KPACK
    t.im2_bytes = kRowsImg * kK2 * 2;
#else
    t.im2_bytes = t.im_bytes;
#endif
    t.l0_bytes = t.im_bytes + t.im2_bytes + 2 * t.im0_bytes;
    t.l1_bytes = 2 * (t.im_bytes + t.im2_bytes);
    t.tab_off = t.l0_bytes + (g.L - 1) * t.l1_bytes;
    t.tab_floats = g.nheads * (2 * 64 + 4);       // per head: Wd[64][2] | bd[2] | pad
    t.img_bytes = t.tab_off + t.tab_floats * 4;
    return t;
}

inline size_t smem_bytes(const Layout& t) {
    // image | head partial sums [2][128] float4 | barriers | tmem slot | work slot (the candidate MMAs over-read <= 1 KB past the images)
    return (size_t)((t.img_bytes + 15) & ~15) + (size_t)2 * (kParts - 1) * kRows * sizeof(float4) + 128;
}

inline bool supported(const GruLayout& g) {
    if (!(g.H == 50 && g.N >= 2 && g.L >= 1 && g.L <= 3)) return false;
    return smem_bytes(make_layout(g)) <= (size_t)kSmemLimit;
}
// Narrower stacks (kMinPadH <= H < 50 units) run on the same kernel zero-padded to 50 units: a padded unit has zero weights and
// biases in and out, so its gates are exactly 1/2, its candidate exactly 0 and its state stays exactly 0 (h' = u (h - c) + c);
// nothing of it reaches a real unit.  The tensor work is that of 50 units, which beats the CUDA-core engine (7.8x slower per flop
// at 50 units) down to about (50 / H)^2 = 4.  Only the local-energy paths use this (the gradient's stash keeps the real width).
constexpr int kMinPadH = 26;
inline GruLayout padded_layout(const GruLayout& g) {
    if (g.H >= 50 || g.H < kMinPadH) return g;
    rnnwf_model m;
    memset(&m, 0, sizeof(m));
    m.cell = RNNWF_CELL_GRU; m.head = g.nheads == 2 ? RNNWF_HEAD_COMPLEX : RNNWF_HEAD_PROB; m.dtype = RNNWF_F32;
    m.num_layers = g.L; m.units = 50; m.n_sites = g.N;
    return make_gru_layout(m);
}
inline bool supported_padded(const GruLayout& g) { return supported(padded_layout(g)); }

// flat TF-order parameters -> shared-memory image.  Weights are pre-scaled so that the gates are 1/(1 + 2^a): r, u rows by
// -log2(e), candidate rows by 2 log2(e); the constant-1 K column carries the biases bg (h part of r, u), bch (h part of the
// candidate) and bci (x part of the candidate).  Layer 0: the x images have K = 16 with the two one-hot rows of the input kernels
// at k = 0, 1 and bci at k = 2.
// scaled weight of output row n (gate-block layout) of the h image (xpart = 0: [r | u | ch]) or the x image (xpart = 1: [cx | r | u])
// of layer l for input unit ku, or (bias) the bias that rides on the constant-1 column.  g is the REAL layout: rows / columns of
// units >= g.H (zero padding to 50 units) are zero.
__device__ __forceinline__ float packed_weight(const GruLayout& g, const float* __restrict__ flat, int l, int xpart, int n, int ku, bool bias) {
    const int H = g.H, d = g.d[l];
    const float kS = -1.4426950408889634f, kC = 2.8853900817779268f;
    const float* Kg = flat + g.flat_off[l];
    const float* bg = Kg + (d + H) * 2 * H;
    const float* Kci = bg + 2 * H;
    const float* Kch = Kci + d * H;
    const float* bci = Kch + H * H;
    const float* bch = bci + H;
    const int blk = n / kBW;
    const int j = blk < 3 ? unit_of_col(n % kBW) : -1;       // output unit of this row (-1: padding)
    if (j < 0 || j >= H) return 0.f;
    if (!bias && (ku < 0 || ku >= H)) return 0.f;
    const bool cand = xpart ? blk == 0 : blk == 2;
    const int gate = xpart ? blk - 1 : blk;                  // 0: r, 1: u (unused for the candidate block)
    if (!xpart) {
        if (!cand) return !bias ? kS * Kg[(d + ku) * 2 * H + gate * H + j] : kS * bg[gate * H + j];
        return !bias ? kC * Kch[ku * H + j] : kC * bch[j];
    }
    if (!bias && ku >= d) return 0.f;                        // layer 0: only the two one-hot rows exist
    if (!cand) return !bias ? kS * Kg[ku * 2 * H + gate * H + j] : 0.f;
    return !bias ? kC * Kci[ku * H + j] : kC * bci[j];
}

__global__ void pack_kernel(GruLayout g, Layout t, const float* __restrict__ flat, unsigned char* __restrict__ img) {
    const int H = g.H;
    const int per_l = 2 * kRowsImg * kKp;               // h image + x image
    const int total = g.L * per_l;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int l = idx / per_l;
        int q = idx % per_l;
        const int xpart = q / (kRowsImg * kKp);
        q %= kRowsImg * kKp;
        const int n = q / kKp, k = q % kKp;
        unsigned char* Lb = img + (l == 0 ? 0 : t.l0_bytes + (l - 1) * t.l1_bytes);
        __half* im1 = reinterpret_cast<__half*>(Lb + (xpart ? t.im_bytes + t.im2_bytes : 0));
        if (xpart && l == 0) {                          // one-hot input: K = 16, rows k = 0, 1 of the input kernels, bci at k = 2
            if (k >= 16) continue;
            const float v = k < 2 ? packed_weight(g, flat, 0, 1, n, k, false) : (k == 2 ? packed_weight(g, flat, 0, 1, n, 0, true) : 0.f);
            const __half hi = __float2half_rn(v);
            im1[core_off(n, k, 2)] = hi;
            reinterpret_cast<__half*>(Lb + t.im_bytes + t.im2_bytes + t.im0_bytes)[core_off(n, k, 2)] = __float2half_rn(v - __half2float(hi));
            continue;
        }
        __half* im2 = reinterpret_cast<__half*>(Lb + (xpart ? t.im_bytes + t.im2_bytes : 0) + t.im_bytes);
#if RNNWF_KPACK
        if (k < kK2 + 2) {                               // units 0..49: hi into B1 at k; lo into B2 (units < 48) or B1 at k + 2 (48, 49)
            const float v = packed_weight(g, flat, l, xpart, n, k, false);
            const __half hi = __float2half_rn(v), lo = __float2half_rn(v - __half2float(hi));
            im1[core_off(n, k, kKC)] = hi;
            if (k < kK2) im2[core_off(n, k, kKC2)] = lo;
            else im1[core_off(n, k + 2, kKC)] = lo;
        } else if (k == 52) {                            // the two halves of the bias against the constant (1, 1)
            const float v = packed_weight(g, flat, l, xpart, n, 0, true);
            const __half hi = __float2half_rn(v);
            im1[core_off(n, 52, kKC)] = hi;
            im1[core_off(n, 53, kKC)] = __float2half_rn(v - __half2float(hi));
        } else if (k == 54 || k == 55) {
            im1[core_off(n, k, kKC)] = __float2half_rn(0.f);
        } else if (k >= 56) {                            // c7: copy of c0 (W_hi of units 0..7 against the first lo halves)
            im1[core_off(n, k, kKC)] = __float2half_rn(packed_weight(g, flat, l, xpart, n, k - 56, false));
        }
#else
        const float v = k == kKOne ? packed_weight(g, flat, l, xpart, n, 0, true) : packed_weight(g, flat, l, xpart, n, k < 50 ? k : -1, false);
        const __half hi = __float2half_rn(v);
        im1[core_off(n, k, kKC)] = hi;
        im2[core_off(n, k, kKC)] = __float2half_rn(v - __half2float(hi));
#endif
    }
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < t.tab_floats; idx += gridDim.x * blockDim.x) {
        const int hd = idx / 132, r = idx % 132;        // per head: [part][slot (26)][2] weights of unit 24 part + slot | pad | bd[2] at 128 | pad
        const float* hw = flat + g.flat_head + hd * (2 * H + 2);
        float v = 0.f;
        if (r < 2 * kUP * kParts) {                     // the first two slots of parts >= 1 duplicate the last two units of the part before: zero weights
            const int part = r / (2 * kUP), slot = (r % (2 * kUP)) / 2, o = r & 1;
            if (!(part >= 1 && slot < kUP - kPU) && kPU * part + slot < H) v = hw[2 * (kPU * part + slot) + o];
        }
        else if (r >= 128 && r < 130) v = hw[2 * H + (r - 128)];
        reinterpret_cast<float*>(img + t.tab_off)[idx] = v;
    }
}

// (hi, lo) FP16 pairs of 2 * NC consecutive values -> NC operand-region columns (two values per column).
// DUP (dense K packing, the column of units 48, 49): hi goes to two adjacent columns (the second one meets W_lo of those units).
template <int NC, bool DUP = false> __device__ __forceinline__ void stage_cols(uint32_t col, const float* h) {
    float hi[NC + 1], lo[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const uint32_t wh = pack_h2(h[2 * c], h[2 * c + 1]);
        const float2 f = unpack_h2(wh);
        hi[c] = __uint_as_float(wh);
#if RNNWF_CAND >= 2
        float d0, d1;
        f2_split(f2_sub(f2_make(h[2 * c], h[2 * c + 1]), f2_make(f.x, f.y)), d0, d1);
        lo[c] = __uint_as_float(pack_h2(d0, d1));
#else
        lo[c] = __uint_as_float(pack_h2(h[2 * c] - f.x, h[2 * c + 1] - f.y));
#endif
    }
    if constexpr (NC == 4) {
        umma::tmem_st4(col, hi);
        umma::tmem_st4(col + kLoCol, lo);
    } else {
        static_assert(NC == 1, "column groups are 4 wide or the single tail column");
        if constexpr (DUP) {
            hi[1] = hi[0];
            umma::tmem_st2(col, hi);
        } else {
            umma::tmem_st1(col, hi);
        }
        umma::tmem_st1(col + kLoCol, lo);
    }
}
// the 13th column of a row thread: units 24, 25 (part 0) or 48, 49 (part 1; with dense K packing hi is stored twice)
__device__ __forceinline__ void stage_tail(uint32_t col, const float* h, int part) {
#if RNNWF_KPACK
    if (part == kParts - 1) stage_cols<1, true>(col, h);
    else
#endif
        stage_cols<1>(col, h);
}
// the 26 units of a row thread -> columns [12 part, 12 part + 13) of an operand region
__device__ __forceinline__ void stage_all(uint32_t reg_part_addr, const float* hp, int part) {
#pragma unroll
    for (int gq = 0; gq < kG8; ++gq) stage_cols<4>(reg_part_addr + 4 * gq, hp + 8 * gq);
    stage_tail(reg_part_addr + 4 * kG8, hp + 8 * kG8, part);
}

struct Args {
    GruLayout g;
    Layout t;
    int Mold, tiles128;
    int64_t rows_total;
    const unsigned char* img;
    const uint8_t* sigT;
    float* hstore;            // BASE: written (every layer, every site); FLIP: restart states
    double *la_sel, *la_oth;  // BASE: written; FLIP: read (at the modified site only)
    float* gstore;            // BASE, optional: backward factors [tile][site][layer]{[unit][M][4], [unit][M]} for the tensor-core BPTT (gru_tc16b.cuh)
    float* la_self;           // FP32 copy of la_sel (the values are FP32 numbers): what the flip chains subtract site by site
    double* lp;               // BASE: sum_n la_sel
    double* delta;            // FLIP: [tile][slot][M]
    int* counter;
    // complex cRNN / J1-J2 exchanges (CPLX instantiations): phases, imaginary parts, slot plan
    double *ph_sel, *ph_oth, *lp_im, *delta_im;
    uint8_t* sampT;           // SAMPLE: drawn configurations [tile][site][128]
    uint64_t seed, sample_offset;   // SAMPLE: Philox key, global id of row 0
    const int* order;         // slots by decreasing chain length
    const double *j1, *j2;    // couplings: slots with a zero coupling are skipped (J1J2/TrainingRNN_J1J2.py:69,84)
    int n_kind1, n_kind2, nslots;
};

// per-row-thread state of one chain
struct Ctx {
    const float* tab;
    float4* zsm;
    uint64_t* bars;
    uint32_t lane_addr;
    int rowi, m, part;
    bool live;
    size_t rowbase;
    int s, t;                 // modified sites (s = -1: none)
    uint32_t g, cda, zc;      // steps done by this CTA so far (phase index of full_ru / full_c / ru_free); next c_done phase index; top-layer steps so far (z_done phase index)
    float4 pz;                // pending head (part 0): partial sums, selected outcome, site, c_done phase and zsm buffer of that step
    int psg, pn;
    uint32_t pph, pbuf;
    double p_la, p_ph;        // base-pass la_sel / ph_sel of the pending site (loaded a step ahead: global latency off the critical path)
    float p_laf, accf, compf; // probability head: the site terms are FP32 numbers, summed with Kahan compensation in FP32 -- FP64
                              // instructions in the site loop cost ~150 cycles each on this part (ncu: stall_math on every DADD)
    int nup;
    bool skew;
    double acc, acc_im;
#ifdef RNNWF_TC16P_DEBUG
    long long w_ru, w_c, t_ru, t_c;   // cycles: waiting for full_ru / full_c, inside G_ru / G_c
#endif
};
#ifdef RNNWF_TC16P_DEBUG
#define TCP_T(...) __VA_ARGS__
#else
#define TCP_T(...)
#endif

template <bool BASE> __device__ __forceinline__ int spin_of(const Args& a, const Ctx& c, int q) {
    int v = a.sigT[(c.rowbase + q) * a.Mold + c.m];
    if (!BASE && (q == c.s || q == c.t)) v = 1 - v;
    return v;
}

// log-softmax / amplitude / phase of the pending top-layer step (part 0 only); part 1's partial sums come through zsm
template <bool BASE, bool CPLX> __device__ __forceinline__ void finish_head(const Args& a, Ctx& c) {
    if (c.pn < 0) return;
    const int pn = c.pn, psg = c.psg, N = a.g.N;
    c.pn = -1;
    umma::mbar_wait(&c.bars[kZDone], c.pph & 1);      // part 1's z_done arrival of that step (acquire of its zsm store)
    if (!c.live) return;
    float4 pz = c.pz;
#pragma unroll
    for (int q = 0; q < kParts - 1; ++q) {
        const float4 o = c.zsm[(c.pbuf * (kParts - 1) + q) * kRows + c.rowi];
        pz.x += o.x; pz.y += o.y; pz.z += o.z; pz.w += o.w;
    }
    const float* tab = c.tab;
    const float f0 = pz.x + tab[128], f1 = pz.y + tab[129];
    // log softmax of the 2-way head in FP32 (log1pf/expf, ~1e-7 relative); the site terms are summed in FP64
    const float dsel = psg ? f0 - f1 : f1 - f0;                             // z_other - z_selected
    double ls = dsel > 30.f ? -(double)dsel : -(double)log1pf(expf(dsel));
    double lo = -dsel > 30.f ? (double)dsel : -(double)log1pf(expf(-dsel));
    double ps = 0.0, po = 0.0;
    if (CPLX) {
        // amplitude = sqrt(softmax) with the zero-magnetisation mask and renormalisation
        // (J1J2/ComplexRNNwavefunction.py:85-93,147-155); phase = pi * softsign (:8-9)
        ls *= 0.5;
        lo *= 0.5;
        if (2 * pn >= N) {
            const int half = N / 2, ndn = pn - c.nup;
            const bool ok_dn = (half - 1 - ndn) >= 0, ok_up = (half - 1 - c.nup) >= 0;
            const bool ok_sel = psg ? ok_up : ok_dn, ok_oth = psg ? ok_dn : ok_up;
            const double ninf = -__longlong_as_double(0x7ff0000000000000LL);
            if (!ok_sel) ls = ninf; else if (!ok_oth) ls = 0.0;
            if (!ok_oth) lo = ninf; else if (!ok_sel) lo = 0.0;
        }
        const float y0 = pz.z + tab[132 + 128], y1 = pz.w + tab[132 + 129];
        const double ys = psg ? (double)y1 : (double)y0, yo = psg ? (double)y0 : (double)y1;
        ps = kPi * ys / (1.0 + fabs(ys));
        po = kPi * yo / (1.0 + fabs(yo));
        c.nup += psg;
    }
    const size_t o_ = (c.rowbase + pn) * a.Mold + c.m;
    if (BASE) {
        if (a.la_sel != nullptr) {             // log psi alone (rnnwf_logpsi) keeps no per-site terms
            a.la_sel[o_] = ls;
            a.la_oth[o_] = lo;
            if (CPLX) { a.ph_sel[o_] = ps; a.ph_oth[o_] = po; }
        }
        c.acc += ls;
        if (CPLX) c.acc_im += ps;
    } else {
        c.acc += ls - c.p_la;
        if (CPLX) c.acc_im += ps - c.p_ph;
    }
}

// probability head without FP64: log-softmax term of the pending top-layer step in FP32 (log1pf / expf, ~1e-7 relative, the same
// numbers finish_head casts to double), accumulated with Kahan compensation.  FLIP: term - base term, both FP32 numbers of nearly
// equal magnitude, so the difference is (almost always) exact.
// LATE: called two steps after the top-layer step instead of one.  No wait is needed: this thread has passed the commit barrier of
// the current step, whose MMAs were issued after every row thread had released the accumulators of the step before (kAccFree cannot
// be over-arrived: a thread drains step g + 1 only after phase g has completed), i.e. after every thread had finished the top-layer
// step (the barrier chain carries the acquire of part 1's zsm store).
template <bool BASE, bool LATE = false> __device__ __forceinline__ void finish_head_f32(const Args& a, Ctx& c) {
    if (c.pn < 0) return;
    const int pn = c.pn, psg = c.psg;
    c.pn = -1;
    if (!LATE) umma::mbar_wait(&c.bars[kZDone], c.pph & 1);      // part 1's z_done arrival of that step (acquire of its zsm store)
    if (!c.live) return;
    float4 o = c.zsm[c.pbuf * (kParts - 1) * kRows + c.rowi];
    if constexpr (kParts == 3) {
        const float4 o2 = c.zsm[(c.pbuf * (kParts - 1) + 1) * kRows + c.rowi];
        o.x += o2.x; o.y += o2.y;
    }
    const float f0 = c.pz.x + o.x + c.tab[128], f1 = c.pz.y + o.y + c.tab[129];
    const float dsel = psg ? f0 - f1 : f1 - f0;                             // z_other - z_selected
    const float ls = dsel > 30.f ? -dsel : -log1pf(expf(dsel));
    float term = ls;
    if (BASE) {
        if (a.la_sel != nullptr) {             // log psi alone (rnnwf_logpsi) keeps no per-site terms
            const float lo = -dsel > 30.f ? dsel : -log1pf(expf(-dsel));
            const size_t o_ = (c.rowbase + pn) * a.Mold + c.m;
            a.la_sel[o_] = (double)ls;
            a.la_oth[o_] = (double)lo;
            a.la_self[o_] = ls;
        }
    } else {
        term = ls - c.p_laf;
    }
    const float y = term - c.compf, t = c.accf + y;
    c.compf = (t - c.accf) - y;
    c.accf = t;
}

// What the backward recurrence of one unit needs from the forward pass, as five factors of d h (gru_tc16b.cuh):
//   u: d h_{n-1} += d h * u;   alpha = (1-u)(1-c^2): d a_c = d h * alpha;   beta = (h_prev - c) u (1-u): d a_u = d h * beta;
//   gamma = alpha * aq * r (1-r): d a_r = d h * gamma;   rho = alpha * r: d aq = d h * rho         (aq = h Kch + bch = dq / (2 log2 e))
// A (tile, site, layer) block of the stash is [unit][row][u, alpha, beta, gamma] followed by [unit][row] rho: a row thread writes a
// unit with one 16-byte and one 4-byte store, a warp's 32 rows are 512 + 128 contiguous bytes (five 4-byte stores per unit into
// [factor][unit][row] cost 21 ms per stash pass at cfg2, 8-byte stores into [factor][row][unit] 50 ms: 32 sectors per instruction).
__device__ __forceinline__ void store_bwd_factors(float* g4, float* grho, float u, float r, float c, float dq, float hprev) {
    const float aq = dq * 0.34657359027997264f;            // 1 / (2 log2 e): the candidate rows of the images are pre-scaled
    const float al = (1.0f - u) * (1.0f - c * c);
    *reinterpret_cast<float4*>(g4) = make_float4(u, al, (hprev - c) * u * (1.0f - u), al * aq * r * (1.0f - r));
    *grho = al * r;
}

// reset / update gates 1/(1 + 2^a) of two units, in place.  RNNWF_GATES selects how many reciprocals are shared (MUFU pipe against
// issue slots): 0: r0, u0, r1, u1 share ONE reciprocal (exponents clamped to 30, so the product of the four denominators stays
// below 2^121): 2.5 MUFU and 11 instructions per unit;  1: one reciprocal per gate (ex2(+big) = inf -> 1/inf = 0, no clamps): 4 MUFU,
// 6 instructions;  2: r and u of a unit share a reciprocal: 3 MUFU, 10 instructions;  3: as 1 with packed additions
__device__ __forceinline__ void ru_one(float& r, float& u) {
#if RNNWF_GATES == 1 || RNNWF_GATES == 3
    r = rcp(1.0f + ex2(r));
    u = rcp(1.0f + ex2(u));
#else
    const float er = 1.0f + ex2(fminf(r, 60.f)), eu = 1.0f + ex2(fminf(u, 60.f));
    const float inv = rcp(er * eu);
    r = inv * eu; u = inv * er;
#endif
}
__device__ __forceinline__ void ru_pair(float& r0, float& u0, float& r1, float& u1) {
#if RNNWF_GATES == 0
#ifdef RNNWF_NOCLAMP   // measurement only: valid when every reset / update pre-activation is known to stay below 30 log 2
    const float er0 = 1.0f + ex2(r0), eu0 = 1.0f + ex2(u0);
    const float er1 = 1.0f + ex2(r1), eu1 = 1.0f + ex2(u1);
#else
    const float er0 = 1.0f + ex2(fminf(r0, 30.f)), eu0 = 1.0f + ex2(fminf(u0, 30.f));
    const float er1 = 1.0f + ex2(fminf(r1, 30.f)), eu1 = 1.0f + ex2(fminf(u1, 30.f));
#endif
    const float p0 = er0 * eu0, p1 = er1 * eu1;
    const float inv = rcp(p0 * p1);
    const float i0 = inv * p1, i1 = inv * p0;                          // 1/p0, 1/p1
    r0 = i0 * eu0; u0 = i0 * er0;
    r1 = i1 * eu1; u1 = i1 * er1;
#elif RNNWF_GATES == 4   // as 0 with the additions and products in packed FP32: 16 instead of 22 instructions per unit pair
    const f2_t one2 = f2_make(1.0f, 1.0f);
    const f2_t er = f2_add(f2_make(ex2(fminf(r0, 30.f)), ex2(fminf(r1, 30.f))), one2);
    const f2_t eu = f2_add(f2_make(ex2(fminf(u0, 30.f)), ex2(fminf(u1, 30.f))), one2);
    float p0, p1;
    f2_split(f2_mul(er, eu), p0, p1);
    const float inv = rcp(p0 * p1);
    const f2_t i01 = f2_mul(f2_make(inv, inv), f2_make(p1, p0));       // (1/p0, 1/p1)
    f2_split(f2_mul(i01, eu), r0, r1);
    f2_split(f2_mul(i01, er), u0, u1);
#elif RNNWF_GATES == 3
    const f2_t one2 = f2_make(1.0f, 1.0f);
    float a0, a1;
    f2_split(f2_add(f2_make(ex2(r0), ex2(r1)), one2), a0, a1);
    r0 = rcp(a0); r1 = rcp(a1);
    f2_split(f2_add(f2_make(ex2(u0), ex2(u1)), one2), a0, a1);
    u0 = rcp(a0); u1 = rcp(a1);
#else
    ru_one(r0, u0);
    ru_one(r1, u1);
#endif
}

// one (site n, layer l) step of a row thread: pull the step's accumulators out of TMEM, release them to the MMA warp, then
// reset / update gates, candidate, new state, head partial sums and restaging from registers.
// hp: this thread's 25 units of h^l (previous site in, this site out).
// SAMPLE (with BASE): the autoregressive sampler -- no teacher forcing, nothing stashed; the top-layer step ends with the draw of
// sigma_n (part 0: head, Philox uniform keyed by (global sample id, site), one-hot input of site n + 1), see row_chain_sample.
template <bool BASE, bool CPLX, int LS = -1, bool SAMPLE = false>
__device__ __forceinline__ void row_step(const Args& a, Ctx& c, int n, int l_dyn, float* hp) {
    constexpr int H = 50;
    const int L = LS >= 0 ? 3 : a.g.L, l = LS >= 0 ? LS : l_dyn;      // LS >= 0: the statically specialised copies of a 3-layer stack
    const int N = a.g.N, Mold = a.Mold, part = c.part;
    const bool top = l == L - 1;
    const uint32_t par = c.g & 1;
    const uint32_t dpart = c.lane_addr + kPU * part;                      // this thread's columns inside a gate block
    // global loads this step will need at its end are issued before the wait (part 0): the spin of site n (one-hot input of
    // (n + 1, 0) / selected outcome of the head) and the base-pass terms of site n
    int spin_n = 0;
    double la_n = 0.0, ph_n = 0.0;
    float la_nf = 0.f;
    // (part 0 finishes the head, part 1 stages the one-hot input: the two warps of an SM sub-partition carry similar extra work)
    float u_n = 0.f;
    if (SAMPLE && top && part == 0) u_n = philox_uniform(a.seed, a.sample_offset + (uint64_t)(c.rowbase / (size_t)N) * Mold + c.m, (uint32_t)n);   // global sample id: the draws do not depend on the sharding
    if (!SAMPLE && c.live && ((part == 0 && top) || (part == 1 && l == 0))) {
        spin_n = a.sigT[(c.rowbase + n) * Mold + c.m];   // raw: the flip of a modified site is applied where the value is used, after
                                                         // the gate math -- nothing before the accumulator drain may wait on this load
        if (!BASE && top && part == 0) {
            if (CPLX) {
                la_n = a.la_sel[(c.rowbase + n) * Mold + c.m];
                ph_n = a.ph_sel[(c.rowbase + n) * Mold + c.m];
            } else {
                la_nf = a.la_self[(c.rowbase + n) * Mold + c.m];
            }
        }
    }
    TCP_T(long long t0 = clock64();)
    umma::mbar_wait(&c.bars[kFull], par);
#if RNNWF_SKEW > 0   // experiment: the part >= 1 warps start a chain RNNWF_SKEW cycles behind the part 0 warps they share a scheduler with
    if (part != 0 && c.skew) {
        const long long ts = clock64();
        while (clock64() - ts < RNNWF_SKEW) {}
        c.skew = false;
    }
#endif
    umma::fence_after_sync();
    TCP_T(long long t1 = clock64(); c.w_ru += t1 - t0;)
    float rr[kUP], uu[kUP], dc[kUP], dq[kUP];
#pragma unroll
    for (int gq = 0; gq < kG8; ++gq) {
        umma::tmem_ld8p(dpart + kColRU + 8 * gq, rr + 8 * gq);
        umma::tmem_ld8p(dpart + kColRU + kBW + 8 * gq, uu + 8 * gq);
    }
    umma::tmem_ld2p(dpart + kColRU + 8 * kG8, rr + 8 * kG8);
    umma::tmem_ld2p(dpart + kColRU + kBW + 8 * kG8, uu + 8 * kG8);
#pragma unroll
    for (int gq = 0; gq < kG8; ++gq) {
        umma::tmem_ld8p(dpart + kColCX + 8 * gq, dc + 8 * gq);
        umma::tmem_ld8p(dpart + kColCH + 8 * gq, dq + 8 * gq);
    }
    umma::tmem_ld2p(dpart + kColCX + 8 * kG8, dc + 8 * kG8);
    umma::tmem_ld2p(dpart + kColCH + 8 * kG8, dq + 8 * kG8);
    umma::wait_ld();
    umma::fence_before_sync();
    umma::mbar_arrive(&c.bars[kAccFree]);              // the accumulators may be overwritten by the next step's MMAs
    TCP_T(long long t2 = clock64(); c.w_c += t2 - t1;)
#ifdef RNNWF_ABL_NOGATES   // ablation (timing / power only, results are garbage): the tensor pipe alone -- drain, release, hand back
    if (top) {
        if (part != 0) umma::mbar_arrive(&c.bars[kZDone]);
        ++c.zc;
    }
    hp[0] += rr[0] + uu[1] + dc[2] + dq[3];
    umma::fence_before_sync();
    umma::mbar_arrive(&c.bars[kCDone]);
    ++c.g;
    ++c.cda;
    return;
#endif
    if (!SAMPLE && part == 0) {
        // the pending head of the last top-layer step.  Three-layer copies: finished in the bottom-layer step, the shortest of the
        // three and the one the top layer's MMAs run under (ncu: the top-layer copy waited 11 % of its time at the commit barrier);
        // the middle-layer step only takes it when no bottom-layer step follows on this anti-diagonal (the last two of a chain)
        if constexpr (CPLX) finish_head<BASE, CPLX>(a, c);
        else if constexpr (LS == 0 && RNNWF_LATE_HEAD) finish_head_f32<BASE, true>(a, c);
        else if constexpr (LS == 1 && RNNWF_LATE_HEAD) { if (n + 1 >= N) finish_head_f32<BASE>(a, c); }
        else finish_head_f32<BASE>(a, c);
    }
    // ---- reset / update gates (see ru_pair).  RNNWF_FUSED: the gates of a unit pair are computed right before its candidate, so
    // that the MUFU-heavy head of one pair and the MUFU-free tail (state update, FP16 split, restaging) of the previous ones sit in
    // the same scheduling window: both warps of an SM sub-partition run the same phase at the same time, and a phase that is all
    // ex2 / rcp is bound by the 4-lane XU pipe while the issue slots idle
#if !RNNWF_FUSED
#pragma unroll
    for (int q = 0; q < kUP; q += 2) ru_pair(rr[q], uu[q], rr[q + 1], uu[q + 1]);
#endif
    TCP_T(long long t3 = clock64(); c.t_ru += t3 - t2;)
    // ---- candidate, new state, head partial sums, restaging
    const uint32_t reg = c.lane_addr + kColR + 64 * l + (kPU / 2) * part;
    const float* tab = c.tab + 2 * kUP * part;
    float* hst = BASE && !SAMPLE ? a.hstore + (((c.rowbase + n) * L + l) * (size_t)H + kPU * part) * Mold + c.m : nullptr;
    float y0 = 0.f, y1 = 0.f;
    f2_t z01 = f2_make(0.f, 0.f);                      // head partial sums (z0, z1), one packed FMA per unit
#pragma unroll
    for (int gq = 0; gq <= kG8; ++gq) {                // the column groups of 4 unit pairs and the last pair
#pragma unroll
        for (int q = 0; q < (gq < kG8 ? 8 : 2); q += 2) {
            const int jl = 8 * gq + q;
#if RNNWF_FUSED
            ru_pair(rr[jl], uu[jl], rr[jl + 1], uu[jl + 1]);
#endif
            // candidate tanh(x) = 1 - 2 / (1 + 2^a), a = 2 log2(e) x.  RNNWF_CAND 0: two units share a reciprocal, 1: one each
#if RNNWF_CAND == 0
            const float ec0 = 1.0f + ex2(fminf(fmaf(rr[jl], dq[jl], dc[jl]), 60.f));
            const float ec1 = 1.0f + ex2(fminf(fmaf(rr[jl + 1], dq[jl + 1], dc[jl + 1]), 60.f));
            const float ic = rcp(ec0 * ec1);
            const float c0 = fmaf(-2.0f, ic * ec1, 1.0f), c1 = fmaf(-2.0f, ic * ec0, 1.0f);
#elif RNNWF_CAND == 1
            const float c0 = fmaf(-2.0f, rcp(1.0f + ex2(fmaf(rr[jl], dq[jl], dc[jl]))), 1.0f);
            const float c1 = fmaf(-2.0f, rcp(1.0f + ex2(fmaf(rr[jl + 1], dq[jl + 1], dc[jl + 1]))), 1.0f);
#endif
#if RNNWF_CAND < 2
            const float h0 = fmaf(uu[jl], hp[jl] - c0, c0), h1 = fmaf(uu[jl + 1], hp[jl + 1] - c1, c1);
#else       // packed: pre-activation, 1 + 2^a, 1 - 2/(.), h - c and the state update are one instruction per unit pair each
            float a0, a1, h0, h1;
            f2_split(f2_fma(f2_make(rr[jl], rr[jl + 1]), f2_make(dq[jl], dq[jl + 1]), f2_make(dc[jl], dc[jl + 1])), a0, a1);
#if RNNWF_CAND == 2
            if (jl / 2 < RNNWF_EX2POLY) f2_split(f2_add(ex2_poly2(a0, a1), f2_make(1.0f, 1.0f)), a0, a1);
            else f2_split(f2_add(f2_make(ex2(a0), ex2(a1)), f2_make(1.0f, 1.0f)), a0, a1);
            const f2_t cc = f2_fma(f2_make(-2.0f, -2.0f), f2_make(rcp(a0), rcp(a1)), f2_make(1.0f, 1.0f));
#elif RNNWF_CAND == 3   // the reciprocals on the FMA pipe
            f2_split(f2_add(f2_make(ex2(fminf(a0, 60.f)), ex2(fminf(a1, 60.f))), f2_make(1.0f, 1.0f)), a0, a1);
            const f2_t cc = f2_fma(f2_make(2.0f, 2.0f), neg_rcp2(a0, a1), f2_make(1.0f, 1.0f));
#else                   // 4: one MUFU reciprocal for the two units
            f2_split(f2_add(f2_make(ex2(fminf(a0, 60.f)), ex2(fminf(a1, 60.f))), f2_make(1.0f, 1.0f)), a0, a1);
            const float ic = rcp(a0 * a1);
            const f2_t cc = f2_fma(f2_make(-2.0f, -2.0f), f2_make(ic * a1, ic * a0), f2_make(1.0f, 1.0f));
#endif
            f2_split(f2_fma(f2_make(uu[jl], uu[jl + 1]), f2_sub(f2_make(hp[jl], hp[jl + 1]), cc), cc), h0, h1);
#endif
            if constexpr (BASE && !SAMPLE) {
                if (a.gstore != nullptr && c.live) {   // gradient's stash pass: the factors of the backward recurrence (hp still holds h_prev)
                    float cs0, cs1;
#if RNNWF_CAND < 2
                    cs0 = c0; cs1 = c1;
#else
                    f2_split(cc, cs0, cs1);
#endif
                    float* blk = a.gstore + ((c.rowbase + n) * L + l) * 5 * (size_t)H * Mold;          // [unit][row][4] | [unit][row]
                    const size_t ur = (size_t)(kPU * part + jl) * Mold + c.m;
                    store_bwd_factors(blk + 4 * ur, blk + 4 * (size_t)H * Mold + ur, uu[jl], rr[jl], cs0, dq[jl], hp[jl]);
                    store_bwd_factors(blk + 4 * (ur + Mold), blk + 4 * (size_t)H * Mold + ur + Mold, uu[jl + 1], rr[jl + 1], cs1, dq[jl + 1], hp[jl + 1]);
                }
            }
            hp[jl] = h0;
            hp[jl + 1] = h1;
            if (top) {
                const float2 w0 = *reinterpret_cast<const float2*>(tab + 2 * jl), w1 = *reinterpret_cast<const float2*>(tab + 2 * jl + 2);
                z01 = f2_fma(f2_make(h0, h0), f2_make(w0.x, w0.y), z01);
                z01 = f2_fma(f2_make(h1, h1), f2_make(w1.x, w1.y), z01);
                if (CPLX) {
                    y0 = fmaf(h0, tab[132 + 2 * jl], y0);
                    y1 = fmaf(h0, tab[132 + 2 * jl + 1], y1);
                    y0 = fmaf(h1, tab[132 + 2 * jl + 2], y0);
                    y1 = fmaf(h1, tab[132 + 2 * jl + 3], y1);
                }
            }
            if (BASE && !SAMPLE && c.live && a.hstore != nullptr) {
                hst[(size_t)jl * Mold] = h0;
                hst[(size_t)(jl + 1) * Mold] = h1;
            }
        }
        // region l takes the new state: it is the h operand of (n + 1, l) and the x operand of (n, l + 1), both visited later
        if (gq < kG8) stage_cols<4>(reg + 4 * gq, hp + 8 * gq);
        else stage_tail(reg + 4 * kG8, hp + 8 * kG8, part);
    }
    if (!BASE && (n == c.s || n == c.t)) spin_n = 1 - spin_n;
    if (!SAMPLE && part == 1 && l == 0 && n + 1 < N) {            // one-hot input of (n + 1, 0): the spin of site n (M(n, 0) has completed)
        const int code = c.live ? spin_n : 2;
        const float oh[1] = {__uint_as_float(pack_h2(code == 0 ? 1.f : 0.f, code == 1 ? 1.f : 0.f))};
        umma::tmem_st1(c.lane_addr + kColR + 64 * L, oh);
    }
    if (top) {   // partial head sums; part 0 finishes the log-softmax at its next step (or after the chain)
        float z0, z1;
        f2_split(z01, z0, z1);
        if constexpr (SAMPLE) {
            // draw sigma_n (1DTFIM/RNNwavefunction.py:65-70; U(1) mask J1J2/ComplexRNNwavefunction.py:85-95): P(0) of the 2-way softmax
            // in FP32, sigma = (u >= P(0)), the convention of gru_sample_kernel and of the oracle
            if (part != 0) {
                c.zsm[(par * (kParts - 1) + (part - 1)) * kRows + c.rowi] = make_float4(z0, z1, y0, y1);
                umma::mbar_arrive(&c.bars[kZDone]);
            } else {
                umma::mbar_wait(&c.bars[kZDone], c.zc & 1);
#pragma unroll
                for (int q = 0; q < kParts - 1; ++q) {
                    const float4 o = c.zsm[(par * (kParts - 1) + q) * kRows + c.rowi];
                    z0 += o.x; z1 += o.y;
                }
                z0 += c.tab[128]; z1 += c.tab[129];
                const float p0 = 1.0f / (1.0f + expf(z1 - z0));
                int sg = u_n >= p0 ? 1 : 0;
                if (CPLX && 2 * n >= N) {
                    const int half = N / 2, ndn = n - c.nup;
                    const bool ok_dn = (half - 1 - ndn) >= 0, ok_up = (half - 1 - c.nup) >= 0;
                    if (!ok_up) sg = 0;
                    else if (!ok_dn) sg = 1;
                }
                c.nup += sg;
                a.sampT[(c.rowbase + n) * Mold + c.m] = (uint8_t)sg;
                if (n + 1 < N) {
                    const float oh[1] = {__uint_as_float(pack_h2(sg == 0 ? 1.f : 0.f, sg == 1 ? 1.f : 0.f))};
                    umma::tmem_st1(c.lane_addr + kColR + 64 * L, oh);
                }
            }
            ++c.zc;
        } else if (part == 0) {
            c.pz = make_float4(z0, z1, y0, y1);
            c.pn = n;
            c.psg = spin_n;
            c.p_la = la_n;
            c.p_ph = ph_n;
            c.p_laf = la_nf;
            c.pph = c.zc;
            c.pbuf = par;
        } else {
            c.zsm[(par * (kParts - 1) + (part - 1)) * kRows + c.rowi] = make_float4(z0, z1, y0, y1);
            umma::mbar_arrive(&c.bars[kZDone]);
        }
        if constexpr (!SAMPLE) ++c.zc;
    }
    umma::wait_st();
    umma::fence_before_sync();
    umma::mbar_arrive(&c.bars[kCDone]);
    TCP_T(c.t_c += clock64() - t3;)
    ++c.g;
    ++c.cda;
}

// all row-thread work of one chain.  kind 0: sigma with site s flipped (TFIM); kind 1 / 2: sites s and t = s + kind exchanged
// (J1-J2); BASE: the unmodified configuration from site 0 (s = -1).  Requires s + 1 < N.
template <bool BASE, bool CPLX>
__device__ __forceinline__ void row_chain(const Args& a, Ctx& c) {
    constexpr int H = 50;
    const int L = a.g.L, N = a.g.N, Mold = a.Mold, s = c.s, part = c.part;
    // hidden states of this thread's 25 units: hA is the layer visited next, hB the one after it, hC the third (rotated by moves)
    float hA[kUP], hB[kUP], hC[kUP];
#pragma unroll
    for (int j = 0; j < kUP; ++j) { hA[j] = 0.f; hB[j] = 0.f; hC[j] = 0.f; }
    if (!BASE && c.live) {   // restart from the base states after site s: hA <- top layer, hB <- the layer below, ...
        const float* src = a.hstore + ((c.rowbase + s) * L * (size_t)H + kPU * part) * Mold + c.m;
#pragma unroll
        for (int j = 0; j < kUP; ++j) {
            hA[j] = src[((size_t)(L - 1) * H + j) * Mold];
            if (L > 1) hB[j] = src[((size_t)(L - 2) * H + j) * Mold];
            if (L > 2) hC[j] = src[(size_t)j * Mold];
        }
    }
    const uint32_t regp = c.lane_addr + kColR + (kPU / 2) * part;
    stage_all(regp + 64 * (L - 1), hA, part);
    if (L > 1) stage_all(regp + 64 * (L - 2), hB, part);
    if (L > 2) stage_all(regp, hC, part);
    c.pn = -1;
    c.nup = 0;
    if (part == 0) {
        const int code = (c.live && s >= 0) ? spin_of<BASE>(a, c, s) : 2;   // input of site s + 1 (the zero vector at site 0)
        if (CPLX && c.live && !BASE) {                                      // up spins among the sites before s + 1
            for (int q = 0; q < s; ++q) c.nup += a.sigT[(c.rowbase + q) * Mold + c.m];
            c.nup += code;
        }
        const float oh[1] = {__uint_as_float(pack_h2(code == 0 ? 1.f : 0.f, code == 1 ? 1.f : 0.f))};
        umma::tmem_st1(c.lane_addr + kColR + 64 * L, oh);
    }
    umma::wait_st();
    umma::fence_before_sync();
    umma::mbar_arrive(&c.bars[kCDone]);                // "step -1": operands staged
    ++c.cda;
    const int n0 = s + 1;
#if RNNWF_UNROLL3
    if (L == 3) {   // one specialised copy of the step per layer: no register rotation, head / one-hot code only where it runs
#pragma unroll 1
        for (int d = n0; d <= N + 1; ++d) {
            if (d - 2 >= n0 && d - 2 < N) row_step<BASE, CPLX, 2>(a, c, d - 2, 2, hA);
            if (d - 1 >= n0 && d - 1 < N) row_step<BASE, CPLX, 1>(a, c, d - 1, 1, hB);
            if (d < N) row_step<BASE, CPLX, 0>(a, c, d, 0, hC);
        }
    } else
#endif
#pragma unroll 1
    for (int d = n0; d <= N - 1 + L - 1; ++d) {        // anti-diagonals, top layer first
#pragma unroll 1
        for (int l = L - 1; l >= 0; --l) {
            const int n = d - l;
            if (n >= n0 && n < N) row_step<BASE, CPLX>(a, c, n, l, hA);
            if (L == 3) {                              // (hA, hB, hC) <- (hB, hC, hA)
#pragma unroll
                for (int j = 0; j < kUP; ++j) { const float tmp = hA[j]; hA[j] = hB[j]; hB[j] = hC[j]; hC[j] = tmp; }
            } else if (L == 2) {
#pragma unroll
                for (int j = 0; j < kUP; ++j) { const float tmp = hA[j]; hA[j] = hB[j]; hB[j] = tmp; }
            }
        }
    }
    asm volatile("bar.sync 2, %0;" ::"n"(kRowThreads) : "memory");   // part 1's partials of the last site
    if (part == 0) {
        if constexpr (CPLX) finish_head<BASE, CPLX>(a, c);
        else {
            finish_head_f32<BASE>(a, c);
            c.acc += (double)c.accf - (double)c.compf;
        }
    }
}

// all row-thread work of the sampler on one tile: sites in order, layers bottom-up (site n + 1 needs the draw of site n, so there
// are no anti-diagonals to walk and every step consumes the output of the step before it)
template <bool CPLX>
__device__ __forceinline__ void row_chain_sample(const Args& a, Ctx& c) {
    const int L = a.g.L, N = a.g.N, part = c.part;
    float hA[kUP], hB[kUP], hC[kUP];       // layers 0, 1, 2
#pragma unroll
    for (int j = 0; j < kUP; ++j) { hA[j] = 0.f; hB[j] = 0.f; hC[j] = 0.f; }
    const uint32_t regp = c.lane_addr + kColR + (kPU / 2) * part;
    stage_all(regp, hA, part);
    if (L > 1) stage_all(regp + 64, hB, part);
    if (L > 2) stage_all(regp + 128, hC, part);
    c.pn = -1;
    c.nup = 0;
    if (part == 0) {                       // the input of site 0 is the zero vector
        const float oh[1] = {0.f};
        umma::tmem_st1(c.lane_addr + kColR + 64 * L, oh);
    }
    umma::wait_st();
    umma::fence_before_sync();
    umma::mbar_arrive(&c.bars[kCDone]);                // "step -1": operands staged
    ++c.cda;
#pragma unroll 1
    for (int n = 0; n < N; ++n) {
        if (L == 3) {
            row_step<true, CPLX, 0, true>(a, c, n, 0, hA);
            row_step<true, CPLX, 1, true>(a, c, n, 1, hB);
            row_step<true, CPLX, 2, true>(a, c, n, 2, hC);
        } else if (L == 2) {
            row_step<true, CPLX, -1, true>(a, c, n, 0, hA);
            row_step<true, CPLX, -1, true>(a, c, n, 1, hB);
        } else {
            row_step<true, CPLX, -1, true>(a, c, n, 0, hA);
        }
    }
}

// the MMA instructions of one (site, layer) step; executed by every lane of the (converged) MMA warp, one elected lane issues.
// x group: D[cx | r | u | 4] = x * X^T (overwrite), h group: D[r | u | ch | junk] += h * H^T, where the very first h
// instruction is split in two because it accumulates onto r, u but must overwrite ch.  Consecutive instructions reuse the A chunk
// where they can (hi x B_hi, hi x B_lo, then lo x B_hi).
__device__ __forceinline__ void issue_step(uint32_t tbase, uint32_t rX, uint32_t rH, uint32_t x_1, uint32_t x_2, uint32_t h_1, uint32_t h_2,
                                           bool k16) {
#ifdef RNNWF_ABL_NOMMA     // ablation (timing / power only, results are garbage): the gate math alone -- no tcgen05.mma, the commit fires at once
    return;
#endif
    constexpr uint32_t idAll = (1u << 4) | ((uint32_t)(kNAll >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // F16 x F16 -> F32, M = 128
    constexpr uint32_t idRU = (1u << 4) | ((uint32_t)(kNRU >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t idC = (1u << 4) | ((uint32_t)(kNC >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t dX = tbase + kColX, dH = tbase + kColRU, dCH = tbase + kColCH;
    if (k16) {   // one-hot input: exact in FP16, no low limb
        umma::mma_f16_ts_elect(dX, rX, umma::smem_desc(x_1, 128, 2 * 128), idAll, 0);
        umma::mma_f16_ts_elect(dX, rX, umma::smem_desc(x_2, 128, 2 * 128), idAll, 1);
    } else {
#if RNNWF_KPACK
        // operand chunk q (columns 8 q ..) against B1 from core column 2 q (q <= 3) or 2 q - 7 (q >= 4); chunks 0..2 also against B2
        const uint64_t b1 = umma::smem_desc(x_1, 128, kKC * 128), b2 = umma::smem_desc(x_2, 128, kKC2 * 128);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            umma::mma_f16_ts_elect(dX, rX + q * 8, b1 + (uint64_t)(q * 16), idAll, q > 0);
            umma::mma_f16_ts_elect(dX, rX + q * 8, b2 + (uint64_t)(q * 16), idAll, 1);
        }
#ifndef RNNWF_ABL_SKIPLO   // ablation (timing / power only): without the lo x W_hi products (3 of the 10 MMAs of an operand group)
#pragma unroll
        for (int q = 3; q < 7; ++q) umma::mma_f16_ts_elect(dX, rX + q * 8, b1 + (uint64_t)((q == 3 ? 6 : 2 * q - 7) * 8), idAll, 1);
#else
        umma::mma_f16_ts_elect(dX, rX + 3 * 8, b1 + (uint64_t)(6 * 8), idAll, 1);
#endif
#else
        const uint64_t bhi = umma::smem_desc(x_1, 128, kKC * 128), blo = umma::smem_desc(x_2, 128, kKC * 128);
#pragma unroll
        for (int ks = 0; ks < kKp / 16; ++ks) {
            umma::mma_f16_ts_elect(dX, rX + ks * 8, bhi + (uint64_t)(ks * 16), idAll, ks > 0);
            umma::mma_f16_ts_elect(dX, rX + ks * 8, blo + (uint64_t)(ks * 16), idAll, 1);
            umma::mma_f16_ts_elect(dX, rX + 32 + ks * 8, bhi + (uint64_t)(ks * 16), idAll, 1);
        }
#endif
    }
    {
        const uint64_t bch = umma::smem_desc(h_1 + 2 * kBW * kKp * 2, 128, kKC * 128);          // candidate rows of the first image
#if RNNWF_KPACK
        const uint64_t b1 = umma::smem_desc(h_1, 128, kKC * 128), b2 = umma::smem_desc(h_2, 128, kKC2 * 128);
        umma::mma_f16_ts_elect(dH, rH, b1, idRU, 1);
        umma::mma_f16_ts_elect(dCH, rH, bch, idC, 0);
        umma::mma_f16_ts_elect(dH, rH, b2, idAll, 1);
#pragma unroll
        for (int q = 1; q < 3; ++q) {
            umma::mma_f16_ts_elect(dH, rH + q * 8, b1 + (uint64_t)(q * 16), idAll, 1);
            umma::mma_f16_ts_elect(dH, rH + q * 8, b2 + (uint64_t)(q * 16), idAll, 1);
        }
#ifndef RNNWF_ABL_SKIPLO
#pragma unroll
        for (int q = 3; q < 7; ++q) umma::mma_f16_ts_elect(dH, rH + q * 8, b1 + (uint64_t)((q == 3 ? 6 : 2 * q - 7) * 8), idAll, 1);
#else
        umma::mma_f16_ts_elect(dH, rH + 3 * 8, b1 + (uint64_t)(6 * 8), idAll, 1);
#endif
#else
        const uint64_t bhi = umma::smem_desc(h_1, 128, kKC * 128), blo = umma::smem_desc(h_2, 128, kKC * 128);
        umma::mma_f16_ts_elect(dH, rH, bhi, idRU, 1);
        umma::mma_f16_ts_elect(dCH, rH, bch, idC, 0);
        umma::mma_f16_ts_elect(dH, rH, blo, idAll, 1);
        umma::mma_f16_ts_elect(dH, rH + 32, bhi, idAll, 1);
#pragma unroll
        for (int ks = 1; ks < kKp / 16; ++ks) {
            umma::mma_f16_ts_elect(dH, rH + ks * 8, bhi + (uint64_t)(ks * 16), idAll, 1);
            umma::mma_f16_ts_elect(dH, rH + ks * 8, blo + (uint64_t)(ks * 16), idAll, 1);
            umma::mma_f16_ts_elect(dH, rH + 32 + ks * 8, bhi + (uint64_t)(ks * 16), idAll, 1);
        }
#endif
    }
}

struct Dbg { long long wru, wc, tru, tc, chain, mw1, mw2, mi; int nch; };

// the persistent work loop of one warp role (ROW: the 8 row warps; otherwise the MMA warp).  Both roles execute the same
// sequence of CTA barriers; they are separate instantiations so that each runs under its own register budget (setmaxnreg).
template <bool BASE, bool CPLX, bool ROW, bool IDLE = false, bool SAMPLE = false>
__device__ __forceinline__ uint32_t work_loop(const Args& a, const float* tab, float4* zsm, uint64_t* bars, int* s_work, uint32_t tbase,
                                              uint32_t lane_addr, uint32_t sB, Dbg& dbg) {
    const Layout& t = a.t;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int L = a.g.L, N = a.g.N, Mold = a.Mold;
    const int part = warp >> 2, rowi = tid & 127;
    const int total = (BASE ? 1 : a.nslots) * a.tiles128;
    bool weights_ready = false;
    uint32_t gstep = 0, cdp = 0, ztop = 0;   // steps done so far; c_done phases used so far (one per step + one per chain); top-layer steps so far
    while (true) {
        if (tid == 0) *s_work = atomicAdd(a.counter, 1);
        __syncthreads();
        const int work = *s_work;
        __syncthreads();
        if (work >= total) break;
        const int tile = work % a.tiles128;
        int slot = 0, s = -1, tt = -1;                                        // modified sites of this chain (none for BASE)
        if (!BASE) {
            slot = a.order ? a.order[work / a.tiles128] : work / a.tiles128;  // decreasing chain length
            if (slot < a.nslots - a.n_kind1 - a.n_kind2) s = slot;
            else if (slot < a.nslots - a.n_kind2) { s = slot - (a.nslots - a.n_kind1 - a.n_kind2); tt = s + 1; }
            else { s = slot - (a.nslots - a.n_kind2); tt = s + 2; }
            if (tt >= 0 && ((tt == s + 1 && a.j1 && a.j1[s] == 0.0) || (tt == s + 2 && a.j2 && a.j2[s] == 0.0))) continue;
        }
        const int64_t R = (int64_t)tile * kRows + rowi;
        const bool live = ROW && R < a.rows_total;
        const int64_t t120 = live ? R / Mold : 0;
        const int m = live ? (int)(R % Mold) : 0;
        const size_t rowbase = (size_t)t120 * N;                            // index of (old tile, site 0)
        const int n0 = s + 1, nsteps = (N - n0) * L;
        double acc = 0.0, acc_im = 0.0;
        if (!BASE && live && part == 0) {
            acc = a.la_oth[(rowbase + s) * Mold + m] - a.la_sel[(rowbase + s) * Mold + m];
            if (CPLX) acc_im = a.ph_oth[(rowbase + s) * Mold + m] - a.ph_sel[(rowbase + s) * Mold + m];
        }
        if (nsteps > 0 && !IDLE) {
            if constexpr (ROW) {
                Ctx c;
                c.tab = tab; c.zsm = zsm; c.bars = bars; c.lane_addr = lane_addr; c.rowi = rowi; c.m = m; c.part = part; c.live = live;
                c.rowbase = rowbase; c.s = s; c.t = tt; c.g = gstep; c.cda = cdp; c.zc = ztop; c.skew = true; c.acc = acc; c.acc_im = acc_im;
                c.pz = make_float4(0.f, 0.f, 0.f, 0.f); c.psg = 0; c.pn = -1; c.pph = 0; c.pbuf = 0; c.nup = 0; c.p_la = 0.0; c.p_ph = 0.0; c.p_laf = 0.f; c.accf = 0.f; c.compf = 0.f;
                TCP_T(c.w_ru = c.w_c = c.t_ru = c.t_c = 0; long long ch0 = clock64();)
                if constexpr (SAMPLE) row_chain_sample<CPLX>(a, c);
                else row_chain<BASE, CPLX>(a, c);
                acc = c.acc; acc_im = c.acc_im;
                TCP_T(dbg.wru += c.w_ru; dbg.wc += c.w_c; dbg.tru += c.t_ru; dbg.tc += c.t_c; dbg.chain += clock64() - ch0; ++dbg.nch;)
            } else {   // MMA warp: all lanes stay converged, one elected lane issues
                if (!weights_ready) { umma::mbar_wait(&bars[kWImg], 0); weights_ready = true; }
                uint32_t g = gstep, cdw = cdp;
                int pn = -1000, pl = -1;
                TCP_T(long long m_w1 = 0, m_w2 = 0, m_i = 0;)
                if constexpr (SAMPLE) {   // sites in order, layers bottom-up: every step waits for the restaged output of the step before it
#pragma unroll 1
                    for (int n = 0; n < N; ++n) {
#pragma unroll 1
                        for (int l = 0; l < L; ++l) {
                            const uint32_t lb = sB + (l == 0 ? 0u : (uint32_t)(t.l0_bytes + (l - 1) * t.l1_bytes));
                            const uint32_t h_hi = lb, h_lo = lb + t.im_bytes, x_hi = lb + t.im_bytes + t.im2_bytes;
                            const uint32_t x_lo = x_hi + (l == 0 ? t.im0_bytes : t.im_bytes);
                            const uint32_t rX = tbase + kColR + 64 * (l == 0 ? L : l - 1), rH = tbase + kColR + 64 * l;
                            umma::mbar_wait(&bars[kAccFree], (g - 1) & 1);
                            umma::mbar_wait(&bars[kCDone], cdw & 1);
                            umma::fence_after_sync();
                            issue_step(tbase, rX, rH, x_hi, x_lo, h_hi, h_lo, l == 0);
                            umma::commit_elect(&bars[kFull]);
                            ++g;
                            ++cdw;
                        }
                    }
                } else
#pragma unroll 1
                for (int d = n0; d <= N - 1 + L - 1; ++d) {
#pragma unroll 1
                    for (int l = L - 1; l >= 0; --l) {
                        const int n = d - l;
                        if (n < n0 || n >= N) continue;
                        // does this step consume the state produced by the step right before it?
                        const bool dep = pn == -1000 || (pn == n && pl == l - 1) || (pn == n - 1 && pl == l);
                        pn = n; pl = l;
                        const uint32_t lb = sB + (l == 0 ? 0u : (uint32_t)(t.l0_bytes + (l - 1) * t.l1_bytes));
                        const uint32_t h_hi = lb, h_lo = lb + t.im_bytes, x_hi = lb + t.im_bytes + t.im2_bytes;
                        const uint32_t x_lo = x_hi + (l == 0 ? t.im0_bytes : t.im_bytes);
                        const uint32_t rX = tbase + kColR + 64 * (l == 0 ? L : l - 1), rH = tbase + kColR + 64 * l;
                        TCP_T(long long q0 = clock64();)
                        umma::mbar_wait(&bars[kAccFree], (g - 1) & 1);             // previous step's accumulators are in registers
                        TCP_T(long long q1 = clock64(); m_w1 += q1 - q0;)
                        if (dep) umma::mbar_wait(&bars[kCDone], cdw & 1);          // previous step's state restaged
                        umma::fence_after_sync();
                        TCP_T(long long q2 = clock64(); m_w2 += q2 - q1;)
                        issue_step(tbase, rX, rH, x_hi, x_lo, h_hi, h_lo, l == 0);
                        umma::commit_elect(&bars[kFull]);
                        TCP_T(m_i += clock64() - q2;)
                        ++g;
                        ++cdw;
                    }
                }
                TCP_T(dbg.mw1 += m_w1; dbg.mw2 += m_w2; dbg.mi += m_i;)
            }
            gstep += (uint32_t)nsteps;
            cdp += (uint32_t)nsteps + 1;
            ztop += (uint32_t)(N - n0);
        }
        if (!SAMPLE && live && part == 0) {
            if (BASE) {
                a.lp[t120 * Mold + m] = acc;
                if (CPLX) a.lp_im[t120 * Mold + m] = acc_im;
            } else {
                a.delta[((size_t)t120 * a.nslots + slot) * Mold + m] = acc;
                if (CPLX) a.delta_im[((size_t)t120 * a.nslots + slot) * Mold + m] = acc_im;
            }
        }
    }
    return gstep;
}

template <bool BASE, bool CPLX, bool SAMPLE = false>
__global__ void __launch_bounds__(kThreads, 1) chain_kernel(const __grid_constant__ Args a) {
    extern __shared__ __align__(128) unsigned char smem_p16[];
    const Layout& t = a.t;
    const float* tab = reinterpret_cast<const float*>(smem_p16 + t.tab_off);
    float4* zsm = reinterpret_cast<float4*>(smem_p16 + ((t.img_bytes + 15) & ~15));
    uint64_t* bars = reinterpret_cast<uint64_t*>(zsm + 2 * (kParts - 1) * kRows);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);
    int* s_work = reinterpret_cast<int*>(tmem_slot + 1);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int L = a.g.L;
    const bool is_row = warp < kMmaWarp;

    if (warp == kMmaWarp) umma::tmem_alloc(tmem_slot, 512);
    if (tid == 0) {
        umma::mbar_init(&bars[kFull], 1);
        umma::mbar_init(&bars[kAccFree], kRowThreads);
        umma::mbar_init(&bars[kCDone], kRowThreads);
        umma::mbar_init(&bars[kWImg], 1);
        umma::mbar_init(&bars[kZDone], kRowThreads - kRows);
        umma::mbar_fence_init();
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    if (tid == 0) {   // the whole weight image stays resident: bulk async copies (TMA), one barrier
        umma::mbar_expect_tx(&bars[kWImg], (uint32_t)t.img_bytes);
        for (uint32_t o = 0; o < (uint32_t)t.img_bytes; o += 32768)
            umma::bulk_g2s(smem_p16 + o, a.img + o, min(32768u, (uint32_t)t.img_bytes - o), &bars[kWImg]);
    }
    const uint32_t tbase = *tmem_slot;
    const uint32_t lane_addr = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    if (warp < 4) {   // zero the operand regions (the staging writes every used column, incl. the constant 1), set the constant of X0
        const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (uint32_t c = 0; c < (uint32_t)(64 * L + 8); c += 8) umma::tmem_st8(lane_addr + kColR + c, z);
        const float one[1] = {__uint_as_float(pack_h2(1.0f, 0.0f))};
        umma::tmem_st1(lane_addr + kColR + 64 * L + 1, one);               // one-hot region: k = 2 is the constant 1
#if RNNWF_KPACK
        const float one2[1] = {__uint_as_float(pack_h2(1.0f, 1.0f))};                            // meets (bias_hi, bias_lo) in B1
        for (int l = 0; l < L; ++l) umma::tmem_st1(lane_addr + kColR + 64 * l + kOneCol, one2);   // state regions (nothing else writes it)
#else
        for (int l = 0; l < L; ++l) umma::tmem_st1(lane_addr + kColR + 64 * l + kKOne / 2, one);   // state regions: k = 50 (nothing else writes it)
#endif
        umma::wait_st();
    }
    if (is_row) umma::mbar_wait(&bars[kWImg], 0);                           // tab is read with ordinary loads
    const uint32_t sB = umma::smem_u32(smem_p16);
    Dbg dbg;
    memset(&dbg, 0, sizeof(dbg));
    TCP_T(const long long k_t0 = clock64();)
    // register budgets per role: the kernel is compiled for 168 registers per thread (3 warps of an SM sub-partition must fit);
    // the MMA warp hands most of its share to the row warps, which keep three layers' hidden states in registers
    uint32_t gstep;
    if (is_row) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRowRegs));
        gstep = work_loop<BASE, CPLX, true, false, SAMPLE>(a, tab, zsm, bars, s_work, tbase, lane_addr, sB, dbg);
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kMmaRegs));
        if (warp == kMmaWarp) gstep = work_loop<BASE, CPLX, false, false, SAMPLE>(a, tab, zsm, bars, s_work, tbase, lane_addr, sB, dbg);
        else gstep = work_loop<BASE, CPLX, false, true, SAMPLE>(a, tab, zsm, bars, s_work, tbase, lane_addr, sB, dbg);   // CTA barriers only
    }
    TCP_T(if (gstep > 0 && !BASE) {
        const long long tot = clock64() - k_t0;
        unsigned smid;
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        if (tid == 0)
            printf("tc16p blk %3d sm %3u: %6u steps %3d chains, %5lld cyc/step; wait_ru %4lld G_ru %4lld wait_c %4lld G_c %4lld; per chain outside steps %lld\n",
                   blockIdx.x, smid, gstep, dbg.nch, tot / gstep, dbg.wru / gstep, dbg.tru / gstep, dbg.wc / gstep, dbg.tc / gstep,
                   (dbg.chain - dbg.wru - dbg.tru - dbg.wc - dbg.tc) / dbg.nch);
        if (tid == kRowThreads && (blockIdx.x % 37) == 0)
            printf("tc16p blk %3d mma: per step wait1 %lld wait2 %lld issue %lld\n", blockIdx.x, dbg.mw1 / gstep, dbg.mw2 / gstep, dbg.mi / gstep);
    })
    umma::fence_before_sync();
    __syncthreads();
    if (warp == kMmaWarp) umma::tmem_dealloc(tbase, 512);
}

template <bool CPLX>
static int launch_chains(Args& a, int sms, bool flips, cudaStream_t s) {
    const int smem = (int)smem_bytes(a.t);
    RNNWF_CHECK(smem <= kSmemLimit, -3, "tensor-core chain kernel needs %d bytes of shared memory", smem);
    {
        RNNWF_CUDA(cudaMemsetAsync(a.counter, 0, sizeof(int), s));
        auto k = chain_kernel<true, CPLX>;
        RNNWF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        prof_count();
        k<<<std::min(a.tiles128, sms), kThreads, smem, s>>>(a);
        RNNWF_CUDA(cudaGetLastError());
    }
    if (flips) {
        RNNWF_CUDA(cudaMemsetAsync(a.counter, 0, sizeof(int), s));
        auto k = chain_kernel<false, CPLX>;
        RNNWF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        const int grid = (int)std::min<int64_t>((int64_t)a.nslots * a.tiles128, sms);
        prof_count();
        prof_mark(0, s);
        k<<<grid, kThreads, smem, s>>>(a);
        prof_mark(1, s);
        RNNWF_CUDA(cudaGetLastError());
    }
    return 0;
}

static Args make_args(const GruLayout& g, int Mold, int tiles, unsigned char* img, const uint8_t* sigT, float* hstore, double* la_sel,
                      double* la_oth, float* la_self, double* lp, double* delta, int* counter, int& sms) {
    int dev = 0;
    sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    Args a;
    memset(&a, 0, sizeof(a));
    a.g = g; a.t = make_layout(g); a.Mold = Mold;
    a.rows_total = (int64_t)tiles * Mold;
    a.tiles128 = (int)cdiv(a.rows_total, kRows);
    a.img = img; a.sigT = sigT; a.hstore = hstore; a.la_sel = la_sel; a.la_oth = la_oth; a.la_self = la_self; a.lp = lp; a.delta = delta; a.counter = counter;
    a.nslots = g.N;
    return a;
}

// base pass + single-flip chains (FP32 pRNN with 50 units)
static int launch_eloc(const GruLayout& greal, int Mold, int tiles, const float* params, unsigned char* img, const uint8_t* sigT, float* hstore,
                       double* la_sel, double* la_oth, float* la_self, double* lp, double* delta, int* counter, bool flips, cudaStream_t s,
                       float* gstore = nullptr) {
    int sms;
    const GruLayout g = padded_layout(greal);      // hstore, the image and the kernel see 50 units; the pack reads the real widths
    Args a = make_args(g, Mold, tiles, img, sigT, hstore, la_sel, la_oth, la_self, lp, delta, counter, sms);
    a.gstore = gstore;
    prof_count(); pack_kernel<<<148, 256, 0, s>>>(greal, a.t, params, img);
    return launch_chains<false>(a, sms, flips, s);
}

// log psi alone: the base pass without the stash of restart states and per-site terms (rnnwf_logpsi)
static int launch_logpsi(const GruLayout& greal, int tiles128, const float* params, unsigned char* img, const uint8_t* sigT, double* lp_re,
                         double* lp_im, int* counter, cudaStream_t s) {
    int sms;
    const GruLayout g = padded_layout(greal);
    Args a = make_args(g, kRows, tiles128, img, sigT, nullptr, nullptr, nullptr, nullptr, lp_re, nullptr, counter, sms);
    a.lp_im = lp_im;
    prof_count(); pack_kernel<<<148, 256, 0, s>>>(greal, a.t, params, img);
    return g.nheads == 2 ? launch_chains<true>(a, sms, false, s) : launch_chains<false>(a, sms, false, s);
}

// autoregressive sampler (probability head, or the cRNN's amplitude head with the U(1) mask): 128-row tiles, one persistent CTA each
static int launch_sample(const GruLayout& greal, int tiles128, const float* params, unsigned char* img, uint8_t* sampT, int* counter,
                         uint64_t seed, uint64_t sample_offset, cudaStream_t s) {
    int sms;
    const GruLayout g = padded_layout(greal);
    Args a = make_args(g, kRows, tiles128, img, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, counter, sms);
    a.sampT = sampT; a.seed = seed; a.sample_offset = sample_offset;
    const int smem = (int)smem_bytes(a.t);
    RNNWF_CHECK(smem <= kSmemLimit, -3, "tensor-core sampler needs %d bytes of shared memory", smem);
    prof_count(); pack_kernel<<<148, 256, 0, s>>>(greal, a.t, params, img);
    RNNWF_CUDA(cudaMemsetAsync(a.counter, 0, sizeof(int), s));
    prof_count();
    if (g.nheads == 2) {
        auto k = chain_kernel<true, true, true>;
        RNNWF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        k<<<std::min(a.tiles128, sms), kThreads, smem, s>>>(a);
    } else {
        auto k = chain_kernel<true, false, true>;
        RNNWF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        k<<<std::min(a.tiles128, sms), kThreads, smem, s>>>(a);
    }
    RNNWF_CUDA(cudaGetLastError());
    return 0;
}

// base pass + NN / NNN exchange chains of the complex cRNN (J1-J2); `order` lists the 2N-3 slots by decreasing chain length
static int launch_j1j2(const GruLayout& greal, int Mold, int tiles, const float* params, unsigned char* img, const uint8_t* sigT, float* hstore,
                       double* la_sel, double* la_oth, double* ph_sel, double* ph_oth, double* lp_re, double* lp_im, double* delta_re,
                       double* delta_im, const int* order, const double* j1, const double* j2, int* counter, cudaStream_t s) {
    int sms;
    const GruLayout g = padded_layout(greal);
    Args a = make_args(g, Mold, tiles, img, sigT, hstore, la_sel, la_oth, nullptr, lp_re, delta_re, counter, sms);
    a.ph_sel = ph_sel; a.ph_oth = ph_oth; a.lp_im = lp_im; a.delta_im = delta_im; a.order = order; a.j1 = j1; a.j2 = j2;
    a.n_kind1 = g.N - 1; a.n_kind2 = g.N - 2; a.nslots = a.n_kind1 + a.n_kind2;
    prof_count(); pack_kernel<<<148, 256, 0, s>>>(greal, a.t, params, img);
    return launch_chains<true>(a, sms, true, s);
}

}  // namespace tc16p
}  // namespace rnnwf
