// gru_engine.cuh — FP32/FP64 FFMA tile engine for stacked cuDNN-style GRU cells (sm_100a).
//
// Replaces the per-site TF op chains of  1DTFIM/RNNwavefunction.py:65-70,107-111  (MultiRNNCell of
// CudnnCompatibleGRUCell + Dense).  One CTA owns a tile of M "rows" (samples or connected
// configurations).  Packed weights of ALL layers stay resident in shared memory for the whole
// launch; the hidden state tile lives in shared memory ([layer][unit][row], row-contiguous) and each
// thread keeps an (SPT rows x 2 units x 4 gate-accumulators) register tile.
#pragma once
#include "common.cuh"

namespace rnnwf {

// ---------------------------------------------------------------------------------------------
// Parameter layouts.  "flat" = TF-variable creation order (what the Python side / Adam sees);
// "packed" = per-thread-column interleaved layout the kernels read.
//   flat, per layer l (input width d):  Kg[(d+H),2H] | bg[2H] | Kci[d,H] | Kch[H,H] | bci[H] | bch[H]
//   flat, heads: Wd[H,2] | bd[2]   ( x2 for the complex RNN: amplitude head then phase head )
//   packed, per layer:  wx_ru[d][CT][4] | wx_c[d][CT][2] | wh_ru[H][CT][4] | wh_c[H][CT][2] | b[CT][8]
//     slot order of a column-thread ct (units j0=2ct, j1=2ct+1):  ru = {r_j0, r_j1, u_j0, u_j1},
//     c = {c_j0, c_j1},  b = {bg_r j0,j1, bg_u j0,j1, bci j0,j1, bch j0,j1}
// ---------------------------------------------------------------------------------------------
struct GruLayout {
    int L, H, CT, N, nheads;
    int d[kMaxLayers];
    int flat_off[kMaxLayers];
    int pk_off[kMaxLayers];
    int o_wx_c[kMaxLayers], o_wh_ru[kMaxLayers], o_wh_c[kMaxLayers], o_b[kMaxLayers], pk_size[kMaxLayers];
    int flat_head, pk_head;
    int P, PK;
};

inline GruLayout make_gru_layout(const rnnwf_model& m) {
    GruLayout g;
    memset(&g, 0, sizeof(g));
    g.L = m.num_layers;
    g.H = m.units;
    g.CT = (m.units + 1) / 2;
    g.N = m.n_sites;
    g.nheads = m.head == RNNWF_HEAD_COMPLEX ? 2 : 1;
    int fo = 0, po = 0;
    for (int l = 0; l < g.L; ++l) {
        int d = l == 0 ? 2 : g.H, H = g.H, CT = g.CT;
        g.d[l] = d;
        g.flat_off[l] = fo;
        fo += (d + H) * 2 * H + 2 * H + d * H + H * H + 2 * H;
        g.pk_off[l] = po;
        int o = align4(d * CT * 4);
        g.o_wx_c[l] = o;
        o += align4(d * CT * 2);
        g.o_wh_ru[l] = o;
        o += align4(H * CT * 4);
        g.o_wh_c[l] = o;
        o += align4(H * CT * 2);
        g.o_b[l] = o;
        o += CT * 8;
        g.pk_size[l] = o;
        po += o;
    }
    g.flat_head = fo;
    g.pk_head = po;
    fo += g.nheads * (2 * g.H + 2);
    po += align4(g.nheads * (2 * g.H + 2));
    g.P = fo;
    g.PK = po;
    return g;
}

// Launch geometry chosen on the host (see choose_gru_launch in gru.cu)
struct GruLaunch {
    int CT, RT, M, Mp;   // column threads, row threads, rows per tile (RT*SPT), M rounded to 16
    int NTc;             // compute threads rounded up to a warp multiple; block = NTc + kHeadThreads
    int w_smem;          // packed weights resident in shared memory
    int ring_kc;         // !w_smem: K rows per chunk of the shared-memory weight ring (0: weights read straight from global / L2)
    int smem_bytes;
};

// Weight ring for stacks whose packed weights do not fit in shared memory (e.g. GRU(100) in float64: 245 KB).  The compute warps
// stream the h-part (and, above layer 0, the x-part) of a layer's weights through two shared-memory buffers of KC K-rows each with
// cp.async: every weight crosses L2 -> SM once per CTA and step instead of once per row-thread group, and its latency is hidden
// behind the FMAs of the previous chunk (ncu of the float64 chain kernel before this: long-scoreboard 4.8 and barrier 3.7 stalls per
// issue, FP64 pipe 25 % busy).  bar.sync 1 synchronises the NTc compute threads only; the head warps never see it.
template <typename T> struct WRing {
    T* buf;      // [2][KC * CT * 6]: per buffer ru rows [KC][CT][4] then c rows [KC][CT][2]
    int KC, nthr, tid;
};
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
template <typename T>
__device__ __forceinline__ void ring_fetch(const WRing<T>& r, int buf, const T* __restrict__ src_ru, const T* __restrict__ src_c, int rows, int CT) {
    unsigned char* d_ru = reinterpret_cast<unsigned char*>(r.buf + (size_t)buf * r.KC * CT * 6);
    unsigned char* d_c = d_ru + (size_t)r.KC * CT * 4 * sizeof(T);
    const unsigned char* s_ru = reinterpret_cast<const unsigned char*>(src_ru);
    const unsigned char* s_c = reinterpret_cast<const unsigned char*>(src_c);
    const int n16 = rows * CT * 4 * (int)sizeof(T) / 16, n8 = rows * CT * 2 * (int)sizeof(T) / 8;
    for (int i = r.tid; i < n16; i += r.nthr) cp_async16(d_ru + 16 * i, s_ru + 16 * i);
    for (int i = r.tid; i < n8; i += r.nthr) cp_async8(d_c + 8 * i, s_c + 8 * i);
    asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int PENDING> __device__ __forceinline__ void ring_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }
__device__ __forceinline__ void ring_sync(int nthr) { asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory"); }

// acc += A[K rows of the tile] x W[K][thread's 6 gate columns], the weights of K-row k at ru + (k*CT + ct)*4 and c + (k*CT + ct)*2
template <typename T>
__device__ __forceinline__ void preact_rows(const T* __restrict__ a_rows, int M, int k0, int rows, const T* __restrict__ ru, const T* __restrict__ cw,
                                            int CT, int ct, T (&ar)[2][VT<T>::SPT], T (&au)[2][VT<T>::SPT], T (&acq)[2][VT<T>::SPT]) {
    constexpr int SPT = VT<T>::SPT;
#pragma unroll 2
    for (int k = 0; k < rows; ++k) {
        T a[SPT], w[4], c[2];
        ldv<SPT>(a, a_rows + (size_t)(k0 + k) * M);
        ldv<4>(w, ru + (k * CT + ct) * 4);
        ldv<2>(c, cw + (k * CT + ct) * 2);
#pragma unroll
        for (int s = 0; s < SPT; ++s) {
            ar[0][s] = fma(a[s], w[0], ar[0][s]);
            ar[1][s] = fma(a[s], w[1], ar[1][s]);
            au[0][s] = fma(a[s], w[2], au[0][s]);
            au[1][s] = fma(a[s], w[3], au[1][s]);
            acq[0][s] = fma(a[s], c[0], acq[0][s]);
            acq[1][s] = fma(a[s], c[1], acq[1][s]);
        }
    }
}
// the same product with the weights streamed through the ring (all r.nthr compute threads call this together)
template <typename T>
__device__ __forceinline__ void preact_ring(const WRing<T>& r, const T* __restrict__ a_rows, int M, int K, const T* __restrict__ ru_g,
                                            const T* __restrict__ c_g, int CT, int ct, T (&ar)[2][VT<T>::SPT], T (&au)[2][VT<T>::SPT],
                                            T (&acq)[2][VT<T>::SPT]) {
    const int KC = r.KC, nch = (K + KC - 1) / KC;
    ring_fetch<T>(r, 0, ru_g, c_g, min(KC, K), CT);
    for (int i = 0; i < nch; ++i) {
        const int k0 = i * KC;
        if (i + 1 < nch) {
            ring_fetch<T>(r, (i + 1) & 1, ru_g + (size_t)(k0 + KC) * CT * 4, c_g + (size_t)(k0 + KC) * CT * 2, min(KC, K - k0 - KC), CT);
            ring_wait<1>();
        } else {
            ring_wait<0>();
        }
        ring_sync(r.nthr);                                   // chunk i has landed for every thread
        const T* b = r.buf + (size_t)(i & 1) * KC * CT * 6;
        preact_rows<T>(a_rows, M, k0, min(KC, K - k0), b, b + (size_t)KC * CT * 4, CT, ct, ar, au, acq);
        ring_sync(r.nthr);                                   // buffer i & 1 may be refilled (chunk i + 2)
    }
}

template <typename T>
__global__ void pack_gru_kernel(GruLayout g, const T* __restrict__ flat, T* __restrict__ pk) {
    const int H = g.H, CT = g.CT;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < g.PK; idx += gridDim.x * blockDim.x) {
        T val = T(0);
        if (idx >= g.pk_head) {
            int i = idx - g.pk_head;
            if (i < g.nheads * (2 * H + 2)) val = flat[g.flat_head + i];
        } else {
            int l = 0;
            while (l + 1 < g.L && idx >= g.pk_off[l + 1]) ++l;
            const int d = g.d[l];
            const T* Kg = flat + g.flat_off[l];
            const T* bg = Kg + (d + H) * 2 * H;
            const T* Kci = bg + 2 * H;
            const T* Kch = Kci + d * H;
            const T* bci = Kch + H * H;
            const T* bch = bci + H;
            int loc = idx - g.pk_off[l];
            if (loc < g.o_wx_c[l]) {
                int k = loc / (CT * 4), ct = (loc % (CT * 4)) / 4, s = loc % 4, j = 2 * ct + (s & 1);
                if (k < d && j < H) val = Kg[k * 2 * H + (s >> 1) * H + j];
            } else if (loc < g.o_wh_ru[l]) {
                int q = loc - g.o_wx_c[l];
                int k = q / (CT * 2), ct = (q % (CT * 2)) / 2, j = 2 * ct + (q & 1);
                if (k < d && j < H) val = Kci[k * H + j];
            } else if (loc < g.o_wh_c[l]) {
                int q = loc - g.o_wh_ru[l];
                int k = q / (CT * 4), ct = (q % (CT * 4)) / 4, s = q % 4, j = 2 * ct + (s & 1);
                if (k < H && j < H) val = Kg[(d + k) * 2 * H + (s >> 1) * H + j];
            } else if (loc < g.o_b[l]) {
                int q = loc - g.o_wh_c[l];
                int k = q / (CT * 2), ct = (q % (CT * 2)) / 2, j = 2 * ct + (q & 1);
                if (k < H && j < H) val = Kch[k * H + j];
            } else {
                int q = loc - g.o_b[l];
                int ct = q / 8, s = q % 8, j = 2 * ct + (s & 1);
                if (j < H) {
                    int which = s >> 1;
                    val = which == 0 ? bg[j] : which == 1 ? bg[H + j] : which == 2 ? bci[j] : bch[j];
                }
            }
        }
        pk[idx] = val;
    }
}

// ---------------------------------------------------------------------------------------------
// One GRU layer on a tile:
//   [r|u] = sigmoid([x,h] Kg + bg);  c = tanh(x Kci + bci + r*(h Kch + bch));  h' = (1-u) c + u h
// Thread (rt, ct) produces h' for rows rt*SPT..+SPT-1 and units 2ct, 2ct+1 into hn[2][SPT].
// Ax == nullptr: the layer input is a one-hot / zero vector given by byte codes in `sig`
// (0,1 = one-hot spin; 2 = the all-zero first input of RNNwavefunction.py:52-55).
// ---------------------------------------------------------------------------------------------
// Pre-activations of one layer on the thread's register tile:
//   ar/au = gate pre-activations (incl. bias), ac = x Kci + bci, aq = h Kch + bch.
template <typename T>
__device__ __forceinline__ void gru_preact(const GruLayout& g, int l, const T* __restrict__ wl,
                                           const T* __restrict__ Ax, const T* __restrict__ Ah,
                                           const uint8_t* __restrict__ sig, int M, int ct, int rt,
                                           T (&ar)[2][VT<T>::SPT], T (&au)[2][VT<T>::SPT], T (&ac)[2][VT<T>::SPT],
                                           T (&aq)[2][VT<T>::SPT], const WRing<T>* ring = nullptr) {
    constexpr int SPT = VT<T>::SPT;
    const int H = g.H, CT = g.CT, d = g.d[l];
    const T* wx_ru = wl;
    const T* wx_c = wl + g.o_wx_c[l];
    const T* wh_ru = wl + g.o_wh_ru[l];
    const T* wh_c = wl + g.o_wh_c[l];
    {
        T b[8];
        ldv<8>(b, wl + g.o_b[l] + ct * 8);
#pragma unroll
        for (int s = 0; s < SPT; ++s) {
            ar[0][s] = b[0]; ar[1][s] = b[1]; au[0][s] = b[2]; au[1][s] = b[3];
            ac[0][s] = b[4]; ac[1][s] = b[5]; aq[0][s] = b[6]; aq[1][s] = b[7];
        }
    }
    const int row0 = rt * SPT;
    if (Ax == nullptr) {
        T w0[4], w1[4], c0[2], c1[2];
        ldv<4>(w0, wx_ru + ct * 4);
        ldv<4>(w1, wx_ru + (CT + ct) * 4);
        ldv<2>(c0, wx_c + ct * 2);
        ldv<2>(c1, wx_c + (CT + ct) * 2);
#pragma unroll
        for (int s = 0; s < SPT; ++s) {
            const int code = sig[row0 + s];
            const T m0 = code == 0 ? T(1) : T(0), m1 = code == 1 ? T(1) : T(0);
            ar[0][s] += m0 * w0[0] + m1 * w1[0];
            ar[1][s] += m0 * w0[1] + m1 * w1[1];
            au[0][s] += m0 * w0[2] + m1 * w1[2];
            au[1][s] += m0 * w0[3] + m1 * w1[3];
            ac[0][s] += m0 * c0[0] + m1 * c1[0];
            ac[1][s] += m0 * c0[1] + m1 * c1[1];
        }
    } else if (ring) {
        preact_ring<T>(*ring, Ax + row0, M, d, wx_ru, wx_c, CT, ct, ar, au, ac);
    } else {
        preact_rows<T>(Ax + row0, M, 0, d, wx_ru, wx_c, CT, ct, ar, au, ac);
    }
    if (ring) preact_ring<T>(*ring, Ah + row0, M, H, wh_ru, wh_c, CT, ct, ar, au, aq);
    else preact_rows<T>(Ah + row0, M, 0, H, wh_ru, wh_c, CT, ct, ar, au, aq);
}

template <typename T>
__device__ __forceinline__ void gru_layer(const GruLayout& g, int l, const T* __restrict__ wl,
                                          const T* __restrict__ Ax, const T* __restrict__ Ah,
                                          const uint8_t* __restrict__ sig, int M, int ct, int rt,
                                          T (&hn)[2][VT<T>::SPT], const WRing<T>* ring = nullptr) {
    constexpr int SPT = VT<T>::SPT;
    const int H = g.H;
    T ar[2][SPT], au[2][SPT], ac[2][SPT], aq[2][SPT];
    gru_preact<T>(g, l, wl, Ax, Ah, sig, M, ct, rt, ar, au, ac, aq, ring);
    const int row0 = rt * SPT;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int j = 2 * ct + u;
        if (j < H) {
            T hold[SPT];
            ldv<SPT>(hold, Ah + j * M + row0);
#pragma unroll
            for (int s = 0; s < SPT; ++s) {
                const T r = sigmoid_(ar[u][s]);
                const T uu = sigmoid_(au[u][s]);
                const T c = tanh_(fma(r, aq[u][s], ac[u][s]));
                hn[u][s] = fma(uu, hold[s] - c, c);
            }
        }
    }
}

// All layers of one site for a tile.  Barrier protocol per layer: compute -> sync -> in-place write of the
// layer's state (and optional stash to HBM) -> sync.  Head warps and idle lanes only take the barriers.
template <typename T, bool STASH>
__device__ __forceinline__ void gru_site(const GruLayout& g, const T* __restrict__ w, T* __restrict__ hbuf,
                                         const uint8_t* __restrict__ sigcur, int M, int ct, int rt,
                                         bool is_compute, T* __restrict__ stash, const WRing<T>* ring = nullptr) {
    constexpr int SPT = VT<T>::SPT;
    const int H = g.H;
    for (int l = 0; l < g.L; ++l) {
        T hn[2][SPT];
        T* hl = hbuf + l * H * M;
        if (ring) {   // the padding lanes of the last compute warp take part in the ring's copies and barriers (row tile clamped, nothing written)
            if (ring->tid < ring->nthr) gru_layer<T>(g, l, w + g.pk_off[l], l == 0 ? nullptr : hl - H * M, hl, sigcur, M, ct, min(rt, M / SPT - 1), hn, ring);
        } else if (is_compute) gru_layer<T>(g, l, w + g.pk_off[l], l == 0 ? nullptr : hl - H * M, hl, sigcur, M, ct, rt, hn);
        __syncthreads();
        if (is_compute) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int j = 2 * ct + u;
                if (j < H) {
                    stv<SPT>(hl + j * M + rt * SPT, hn[u]);
                    if (STASH) stv<SPT>(stash + (l * H + j) * M + rt * SPT, hn[u]);
                }
            }
        }
        __syncthreads();
    }
}

// Dense head logits of row m from the top layer state: z[o] = bd[o] + sum_j h[j][m] Wd[j][o]
// (tf.layers.Dense(2), 1DTFIM/RNNwavefunction.py:33,67).  `wd` points at Wd[H][2] | bd[2].
template <typename T>
__device__ __forceinline__ void dense2(const T* __restrict__ htop, const T* __restrict__ wd, int H, int M, int m,
                                       T& z0, T& z1) {
    T a0 = wd[2 * H], a1 = wd[2 * H + 1];
#pragma unroll 4
    for (int j = 0; j < H; ++j) {
        const T h = htop[j * M + m];
        a0 = fma(h, wd[2 * j], a0);
        a1 = fma(h, wd[2 * j + 1], a1);
    }
    z0 = a0;
    z1 = a1;
}

}  // namespace rnnwf
