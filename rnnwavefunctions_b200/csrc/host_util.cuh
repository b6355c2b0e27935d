// host_util.cuh — workspace bump allocator and launch-geometry selection (host side).
#pragma once
#include "gru_engine.cuh"

namespace rnnwf {

// Bump allocator over the caller-provided workspace.  With base == nullptr it only measures.
struct Ws {
    unsigned char* base;
    size_t used;
    size_t cap;
    Ws(void* p, size_t c) : base(reinterpret_cast<unsigned char*>(p)), used(0), cap(c) {}
    template <typename U> U* take(size_t n) {
        used = (used + 255) & ~(size_t)255;
        U* r = base ? reinterpret_cast<U*>(base + used) : nullptr;
        used += n * sizeof(U);
        return r;
    }
    bool ok() const { return base == nullptr || used <= cap; }
};

// Pick (RT, M) so that CT*RT fills whole groups of 4 warps (one per SM sub-partition) as tightly as
// possible, the hidden-state tile + resident weights fit in 227 KB of shared memory, and M is as large
// as possible among near-ties.  Block = compute threads (<= 384) + 4 head warps.
// With rows_hint > 0 (one-CTA-per-tile kernels: sampler, log psi, gradient) the tile is instead sized so that the
// tiles fill the SMs in as few waves as possible: cost = waves * (M + 24) (per-site time of a tile = fixed part + rows).
inline double wave_cost(int64_t rows_hint, int ndir, int M, int sms = 148) {
    const int64_t tiles = ndir * ((rows_hint + M - 1) / M);
    return (double)((tiles + sms - 1) / sms) * (M + 24);   // + a fixed per-site cost worth ~24 rows (measured: tiny tiles are inefficient)
}

constexpr int kRingKC = 10;   // K rows per weight-ring chunk (even: keeps the 8- and 16-byte cp.async pieces aligned for float and double)

template <typename T> inline GruLaunch choose_gru_launch(const GruLayout& g, int64_t rows_hint = 0, int ndir = 1) {
    constexpr int SPT = VT<T>::SPT;
    GruLaunch best;
    memset(&best, 0, sizeof(best));
    for (int wsm = 1; wsm >= 0; --wsm) {
        double best_eff = -1.0, best_cost = 1e300;
        for (int RT = 1; RT <= 64; ++RT) {
            const int nt = g.CT * RT, M = RT * SPT;
            if (nt > 384 || M > 2 * kHeadThreads) break;
            const int Mp = (M + 15) & ~15;
            // weights resident (wsm) or streamed through a two-buffer ring of kRingKC K-rows (see WRing in gru_engine.cuh)
            const size_t ring = wsm ? 0 : 2 * (size_t)kRingKC * g.CT * 6 * sizeof(T) + 16;
            size_t smem = (wsm ? (((size_t)g.PK * sizeof(T) + 15) & ~(size_t)15) : 0) + (size_t)g.L * g.H * M * sizeof(T) +
                          2 * (size_t)Mp + 64 + ring;
            if (smem > (size_t)kSmemLimit) break;
            bool take;
            if (rows_hint > 0) {
                const double cost = wave_cost(rows_hint, ndir, M);
                take = cost < best_cost || (cost == best_cost && best.RT > 0 && M > best.M);
                if (take) best_cost = cost;
            } else {
                const double eff = (double)nt / (128.0 * (double)((nt + 127) / 128));
                take = eff >= best_eff - 0.03;   // near-tie: prefer the larger tile
                if (eff > best_eff) best_eff = eff;
            }
            if (take) {
                best.CT = g.CT; best.RT = RT; best.M = M; best.Mp = Mp;
                best.NTc = (nt + 31) & ~31; best.w_smem = wsm; best.ring_kc = wsm ? 0 : kRingKC; best.smem_bytes = (int)smem;
            }
        }
        if (best.RT > 0) return best;
    }
    return best;
}

}  // namespace rnnwf
