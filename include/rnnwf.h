/*
 * rnnwf.h — C ABI of librnnwf_b200.so: the B200-native (sm_100a) implementation of the RNN-wavefunction
 * VMC hot path (sample -> local energy -> VMC gradient).
 *
 * The reference (MatteoMartinelli97/RNNWavefunctions) has NO native/FFI interface: its only boundary is
 * the Python surface that ends in `sess.run(...)`.  Each entry point below therefore names the reference
 * Python call (file:line, relative to the reference root) whose device work it replaces.
 *
 * Conventions
 *   - extern "C"; plain pointers and sizes only.  Every function returns 0 on success, <0 on error;
 *     rnnwf_last_error() returns a thread-local message.  No exceptions cross the ABI.
 *   - All buffer arguments are DEVICE pointers owned by the caller unless the name ends in `_host`.
 *     The library keeps no global device state; scratch comes from the caller (`ws`, `ws_bytes`,
 *     size from rnnwf_workspace_bytes()).  All work is enqueued asynchronously on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream).
 *   - Samples are uint8 [ns, N] row-major with values 0/1 (site index = draw order, as
 *     1DTFIM/RNNwavefunction.py:72).  2-D RNN samples are uint8 [ns, Nx, Ny] indexed [b][x][y]
 *     (2DTFIM_2DRNN/RNNwavefunction.py:116).
 *   - Parameters are one flat buffer (float or double according to model.dtype) in TF-variable
 *     creation order (names, shapes and order: rnnwavefunctions_b200/params.py).
 */
#ifndef RNNWF_H
#define RNNWF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RNNWF_ABI_VERSION 1

enum { RNNWF_CELL_GRU = 0, RNNWF_CELL_MDRNN = 1 };
enum { RNNWF_HEAD_PROB = 0,      /* Dense(2)+softmax                    1DTFIM/RNNwavefunction.py:33           */
       RNNWF_HEAD_COMPLEX = 1 }; /* sqrt-softmax amplitude + pi*softsign phase, U(1) mask  J1J2/ComplexRNNwavefunction.py:42-43,85-93 */
enum { RNNWF_F32 = 0, RNNWF_F64 = 1 };

/* Flags for the TFIM entry points */
enum { RNNWF_PARITY_SYM = 1 };   /* log(0.5(P(s)+P(reversed s)))        1DTFIM/RNNwavefunction_paritysym.py:125,145 */

/* Workspace selectors */
enum { RNNWF_OP_SAMPLE = 0, RNNWF_OP_LOGPSI = 1, RNNWF_OP_TFIM_ELOC = 2, RNNWF_OP_VMC_GRAD = 3, RNNWF_OP_J1J2_ELOC = 4 };

typedef struct rnnwf_model {
    int32_t cell;        /* RNNWF_CELL_*                                                                */
    int32_t head;        /* RNNWF_HEAD_*                                                                */
    int32_t dtype;       /* RNNWF_F32 (1DTFIM, J1J2) or RNNWF_F64 (both 2-D apps use float64)           */
    int32_t num_layers;  /* stacked GRU layers (MultiRNNCell); MDRNN: 1                                 */
    int32_t units;       /* hidden units per layer (the run_* drivers always use equal widths)          */
    int32_t n_sites;     /* N for chains; Nx*Ny for lattices                                            */
    int32_t nx, ny;      /* lattice shape for the 2-D RNN path / 2-D bonds; 0,0 for chains              */
} rnnwf_model;

const char* rnnwf_last_error(void);
int rnnwf_abi_version(void);

/* Number of parameters of `m` (matches the reference's printed count, 1DTFIM/TrainingRNN_1DTFIM.py:127-136). */
int64_t rnnwf_param_count(const rnnwf_model* m);

/* Scratch bytes needed by operation `op` (RNNWF_OP_*) for `ns` samples. */
size_t rnnwf_workspace_bytes(const rnnwf_model* m, int op, int64_t ns, int flags);

/* ---- K1: autoregressive sampling -------------------------------------------------------------------
 * Replaces sess.run(samples_) of RNNwavefunction.sample (1DTFIM/RNNwavefunction.py:35-74,
 * 2DTFIM_1DRNN/RNNwavefunction.py:40-84, 2DTFIM_2DRNN/RNNwavefunction.py:35-118,
 * J1J2/ComplexRNNwavefunction.py:45-103 incl. the zero-magnetisation mask).
 * Draw for (sample id, site) uses Philox4x32-10 with counter (sample_offset+row, site) and key `seed`,
 * so the union over ranks is independent of how samples are sharded.                                   */
int rnnwf_sample(const rnnwf_model* m, const void* params, int64_t ns, uint64_t seed, uint64_t sample_offset,
                 uint8_t* samples_out, void* ws, size_t ws_bytes, void* stream);

/* ---- log psi -----------------------------------------------------------------------------------------
 * Replaces sess.run(log_probs_tensor) of .log_probability (1DTFIM/RNNwavefunction.py:76-118; parity
 * variant RNNwavefunction_paritysym.py:80-145 with flags&RNNWF_PARITY_SYM; 2-D variants) and
 * .log_amplitude (J1J2/ComplexRNNwavefunction.py:105-169).
 * out: double[ns] (HEAD_PROB: log-probability) or double[2*ns] interleaved (re, im) (HEAD_COMPLEX).      */
int rnnwf_logpsi(const rnnwf_model* m, const void* params, const uint8_t* samples, int64_t ns, int flags,
                 double* out, void* ws, size_t ws_bytes, void* stream);

/* ---- K2: fused log-prob + TFIM local energies ---------------------------------------------------------
 * Replaces Ising_local_energies (1DTFIM/TrainingRNN_1DTFIM.py:13-75) and Ising2D_local_energies
 * (2DTFIM_1DRNN/Training1DRNN_2DTFIM.py:13-81, 2DTFIM_2DRNN/Training2DRNN_2DTFIM.py:13-83) including the
 * (N+1)*ns log-probability evaluations behind them.  Every single-flip configuration is evaluated by
 * re-running only the RNN steps after the flipped site (prefix reuse).
 * jz: double[N-1 used of N] (chain) or double[Nx*Ny] row-major [x][y] (lattice, m->nx,ny > 0).
 * eloc_out: double[ns]; logp_out: double[ns] or NULL (log-probability of the unflipped samples).        */
int rnnwf_tfim_eloc(const rnnwf_model* m, const void* params, const uint8_t* samples, int64_t ns,
                    const double* jz, double bx, int flags, double* eloc_out, double* logp_out,
                    void* ws, size_t ws_bytes, void* stream);

/* As rnnwf_tfim_eloc, additionally returning the amplitude ratio of every single-flip configuration:
 * ratios_out: double[ns * N], ratios_out[s * N + k] = psi(sigma_s with site k flipped) / psi(sigma_s) = exp((lp_{k+1} - lp_0) / 2), the
 * terms the reference sums at 1DTFIM/TrainingRNN_1DTFIM.py:70-74 (2-D: slot order of Training2DRNN_2DTFIM.py:55-61).  Their
 * sample mean is <sigma^x_k>: the per-site observable behind README.md:3 ("correlation functions").  Requires bx != 0.          */
int rnnwf_tfim_flip_ratios(const rnnwf_model* m, const void* params, const uint8_t* samples, int64_t ns,
                           const double* jz, double bx, int flags, double* eloc_out, double* logp_out, double* ratios_out,
                           void* ws, size_t ws_bytes, void* stream);

/* Which chain kernel rnnwf_tfim_eloc runs for `m`: 3 = tcgen05 kind::f16 with 3xFP16 operands, resident weights and the
 * MMA / gate-math software pipeline (FP32 probability-head GRU with 50 units, <= 3 layers), 2 = the same arithmetic without the
 * pipeline, 1 = tcgen05 kind::tf32 with 3xTF32 operands, 0 = CUDA-core FFMA tile engine (every other shape / dtype).  The
 * environment variable RNNWF_CHAIN = ffma | tc32 | tc16 | tc16p overrides for A/B measurements. */
int rnnwf_tfim_chain_mode(const rnnwf_model* m);

/* Diagonal part only (bit-exact with the reference's f64 accumulation order, :31-38 / 2-D :33-49). */
int rnnwf_tfim_diag(const rnnwf_model* m, const uint8_t* samples, int64_t ns, const double* jz,
                    double* diag_out, void* stream);

/* Materialise the reference's queue: int32 [(N+1), ns, N]; slot 0 = samples, slot i+1 = site i flipped
 * (1DTFIM/TrainingRNN_1DTFIM.py:40-48; 2-D slot order i*Ny+j+1, Training2DRNN_2DTFIM.py:55-61).
 * Only used for API compatibility / enumeration tests: the fused path never builds the queue.           */
int rnnwf_tfim_enumerate(const uint8_t* samples, int64_t ns, int32_t n_sites, int32_t* queue_out, void* stream);

/* ---- J1-J2 ---------------------------------------------------------------------------------------------
 * rnnwf_j1j2_enumerate replaces J1J2MatrixElements / J1J2Slices (J1J2/TrainingRNN_J1J2.py:12-93, :95-127):
 * fixed-slot layout, row r of sample s lives at s*(2N+1)+r; order: diagonal, NN exchanges (ascending site),
 * NNN exchanges (ascending site).  counts_out[s] = number of valid rows (the reference's `num`).
 * sigmas_out: int32 [ns, 2N+1, N] or NULL; elements_out: float [ns, 2N+1].                               */
int rnnwf_j1j2_enumerate(const uint8_t* samples, int64_t ns, int32_t n_sites, const double* j1, const double* j2,
                         const double* bz, int periodic, int marshall_sign, int32_t* sigmas_out,
                         float* elements_out, int32_t* counts_out, void* stream);

/* Fused local energies E_s = sum_{s'} H_{ss'} exp(log psi(s') - log psi(s)) (J1J2/TrainingRNN_J1J2.py:255-279)
 * for open chains; exchanges are evaluated with prefix reuse.  eloc_out: double[2*ns] (re, im);
 * logpsi_out: double[2*ns] or NULL.                                                                      */
int rnnwf_j1j2_eloc(const rnnwf_model* m, const void* params, const uint8_t* samples, int64_t ns,
                    const double* j1, const double* j2, const double* bz, int marshall_sign,
                    double* eloc_out, double* logpsi_out, void* ws, size_t ws_bytes, void* stream);

/* ---- K3: VMC gradient -----------------------------------------------------------------------------------
 * Replaces optimizer.compute_gradients(cost) (1DTFIM/TrainingRNN_1DTFIM.py:156-160; complex form
 * J1J2/TrainingRNN_J1J2.py:197-201): grad = sum_s [ w_re[s] * d(Re log psi_s) + w_im[s] * d(Im log psi_s) ].
 * HEAD_PROB: weights double[ns] (w_s = (E_s - mean E)/ns).  HEAD_COMPLEX: weights double[2*ns] (re, im) with
 * w = 2 (E_s - mean E)/ns.  With RNNWF_PARITY_SYM the weight is applied to the symmetrised log-probability.
 * grad_out: double[P] in parameter order (always f64; cast by the caller).                                */
int rnnwf_vmc_grad(const rnnwf_model* m, const void* params, const uint8_t* samples, int64_t ns,
                   const double* weights, int flags, double* grad_out, void* ws, size_t ws_bytes, void* stream);

/* ---- TF1 Adam (tf.train.AdamOptimizer.apply_gradients, 1DTFIM/TrainingRNN_1DTFIM.py:113,164) -----------
 * theta, m, v: dtype of the model; grad: double[P]; step t is 1-based (after increment).                 */
int rnnwf_adam_step(int dtype, int64_t n, void* theta, void* mom, void* vel, const double* grad, double grad_scale,
                    double lr, double beta1, double beta2, double eps, int64_t t, void* stream);

/* Mean / population variance of E_loc: stats_out = {sum, sum of squares, count} (double[3]).             */
int rnnwf_energy_moments(const double* eloc, int64_t ns, int stride, double* stats_out, void* stream);

/* ---- measurement hooks (bench.py) --------------------------------------------------------------------------
 * Between rnnwf_profile_begin() and rnnwf_profile_end() the library counts its kernel launches and brackets
 * the dominant kernel of each call (the prefix-reuse chain kernel of rnnwf_tfim_eloc / rnnwf_j1j2_eloc) with
 * CUDA events on the caller's stream.  rnnwf_profile_end() waits for those events and returns the launch count,
 * the number of dominant-kernel launches and their summed device time.  Results of the compute entry points
 * are unaffected; the counters are per calling host thread (like rnnwf_last_error), nothing is shared between threads.
 * (The reference has only commented-out time.time() prints, 1DTFIM/TrainingRNN_1DTFIM.py:53-54.) */
int rnnwf_profile_begin(void);
int rnnwf_profile_end(int64_t* launches_out, int64_t* dominant_launches_out, double* dominant_ms_out);

/* FP32 FFMA throughput of the device (TFLOP/s) measured with a register-resident FMA kernel on `stream`:
 * the denominator of the compute roofline the recurrence kernels are held against.        */
int rnnwf_ffma_peak(int iters, double* tflops_out, void* stream);

/* FP64 throughput of the device (TFLOP/s): mode 0 = DFMA (CUDA cores), mode 1 = mma.sync.m8n8k4.f64 (DMMA).  The roofline
 * denominators of the float64 models (2DTFIM_1DRNN/RNNwavefunction.py:26-38, 2DTFIM_2DRNN/MDRNNcell.py:21-35 are float64). */
int rnnwf_fp64_peak(int mode, int iters, double* tflops_out, void* stream);

/* Known-answer check of the tcgen05 / TMEM plumbing the tensor-core recurrence kernels are built on:
 * d[128][n] = a[128][k] * b[n][k]^T on one CTA (A staged in TMEM, B in shared memory, `passes` = 1: TF32, 3: 3xTF32;
 * passes = -(p + 4*dcol): the FP16 path with p in {1, 3} passes and the accumulator at TMEM column dcol). */
int rnnwf_umma_selftest(int n, int k, const float* a, const float* b, float* d, int passes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RNNWF_H */
