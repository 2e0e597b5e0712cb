/* rmhmc_b200.h -- C ABI of the B200-native batched RMHMC / HMC hot path.
 *
 * The reference (emilemathieu/RiemannHamiltonianMonteCarlo) has no FFI: its boundary is the
 * Python call convention of code/rmhmc.py:13 (RMHMC), code/hmc.py:12 (HMC) and
 * code/tools.py:10-74 (LogNormPDF, nextpow2, ac, CalculateESS), driven from code/main.py:49-53,71.
 * This header is what a ctypes binding of that path binds instead (INTEGRATION.md shows the stub);
 * each entry point names the reference lines it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, a negative RMHMC_E_* code on failure; the message is
 *     available from rmhmc_last_error(handle) (or rmhmc_last_error(NULL) for create failures);
 *   - nothing throws, nothing frees or allocates caller-visible memory;
 *   - all data pointers are DEVICE pointers on the handle's device (float64 unless stated),
 *     C-contiguous, never mutated when const; work is enqueued on the handle's CUDA stream and the
 *     call returns without synchronising unless stated;
 *   - one handle per (device, stream); a handle is not re-entrant, different handles may be used
 *     from different host threads;
 *   - D (number of parameters incl. intercept) must be <= 128: D <= 32 runs the warp-per-chain
 *     kernels the benchmark configurations use, 32 < D <= 128 a CTA-per-chain path.
 *
 * Library: librmhmc_b200.so, built for sm_100a only.  There is no CPU fallback.
 */
#ifndef RMHMC_B200_H
#define RMHMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rmhmc_handle rmhmc_handle;

enum {
    RMHMC_OK = 0,
    RMHMC_E_INVALID = -1,   /* bad argument */
    RMHMC_E_CUDA = -2,      /* CUDA runtime error (message has the detail) */
    RMHMC_E_STATE = -3,     /* call sequence error (e.g. advance before chains_init) */
    RMHMC_E_UNSUPPORTED = -4
};

const char* rmhmc_version(void);

/* Bind a data set: XX (n_rows x dim, row-major) and t (n_rows, values 0/1), alpha = prior variance
 * (the reference hard-codes alpha = 100: rmhmc.py:19, hmc.py:18).  The data are copied into the
 * handle's padded device layout; the caller's buffers are not referenced afterwards. */
int rmhmc_create(rmhmc_handle** out, int device, int64_t n_rows, int dim, double alpha,
                 const double* xx_dev, const double* t_dev);
void rmhmc_destroy(rmhmc_handle* h);
/* Re-upload XX / t of the same shape into the handle (chains keep their state); enqueued on the
 * handle's stream.  This is what a caller that owns host buffers does before each batch of rounds. */
int rmhmc_update_data(rmhmc_handle* h, const double* xx_dev, const double* t_dev);
const char* rmhmc_last_error(const rmhmc_handle* h);
/* How the engine evaluates the metric partials dG/dw_d = X^T diag(v (1-2p) x_d) X (rmhmc.py:64-77,142-161):
 *   TENSOR       build the packed symmetric tensor T (all D partials) per chain and contract it per chain --
 *                what the reference does, with the symmetry exploited (O(N D^3) per chain and leapfrog step);
 *   MATRIX_FREE  never form them: tr(G^-1 dG_d) = sum_n c_n x_nd (x_n^T G^-1 x_n) and
 *                p^T G^-1 dG_d G^-1 p = sum_n c_n x_nd (x_n . G^-1 p)^2 as passes over the data (O(N D^2));
 *                same numbers up to summation order.  Default when KR2(X)^T fits in device memory.
 * Frees the handle's chains (call before *_chains_init).  rmhmc_metric_partials always builds the tensor. */
enum { RMHMC_PARTIALS_TENSOR = 0, RMHMC_PARTIALS_MATRIX_FREE = 1 };
int rmhmc_set_partials_mode(rmhmc_handle* h, int mode);
int rmhmc_get_partials_mode(const rmhmc_handle* h);
/* How the engine contracts the Fisher metric G = X^T diag(v) X + I/alpha of every chain (rmhmc.py:51-57,
 * :116-119, :134-137; the position fixed-point iterates and the closing build of a leapfrog step):
 *   FP64_DMMA      fused FP64 kernel on the legacy tensor path (mma.sync DMMA.8x8x4), summation in FP64;
 *   INT8_TCGEN05   Blackwell tensor cores: v and KR2(X) are split into 5 (RMHMC_I8_SLICES=6: 6) balanced base-256
 *                  digits, the digit products accumulate EXACTLY in int32 TMEM accumulators (tcgen05.mma.kind::i8,
 *                  operands staged by tensor-map TMA) and are recombined in int64 / FP64.  Same G up to the digit
 *                  truncation (max relative error 1-5e-12 with 5 digits, 1e-14 with 6); gradient, log-likelihood and
 *                  c_n are evaluated in FP64 as before.  Default whenever the digit planes of KR2(X) (5 bytes per
 *                  entry) fit a third of the free device memory.  More than 16384 rows: K is split so that every int32
 *                  accumulation stays exact, partial sums added in FP64.  32 < dim: v goes through HBM (FP64 kernel ->
 *                  digit kernel -> GEMM) and needs the MATRIX_FREE partials (else the FP64 build runs).
 * Frees the handle's chains (call before *_chains_init).  The seam rmhmc_metric follows the mode. */
enum { RMHMC_METRIC_FP64_DMMA = 0, RMHMC_METRIC_INT8_TCGEN05 = 1 };
int rmhmc_set_metric_mode(rmhmc_handle* h, int mode);
int rmhmc_get_metric_mode(const rmhmc_handle* h);
/* Several stages have two kernel variants chosen by the chain count (few chains: variants that fill the SMs with
 * smaller tiles / row splits; many chains: the throughput variants the benchmark runs).  AUTO picks by count; SMALL /
 * LARGE pin the choice, e.g. to run the benchmark's kernels on a test-sized batch.  A chain's trajectory is
 * bit-reproducible within one regime; across regimes it differs by summation order (~1e-15). */
enum { RMHMC_REGIME_AUTO = 0, RMHMC_REGIME_SMALL = 1, RMHMC_REGIME_LARGE = 2 };
int rmhmc_set_launch_regime(rmhmc_handle* h, int regime);
/* Row-sharded data (very large N): every rank binds ITS rows with rmhmc_create and runs ALL chains;
 * each metric / partials build then ends in one NCCL all-reduce (sum) of the partial
 * G | X^T(t-p) | log-likelihood block resp. of T, after which the per-chain stages run replicated
 * (bit-identical on every rank).  rank 0 creates the id, the caller distributes its 128 bytes
 * (e.g. torch.distributed.broadcast), then every rank calls rmhmc_comm_init before chains_init.
 * NCCL (libnccl.so.2) is loaded at run time; not needed otherwise. */
int rmhmc_comm_unique_id(char* out128);
int rmhmc_comm_init(rmhmc_handle* h, int world, int rank, const char* id128);
/* Chain-sharded runs (BASELINE.json configs[3]: every rank owns its share of the chains, X replicated, no collective
 * on the data path): a communicator used ONLY to combine the end-of-run statistics.  Same id protocol as above. */
int rmhmc_stats_comm_init(rmhmc_handle* h, int world, int rank, const char* id128);
/* Combine statistics over all ranks of the statistics communicator (one ncclAllReduce; purely local without one):
 *   ess       (n_chains x dim, from blr_ess_batched / blr_ess_ragged; NaN = frozen chain counts as 0)
 *             -> ess_sum (dim): sum over ALL chains of all ranks (main.py:70-79 reports min/median/max of it)
 *   samples   (n_chains x n_samples x dim, strides in doubles) -> rhat (dim): Gelman-Rubin over ALL chains of all
 *             ranks (every rank passes the same n_samples)
 *   scalars   (n_scalars doubles, may be NULL): summed in place (leapfrog / iteration counters)
 * All pointers are device pointers; outputs are identical on every rank.  Synchronises the stream. */
int rmhmc_stats_gather(rmhmc_handle* h, const double* ess, int64_t n_chains, const double* samples, int64_t n_samples,
                       int64_t chain_stride, int64_t row_stride, double* ess_sum, double* rhat, double* scalars,
                       int n_scalars);
/* cudaStream_t to enqueue on (0 = legacy default stream). */
int rmhmc_set_stream(rmhmc_handle* h, void* cuda_stream);

/* ---- parity seams: the model pieces the reference inlines in rmhmc.py ----------------------- */

/* For each of n_chains positions theta[c] (n_chains x dim):
 *   G        (n_chains x dim x dim)  Fisher metric X^T diag(v) X + I/alpha     rmhmc.py:51-57
 *   grad     (n_chains x dim)        X^T (t - sigma(X theta)) - theta/alpha    rmhmc.py:100
 *   logjoint (n_chains)              f^T t - sum log(1+e^f) + log N(theta;0,alpha I)  rmhmc.py:31-34
 * Any output may be NULL.  Synchronises the stream. */
int rmhmc_metric(rmhmc_handle* h, int64_t n_chains, const double* theta, double* G, double* grad,
                 double* logjoint);

/* dG (n_chains x dim x dim x dim): dG[c][d] = X^T diag(v (1-2p) x_d) X         rmhmc.py:64-75
 * trace (n_chains x dim): tr(G^-1 dG_d)                                       rmhmc.py:76-77
 * Either may be NULL.  Synchronises the stream. */
int rmhmc_metric_partials(rmhmc_handle* h, int64_t n_chains, const double* theta, double* dG,
                          double* trace);

/* Lower Cholesky factor, inverse and sum(log(diag(L))) of n_chains dense SPD matrices
 * (np.linalg.cholesky / inv / the log-det of rmhmc.py:58-60,171).  Outputs may be NULL. */
int rmhmc_chol_logdet(rmhmc_handle* h, int64_t n_chains, const double* G, double* L, double* Ginv,
                      double* logdet);

/* n generalized leapfrog steps (rmhmc.py:96-163) from caller-supplied states, no randomness: for each
 * chain c, start at theta[c], momentum mom[c], integrate nsteps[c] steps in direction dir[c] (+1/-1)
 * with StepSize = step_size and NumOfNewtonSteps = n_fixed.  Outputs (any may be NULL): final
 * position / momentum (n_chains x dim) and the Hamiltonian (rmhmc.py:172,176) at the start and at
 * the end.  Re-initialises the handle's chains; synchronises. */
int rmhmc_leapfrog(rmhmc_handle* h, int64_t n_chains, const double* theta, const double* mom,
                   const int32_t* dir, const int32_t* nsteps, double step_size, int n_fixed,
                   double* out_theta, double* out_mom, double* out_h_start, double* out_h_end);

/* ---- the sampler engine: rmhmc.py:37-191 batched over independent chains -------------------- */

/* Allocate state for n_chains chains and evaluate it at theta0 (n_chains x dim; NULL = the
 * reference's start 1e-3, rmhmc.py:27).  Resets iteration counters. */
int rmhmc_chains_init(rmhmc_handle* h, int64_t n_chains, const double* theta0);

/* NumOfLeapFrogSteps, StepSize, NumOfNewtonSteps of rmhmc.py:13. */
int rmhmc_configure(rmhmc_handle* h, int n_leapfrog, double step_size, int n_fixed);

/* Randomness, one of:
 *  tape   -- host-supplied draws in the reference's consumption order (rmhmc.py:80,89,90,181) for
 *            iterations [it_base, it_base + n_window): z (n_window x n_chains x dim),
 *            u_step, z_dir, u_acc (n_window x n_chains).  u_acc is consumed only if Ratio <= 0.
 *  philox -- counter-based Philox4x32-10 keyed by (seed, chain_offset + chain, iteration). */
int rmhmc_set_tape(rmhmc_handle* h, int64_t it_base, int64_t n_window, const double* z,
                   const double* u_step, const double* z_dir, const double* u_acc);
int rmhmc_set_philox(rmhmc_handle* h, uint64_t seed, int64_t chain_offset);
/* Kinetic energy of the RMHMC engine.  GAUSSIAN: rmhmc.py (K = p' G^-1 p / 2).  STUDENT_T: the MATLAB original
 * authors_code/Bayes_Log_Reg/MCMC/BLR_RMHMC_StudentT.m:205-414 -- K = (1+D)/2 log(1 + p' G^-1 p), momentum drawn as
 * mvtrnd(G, 1)' = diag(G)^-1/2 L z / sqrt(chi2_1) (:265), LastTerm scaled by (1+D)/2 / (1 + p' G^-1 p) (:296,:370),
 * position update weights (1+D) / (1 + p' G^-1 p) (:311-326), no renormalisation hacks, samples stored from iteration
 * BurnIn on (:403-405).  Needs the MATRIX_FREE partials, dim <= 32, unsharded data; the implicit momentum iterates then run
 * as one pass + one per-chain kernel each.  MATLAB only: parity is against a port (oracle/blr_oracle.py), unpinned.
 * Under a host tape the chi-square draw is z_chi^2 with z_chi (n_window x n_chains) set by rmhmc_set_tape_chi after
 * rmhmc_set_tape. */
enum { RMHMC_MOMENTUM_GAUSSIAN = 0, RMHMC_MOMENTUM_STUDENT_T = 1 };
int rmhmc_set_momentum_family(rmhmc_handle* h, int family);
int rmhmc_set_tape_chi(rmhmc_handle* h, const double* z_chi);

/* Sample store (rmhmc.py:190-191): samples (n_chains x capacity x dim); the state after
 * iteration `it` goes to row it - burn_in for it > burn_in (row 0 is never written, as in the
 * reference).  NULL disables storing. */
int rmhmc_set_samples(rmhmc_handle* h, double* samples, int64_t capacity, int64_t burn_in);

/* Optional per-step trace of the first n_iters iterations (parity tests).  Shapes:
 * theta_steps (n_chains x n_iters x n_leapfrog x dim): position after each leapfrog step;
 * mom_end, theta_end, mom0 (n_chains x n_iters x dim); h_current, h_proposed (n_chains x n_iters);
 * flags int32 (n_chains x n_iters): bit0 accepted, bit1 uniform consumed, bit4 direction > 0,
 * bits 8.. RandomStep.  All-or-nothing: pass NULL for theta_steps to disable. */
int rmhmc_set_trace(rmhmc_handle* h, int64_t n_iters, double* theta_steps, double* mom_end,
                    double* theta_end, double* mom0, double* h_current, double* h_proposed,
                    int32_t* flags);

/* Enqueue n_rounds rounds.  One round = one generalized leapfrog step (rmhmc.py:96-163) for every
 * chain that has completed fewer than it_stop iterations, plus that chain's accept/reject and next
 * momentum draw when its trajectory ends.  Does not synchronise. */
int rmhmc_advance(rmhmc_handle* h, int64_t n_rounds, int64_t it_stop);

/* Run rounds until every chain has completed it_stop iterations; synchronises.  rounds_done (host
 * pointer, may be NULL) receives the number of rounds executed. */
int rmhmc_run(rmhmc_handle* h, int64_t it_stop, int64_t* rounds_done);

/* Copy out per-chain state (device pointers, any may be NULL): theta (n_chains x dim) current
 * position, iters / accepted / leapfrogs int64 (n_chains), renorm_mom / renorm_pos int32. */
int rmhmc_read_state(rmhmc_handle* h, double* theta, int64_t* iters, int64_t* accepted,
                     int64_t* leapfrogs, int32_t* renorm_mom, int32_t* renorm_pos);

/* A handle owns ONE chain set: every *_chains_init, rmhmc_leapfrog, rmhmc_set_partials_mode and rmhmc_set_metric_mode
 * replaces (or frees) it.  The generation counter changes whenever that happens, so that a caller holding a sampler
 * object can detect that its chains are gone (the Python samplers check it and raise). */
int64_t rmhmc_chain_generation(const rmhmc_handle* h);
/* Number of kernels this handle has launched so far (bench.py's gpu_launches). */
int64_t rmhmc_launch_count(const rmhmc_handle* h);

/* CUDA-event timing of the engine's kernels: when enabled, every launch of a given kernel class is
 * bracketed by events on the handle's stream.  kind: 0 metric build (position iterates),
 * 1 metric build (closing), 2 partials build (tensor mode), 3 per-chain turn (end of one leapfrog step +
 * start of the next; matrix-free: also the momentum iterates), 4 per-chain position solve / factorisation,
 * 5 quadratic-form pass, 6 leverage GEMM, 7 trace pass (matrix-free mode), 8 / 9 the two kernels of the INT8 metric
 * build of a position iterate (k_i8_vslice_mma: f = X theta and the digits of v; k_i8_gemm: tcgen05 digit GEMM, all its
 * launches; the sum of a build's two kernels is also counted under 0 / 1), 10 the NCCL all-reduces of the row-sharded
 * mode, 11 the digit kernel of the closing build (k_i8_vslice_mma_closing: also X^T (t - p), log-likelihood, c_n),
 * 12 k_i8_qdigits (digits of the packed G^-1 for the leverage GEMM).  Returns accumulated
 * milliseconds and launch count since the last reset; synchronises. */
int rmhmc_profile_enable(rmhmc_handle* h, int enable);
int rmhmc_profile_read(rmhmc_handle* h, int kind, double* ms, int64_t* launches);

/* ---- Euclidean HMC: hmc.py:38-89 batched over chains ---------------------------------------- */
int hmc_chains_init(rmhmc_handle* h, int64_t n_chains, const double* theta0);  /* NULL = zeros, hmc.py:27 */
int hmc_configure(rmhmc_handle* h, int n_leapfrog, double step_size);
/* tape without z_dir (hmc.py:41,48,78), or rmhmc_set_philox; samples via rmhmc_set_samples. */
int hmc_set_tape(rmhmc_handle* h, int64_t it_base, int64_t n_window, const double* z,
                 const double* u_step, const double* u_acc);
int hmc_run(rmhmc_handle* h, int64_t it_stop, int64_t* rounds_done);
/* n_rounds free-running rounds without synchronising; one HMC round = one leapfrog step (hmc.py:51-62) of every chain. */
int hmc_advance(rmhmc_handle* h, int64_t n_rounds, int64_t it_stop);

/* ---- manifold MALA / simplified manifold MALA -------------------------------------------------
 * MATLAB-only in the reference: code/authors_code/Bayes_Log_Reg/MCMC/BLR_mMALA.m:159-330 and
 * BLR_mMALA_Simp.m:170-290.  One iteration = one metric evaluation at the proposal
 *   w' = Mean + (z chol(eps G^-1))',  Mean = w + eps/2 G^-1 grad - eps sum_d G^-1 dG_d G^-1 e_d + eps/2 G^-1 tr
 * (simplified != 0: Mean = w + eps/2 G^-1 grad), accepted with the Metropolis-Hastings ratio of BLR_mMALA.m:282.
 * theta0 NULL = zeros (BLR_mMALA.m:165).  step_size is the reference's StepSize (= eps, default 1); it is fixed at
 * chains_init because the cached drift and proposal factor of the current state depend on it.  The tape holds
 * z (n_window x n_chains x dim) and u_acc (n_window x n_chains; consumed only if Ratio <= 0); rmhmc_set_philox,
 * rmhmc_set_samples (row it - burn_in for it >= burn_in, as in the MATLAB loop), rmhmc_set_trace (theta_end = the
 * proposal, h_current = log q(new|old), h_proposed = Ratio, flags bit0 accepted / bit1 uniform consumed) and
 * rmhmc_read_state apply.  dim <= 32.  The full drift runs on the MATRIX_FREE partials (the handle is switched). */
/* simplified == 2 selects the IWLS proposal of code/iwls.py:13-89 (Gamerman's iterated weighted least squares): the same
 * Metropolis-Hastings loop with proposal N(w + G^-1 grad, G^-1) -- algebraically cov . X^T W z of iwls.py:31-35 -- and the
 * density's log-determinant taken from chol(cov + 1e-6 I) as in iwls.py:64,68; step_size is ignored.  Pinned: the
 * unmodified iwls.py runs under a tape with np.random.multivariate_normal replaced by mean + chol(cov) z
 * (tests/golden/iwls_*.npz). */
int mmala_chains_init(rmhmc_handle* h, int64_t n_chains, const double* theta0, int simplified, double step_size);
int mmala_set_tape(rmhmc_handle* h, int64_t it_base, int64_t n_window, const double* z, const double* u_acc);
int mmala_run(rmhmc_handle* h, int64_t it_stop, int64_t* rounds_done);
/* Current proposal distribution N(mean, L L^T) of every chain: mean (n_chains x dim), L lower (n_chains x dim x dim);
 * either may be NULL.  Lets a host that owns the random stream (the single-chain iwls() drop-in) draw the proposal itself
 * and pass the engine z = L^-1 (w' - mean).  Synchronises. */
int mmala_read_proposal(rmhmc_handle* h, double* mean, double* chol_lower);
/* HMC launch shape (default fused = 1, 64 rounds per launch; env RMHMC_HMC_FUSED=0): fused runs up to rounds_per_launch
 * leapfrog rounds of hmc.py:51-62 per kernel launch with the chain state in registers (D <= 32, not row-sharded);
 * 0 = three launches per round.  The two can be mixed on one chain set.  rounds_per_launch = 0 keeps the current value. */
int hmc_set_fused(rmhmc_handle* h, int fused, int rounds_per_launch);
int mmala_advance(rmhmc_handle* h, int64_t n_rounds, int64_t it_stop);     /* one round = one iteration of every chain */

/* ---- tools.py:32-74 batched ------------------------------------------------------------------ */
/* ESS of every (chain, parameter) series: samples (n_chains x n_samples x dim) with the given
 * strides in doubles; ess (n_chains x dim).  Exactly tools.CalculateESS(series, max_lag) incl. the
 * nFFT = nextpow2(n)+1 circular aliasing; max_lag <= nFFT - 1 as in the reference.  Series of more than 24000 samples
 * use a global scratch copy instead of shared memory (and synchronise).  Does not need a handle. */
int blr_ess_batched(int device, void* cuda_stream, const double* samples, int64_t n_chains,
                    int64_t n_samples, int dim, int64_t chain_stride, int64_t row_stride,
                    int64_t max_lag, double* ess);

/* tools.ac (tools.py:21-30) for n_series contiguous series of n_samples: normalised circular
 * autocorrelation (period nextpow2(n)+1) at lags 0..n_lag; acf is (n_series x (n_lag+1)). */
int blr_autocorr(int device, void* cuda_stream, const double* series, int64_t n_series,
                 int64_t n_samples, int64_t n_lag, double* acf);

/* Ragged variant for free-running chains: chain c uses rows [starts[c], starts[c] + counts[c]) of its
 * block (int64 device arrays) with max_lag = counts[c] - 1; max_samples bounds counts[c]. */
int blr_ess_ragged(int device, void* cuda_stream, const double* samples, int64_t n_chains,
                   int64_t max_samples, int dim, int64_t chain_stride, int64_t row_stride,
                   const int64_t* starts, const int64_t* counts, double* ess);

/* Issue peaks of the two tensor paths of this library, measured live on `device` (host pointers, either may be NULL):
 * FP64 DMMA.8x8x4 in TFLOP/s and tcgen05.mma.kind::i8 (128x256x32, TMEM accumulators) in TOP/s.  bench.py's roofline
 * denominators (MEASURED_PEAKS.json has neither).  Synchronises. */
int blr_device_peaks(int device, void* cuda_stream, double* fp64_dmma_tflops, double* int8_tcgen05_tops);

/* Gelman-Rubin Rhat per parameter over n_chains >= 2 chains of n_samples >= 2 samples (same layout as above); rhat (dim).
 * Not in the reference (its main.py:70-79 only reports ESS); SURVEY.md 8c: classic estimator, W = mean of the chain
 * variances (ddof 1), B/S = variance of the chain means (ddof 1), Rhat = sqrt(((S-1)/S W + B/S) / W). */
int blr_rhat(int device, void* cuda_stream, const double* samples, int64_t n_chains, int64_t n_samples,
             int dim, int64_t chain_stride, int64_t row_stride, double* rhat);

#ifdef __cplusplus
}
#endif
#endif /* RMHMC_B200_H */
