/*
 * gogp_b200 -- C-ABI of the B200-native GP hot path.
 *
 * This is the drop-in boundary for GoGP's gp.GP (reference gp/gp.go): the Go
 * package keeps its public surface (gp.GP{NDim, Simil, Noise}, Observe,
 * Gradient, Absorb, LML, Produce; gp.Model) and reaches the GPU only through
 * the entry points below (cgo binding: INTEGRATION.md, go/gp/gp.go).
 *
 * Conventions
 *   - plain C, no exceptions cross the boundary; every call returns a
 *     gogp_status; gogp_last_error(h) has the text of the last failure;
 *   - every pointer is a HOST pointer to float64 data owned by the caller;
 *     nothing is retained after return (cgo pointer rules);
 *   - inputs X, Z are row-major, one point per row (the flattening of Go's
 *     [][]float64);
 *   - a handle is bound to one CUDA device and owns one stream; handles are
 *     not thread-safe (neither is gp.GP) but distinct handles may be used
 *     concurrently from different OS threads (multi-start restarts);
 *   - there is no CPU fallback: without a CUDA device gogp_create fails with
 *     GOGP_CUDA_ERROR.
 */
#ifndef GOGP_B200_H
#define GOGP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    GOGP_OK = 0,
    GOGP_BAD_ARGUMENT = 1,          /* Observe's panic("len(x)"), gp/gp.go:398-400 */
    GOGP_NOT_POSITIVE_DEFINITE = 2, /* Factorize(K) == false, gp/gp.go:228-230 */
    GOGP_ILL_CONDITIONED = 3,       /* gonum's Condition error (cond > 1e16) from SolveVecTo, gp/gp.go:233-236: the
                                       factor, alpha and LML are complete, as there; estimated from below */
    GOGP_CUDA_ERROR = 4,
    GOGP_NCCL_ERROR = 5,
    GOGP_OUT_OF_MEMORY = 6,
    GOGP_NOT_READY = 7,             /* Gradient/LML/Produce before Observe/Absorb */
    GOGP_UNSUPPORTED = 8            /* kernel expression too large for the descriptor */
} gogp_status;

/*
 * Kernel-expression descriptor: a postfix program over a value stack.  It
 * replaces the per-element dynamic dispatch into gp.Kernel.Observe
 * (gp/gp.go:14-17, :110-111, :134-135) plus the AD backward pass
 * (model.Gradient, gp/gp.go:113,137): a user Simil written in Go cannot run on
 * the device, so the stock kernels of kernel/kernel.go and kernel/noise.go and
 * their scaled sums and products are described, not called.
 *
 * Parameter indices refer to the kernel's own theta vector (ThetaSimil for the
 * similarity program, ThetaNoise for the noise program); a parameter enters a
 * leaf as scale * theta[param] (tutorial/hyperpriors/kernel/kernel.go:24 uses
 * 10*x[p]).  `dim` selects the input coordinate the 1-D stock kernel acts on.
 */
typedef enum {
    GOGP_OP_CONST = 0,           /* push `constant` */
    GOGP_OP_PARAM = 1,           /* push scale[0] * theta[param[0]] */
    GOGP_OP_ADD = 2,             /* pop b, pop a, push a + b */
    GOGP_OP_MUL = 3,             /* pop b, pop a, push a * b */
    GOGP_OP_NORMAL = 4,          /* kernel.Normal.Cov(l, xa, xb)       kernel/kernel.go:23-26 */
    GOGP_OP_PERIODIC = 5,        /* kernel.Periodic.Cov(l, p, xa, xb)  kernel/kernel.go:44-47 */
    GOGP_OP_MATERN32 = 6,        /* kernel.Matern32.Cov(l, xa, xb)     kernel/kernel.go:70-73 */
    GOGP_OP_MATERN52 = 7,        /* kernel.Matern52.Cov as shipped: 5/3 == 1, kernel/kernel.go:89-92 */
    GOGP_OP_MATERN52_TEXTBOOK = 8, /* (1 + sqrt5 d + 5/3 d^2) exp(-sqrt5 d); not in the reference */
    GOGP_OP_EVENTS = 9           /* tutorial/events/kernel/kernel.go:33-44: push the discount of the first event
                                    (gogp_set_events) whose from- or to-boundary separates xa and xb on `dim`, else 1 */
} gogp_op_kind;

typedef struct {
    uint8_t kind;      /* gogp_op_kind */
    uint8_t dim;       /* input coordinate for leaves, 0 <= dim < ndim */
    int16_t param[2];  /* theta indices: [0] = length scale l (or the PARAM), [1] = period p */
    double scale[2];   /* constant multipliers of those parameters (1.0 when unused) */
    double constant;   /* GOGP_OP_CONST value */
} gogp_op;

/* The noise program may only use CONST, PARAM, ADD, MUL (every noise kernel the
 * reference ships is input-independent: kernel/noise.go:21-53).  UniformNoise is
 * PARAM(0) PARAM(0) MUL; ConstantNoise(c) is CONST(c*c) with ntheta_noise 0;
 * tutorial/anynoise's noise is CONST(1e-5) with ntheta_noise 1. */

typedef struct gogp_handle gogp_handle;

/* Phases reported by gogp_phase_times (milliseconds of device time, CUDA events,
 * for the most recent Observe/Absorb, Gradient and Produce). */
enum {
    GOGP_PHASE_UPLOAD = 0, /* host->device copies of theta/X/Y */
    GOGP_PHASE_BUILD = 1,  /* covariance build K(X,X)+noise           gp/gp.go:109-225 */
    GOGP_PHASE_POTRF = 2,  /* Cholesky                                 gp/gp.go:228 */
    GOGP_PHASE_SOLVE = 3,  /* alpha = K^-1 y, logdet, y.alpha          gp/gp.go:232-236,250-251 */
    GOGP_PHASE_POTRI = 4,  /* K^-1 from L (gradient only)              gp/gp.go:454,480 */
    GOGP_PHASE_TRACE = 5,  /* fused trace 0.5 tr((aa^T-K^-1) dK)       gp/gp.go:434-486 */
    GOGP_PHASE_PREDICT = 6,/* Produce                                  gp/gp.go:258-360 */
    GOGP_NPHASE = 7
};

/* gp.GP{NDim, Simil, Noise} (gp/gp.go:20-23).  noise == NULL / n_noise_ops == 0
 * selects the reference default ConstantNoise(1e-5) (gp/gp.go:43-48). */
gogp_status gogp_create(int ndim,
                        const gogp_op* simil, int n_simil_ops, int ntheta_simil,
                        const gogp_op* noise, int n_noise_ops, int ntheta_noise,
                        int device, gogp_handle** out);
void gogp_destroy(gogp_handle* h);

/* The event table of tutorial/events' Simil{Events} (tutorial/events/kernel/kernel.go:10-12): n rows of
 * (from, to, discount), n <= 16, used by GOGP_OP_EVENTS leaves. */
gogp_status gogp_set_events(gogp_handle* h, const double* events, int n);

/* Assigning the fields gp.X, gp.Y (tutorial/tutorial.go:114-115) for the
 * hyper-parameters-only mode of Observe.  X is N x ndim row-major. */
gogp_status gogp_set_data(gogp_handle* h, const double* X, const double* Y, int64_t N);

/* gp.GP.Observe (gp/gp.go:374-413).  log_theta holds ntheta_simil+ntheta_noise
 * log-scale hyper-parameters.  with_obs != 0 is the "inputs are inferred too"
 * layout: X (N x ndim) and Y (N) are the tail of Observe's argument and the
 * following gogp_gradient returns the full [theta | X | Y] gradient.  With
 * with_obs == 0, X/Y may be NULL (use the data set by gogp_set_data) or given
 * (equivalent to gogp_set_data followed by the call).  Returns the log marginal
 * likelihood in *lml.  GOGP_NOT_POSITIVE_DEFINITE is where the reference panics. */
gogp_status gogp_observe(gogp_handle* h, const double* log_theta, int with_obs,
                         const double* X, const double* Y, int64_t N, double* lml);

/* gp.GP.Gradient (gp/gp.go:418-499): gradient of the last Observe with respect
 * to its argument, layout [log theta_simil | log theta_noise | X flat | Y];
 * len must be ntheta (hyper-parameters only) or ntheta + N*(ndim+1) (with_obs).
 * K^-1 is computed here, lazily, so value-only Observe calls cost one Cholesky. */
gogp_status gogp_gradient(gogp_handle* h, double* grad, int64_t len);

/* gp.GP.Absorb (gp/gp.go:80-87): natural-scale parameters, no gradient. */
gogp_status gogp_absorb(gogp_handle* h, const double* theta_simil, const double* theta_noise,
                        const double* X, const double* Y, int64_t N);

/* The expanding window of tutorial.Evaluate (tutorial/tutorial.go:91-179): append m
 * observations to the absorbed ones WITH THE HYPER-PARAMETERS UNCHANGED.  The factor of the first N observations is
 * the leading block of the new one, so only the block rows from the last (partial) 128-tile on are rebuilt and
 * factored: O(N^2 (m + 128)) instead of O((N + m)^3); alpha and the LML follow.  Equivalent to gogp_absorb on all
 * N + m observations (results agree to rounding); needs the hyper-parameters-only form (no with_obs). */
gogp_status gogp_extend(gogp_handle* h, const double* Xnew, const double* Ynew, int64_t m, double* lml);

/* gp.GP.LML (gp/gp.go:244-253) of the absorbed observations; 0 when N == 0. */
gogp_status gogp_lml(gogp_handle* h, double* lml);

/* gp.GP.Produce (gp/gp.go:258-360): posterior mean and standard deviation of
 * the latent function at M points Z (M x ndim).  N == 0 gives the prior.  The
 * radicand is clamped at 0 (the reference takes sqrt of a possibly tiny
 * negative number, gp/gp.go:356; its own goldens expect 0 there). */
gogp_status gogp_produce(gogp_handle* h, const double* Z, int64_t M, double* mu, double* sigma);

/* The stored results the reference lets a user keep (gp/gp.go:255-257):
 * alpha (N) and, optionally, the N x N row-major lower Cholesky factor. */
gogp_status gogp_get_alpha(gogp_handle* h, double* alpha, int64_t N);
gogp_status gogp_get_factor(gogp_handle* h, double* L, int64_t N);

/* "Produce on stored results": the reference exports L and Alpha so that a user can keep them and restore them
 * before Produce (gp/gp.go:35-36, 255-257).  Restores what gogp_get_alpha / gogp_get_factor returned together
 * with the natural-scale parameters and the inputs; gogp_produce then works as after Absorb.  No Y is stored,
 * so gogp_lml and gogp_gradient report GOGP_NOT_READY until the next Observe / Absorb. */
gogp_status gogp_set_state(gogp_handle* h, const double* theta_simil, const double* theta_noise, const double* X,
                           int64_t N, const double* alpha, const double* L);

const char* gogp_last_error(const gogp_handle* h);
const char* gogp_status_string(gogp_status s);
gogp_status gogp_phase_times(const gogp_handle* h, double* ms /* GOGP_NPHASE */);

/* Number of kernels this library has launched on the handle's stream since
 * creation (bench.py's gpu_launches). */
int64_t gogp_launch_count(const gogp_handle* h);

/* Device-side stopwatch on the handle's stream (CUDA events): start records an
 * event after all work queued so far, stop records another, waits for it and
 * returns the milliseconds in between.  bench.py brackets its timed region with
 * these because the handle's stream is not torch's current stream. */
gogp_status gogp_timer_start(gogp_handle* h);
gogp_status gogp_timer_stop(gogp_handle* h, double* ms);

/* Per-launch accounting of the dominant kernel (the DMMA GEMM): while enabled,
 * every GEMM launch is bracketed by CUDA events.  gogp_profile_read waits for the
 * stream and returns, since the last enable, the summed device time of the GEMM
 * launches, the FP64 flops they executed (2*128*128*k per tile, triangular k
 * ranges counted as executed) and their number. */
gogp_status gogp_profile_enable(gogp_handle* h, int on);
gogp_status gogp_profile_read(gogp_handle* h, double* gemm_ms, double* gemm_flops, int64_t* gemm_launches);

/* ---- Device-level building blocks for the multi-GPU block-cyclic factorisation ------------
 * (orchestrated one process per GPU by gogp_b200/dist_chol.py with
 * torch.distributed/NCCL panel broadcasts).  Every pointer below is a DEVICE pointer on the
 * handle's device, `stream` is the caller's cudaStream_t (NULL: the legacy default stream), all sizes are
 * multiples of 128, calls are asynchronous on that stream. */

/* Upload inputs once: X (N x ndim, host) -> the handle's dimension-major device copy, zero-padded
 * to a multiple of `block` (a multiple of 128: the distribution block size). */
gogp_status gogp_dev_set_inputs(gogp_handle* h, const double* X, int64_t N, int64_t block);
/* One block of K(X,X)+noise for natural-scale parameters: rows [row0, row0+rows), columns
 * [col0, col0+cols) into out (leading dimension ld).  diagonal != 0 (row0 == col0, rows == cols):
 * lower tiles only, noise on the diagonal, identity beyond N; otherwise full rectangle, zero beyond N. */
gogp_status gogp_dev_cov_block(gogp_handle* h, const double* theta_simil, const double* theta_noise,
                               int64_t row0, int64_t rows, int64_t col0, int64_t cols, int diagonal,
                               double* out, int64_t ld, void* stream);
/* Cholesky of an n x n block in place (lower) + inverses of its 128 x 128 diagonal tiles into winv
 * ([n/128][128][128]); *info (device int, zero it first) gets base+1-based index of a bad pivot. */
gogp_status gogp_dev_potrf(gogp_handle* h, double* A, int64_t ld, int64_t n, double* winv, int* info, int base,
                           void* stream);
/* B (m x n, ldb) <- B L^-T with the factored block L (ldl) and its tile inverses. */
gogp_status gogp_dev_trsm(gogp_handle* h, double* B, int64_t ldb, int64_t m, const double* L, int64_t ldl,
                          int64_t n, const double* winv, void* stream);
/* C = beta C + alpha A B^T (m x n, inner k); lower != 0: only tiles on or below the diagonal. */
gogp_status gogp_dev_gemm(gogp_handle* h, double* C, int64_t ldc, const double* A, int64_t lda, const double* B,
                          int64_t ldb, int64_t m, int64_t n, int64_t k, double alpha, double beta, int lower,
                          void* stream);
/* The same GEMM on a window of one rank's local matrix of a pr x pc block-cyclic distribution with tb x tb tiles
 * per distribution block: tile (ti, tj) of C belongs to global block (r0 + pr (ti / tb), c0 + pc (tj / tb)) and is
 * computed only where it meets the lower triangle of the global matrix (tb == 0: no mask).  One launch per
 * trailing update instead of one per owned block row (gogp_b200/csrc/grid.hpp).  ktri != 0: A is upper
 * triangular with respect to its own origin (a diagonal block of L^-T), the k range of row tile ti starts at
 * ti * 128. */
gogp_status gogp_dev_gemm_bc(gogp_handle* h, double* C, int64_t ldc, const double* A, int64_t lda, const double* B,
                             int64_t ldb, int64_t m, int64_t n, int64_t k, double alpha, double beta, int tb, int r0,
                             int pr, int c0, int pc, int ktri, void* stream);
/* Pre-allocate the scratch gogp_dev_trsm needs for right-hand sides of up to `rows` rows, so that no allocation
 * happens between collectives. */
gogp_status gogp_dev_reserve(gogp_handle* h, int64_t rows);
/* out[0] = sum_{i < nvalid} log L_ii of a factored block (out: 4 doubles; [2], [3] = min, max L_ii). */
gogp_status gogp_dev_sumlogdiag(gogp_handle* h, const double* L, int64_t ld, int64_t nvalid, double* out,
                                void* stream);
/* Block forward-substitution pieces for z = L^-1 y:  acc[r] -= sum_c B[r][c] v[c]  (rows x cols block),
 * and  z = Lkk^-1 rhs  for one factored diagonal block (rhs is destroyed; n multiple of 128). */
gogp_status gogp_dev_gemv_sub(gogp_handle* h, const double* B, int64_t ld, int64_t rows, int64_t cols,
                              const double* v, double* acc, double* scratch, void* stream);
gogp_status gogp_dev_trsv(gogp_handle* h, const double* L, int64_t ld, const double* winv, double* rhs, double* z,
                          int64_t n, void* stream);

/* ---- hyper-parameter optimisation without leaving the process ----
 * The tutorial's MLE drivers (tutorial/tutorial.go:124-175) over the data set by gogp_set_data:
 * method 0 = the infer.Adam loop (:156-168), method 1 = L-BFGS as optimize.Minimize is used
 * (:131-155).  X and Y stay resident in HBM; per evaluation only the parameters go down and
 * LML + gradient come back.  The objective is LML(log_theta) + prior(log_theta), maximised;
 * `prior` may be NULL (plain MLE) or a host callback that returns the log prior at x and ADDS
 * its gradient to grad (gp.Model's Priors, gp/model.go:17-27; in Go an exported cgo function).
 * log_theta (ntheta_simil + ntheta_noise values) is updated in place.  A point where the
 * covariance is not positive definite is a rejected step, not an error; GOGP_NOT_POSITIVE_DEFINITE
 * is returned only if the starting point itself fails. */
typedef struct {
    int method;        /* 0 adam, 1 lbfgs */
    int max_iters;     /* ITERS */
    double threshold;  /* THRESHOLD: stop when every |gradient_i| is below it */
    double rate;       /* Adam RATE */
    double beta1, beta2, eps; /* Adam moments; 0 selects 0.9, 0.999, 1e-8 */
    int history;       /* L-BFGS pairs kept; 0 selects 15 */
} gogp_opt_settings;
typedef struct {
    int iters;         /* Adam steps / L-BFGS directions taken */
    int evals;         /* LML evaluations spent (Observe: build + Cholesky + solves) */
    double lml0;       /* objective at the starting point */
    double lml;        /* objective at the returned point */
    int converged;     /* 1: gradient threshold met */
    int grads;         /* gradient evaluations spent (K^-1 + trace); <= evals: a trial step the line search
                          rejects on its value alone never computes K^-1 */
} gogp_opt_result;
typedef double (*gogp_prior_fn)(void* ctx, const double* x, int64_t n, double* grad);
gogp_status gogp_optimize(gogp_handle* h, const gogp_opt_settings* settings, double* log_theta,
                          gogp_prior_fn prior, void* ctx, gogp_opt_result* result);

/* Pieces of the distributed K^-1 and gradient, same conventions:
 * out (n x n, upper triangular, same ld as L) = L^-T for one factored diagonal block; */
gogp_status gogp_dev_trtri_t(gogp_handle* h, const double* L, int64_t ld, int64_t n, const double* winv, double* out,
                             void* stream);
/* the fused gradient trace over one rows x cols block of K^-1 whose origin is element (row0, col0)
 * of the global matrix (only elements with global row >= global column count; alpha is the full
 * padded vector; inputs from gogp_dev_set_inputs).  acc[0 .. ntheta_simil] is accumulated into:
 * the similarity parameters' 0.5 tr(W dK/dlog theta) sums and, last, tr(W) of the block.
 * scratch: (rows/128)(cols/128)(ntheta_simil+1) doubles. */
gogp_status gogp_dev_trace_block(gogp_handle* h, const double* theta_simil, const double* alpha, const double* kinv,
                                 int64_t ld, int64_t row0, int64_t rows, int64_t col0, int64_t cols, double* acc,
                                 double* scratch, void* stream);
/* The same trace in ONE launch over a rank's whole local matrix (rows x cols at kinv, ld) of a pr x pc
 * block-cyclic distribution with tb x tb tiles per block, the first local block row / column being global block
 * r0 / c0; blocks above the global diagonal are skipped.  scratch: (rows/128)(cols/128)(ntheta_simil+1) doubles. */
gogp_status gogp_dev_trace_local(gogp_handle* h, const double* theta_simil, const double* alpha, const double* kinv,
                                 int64_t ld, int64_t rows, int64_t cols, int tb, int r0, int pr, int c0, int pc,
                                 double* acc, double* scratch, void* stream);
/* Host arithmetic on the (input-independent) noise program: variance and d variance / d log theta_n
 * at natural-scale theta_noise; the noise gradient is 0.5 tr(W) dlog[q] (gp/gp.go:133-150). */
gogp_status gogp_noise_eval(gogp_handle* h, const double* theta_noise, double* variance, double* dlog);

/* ---- gp.GP across the GPUs of one box: 2D block-cyclic K, NCCL inside the library ---------------------------
 * (SURVEY.md section 8e / 8b `gogp_create_grid`; BASELINE configs[4], N = 131072.)  The covariance matrix no longer
 * fits one GPU: K is cut into block x block pieces dealt round-robin to a pr x pc process grid, every rank builds
 * its own pieces from the replicated inputs, and Observe / Gradient (gp/gp.go:374-413, 418-499) run as a
 * right-looking block Cholesky plus one fused pass for V = L^-T and K^-1 = V V^T.  The panels travel over NVLink
 * peer memory, pulled by the copy engines so that no SM leaves the GEMMs (gogp_b200/csrc/peer_bcast.hpp; CUDA IPC
 * between processes); the small reductions, and the panels too when a peer mapping is unavailable, go through NCCL
 * (gogp_b200/csrc/grid.hpp, grid.cu).  libnccl.so.2 is bound at the first grid call (dlopen:
 * the copy already loaded in the process, e.g. PyTorch's, else the system one); a library without NCCL still
 * serves every single-GPU entry point, and a grid call then fails with GOGP_NCCL_ERROR.
 *
 * Two ways to form the grid:
 *   - gogp_create_grid: ONE process drives all GPUs (the Go host of north_star: one gp.GP value, one cgo call
 *     per Observe).  The library runs one host thread per device for the duration of each call.
 *   - gogp_grid_unique_id + gogp_grid_create_rank: SPMD, one process (or thread) per GPU, e.g. under
 *     torchrun; every rank makes the same calls with the same arguments, like MPI.  The 128-byte id is created
 *     by one rank and distributed by the host (any channel).
 * Results (LML, gradient) are replicated: every rank returns the same numbers.  pr * pc must equal the number of
 * ranks (pr = pc = 0: the default grid, pr >= pc, powers of two); block is a multiple of 128 (0: 2048).
 * Hyper-parameters-only mode (gp.X, gp.Y assigned as fields, tutorial/tutorial.go:114-115); the with_obs input
 * gradient and Produce stay single-GPU (gogp_observe / gogp_produce). */
typedef struct gogp_grid gogp_grid;
#define GOGP_GRID_ID_BYTES 128
enum {
    GOGP_GRID_BUILD = 0,  /* covariance build of the owned blocks */
    GOGP_GRID_FACTOR = 1, /* block-cyclic Cholesky */
    GOGP_GRID_SOLVE = 2,  /* z = L^-1 y, log det, LML */
    GOGP_GRID_SWEEP = 3,  /* V = L^-T and K^-1 = V V^T, one pass */
    GOGP_GRID_ALPHA = 4,  /* alpha = V z */
    GOGP_GRID_TRACE = 5,  /* fused gradient trace + all-reduce */
    GOGP_GRID_NPHASE = 6
};
gogp_status gogp_create_grid(int ndim, const gogp_op* simil, int n_simil_ops, int ntheta_simil,
                             const gogp_op* noise, int n_noise_ops, int ntheta_noise,
                             const int* devices, int ndev, int pr, int pc, int64_t block, gogp_grid** out);
gogp_status gogp_grid_unique_id(unsigned char id[GOGP_GRID_ID_BYTES]);
gogp_status gogp_grid_create_rank(int ndim, const gogp_op* simil, int n_simil_ops, int ntheta_simil,
                                  const gogp_op* noise, int n_noise_ops, int ntheta_noise,
                                  int device, int rank, int world, int pr, int pc, int64_t block,
                                  const unsigned char id[GOGP_GRID_ID_BYTES], gogp_grid** out);
void gogp_grid_destroy(gogp_grid* g);
/* gp.X, gp.Y (replicated on every rank); allocates the rank's share of K: about N^2 * 8 / (pr pc) bytes plus
 * two panels of N * block * 8 bytes.  GOGP_OUT_OF_MEMORY when it does not fit. */
gogp_status gogp_grid_set_data(gogp_grid* g, const double* X, const double* Y, int64_t N);
/* gp.GP.Observe, hyper-parameters only: log marginal likelihood at log_theta (ntheta_simil + ntheta_noise). */
gogp_status gogp_grid_observe(gogp_grid* g, const double* log_theta, double* lml);
/* gp.GP.Gradient of the last gogp_grid_observe with respect to log_theta; len = ntheta_simil + ntheta_noise. */
gogp_status gogp_grid_gradient(gogp_grid* g, double* grad, int64_t len);
/* gp.GP.Absorb with the data of gogp_grid_set_data (natural-scale parameters) and gp.GP.LML. */
gogp_status gogp_grid_absorb(gogp_grid* g, const double* theta_simil, const double* theta_noise);
gogp_status gogp_grid_lml(gogp_grid* g, double* lml);
/* alpha = K^-1 y (N values), available after gogp_grid_gradient. */
gogp_status gogp_grid_get_alpha(gogp_grid* g, double* alpha, int64_t N);
/* Device milliseconds per phase of the last Observe and Gradient (max over the ranks this process drives), and
 * the part of each phase the priority stream spent inside NCCL calls (it includes waiting for peers). */
gogp_status gogp_grid_phase_times(const gogp_grid* g, double* ms /* GOGP_GRID_NPHASE */,
                                  double* comm_ms /* GOGP_GRID_NPHASE or NULL */);
/* stats[0] bytes received by the first local rank through collectives since creation, [1] kernels launched by it,
 * [2] pr, [3] pc, [4] block, [5] bytes of device memory held by it, [6] NCCL version code, [7] ranks, [8] the part of
 * [0] that was pulled over peer memory by the copy engines (panel broadcasts) instead of NCCL, [9] device
 * milliseconds of the last Observe + Gradient on the slowest local rank (its own sum of phases). */
gogp_status gogp_grid_stats(const gogp_grid* g, double* stats /* 12 */);
const char* gogp_grid_last_error(const gogp_grid* g);

/* Test/diagnostic access to device state: what = 0 K (before factorisation is
 * not kept; returns the factor buffer), 1 L, 2 K^-1 (after gogp_gradient).
 * out is N x N row-major, lower triangle valid, upper mirrored. */
gogp_status gogp_debug_fetch(gogp_handle* h, int what, double* out, int64_t N);

/* Covariance build only (no factorisation): fills out (N x N row-major,
 * symmetric) with K(X,X)+noise for natural-scale parameters.  Test hook for
 * the build kernel and the descriptor interpreter. */
gogp_status gogp_debug_build(gogp_handle* h, const double* theta_simil, const double* theta_noise,
                             const double* X, int64_t N, double* out);

/* FP64 throughput microbenchmarks on the handle's device: which = 0 DMMA
 * (mma.sync m8n8k4 f64), 1 DFMA.  Returns TFLOP/s in *tflops. */
gogp_status gogp_debug_fp64_peak(gogp_handle* h, int which, double* tflops);

/* Device time of the 128 x 128 tile Cholesky+inverse kernel in microseconds (variant 0: the
 * shipped blocked kernel; 1: the first, column-by-column version kept as the timing baseline). */
gogp_status gogp_debug_leaf(gogp_handle* h, int variant, int iters, double* usec);

/* Test hook for the tile kernel alone: A (128 x 128 row-major, lower used) -> L (lower, upper
 * zeroed) and W = L^-1 (lower); *info = 0 or the 1-based index of the first non-positive pivot. */
gogp_status gogp_debug_leaf_run(gogp_handle* h, int variant, const double* A, double* L, double* W, int* info);

/* One C -= A B^T of size n (tiles of the trailing update), timed; returns
 * TFLOP/s.  mode 0 full, 1 lower-triangular (SYRK). */
gogp_status gogp_debug_gemm(gogp_handle* h, int64_t n, int64_t k, int mode, int iters, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* GOGP_B200_H */
