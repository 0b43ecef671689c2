/*
 * bagpu.h -- C ABI of libbagpu.so, the B200-native (sm_100a, FP64) bundle-adjustment hot path.
 *
 * Drop-in boundary for CelestineAngla/BundleAdjustment.jl: every entry point replaces one call
 * the reference makes on its BALNLPModel / Levenberg_Marquardt surface (file:line below are
 * relative to the reference repository).  Julia reaches them with `ccall`;
 * plain pointers and sizes only, no exceptions cross the boundary: every function returns an
 * int status (0 = ok) and ba_last_error() gives the message.
 *
 * Layout contract (identical to the reference, src/ReadFiles.jl:9-53, src/BALNLPModels.jl:79-88):
 *   x      = [X_1..X_npnts (3 each) ; C_1..C_ncams (9 each)],  C = (r1 r2 r3 t1 t2 t3 k1 k2 f)
 *   cam_idx, pnt_idx : 1-based Int64, one per observation;  pt2d : interleaved (x, y)
 *   cx     : 2*nobs  (observation k at 2k-1, 2k, 1-based)
 *   rows, cols, vals : 24*nobs COO triplets, per observation row 1 (12) then row 2 (12),
 *                      column order [X(3) r(3) t(3) k1 k2 f]; rows/cols 1-based Int64.
 * Host-pointer calls copy in/out on the handle's stream and return when the result is in the
 * caller's buffer.  *_dev variants take device pointers (same layouts) and only enqueue work.
 * A handle is not re-entrant: one call at a time per handle.
 */
#ifndef BAGPU_H
#define BAGPU_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#define BA_API __attribute__((visibility("default")))
#else
#define BA_API
#endif

typedef struct ba_handle ba_handle;

enum {
  BA_OK = 0,
  BA_ERR_ARG = 1,       /* bad argument (null pointer, index out of range, ...) */
  BA_ERR_CUDA = 2,      /* CUDA runtime failure, see ba_last_error */
  BA_ERR_UNSORTED = 3,  /* LM entry points need point-major observation order (BAL file order) */
  BA_ERR_COMM = 4,      /* NCCL failure / communicator not initialised */
  BA_ERR_NUMERIC = 5    /* PCG breakdown (non-finite or non-positive curvature), non-positive Cholesky pivot */
};

/* How the damped system (J'J + lambda I) delta = -J'r is solved once the points are eliminated.  The reference's
 * facto x perm choices (src/lm.jl:15-19: :LDL / :QR with AMD / Metis) all give the exact solution of that system;
 * BA_SOLVER_EXACT is their counterpart: the reduced camera system assembled explicitly and factorised by a dense
 * FP64 Cholesky (what ldl_factorize + ldl_solve!, src/ldl_aux.jl:122-201,4-42, amount to after the ordering has
 * eliminated residual rows and points) plus refinement steps with the matrix-free FP64 residual.  BA_SOLVER_PCG is
 * the matrix-free preconditioned CG (stopped at pcg_tol).  AUTO: a dense solve up to 2048 cameras (MIXED from 8192
 * camera unknowns on -- except on 8 or more ranks with peer access, where the distributed FP64 factorisation of EXACT
 * is the faster one -- EXACT below), PCG above.
 * BA_SOLVER_MIXED is the reference's mixed-precision mode (src/lm.jl:92-98,165-173: facto_type below the model type,
 * "factorise in Float32, everything else in Float64"): the same explicit reduced camera
 * system, factorised in FP32 storage on the tensor cores (three TF32 MMAs per product: FP32-level accuracy), and that
 * factor preconditions FP64 CG on the FP64 matrix-free operator -- a handful of iterations to pcg_tol, so the step
 * keeps the FP64 accuracy of the other solvers.  When the FP32 factor is no usable preconditioner (non-positive
 * pivot, or no convergence in 30 iterations) the solve falls back to the FP64 factorisation (ba_lm_stats.mixed_fallbacks). */
enum { BA_SOLVER_AUTO = 0, BA_SOLVER_PCG = 1, BA_SOLVER_EXACT = 2, BA_SOLVER_MIXED = 3 };

/* ---- model lifetime: BALNLPModel(filename) ctor, src/BALNLPModels.jl:91-106 ---------------- */
/* Copies indices and pt2d to the device (caller buffers are not retained).  device = CUDA
 * ordinal.  The handle owns a stream and all device storage; free with ba_destroy (from a
 * Julia finalizer). */
BA_API int ba_create(int64_t ncams, int64_t npnts, int64_t nobs, const int64_t* cam_idx_1based,
              const int64_t* pnt_idx_1based, const double* pt2d, int device, ba_handle** out);
/* Observation-sharded variant for one-process-per-GPU runs: rank r of nranks keeps a contiguous
 * range of observations cut on point boundaries (needs point-major order).  All arrays passed
 * are the FULL problem; outputs of the per-observation calls are the LOCAL slices, whose global
 * ranges ba_shard_range reports (0-based, half-open). */
BA_API int ba_create_sharded(int64_t ncams, int64_t npnts, int64_t nobs, const int64_t* cam_idx_1based,
                      const int64_t* pnt_idx_1based, const double* pt2d, int device, int rank,
                      int nranks, ba_handle** out);
/* Several GPUs behind ONE handle in ONE process -- the mode a single-process caller (the reference's Julia session,
 * src/main.jl:27-30) uses: `ngpus` devices (devices[0..ngpus), or 0..ngpus-1 when devices is NULL; ngpus <= 0 = all
 * visible devices).  The library shards the observations over the devices exactly like ba_create_sharded, runs one
 * host thread per device, sets up NCCL and the peer-memory exchange itself, and presents FULL-LENGTH arrays to the
 * caller: every host-pointer entry point (ba_residual ... ba_jtprod, ba_lm_step, ba_lm_solve) behaves as on a
 * single-GPU handle.  Device-pointer variants, ba_set_stream and ba_set_profiling are per GPU and return BA_ERR_ARG. */
BA_API int ba_create_multi(int64_t ncams, int64_t npnts, int64_t nobs, const int64_t* cam_idx_1based,
                           const int64_t* pnt_idx_1based, const double* pt2d, int ngpus, const int* devices,
                           ba_handle** out);
BA_API int ba_destroy(ba_handle* h);
BA_API int ba_shard_range(const ba_handle* h, int64_t* obs0, int64_t* obs1, int64_t* pnt0, int64_t* pnt1);
/* Host-side partition used by ba_create_sharded (no GPU needed): cuts[r]..cuts[r+1] is rank r's
 * observation range; cuts has nranks+1 entries.  Returns BA_ERR_UNSORTED if not point-major. */
BA_API int ba_partition_observations(int64_t nobs, const int64_t* pnt_idx_1based, int nranks, int64_t* cuts);
BA_API const char* ba_last_error(const ba_handle* h);
BA_API const char* ba_version(void);
/* FP64 FMA throughput of `device` in TFLOP/s, measured with a register-resident FMA kernel (the secondary ceiling
 * of this path next to HBM bandwidth; not part of the reference's surface). */
BA_API int ba_measure_fp64_peak(int device, double* tflops);
/* The same for the FP64 tensor-core path (mma.sync.m8n8k4.f64, DMMA): the ceiling of the dense Cholesky of
 * BA_SOLVER_EXACT. */
BA_API int ba_measure_fp64_mma_peak(int device, double* tflops);
/* Run this handle's work on an existing stream (cudaStream_t passed as void*), e.g. the
 * caller's current stream so that its own CUDA events bracket the kernels. */
BA_API int ba_set_stream(ba_handle* h, void* cuda_stream);
/* Pinned host buffers for callers that want full PCIe rate on the host-pointer calls. */
BA_API int ba_alloc_pinned(uint64_t bytes, void** out);
BA_API int ba_free_pinned(void* p);

/* ---- NLPModels surface ---------------------------------------------------------------------- */
/* NLPModels.cons!(nlp, x, cx), src/BALNLPModels.jl:115-122 (== residual! of FeasibilityResidual,
 * call sites src/lm.jl:39,252,268).  NaN/Inf are left in cx like the reference. */
BA_API int ba_residual(ba_handle* h, const double* x, double* cx);
/* NLPModels.jac_structure!(nlp, rows, cols), src/BALNLPModels.jl:125-158 (src/lm.jl:53). */
BA_API int ba_jac_structure(ba_handle* h, int64_t* rows, int64_t* cols);
/* NLPModels.jac_coord!(nlp, x, vals), src/BALNLPModels.jl:161-206 (src/lm.jl:54,341). */
BA_API int ba_jac_coord(ba_handle* h, const double* x, double* vals);
/* cons! + jac_coord! in one pass over the observations (the headline metric). */
BA_API int ba_residual_jac(ba_handle* h, const double* x, double* cx, double* vals);
/* Jv = J(x) v  (2*nobs)  and  Jtv = J(x)' v  (nvar): the products mul_sparse!(…) forms from
 * (rows, cols, vals), src/lma_aux.jl:194-212, call sites src/lm.jl:57,356,370; here matrix-free. */
BA_API int ba_jprod(ba_handle* h, const double* x, const double* v, double* Jv);
BA_API int ba_jtprod(ba_handle* h, const double* x, const double* v, double* Jtv);

/* device-pointer variants (x_dev has the full nvar layout; outputs are local slices) */
BA_API int ba_residual_dev(ba_handle* h, const double* x_dev, double* cx_dev);
BA_API int ba_jac_structure_dev(ba_handle* h, int64_t* rows_dev, int64_t* cols_dev);
BA_API int ba_jac_coord_dev(ba_handle* h, const double* x_dev, double* vals_dev);
BA_API int ba_residual_jac_dev(ba_handle* h, const double* x_dev, double* cx_dev, double* vals_dev);
BA_API int ba_jprod_dev(ba_handle* h, const double* x_dev, const double* v_dev, double* Jv_dev);
BA_API int ba_jtprod_dev(ba_handle* h, const double* x_dev, const double* v_dev, double* Jtv_dev);
BA_API int ba_sync(ba_handle* h);
/* PCG preconditioner: block-Jacobi (exact 9x9 diagonal blocks of the reduced camera system) plus an additive
 * coarse level over `n` clusters of consecutive cameras (piecewise-constant interpolation of the six pose
 * components, n <= 24; default 16; 0 = plain block-Jacobi).  Changes the iteration count, not the solution. */
BA_API int ba_set_coarse_clusters(ba_handle* h, int n);
/* PCG deflation: extend the coarse level by up to `k` (<= 32; default 32; 0 = off) base vectors plus up to 16
 * refreshed ones, all harvested from the PCG solves themselves (CG is a Lanczos process: Ritz vectors of the
 * preconditioned reduced camera system come without extra products; ba_lm.cu).  Only used
 * for camera systems above the single-CTA threshold.  Changes the iteration count, not the solution. */
BA_API int ba_set_deflation(ba_handle* h, int k);
/* Choose the damped solve (BA_SOLVER_*); takes effect from the next LM call on.  Env BAGPU_SOLVER=pcg|exact|mixed
 * sets the default of new handles. */
BA_API int ba_set_solver(ba_handle* h, int solver);
/* Outcome of the last damped solve on this handle: solver used (BA_SOLVER_PCG / BA_SOLVER_EXACT / BA_SOLVER_MIXED;
 * EXACT after a mixed solve fell back), whether it converged (PCG: reached pcg_tol before pcg_max_iter; exact:
 * factorisation succeeded), the relative residual it stopped at (PCG: sqrt(r'M^-1 r / r0'M^-1 r0); exact:
 * ||b - S x|| / ||b|| of the direct solve, measured matrix-free before the refinement step; mixed: the same true
 * residual of the returned solution) and its iteration count (PCG / CG iterations, refinement steps). */
BA_API int ba_last_solve_info(const ba_handle* h, int32_t* solver, int32_t* converged, double* rel, int32_t* iters);
/* Profiling: with it on, the per-observation evaluation kernel (k_eval: cons!/jac_coord!/fused) is
 * bracketed by CUDA events on the handle's stream and ba_last_eval_ms returns its device time alone
 * (waits for that kernel); for roofline reporting.  Off by default: the events would sit between the
 * camera-precompute kernel and its programmatic dependent launch. */
BA_API int ba_set_profiling(ba_handle* h, int on);
BA_API int ba_last_eval_ms(ba_handle* h, float* ms);

/* ---- Levenberg-Marquardt, src/lm.jl:15-418 --------------------------------------------------- */
typedef struct ba_lm_params {
  double restol, satol, srtol, oatol, ortol, atol, rtol; /* src/lm.jl:21-24 */
  double nu_d, nu_m, lambda, delta_d;                    /* src/lm.jl:25 */
  int64_t ite_max;                                       /* src/lm.jl:26 */
  int32_t linesearch;                                    /* positional arg, src/lm.jl:19 */
  int32_t pcg_max_iter;                                  /* cap per damped solve */
  double pcg_tol;  /* stop when sqrt(r'M^-1 r / r0'M^-1 r0) <= pcg_tol */
  int32_t solver;  /* BA_SOLVER_*; AUTO = the handle's setting (ba_set_solver) */
  int32_t reserved;
} ba_lm_params;

typedef struct ba_lm_row { /* one log_row of src/lm.jl:304, plus solver counters */
  int64_t iter;
  double f, df, dfeas, lambda, delta_norm, rho;
  int32_t accepted; /* branch taken at src/lm.jl:306 */
  int32_t acc_str;  /* the logged "acc"/"rej" string rule of src/lm.jl:260 */
  int32_t pcg_iters; /* PCG iterations (refinement steps of the exact solve) */
  int32_t ntimes;   /* back-tracking halvings, src/lm.jl:262-295 */
  int32_t solver;   /* BA_SOLVER_PCG / BA_SOLVER_EXACT */
  int32_t converged; /* 0: the solve stopped at pcg_max_iter with solve_rel > pcg_tol (the step is inexact) */
  double solve_rel; /* relative residual the solve stopped at (see ba_last_solve_info) */
} ba_lm_row;

typedef struct ba_lm_stats {
  int32_t status; /* 0 unknown 1 small_step 2 first_order 3 small_residual 4 acceptable
                     5 neg_pred 6 exception 7 max_iter   (precedence of src/lm.jl:391-405) */
  int32_t pad;
  int64_t iter;
  double objective, dual_feas, lambda_final, elapsed_s;
  int64_t pcg_iters_total;
  double t_eval_ms, t_assemble_ms, t_pcg_ms, t_backsub_ms; /* CUDA-event phase totals */
  int64_t capped_solves;   /* damped solves that hit pcg_max_iter before pcg_tol (their steps are inexact) */
  double worst_solve_rel;  /* largest solve_rel over the iterations */
  double t_prepare_ms;     /* one-off schedule construction + allocations (first LM call on a handle) */
  /* BA_SOLVER_EXACT only (both are part of t_assemble_ms): explicit assembly of the reduced camera system, and
   * its dense Cholesky factorisations (chol_n^3 / 3 flops each, chol_count of them) */
  double t_schur_ms, t_chol_ms;
  int64_t chol_n, chol_count;
  int64_t mixed_fallbacks; /* BA_SOLVER_MIXED: damped solves that needed the FP64 factorisation after all */
} ba_lm_stats;

typedef void (*ba_iter_cb)(const ba_lm_row* row, void* user);

BA_API void ba_lm_default_params(ba_lm_params* p);
/* One damped solve (J'J + lambda I) delta = -J'r at x (what ldl_factorize + ldl_solve! /
 * myqr + solve_qr! deliver, src/lm.jl:138-152,175-229): delta (nvar), dr2 = 1/2 ||J delta + r||^2,
 * optional jtr (nvar, may be NULL) = J'r, obj = 1/2 ||r||^2. */
BA_API int ba_lm_step(ba_handle* h, const double* x, double lambda, double pcg_tol, int32_t pcg_max_iter,
               double* delta, double* dr2, double* obj, double* jtr, int32_t* pcg_iters);
/* Levenberg_Marquardt(model, facto, perm, normalize, linesearch; x, tolerances...) on device;
 * x_inout: x0 in, solution out (GenericExecutionStats.solution). */
BA_API int ba_lm_solve(ba_handle* h, double* x_inout, const ba_lm_params* p, ba_lm_stats* st,
                ba_iter_cb cb, void* user);

/* ---- multi-GPU (one process per GPU): NCCL communicator over NVLink/NVSwitch ---------------- */
/* rank 0 calls ba_comm_unique_id and broadcasts the 128 bytes by any means (torch.distributed,
 * MPI, a file); every rank then calls ba_comm_init on its sharded handle. */
BA_API int ba_comm_unique_id(uint8_t id128[128]);
BA_API int ba_comm_init(ba_handle* h, const uint8_t id128[128]);
/* Optional peer-memory path for the one exchange that happens every PCG iteration (sum over ranks of the
 * camera-sized vector  sum_k B'w, 9*ncams doubles): each rank exports a mailbox with CUDA IPC
 * (ba_comm_ipc_export, 64 bytes), the handles of ALL ranks in rank order are passed to ba_comm_ipc_import,
 * and from then on the vector kernel that consumes the sum reads every peer's partial directly over
 * NVLink/NVSwitch (flag handshake, fixed rank order => bit-identical on all ranks) instead of calling
 * ncclAllReduce.  The other, per-LM-iteration collectives stay on NCCL. */
BA_API int ba_comm_ipc_export(ba_handle* h, uint8_t handle64[64]);
BA_API int ba_comm_ipc_import(ba_handle* h, const uint8_t* handles64_by_rank);
/* Return to NCCL for that exchange, e.g. when ba_comm_ipc_import failed on some rank (peer access
 * unavailable): every rank must take the same path, so call it on all ranks or on none. */
BA_API int ba_comm_ipc_disable(ba_handle* h);

/* Debug / benchmark entry of the dense FP64 Cholesky used by BA_SOLVER_EXACT: factorises the caller's SPD matrix
 * (n x n, row-major, host) on `device`, solves A x = b, optionally returns L (n x n row-major, lower triangle
 * live) and the device times of the two phases.  Not part of the reference's surface. */
BA_API int ba_dbg_chol(int device, int64_t n, const double* A_rowmajor, const double* b, double* x, double* L_out,
                       float* factor_ms, float* solve_ms);
/* The same for the mixed-precision factor of BA_SOLVER_MIXED: A is rounded to FP32 and factorised in FP32 storage with
 * three-TF32-term tensor-core products; x = (L32 L32')^-1 b by FP64 sweeps over the FP32 factor (a preconditioner
 * application, accurate to about cond(A) * 1e-7); L_out receives the FP32 factor widened to doubles. */
BA_API int ba_dbg_chol32(int device, int64_t n, const double* A_rowmajor, const double* b, double* x, double* L_out,
                         float* factor_ms, float* solve_ms);

/* Development probe of the tensor / FMA rates the mixed-precision factorisation could build on (kind 0: TF32
 * mma.sync.m16n8k8, 1: BF16 mma.sync.m16n8k16, 2: FP32 FMA), TFLOP/s; scripts/probe_peaks.py. */
BA_API int ba_dbg_probe_peak(int device, int kind, double* tflops);

/* ---- host-only helpers of the PCG deflation space (no GPU; exported so that they can be unit-tested) ------ */
/* Eigenpairs of the Lanczos tridiagonal defined by the CG coefficients alpha[0..m), beta[0..m-1):
 * evals ascending (m), evecs column-major (m x m). */
BA_API int ba_dbg_tridiag_eig(const double* alpha, const double* beta, int32_t m, double* evals,
                              double* evecs_colmajor);
/* The k smallest eigenpairs only (evecs m x k column-major); matrices larger than full_below use QL for the
 * values and inverse iteration for the vectors. */
BA_API int ba_dbg_tridiag_smallest(const double* alpha, const double* beta, int32_t m, int32_t k, int32_t full_below,
                                   double* evals, double* evecs_colmajor);
/* From the Gram matrix (n x n, row-major) of n candidate vectors in preference order keep up to k that are
 * numerically independent (remainder >= tol of their norm) and return coefficients (n x kept, row-major) that
 * orthonormalise them. */
BA_API int ba_dbg_select_columns(const double* gram_rowmajor, int32_t n, int32_t k, double tol,
                                 double* coeff_rowmajor, int32_t* kept);

#ifdef __cplusplus
}
#endif
#endif /* BAGPU_H */
