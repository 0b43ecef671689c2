// ba_internal.h -- handle layout and launcher prototypes shared by the .cu files of libbagpu.
#pragma once
#include <cstdint>
#include <cstdio>
#include <functional>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/bagpu.h"
#include "ba_chol.h"

struct ncclComm;
struct ba_group;  // ba_group.cu: the per-device sub-handles of a multi-GPU handle

#define BA_CUDA(call)                                                                          \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) {                                                                   \
      char b_[512];                                                                            \
      snprintf(b_, sizeof b_, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      h->err = b_;                                                                             \
      return BA_ERR_CUDA;                                                                      \
    }                                                                                          \
  } while (0)

// Device-resident state of the LM solver (allocated on first use by ba::lm_prepare).
struct ba_lm_state {
  bool ready = false;
  void* d_slab = nullptr;  // the one device allocation behind every pointer lm_prepare hands out
  // ---- schedules built once per problem on the host -----------------------------------------
  int32_t* d_tstart = nullptr;   // point-major warp tasks: observation offsets, ntasks + 1
  int64_t ntasks = 0;
  int32_t* d_pstart = nullptr;   // first local observation of each local point, npl + 1
  int32_t* d_cperm = nullptr;    // camera-major position -> local observation id
  int32_t* d_ctask_beg = nullptr;  // camera-major tasks: [beg, end) positions of one camera
  int32_t* d_ctask_end = nullptr;
  int32_t* d_cam_t0 = nullptr;   // first task of each camera, ncams + 1
  int32_t* d_ctask_cam = nullptr;  // camera of each task
  int32_t* d_cam_cnt = nullptr;  // per camera: tasks finished in the current pass (ordered two-level sums)
  int64_t nctasks = 0;
  int32_t* d_empty_cams = nullptr;  // cameras without observations on this rank
  int64_t nempty = 0;
  // ---- per observation ------------------------------------------------------------------------
  double2* d_Jp = nullptr;   // 12 planes x nl, point-major: plane j = (row1[j], row2[j]) of the 2x12 block
  double2* d_F = nullptr;    // nl residuals
  int32_t* d_pntc = nullptr; // nl: point id of the observation at each camera-major position
  double2* d_x4 = nullptr;   // 2 x npnts: points of the current iterate padded to 32 bytes
  double2* d_w = nullptr;    // nl: per-observation 2-vector exchanged between the two passes
  double* d_T = nullptr;     // 3 planes x nl: A (V + lambda I)^-1 A^T (symmetric 2x2)
  double2* d_dr = nullptr;   // nl: -(J delta + r), only with linesearch (src/lm.jl:277-279)
  // ---- per local point ------------------------------------------------------------------------
  double* d_V = nullptr;     // 6: symmetric A'A
  double* d_gp = nullptr;    // 3: -A'F
  double* d_Vinv = nullptr;  // 6: (V + lambda I)^-1
  double* d_wp = nullptr;    // 3: Vinv gp
  // ---- per camera -----------------------------------------------------------------------------
  double* d_taskpart = nullptr;  // nctasks x 54 partial sums
  double* d_Ug = nullptr;    // ncams x 54: [U (45, symmetric) | gc (9)]
  double* d_Cr = nullptr;    // ncams x 54: [W Vinv W' diagonal block (45) | W Vinv gp (9)]
  double* d_H = nullptr;     // ncams x 81: U + lambda I
  double* d_Minv = nullptr;  // ncams x 81: inverse Schur diagonal block (block-Jacobi preconditioner)
  double* d_pcg = nullptr;   // 6 vectors of 9 ncams: b, xc, r, z, p, q
  double* d_pcgpart = nullptr;  // per-CTA partials of the two PCG dot products
  // two-level preconditioner (coarse space of piecewise-constant camera clusters)
  int ncl = 0, ctas_per_cluster = 1, mc = 0;  // clusters, vector-kernel CTAs (28 cameras) per cluster, 9 ncl
  double* d_Ac = nullptr;     // mc x mc: P' S P
  double* d_Aci = nullptr;    // its inverse
  double* d_yc = nullptr;     // mc: coarse correction of the current residual
  double* d_cpart = nullptr;  // 9 per vector-kernel CTA: restriction partials
  long long* d_Acq = nullptr; // mc x mc: fixed-point sums of the Schur part of P' S P (order-independent)
  double* d_cdiag = nullptr;  // mc: sqrt of the diagonal of P' (U + lambda I) P (normalisation of d_Acq)
  // deflation vectors harvested from the PCG solves (coarse space [P | Z], mc + kz <= 144 unknowns)
  int kz = 0, kz_base = 0;      // live columns of Z; the first kz_base come from the first long solve
  int kz_base_max = 0, kz_max = 0;  // limits for this problem (0: deflation not in use)
  int hcap = 0;                 // harvest capacity (Lanczos vectors)
  double* d_Z = nullptr;        // n9 x kz_max, column-major, Euclidean-orthonormal columns
  double* d_Zcand = nullptr;    // n9 x 64 candidates
  double* d_harv = nullptr;     // n9 x hcap: z_j / sqrt(r_j.z_j)
  double* d_hcoef = nullptr;    // [alpha (hcap) | beta (hcap)] of the harvested solve
  double* d_zpart = nullptr;    // per vector-kernel CTA: kz partial dot products Z_j . r
  double* d_dsmall = nullptr;   // hcap x 64 coefficient matrix / 64 x 64 Gram matrix
  double2* d_w4 = nullptr;      // 4 x nl: per-observation exchange of the 4-vector Schur product (coarse setup)
  double* d_q4 = nullptr;       // 4 x n9: its results
  int pcg_graph_kz = -1;        // kz the captured PCG graph was built for
  int z_gen = 0, coarse_gen = 0;  // version of Z, and the version the current Ac^-1 was built for
  int defl_iters_first = 0;     // PCG iterations of the solve the base vectors came from
  // ---- exact solve: explicit reduced camera system + dense Cholesky (ba_chol.cu) ---------------------
  bool exact = false;         // decided in lm_prepare from ba_handle::solver and the problem size
  bool mixed = false;         // exact, with the FP32 tensor-core factor as preconditioner of FP64 CG (BA_SOLVER_MIXED)
  bool factor64 = false;      // the current factor in d_S is the FP64 one (exact mode, or mixed after a fall-back)
  int64_t mixed_fallbacks = 0;  // damped solves of this LM run that needed the FP64 factorisation after all
  int64_t cn = 0;             // 9 ncams padded to a multiple of 128
  double* d_S = nullptr;      // cn x cn row-major: the scaled matrix, then its factor L
  long long* d_Sq = nullptr;  // packed lower-triangular tiles: fixed-point sums of the off-diagonal blocks
  double* d_Yh = nullptr;     // 27 per local observation: D_c^-1 (B'A) L_p (borrows d_S's storage when it fits)
  bool yh_aliased = false;
  double* d_cd = nullptr;     // 9 ncams: sqrt(diag(U + lambda I)), the Jacobi scaling
  double* d_ex = nullptr;     // 2 vectors of cn: scaled right-hand side / residual, scaled solution
  ba::chol_plan chol;
  bool attrs_set = false;     // cudaFuncSetAttribute is per device: done once per handle
  double t_schur_ms = 0.0, t_chol_ms = 0.0;  // accumulated phase times of the exact factor (CUDA events)
  int64_t chol_count = 0;
  // outcome of the last damped solve (ba_last_solve_info)
  int last_solver = 0, last_converged = 0, last_iters = 0;
  double last_rel = 0.0;
  // ---- iterates -------------------------------------------------------------------------------
  double* d_x = nullptr;      // current iterate (nvar; only this rank's point slice + cameras are live)
  double* d_xt = nullptr;     // trial iterate
  double* d_delta = nullptr;  // step
  double* d_camt = nullptr;   // camera records of the trial iterate
  // ---- reductions -----------------------------------------------------------------------------
  double* d_part = nullptr;   // per-block partial sums (deterministic two-level reductions)
  int64_t npart = 0;
  double* d_scal = nullptr;   // device scalars (layout in ba_lm.cu)
  double* h_scal = nullptr;   // pinned mirror
  void* pcg_graph = nullptr;  // cudaGraphExec_t: PCG_POLL iterations captured once (ba_lm.cu)
  double pcg_graph_tol = 0.0;
  bool pcg_graph_off = false;  // capture not possible on this stream
  cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

// Peer-memory exchange for the per-PCG-iteration camera vector (one process per GPU, CUDA IPC over
// NVLink/NVSwitch): every rank owns one exported block [flags (16 x u64) | pad | mail[2][9 ncams]].
struct ba_p2p_state {
  bool ready = false;
  bool ipc = false;                      // peer blocks were opened with CUDA IPC (else: raw in-process peer pointers)
  int nranks = 0, rank = 0;
  void* block = nullptr;                 // own allocation (cudaMalloc, exported)
  void* peer_block[16] = {nullptr};      // opened peer blocks (own entry = block)
  double** d_mail = nullptr;             // device array [nranks]: base of each rank's mail[2][n9]
  unsigned long long** d_flags = nullptr;  // device array [nranks]: each rank's flag slots
  unsigned long long* d_seq = nullptr;   // exchange sequence number (device, local)
  size_t mail_off = 256;                 // byte offset of mail inside a block
};

struct ba_handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int64_t ncams = 0, npnts = 0, nobs = 0;            // global problem
  int64_t obs0 = 0, obs1 = 0, pnt0 = 0, pnt1 = 0;    // this rank's shard (0-based, half open)
  int rank = 0, nranks = 1;
  bool sorted = false;                               // point-major order
  int32_t* d_cam = nullptr;   // local obs: 0-based camera id
  int32_t* d_pnt = nullptr;   // local obs: 0-based GLOBAL point id
  double2* d_pt2d = nullptr;  // local obs
  std::vector<int32_t> h_cam, h_pnt;  // host copies kept for schedule construction
  double* d_x = nullptr;      // staging for host-pointer calls (nvar)
  double* d_camtab = nullptr; // ncams * 16 (128-byte records, ba_math.cuh)
  double* d_cx = nullptr;     // staging: 2*nobs_l
  double* d_vals = nullptr;   // staging: 24*nobs_l
  double* d_v = nullptr;      // staging for jprod/jtprod inputs/outputs
  double* d_w = nullptr;
  int64_t* d_rows = nullptr;  // staging
  int64_t* d_cols = nullptr;
  cudaEvent_t ev_eval0 = nullptr, ev_eval1 = nullptr;  // bracket the last k_eval launch (profiling only)
  bool profile = false;
  int coarse_clusters = 16;  // two-level PCG preconditioner: target number of camera clusters (0 = off)
  int deflate = 32;          // PCG deflation: base Ritz vectors wanted (0 = off)
  int solver = BA_SOLVER_AUTO;  // damped solve: auto / PCG / exact (ba_set_solver)
  int exact_refine = 1;      // refinement steps of the exact solve (matrix-free FP64 residual)
  int mixed_max_cg = 30;     // CG iterations of the mixed-precision solve before it falls back to the FP64 factor
                             // (an FP64 factorisation costs about as much as 30 of them on Venice-1778)
  ba_lm_state lm;
  ncclComm* comm = nullptr;
  ba_p2p_state p2p;
  ba_group* group = nullptr;  // non-null: this handle only fans out to one sub-handle per GPU (ba_create_multi)
  mutable std::string err;
  int64_t nvar() const { return 9 * ncams + 3 * npnts; }
  int64_t nobs_l() const { return obs1 - obs0; }
  int64_t npnts_l() const { return pnt1 - pnt0; }
};

namespace ba {
// ---- ba_eval.cu -----------------------------------------------------------------------------
void launch_cam_precompute(const double* x, int64_t npnts, int64_t ncams, double* camtab, cudaStream_t s);
void launch_eval(const ba_handle* h, const double* x, const double* camtab, double* cx, double* vals,
                 cudaStream_t s);
void launch_jac_structure(const ba_handle* h, int64_t* rows, int64_t* cols, cudaStream_t s);
void launch_jprod(const ba_handle* h, const double* x, const double* camtab, const double* v, double* Jv,
                  cudaStream_t s);
void launch_jtprod(const ba_handle* h, const double* x, const double* camtab, const double* v, double* Jtv,
                   bool with_cameras, cudaStream_t s);
// ---- ba_lm.cu -------------------------------------------------------------------------------
int lm_prepare(ba_handle* h);
int lm_exact_workspace(ba_handle* h);  // dense matrix + (distributed) factorisation workspace of the exact solve
// camera part of J(x)'v by the ordered camera-major pass (needs point-major observations; h->d_camtab must
// hold the records of x): Jtv_cams = sum_k B_k' v_k, 9 per camera
int lm_jtprod_cams(ba_handle* h, const double* x, const double* v, double* Jtv_cams);
void lm_release(ba_handle* h);
// ---- ba_group.cu / ba_capi.cu -----------------------------------------------------------------
int create_impl(int64_t ncams, int64_t npnts, int64_t nobs, const int64_t* cam, const int64_t* pnt, const double* pt2d,
                int device, int rank, int nranks, ba_handle** out);
void group_release(ba_handle* h);
int group_residual(ba_handle* h, const double* x, double* cx, double* vals);
int group_jac_structure(ba_handle* h, int64_t* rows, int64_t* cols);
int group_jprod(ba_handle* h, const double* x, const double* v, double* Jv);
int group_jtprod(ba_handle* h, const double* x, const double* v, double* Jtv);
int group_lm_step(ba_handle* h, const double* x, double lambda, double pcg_tol, int32_t pcg_max_iter, double* delta,
                  double* dr2, double* obj, double* jtr, int32_t* pcg_iters);
int group_lm_solve(ba_handle* h, double* x_inout, const ba_lm_params* p, ba_lm_stats* st, ba_iter_cb cb, void* user);
int group_apply(ba_handle* h, const std::function<int(ba_handle*)>& f);  // setters: the same call on every sub-handle
const ba_handle* group_first(const ba_handle* h);
// ---- ba_hostio.cu ---------------------------------------------------------------------------
// dst (host, pageable or page-locked) <- src (device); ordered after the handle's stream; returns when complete
int copy_to_host(ba_handle* h, void* dst, const void* src, size_t bytes);
// ---- ba_comm.cu -----------------------------------------------------------------------------
int allreduce_sum(ba_handle* h, double* buf, size_t n);
int allreduce_sum_i64(ba_handle* h, long long* buf, size_t n);
int allgather_host(ba_handle* h, const void* send, void* recv, size_t bytes);
void comm_release(ba_handle* h);
}  // namespace ba
