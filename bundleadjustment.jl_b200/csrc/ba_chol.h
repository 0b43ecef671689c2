// ba_chol.h -- dense Cholesky of the reduced camera system on the device (ba_chol.cu).
//
// Stands in for what ldl_factorize / ldl_solve! deliver on the camera block once the points are
// eliminated (src/ldl_aux.jl:122-201 numeric factorisation, :4-42 the three solve sweeps): an exact
// factorisation, here of the explicitly assembled Schur complement.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

struct ba_handle;

namespace ba {

constexpr int CHOL_TILE = 128;  // block size of the factorisation; matrices are padded to a multiple of it

inline int64_t chol_padded(int64_t n) { return (n + CHOL_TILE - 1) / CHOL_TILE * CHOL_TILE; }

constexpr int CHOL_NBMAX = 160;   // tile rows at most (cn <= 20480)
constexpr int CHOL_RMAX = 16;     // ranks at most in the distributed factorisation

// Peer view of the distributed factorisation: every rank's copy of the matrix, of the diagonal-block inverses and
// of the control block (flags), addressable from this rank's device (CUDA IPC mappings or raw in-process pointers).
// Control block of a rank (unsigned long long): [k] = epoch once Linv_kk / L_kk of step k may be fetched from the
// owner of tile row k; [CHOL_NBMAX + 16 k + r] = epoch once rank r's tiles of panel k have landed in this rank's matrix.
struct chol_peers {
  void* S[CHOL_RMAX];   // the matrix (double, or float for the mixed-precision factor: same storage)
  void* D[CHOL_RMAX];
  unsigned long long* ctl[CHOL_RMAX];
  int R, q;
  unsigned long long epoch;
};

// Workspace of one factorisation: the matrix itself is owned by the caller.
struct chol_plan {
  int64_t cn = 0;            // padded order (multiple of CHOL_TILE)
  double* d_Dinv = nullptr;  // cn/128 blocks of 128 x 128: inverses of the diagonal blocks of L (lower, row-major)
                             // (the FP32 factorisation keeps its FP32 inverses in the same storage)
  double* d_y = nullptr;     // cn: forward-substitution result
  double* d_w = nullptr;     // cn: right-hand side being consumed
  double* d_x = nullptr;     // cn: solution of the sweeps
  int* d_info = nullptr;     // 0 ok, j + 1 = first non-positive pivot
  long long* d_prof = nullptr;  // BAGPU_POTRF_PROF: phase cycle counts of the first diagonal block
  cudaStream_t side = nullptr;   // panel stream (look-ahead)
  cudaEvent_t ev_col = nullptr, ev_panel = nullptr, ev_join = nullptr;
  void* solve_graph = nullptr;   // cudaGraphExec_t of the 2 cn/128 substitution steps
  const void* solve_graph_A = nullptr;
  bool solve_graph_32 = false;
  bool graph_off = false;        // capture not possible on the caller's stream
  bool attrs_set = false;
  // persistent sweep kernel (both substitution sweeps in one cooperative launch): y and x as {value, flag} elements
  // of 16 bytes (2 x cn of them), the flag being the launch's epoch
  void* d_vq = nullptr;
  unsigned sweep_epoch = 0;
  int sm_count = 0;
  bool sweep_off = false;        // cooperative launch not possible here: per-step kernels
  // distributed factorisation over the ranks of a sharded handle (tile row i belongs to rank i mod R)
  bool dist_ready = false;
  chol_peers peers = {};
  unsigned long long* d_ctl = nullptr;  // own control block
  int* d_cnt = nullptr;                 // CTAs of the current panel solve that have finished (last one signals)
  bool peer_ipc[CHOL_RMAX] = {};
  void* peer_open[3 * CHOL_RMAX] = {};  // IPC mappings to close
};

int chol_plan_init(ba_handle* h, chol_plan& P, int64_t cn);
void chol_plan_release(chol_plan& P);
// A (cn x cn, row-major, leading dimension cn, lower triangle) <- L with A = L L'.  Returns BA_OK and leaves
// *info_host = 0, or the 1-based index of the first non-positive pivot (BA_ERR_NUMERIC).
int chol_factor(ba_handle* h, chol_plan& P, double* A, cudaStream_t s, int* info_host);
// Collective over the ranks of a sharded handle (needs its NCCL communicator): exchange the addresses of A, the
// diagonal-block inverses and the control blocks (CUDA IPC between processes, raw pointers + peer access inside one
// process).  Leaves P.dist_ready false (replicated factorisation) when peer access is not available on every rank.
int chol_dist_setup(ba_handle* h, chol_plan& P, void* A);
// Right-looking factorisation distributed over the ranks: each rank updates its own tile rows; panels travel by
// peer-memory stores fused into the panel-solve kernel, diagonal blocks are pulled from their owner; flags over
// NVLink order the steps.  On return every rank holds the complete factor (the sweeps run replicated).
int chol_factor_dist(ba_handle* h, chol_plan& P, double* A, cudaStream_t s);
// x <- (L L')^-1 b for one right-hand side (b and x: cn doubles on the device, may alias).
int chol_solve(ba_handle* h, chol_plan& P, const double* L, const double* b, double* x, cudaStream_t s);

// Mixed precision (src/lm.jl:92-98,165-173: facto_type below the model type): the same factorisation with FP32 storage
// and three-TF32-term tensor-core products (FP32-level accuracy); A32 is cn x cn floats, leading dimension cn.  The
// sweeps read the FP32 factor and compute in FP64 (right-hand side and solution are doubles): a preconditioner.
// Replicated on every rank of a sharded handle (a distributed variant was measured slower).
int chol_factor32(ba_handle* h, chol_plan& P, float* A32, cudaStream_t s, int* info_host);
int chol_solve32(ba_handle* h, chol_plan& P, const float* L32, const double* b, double* x, cudaStream_t s);

}  // namespace ba
