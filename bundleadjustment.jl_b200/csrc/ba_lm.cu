// ba_lm.cu -- device-resident Levenberg-Marquardt (K5-K13): block JtJ / Jtr assembly, Schur
// complement onto the reduced camera system, its solve -- PCG (block-Jacobi + a coarse level over camera clusters and
// deflation vectors harvested from the PCG solves themselves), or the explicitly assembled system factorised by
// ba_chol.cu: in FP64 (exact solver) or in FP32 on the tensor cores as the preconditioner of FP64 CG (mixed solver,
// Solver::mixed_solve) -- back-substitution, and the LM loop.
//
// Reference semantics: src/lm.jl:15-418 (control flow, kept decision for decision) with the damped
// solve  (J'J + lambda I) delta = -J'r  that src/lm.jl:61-100,138-152,175-229 obtains from
// ldl_aux.jl / qr_aux.jl factorisations of the augmented matrix (SURVEY.md 3.4 shows every
// facto x perm x normalize combination solves this same system).
//
// With x = [points; cameras] (src/ReadFiles.jl:29-47) and J_k = [A_k | B_k] per observation k=(c,p):
//     V_p = sum A'A (3x3)   U_c = sum B'B (9x9)   g_p = -sum A'F   g_c = -sum B'F
//     S = U + lambda I - W (V + lambda I)^-1 W',   W_cp = B_k' A_k,   S dc = g_c - W (V+lI)^-1 g_p
//     dp = (V + lambda I)^-1 (g_p - W' dc),        dr2 = 1/2 || J delta + r ||^2
// S is never formed: S v is applied with two passes -- a point-major pass over the stored J blocks (one
// warp per <=32 observations of whole points: y = B v_c, segmented warp sums per point,
// u = (V+lI)^-1 sum A'y, w = A u; 192 B/obs streamed, coalesced) and a camera-major pass that RECOMPUTES
// the camera part of each block instead of keeping a second, camera-ordered copy of it
// (out_c = (U+lI) v_c - sum B'w; the camera record is warp-uniform, points are gathered as one 32-byte
// sector each; measured faster on B200 than streaming 144 B/obs, profiles/r01_pcg_*).  All sums are two-level and
// ordered (per-task partials, then a fixed-order gather by the last task of a camera): results are
// reproducible run to run and no FP64 atomics are used.  The camera-sized vector updates of PCG are
// a few small multi-CTA kernels (one fused CTA for up to 128 cameras); its dot products are per-CTA partials
// summed in fixed order.  Eight PCG iterations form one CUDA graph; convergence is decided on the device.
// The coarse level solves with [P | Z]' S [P | Z]: P = cluster indicators of the pose components, Z = Ritz
// vectors of the preconditioned system taken from the Lanczos process that CG is (defl_update, ba_ritz.h).
// Sharded over ranks (one process per GPU), the sum over ranks inside the PCG iteration is fused into the
// kernel that consumes it and read from the peers' IPC-mapped memory over NVLink (ba_comm.cu); the
// per-LM-iteration collectives use NCCL.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <future>
#include <thread>
#include "ba_internal.h"
#include "ba_math.cuh"
#include "ba_ritz.h"


#include "ba_lm_kernels.cuh"

namespace ba {

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
namespace {

template <class T>
int dmalloc(ba_handle* h, T** p, size_t n) {
  BA_CUDA(cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(n, 1) * sizeof(T)));
  return BA_OK;
}

inline unsigned nblk(int64_t n, int per) { return (unsigned)std::max<int64_t>((n + per - 1) / per, 1); }

// camera systems of up to this many rows run the vector half of a PCG iteration as one CTA (k_pcg_small)
inline int64_t small_rows_max() {
  static const int64_t v = getenv("BAGPU_VEC_SMALL_MAX") ? atoll(getenv("BAGPU_VEC_SMALL_MAX")) : 1152;
  return v;
}
inline bool trace_on() {
  static const bool v = getenv("BAGPU_TRACE") != nullptr;
  return v;
}
inline int64_t exact_auto_cams() {
  static const int64_t v = getenv("BAGPU_EXACT_AUTO_CAMS") ? atoll(getenv("BAGPU_EXACT_AUTO_CAMS")) : 2048;
  return v;
}
// AUTO: dense systems of at least this many rows are factorised in FP32 on the tensor cores (BA_SOLVER_MIXED), smaller
// ones in FP64 (BA_SOLVER_EXACT): below ~1000 cameras the chain of diagonal blocks dominates both factorisations and
// the extra CG iterations of the mixed solve cost more than the factorisation saves
inline int64_t mixed_auto_rows() {
  static const int64_t v = getenv("BAGPU_MIXED_AUTO_ROWS") ? atoll(getenv("BAGPU_MIXED_AUTO_ROWS")) : 8192;
  return v;
}
inline double exact_max_bytes() {
  static const double v = (getenv("BAGPU_EXACT_MAX_GB") ? atof(getenv("BAGPU_EXACT_MAX_GB")) : 16.0) * 1e9;
  return v;
}
constexpr int DEFL_ROLL = 16;   // deflation vectors refreshed after every solve (on top of the base ones)
constexpr int DEFL_CAND = 64;   // candidates per selection (= DG_N, the widest k_defl_gemm output)
constexpr int DEFL_HCAP = 1024; // Lanczos vectors kept per solve at most
constexpr int DEFL_NR = 4;      // vectors per pass of the multi-vector Schur product (coarse setup)

}  // namespace

// Decide whether the damped solve of this handle is the exact one and, if so, allocate the dense matrix and the
// factorisation workspace; on a sharded handle with a communicator also exchange the peer addresses of the
// distributed factorisation (collective over the ranks).  Called from lm_prepare, and ahead of time from
// ba_comm_init / ba_create_multi so that this one-off setup (CUDA IPC mappings of every peer's matrix) is part of
// creating the communicator rather than of the first Levenberg_Marquardt call.
int lm_exact_workspace(ba_handle* h) {
  ba_lm_state& S = h->lm;
  const int64_t ncams = h->ncams;
  if (S.d_S) return BA_OK;
  BA_CUDA(cudaSetDevice(h->device));
  // auto = up to EXACT_AUTO_CAMS cameras (the factorisation is n^3/3 flops)
  S.exact = h->sorted && ncams > 0 &&
            (h->solver == BA_SOLVER_EXACT || h->solver == BA_SOLVER_MIXED ||
             (h->solver == BA_SOLVER_AUTO && ncams <= exact_auto_cams()));
  S.mixed = S.exact && (h->solver == BA_SOLVER_MIXED || (h->solver == BA_SOLVER_AUTO && 9 * ncams >= mixed_auto_rows()));
  S.cn = chol_padded(9 * ncams);
  if (S.exact && (double)S.cn * (double)S.cn * 8.0 > exact_max_bytes()) {
    if (h->solver == BA_SOLVER_EXACT || h->solver == BA_SOLVER_MIXED) {
      h->err = "BA_SOLVER_EXACT / MIXED: the dense reduced camera system does not fit (raise BAGPU_EXACT_MAX_GB or use PCG)";
      return BA_ERR_ARG;
    }
    S.exact = S.mixed = false;
  }
  if (!S.exact) return BA_OK;
  int rc;
  if ((rc = dmalloc(h, &S.d_S, (size_t)(S.cn * S.cn)))) return rc;
  {
    const int64_t nbt = S.cn / CHOL_TILE;
    if ((rc = dmalloc(h, &S.d_Sq, (size_t)(nbt * (nbt + 1) / 2 * CHOL_TILE * CHOL_TILE)))) return rc;
  }
  // Yh (27 doubles per observation, 1.08 GB on Venice) is dead once k_exact_assemble has run, and the dense matrix is
  // only written after that (k_exact_finish): when it fits, Yh borrows the storage of the matrix (one GB less to
  // allocate; device allocations of this size cost tens of milliseconds each)
  if (27 * h->nobs_l() <= S.cn * S.cn) {
    S.d_Yh = S.d_S;
    S.yh_aliased = true;
  } else if ((rc = dmalloc(h, &S.d_Yh, (size_t)(27 * h->nobs_l())))) {
    return rc;
  }
  if ((rc = dmalloc(h, &S.d_cd, (size_t)(9 * ncams)))) return rc;
  if ((rc = dmalloc(h, &S.d_ex, (size_t)(2 * S.cn)))) return rc;
  if ((rc = chol_plan_init(h, S.chol, S.cn))) return rc;
  // sharded: distribute the factorisation over the ranks (falls back to the replicated one without peer access)
  if (h->nranks > 1 && h->comm) {
    if ((rc = chol_dist_setup(h, S.chol, S.d_S))) return rc;
    // AUTO on many ranks: the FP32 factorisation is replicated (17.5 ms on Venice whatever N), the FP64 one is
    // distributed and overtakes it from 8 ranks on (8 GPUs: 34.7 against 31.3 LM iterations/s) -- every rank decides
    // alike (dist_ready is agreed among the ranks)
    static const int many = getenv("BAGPU_MIXED_AUTO_MAX_RANKS") ? atoi(getenv("BAGPU_MIXED_AUTO_MAX_RANKS")) : 8;
    if (S.mixed && h->solver == BA_SOLVER_AUTO && h->nranks >= many && S.chol.dist_ready) S.mixed = false;
    // NCCL sets up its channels for a message-size class inside the first collective of that class (tens of ms for the
    // ~1 GB integer allreduce of the assembly): pay that here, with the communicator, not in the first LM iteration
    const int64_t nbt = S.cn / CHOL_TILE, packed = nbt * (nbt + 1) / 2 * CHOL_TILE * CHOL_TILE;
    BA_CUDA(cudaMemsetAsync(S.d_Sq, 0, sizeof(long long) * (size_t)packed, h->stream));
    for (int rep = 0; rep < 2; ++rep)
      if ((rc = allreduce_sum_i64(h, S.d_Sq, (size_t)packed))) return rc;
    BA_CUDA(cudaStreamSynchronize(h->stream));
  }
  return BA_OK;
}

int lm_prepare(ba_handle* h) {
  ba_lm_state& S = h->lm;
  if (S.ready) return BA_OK;
  if (!h->sorted) {
    h->err = "the LM entry points need point-major observation order (the order of BAL files)";
    return BA_ERR_UNSORTED;
  }
  BA_CUDA(cudaSetDevice(h->device));
  const int64_t nl = h->nobs_l(), npl = h->npnts_l(), ncams = h->ncams;
  const bool trace = getenv("BAGPU_TRACE") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms_since = [&](std::chrono::steady_clock::time_point t) {
    return std::chrono::duration<double>(now() - t).count() * 1e3;
  };
  const auto tp0 = now();
  // ---- point-major warp tasks: whole points, <= 32 observations; a longer point is a task of its own
  std::vector<int32_t> tstart, pstart((size_t)npl + 1, 0);
  std::thread point_tasks([&] {  // (independent of the camera-major sort below: the two overlap)
    int64_t cur_start = 0, cur_len = 0, a = 0;
    int64_t next_point = 0;  // local points whose pstart is not set yet
    while (a < nl) {
      int64_t b = a + 1;
      while (b < nl && h->h_pnt[(size_t)b] == h->h_pnt[(size_t)a]) ++b;
      const int64_t pl = h->h_pnt[(size_t)a] - h->pnt0;
      for (; next_point <= pl; ++next_point) pstart[(size_t)next_point] = (int32_t)a;
      const int64_t len = b - a;
      if (len > 32) {
        if (cur_len > 0) tstart.push_back((int32_t)cur_start);
        tstart.push_back((int32_t)a);
        cur_start = b;
        cur_len = 0;
      } else if (cur_len + len > 32) {
        tstart.push_back((int32_t)cur_start);
        cur_start = a;
        cur_len = len;
      } else {
        if (cur_len == 0) cur_start = a;
        cur_len += len;
      }
      a = b;
    }
    if (cur_len > 0) tstart.push_back((int32_t)cur_start);
    for (; next_point <= npl; ++next_point) pstart[(size_t)next_point] = (int32_t)nl;
    S.ntasks = (int64_t)tstart.size();
    tstart.push_back((int32_t)nl);
  });
  // ---- camera-major order (stable counting sort) and tasks.  The sort runs on a few host threads: thread t
  // counts the cameras of its contiguous chunk of observations, the (camera, chunk) counts are turned into start
  // positions, and every thread scatters its own chunk -- stable, hence deterministic.
  std::vector<int32_t> cam_start((size_t)ncams + 1, 0), cperm((size_t)nl), tb, te, tc, cam_t0((size_t)ncams + 1, 0);
  {
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)std::thread::hardware_concurrency(), 8, nl / 200000 + 1}));
    std::vector<std::vector<int32_t>> cnt((size_t)nt, std::vector<int32_t>((size_t)ncams, 0));
    auto chunk = [&](int t) { return std::pair<int64_t, int64_t>(nl * t / nt, nl * (t + 1) / nt); };
    auto par = [&](const std::function<void(int)>& f) {
      std::vector<std::thread> th;
      for (int t = 1; t < nt; ++t) th.emplace_back(f, t);
      f(0);
      for (auto& x : th) x.join();
    };
    par([&](int t) {
      const auto r = chunk(t);
      int32_t* c = cnt[(size_t)t].data();
      for (int64_t k = r.first; k < r.second; ++k) c[h->h_cam[(size_t)k]]++;
    });
    int32_t run = 0;
    for (int64_t c = 0; c < ncams; ++c) {
      cam_start[(size_t)c] = run;
      for (int t = 0; t < nt; ++t) {
        const int32_t n = cnt[(size_t)t][(size_t)c];
        cnt[(size_t)t][(size_t)c] = run;  // becomes the write cursor of (chunk t, camera c)
        run += n;
      }
    }
    cam_start[(size_t)ncams] = run;
    par([&](int t) {
      const auto r = chunk(t);
      int32_t* cur = cnt[(size_t)t].data();
      for (int64_t k = r.first; k < r.second; ++k) cperm[(size_t)cur[h->h_cam[(size_t)k]]++] = (int32_t)k;
    });
  }
  point_tasks.join();
  int tsz = 32;  // observations per camera task: enough tasks to fill 148 SMs, at most 256 per warp
  while (tsz < 256 && nl / tsz > 148 * 64) tsz *= 2;
  for (int64_t c = 0; c < ncams; ++c) {
    cam_t0[(size_t)c] = (int32_t)tb.size();
    for (int32_t b = cam_start[(size_t)c]; b < cam_start[(size_t)c + 1]; b += tsz) {
      tb.push_back(b);
      tc.push_back((int32_t)c);
      te.push_back(std::min<int32_t>(b + tsz, cam_start[(size_t)c + 1]));
    }
  }
  cam_t0[(size_t)ncams] = (int32_t)tb.size();
  std::vector<int32_t> empty_cams;
  for (int64_t c = 0; c < ncams; ++c)
    if (cam_start[(size_t)c] == cam_start[(size_t)c + 1]) empty_cams.push_back((int32_t)c);
  S.nempty = (int64_t)empty_cams.size();
  S.nctasks = (int64_t)tb.size();

  const double t_host = ms_since(tp0);
  const auto tp1 = now();
  // one device allocation for the whole LM state of this handle (cudaMalloc is a synchronising driver call whose
  // cost on a shared host varies from microseconds to tens of milliseconds: ~50 of them were a measurable part of the
  // first Levenberg_Marquardt call): the requests are recorded, one slab is allocated, the pointers are carved out
  struct slab_req { void** p; size_t bytes; };
  std::vector<slab_req> reqs;
#define ALLOC(ptr, n)                                                                                      \
  reqs.push_back(slab_req{reinterpret_cast<void**>(&(ptr)),                                                \
                          (std::max<size_t>((size_t)(n), 1) * sizeof(*(ptr)) + 255) & ~(size_t)255})
  ALLOC(S.d_tstart, tstart.size());
  ALLOC(S.d_pstart, pstart.size());
  ALLOC(S.d_cperm, nl);
  ALLOC(S.d_ctask_beg, tb.size());
  ALLOC(S.d_ctask_end, te.size());
  ALLOC(S.d_cam_t0, cam_t0.size());
  ALLOC(S.d_ctask_cam, tc.size());
  ALLOC(S.d_cam_cnt, ncams);
  ALLOC(S.d_empty_cams, empty_cams.size());
  ALLOC(S.d_Jp, 12 * nl);
  ALLOC(S.d_F, nl);
  ALLOC(S.d_pntc, nl);
  ALLOC(S.d_x4, 2 * h->npnts);
  ALLOC(S.d_w, nl);
  ALLOC(S.d_T, 3 * nl);
  ALLOC(S.d_V, 6 * npl);
  ALLOC(S.d_gp, 3 * npl);
  ALLOC(S.d_Vinv, 6 * npl);
  ALLOC(S.d_wp, 3 * npl);
  ALLOC(S.d_taskpart, S.nctasks * NV);
  ALLOC(S.d_Ug, ncams * NV);
  ALLOC(S.d_Cr, ncams * NV);
  ALLOC(S.d_H, ncams * 81);
  ALLOC(S.d_Minv, ncams * 81);
  ALLOC(S.d_pcg, 6 * 9 * ncams);
  ALLOC(S.d_pcgpart, 2 * (int64_t)nblk(9 * ncams, VEC_ROWS));
  {  // two-level preconditioner: clusters of whole vector-kernel CTAs (28 cameras each), at most 16 clusters
    const int nvb = (int)nblk(9 * ncams, VEC_ROWS);
    const int want = std::max(0, std::min(144 / CDOF, h->coarse_clusters));
    S.ncl = 0;
    if (want > 0 && ncams > 0) {
      S.ctas_per_cluster = (nvb + want - 1) / want;
      S.ncl = (nvb + S.ctas_per_cluster - 1) / S.ctas_per_cluster;
    }
    S.mc = CDOF * S.ncl;
    // deflation vectors share the coarse solve (mc + kz <= 144); small systems use the fused kernel without them
    S.kz_max = S.kz_base_max = 0;
    if (h->deflate > 0 && 9 * ncams > small_rows_max()) {
      S.kz_max = std::min(std::min(144 - S.mc, KZ_LIMIT), h->deflate + DEFL_ROLL);
      S.kz_base_max = std::min(h->deflate, S.kz_max);
    }
    const int64_t mt = S.mc + S.kz_max;
    ALLOC(S.d_Ac, mt * mt);
    ALLOC(S.d_Aci, mt * mt);
    ALLOC(S.d_yc, mt);
    ALLOC(S.d_cpart, 9 * (int64_t)nvb);
    ALLOC(S.d_Acq, S.mc * S.mc);
    ALLOC(S.d_cdiag, S.mc);
    if (S.kz_max > 0) {
      ALLOC(S.d_Z, 9 * ncams * S.kz_max);
      ALLOC(S.d_Zcand, 9 * ncams * DEFL_CAND);
      ALLOC(S.d_zpart, (int64_t)nvb * S.kz_max);
      ALLOC(S.d_w4, DEFL_NR * nl);
      ALLOC(S.d_q4, DEFL_NR * 9 * ncams);
    }
  }
  ALLOC(S.d_x, h->nvar());
  ALLOC(S.d_xt, h->nvar());
  ALLOC(S.d_delta, h->nvar());
  ALLOC(S.d_camt, ncams * CAM_REC);
  S.npart = 4 * (int64_t)std::max<int64_t>(nblk(std::max(nl, npl), PT_THREADS), 1024);
  ALLOC(S.d_part, S.npart + 16);  // + scratch for the four step norms
  ALLOC(S.d_scal, S_COUNT);
  // (the exact solve's workspace -- the dense matrix, 2-3 GB on Venice -- is NOT part of this: it is allocated by
  //  lm_exact_workspace when a damped solve is actually requested, so that ba_jtprod, which only needs the
  //  camera-major schedules, stays light)
  const double t_work = 0.0;
  {
    size_t total = 0;
    for (const slab_req& r : reqs) total += r.bytes;
    BA_CUDA(cudaMalloc(&S.d_slab, total));
    size_t off = 0;
    for (const slab_req& r : reqs) {
      *r.p = static_cast<char*>(S.d_slab) + off;
      off += r.bytes;
    }
  }
#undef ALLOC
  // opt-in shared-memory sizes are per device: set them for this handle's device
  BA_CUDA(cudaFuncSetAttribute(k_coarse_assemble, cudaFuncAttributeMaxDynamicSharedMemorySize, 144 * 144 * 8));
  BA_CUDA(cudaFuncSetAttribute(k_coarse_invert, cudaFuncAttributeMaxDynamicSharedMemorySize, 144 * 144 * 8));
  BA_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&S.h_scal), S_COUNT * sizeof(double), cudaHostAllocDefault));
  for (auto& e : S.ev) BA_CUDA(cudaEventCreate(&e));
  const double t_alloc = ms_since(tp1);
  const auto tp2 = now();
  auto up = [&](void* d, const void* s, size_t bytes) {
    return bytes ? cudaMemcpyAsync(d, s, bytes, cudaMemcpyHostToDevice, h->stream) : cudaSuccess;
  };
  BA_CUDA(up(S.d_tstart, tstart.data(), tstart.size() * 4));
  BA_CUDA(up(S.d_pstart, pstart.data(), pstart.size() * 4));
  BA_CUDA(up(S.d_cperm, cperm.data(), cperm.size() * 4));
  if (nl) k_gather_i32<<<nblk(nl, 256), 256, 0, h->stream>>>(nl, S.d_cperm, h->d_pnt, S.d_pntc);  // point id at each camera-major position
  BA_CUDA(up(S.d_ctask_beg, tb.data(), tb.size() * 4));
  BA_CUDA(up(S.d_ctask_end, te.data(), te.size() * 4));
  BA_CUDA(up(S.d_cam_t0, cam_t0.data(), cam_t0.size() * 4));
  BA_CUDA(up(S.d_ctask_cam, tc.data(), tc.size() * 4));
  BA_CUDA(up(S.d_empty_cams, empty_cams.data(), empty_cams.size() * 4));
  BA_CUDA(cudaMemsetAsync(S.d_cam_cnt, 0, sizeof(int32_t) * (size_t)std::max<int64_t>(ncams, 1), h->stream));
  BA_CUDA(cudaMemsetAsync(S.d_scal, 0, S_COUNT * sizeof(double), h->stream));
  BA_CUDA(cudaMemsetAsync(S.d_delta, 0, sizeof(double) * (size_t)h->nvar(), h->stream));
  BA_CUDA(cudaStreamSynchronize(h->stream));  // the host vectors go out of scope
  if (trace)
    fprintf(stderr, "[bagpu] lm_prepare: schedules on the host %.1f ms, device allocation %.1f ms, uploads %.1f ms\n",
            t_host, t_alloc + t_work, ms_since(tp2));
  S.ready = true;
  return BA_OK;
}

// First LM call on a handle: the schedules (host work, then one slab allocation) and the dense workspace of the exact /
// mixed solve (2-3 GB of device allocations: tens of milliseconds of driver time) are set up CONCURRENTLY, the latter
// on a helper thread -- on one rank only: with a communicator the workspace setup is collective and was done with it.
int lm_prepare_all(ba_handle* h) {
  ba_lm_state& S = h->lm;
  if (S.ready || S.d_S || h->nranks > 1) {
    int rc = lm_prepare(h);
    return rc ? rc : lm_exact_workspace(h);
  }
  std::future<int> ws = std::async(std::launch::async, [h] { return lm_exact_workspace(h); });
  const int rc = lm_prepare(h);
  const int rc2 = ws.get();
  return rc ? rc : rc2;
}

int lm_jtprod_cams(ba_handle* h, const double* x, const double* v, double* out) {
  int rc = lm_prepare(h);
  if (rc) return rc;
  ba_lm_state& S = h->lm;
  cudaStream_t s = h->stream;
  const int64_t nl = h->nobs_l(), n9 = 9 * h->ncams;
  BA_CUDA(cudaMemsetAsync(S.d_scal + S_DONE, 0, sizeof(double), s));  // the pass honours the PCG done flag
  k_pad_points<<<(unsigned)std::max<int64_t>((h->npnts_l() + 255) / 256, 1), 256, 0, s>>>(x, h->pnt0, h->pnt1, S.d_x4);
  if (S.nctasks)
    k_cam_pass<2><<<(unsigned)((S.nctasks + PT_THREADS / 32 - 1) / (PT_THREADS / 32)), PT_THREADS, 0, s>>>(
        S.d_ctask_beg, S.d_ctask_end, S.nctasks, S.d_ctask_cam, S.d_cam_t0, S.d_cam_cnt, S.d_cperm, S.d_pntc, nl,
        h->d_camtab, S.d_x4, S.d_F, reinterpret_cast<const double2*>(v), S.d_T, S.d_taskpart, out, S.d_scal, nullptr,
        nullptr, n9);
  if (S.nempty)
    k_zero_cams<<<(unsigned)((S.nempty * 9 + 255) / 256), 256, 0, s>>>(S.d_empty_cams, (int)S.nempty, 9, out, nullptr,
                                                                      nullptr, n9);
  BA_CUDA(cudaGetLastError());
  return BA_OK;
}

void lm_release(ba_handle* h) {
  ba_lm_state& S = h->lm;
  // everything lm_prepare allocates lives in one slab; the rest are the buffers allocated on demand
  void* ptrs[] = {S.d_slab, S.d_dr, S.d_harv, S.d_hcoef, S.d_dsmall, S.d_S, S.d_Sq, S.yh_aliased ? nullptr : S.d_Yh,
                  S.d_cd, S.d_ex};
  for (void* p : ptrs) cudaFree(p);
  chol_plan_release(S.chol);
  if (S.h_scal) cudaFreeHost(S.h_scal);
  if (S.pcg_graph) cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(S.pcg_graph));
  for (auto& e : S.ev)
    if (e) cudaEventDestroy(e);
  S = ba_lm_state();
}

namespace {

struct Solver {
  ba_handle* h;
  ba_lm_state& S;
  cudaStream_t s;
  int64_t nl, npl, ncams, n9;
  double *b, *xc, *r, *z, *p, *q;
  explicit Solver(ba_handle* hh) : h(hh), S(hh->lm), s(hh->stream) {
    nl = h->nobs_l(); npl = h->npnts_l(); ncams = h->ncams; n9 = 9 * ncams;
    b = S.d_pcg; xc = b + n9; r = xc + n9; z = r + n9; p = z + n9; q = p + n9;
  }
  // BAGPU_DEBUG_SYNC=1: synchronise before every check so that a faulting kernel is reported at the
  // check that follows its launch (file:line in ba_last_error) instead of at the next host sync
  int check_at(int line) {
    static const bool dbg = getenv("BAGPU_DEBUG_SYNC") != nullptr;
    cudaError_t e = dbg ? cudaStreamSynchronize(s) : cudaSuccess;
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
      char b_[256];
      snprintf(b_, sizeof b_, "ba_lm.cu:%d: %s", line, cudaGetErrorString(e));
      h->err = b_;
      return BA_ERR_CUDA;
    }
    return BA_OK;
  }
#define check() check_at(__LINE__)
  int reduce_to_at(int line, int64_t nblocks, int K, int slot0) {
    int rc = check_at(line);  // the kernel that produced the partials
    if (rc) return rc;
    k_reduce_parts<<<K, RED_THREADS, 0, s>>>(S.d_part, nblocks, K, S.d_scal, slot0);
    return check_at(line);
  }
#define reduce_to(...) reduce_to_at(__LINE__, __VA_ARGS__)
  // camera-major pass + ordered gather (+ allreduce over ranks) into out (ncams x nacc)
  template <int MODE>
  int cam_pass(double* out) {
    constexpr int nacc = (MODE == 2) ? 9 : NV;
    if (S.nctasks)
      k_cam_pass<MODE><<<nblk(S.nctasks, PT_THREADS / 32), PT_THREADS, 0, s>>>(
          S.d_ctask_beg, S.d_ctask_end, S.nctasks, S.d_ctask_cam, S.d_cam_t0, S.d_cam_cnt, S.d_cperm, S.d_pntc, nl,
          h->d_camtab, S.d_x4, S.d_F, S.d_w, S.d_T, S.d_taskpart, out, S.d_scal, nullptr, nullptr, n9);
    if (S.nempty)
      k_zero_cams<<<nblk((int64_t)S.nempty * nacc, 256), 256, 0, s>>>(S.d_empty_cams, (int)S.nempty, nacc, out,
                                                                     nullptr, nullptr, n9);
    int rc = check();
    if (rc) return rc;
    return allreduce_sum(h, out, (size_t)(ncams * nacc));
  }
  // F, J blocks, V, g_p, U, g_c at x; leaves sum F^2, sum g_p^2 (allreduced) and sum g_c^2 in the scalars
  int build(const double* x) {
    launch_cam_precompute(x, h->npnts, ncams, h->d_camtab, s);
    const unsigned nb = nblk(nl, PT_THREADS), npb = nblk(npl, PT_THREADS);
    int rc0 = check();
    if (rc0) return rc0;
    k_pad_points<<<nblk(npl, 256), 256, 0, s>>>(x, h->pnt0, h->pnt1, S.d_x4);
    if ((rc0 = check())) return rc0;
    k_lm_build<<<nb, PT_THREADS, 0, s>>>(h->d_cam, h->d_pnt, h->d_pt2d, x, h->d_camtab, S.d_Jp, S.d_F, S.d_part, nl);
    int rc = reduce_to(nb, 1, S_F2);
    if (rc) return rc;
    k_point_assemble<<<npb, PT_THREADS, 0, s>>>(S.d_pstart, npl, nl, S.d_Jp, S.d_F, S.d_V, S.d_gp, S.d_part);
    if ((rc = reduce_to(npb, 1, S_GP2))) return rc;
    if ((rc = cam_pass<0>(S.d_Ug))) return rc;
    k_gc_norm<<<1, RED_THREADS, 0, s>>>(ncams, S.d_Ug, S.d_scal);
    if ((rc = check())) return rc;
    return allreduce_sum(h, S.d_scal + S_F2, 2);
  }
  // everything that depends on lambda: Vinv, Schur diagonal blocks -> Minv, right-hand side b, H
  int factor(double lambda) {
    k_point_inv<<<nblk(npl, PT_THREADS), PT_THREADS, 0, s>>>(npl, lambda, S.d_V, S.d_gp, S.d_Vinv, S.d_wp);
    int rc = check();
    if (rc) return rc;
    k_point_prep<<<nblk(nl, PT_THREADS), PT_THREADS, 0, s>>>(h->d_pnt, h->pnt0, nl, S.d_Jp, S.d_Vinv, S.d_wp, S.d_w, S.d_T);
    if ((rc = check())) return rc;
    if ((rc = cam_pass<1>(S.d_Cr))) return rc;
    k_cam_finish<<<nblk(ncams, 64), 64, 0, s>>>(ncams, lambda, S.d_Ug, S.d_Cr, S.d_H, S.d_Minv, b, S.d_scal);
    if ((rc = check())) return rc;
    return S.exact ? exact_factor(lambda) : coarse_setup();
  }
  // Exact solve, factor phase: assemble the Jacobi-scaled reduced camera system explicitly (fixed-point sums of
  // the off-diagonal blocks: order-independent, summed over ranks as integers) and factorise it (ba_chol.cu).
  int exact_factor(double lambda) {
    int rc;
    const int64_t cn = S.cn;
    const bool tr = trace_on();
    cudaEventRecord(S.ev[5], s);
    k_exact_diag<<<nblk(n9, 256), 256, 0, s>>>(n9, lambda, S.d_Ug, S.d_cd);
    k_exact_y<<<nblk(nl, PT_THREADS), PT_THREADS, 0, s>>>(h->d_cam, h->d_pnt, h->pnt0, nl, S.d_Jp, S.d_Vinv, S.d_cd,
                                                          S.d_Yh);
    const int64_t nbt = cn / CHOL_TILE, packed = nbt * (nbt + 1) / 2 * CHOL_TILE * CHOL_TILE;
    BA_CUDA(cudaMemsetAsync(S.d_Sq, 0, sizeof(long long) * (size_t)packed, s));
    if ((rc = check())) return rc;
    if (nl > 0)
      k_exact_assemble<<<(unsigned)std::min<int64_t>(nblk(nl, 8), 148 * 16), 256, 0, s>>>(
          S.d_pstart, h->d_pnt, h->pnt0, h->d_cam, nl, S.d_Yh, reinterpret_cast<unsigned long long*>(S.d_Sq));
    if ((rc = check())) return rc;
    if ((rc = allreduce_sum_i64(h, S.d_Sq, (size_t)packed))) return rc;
    int info = 0;
    S.factor64 = !S.mixed;
    if (S.mixed) {
      // FP32 matrix (same storage), FP32 tensor-core factorisation: the preconditioner of the FP64 CG in mixed_solve
      if ((rc = dense_factor<float>(&info))) return rc;
      if (info > 0) {  // a non-positive pivot in FP32: the FP64 factorisation decides
        if (tr) fprintf(stderr, "[bagpu] mixed solve: FP32 pivot %d not positive, FP64 factorisation instead\n", info);
        S.mixed_fallbacks += 1;
        S.factor64 = true;
      }
    }
    if (S.factor64 && (rc = dense_factor<double>(&info))) return rc;
    cudaEventRecord(S.ev[7], s);
    if ((rc = read_scalars())) return rc;  // S_ERR
    if (info == -2) {
      h->err = "distributed Cholesky: a peer's flag never arrived (a rank is missing or stalled)";
      return BA_ERR_COMM;
    }
    if (info != 0) {
      h->err = "Cholesky of the reduced camera system: non-positive pivot";
      return BA_ERR_NUMERIC;
    }
    float ta = 0, tc = 0;
    cudaEventElapsedTime(&ta, S.ev[5], S.ev[6]);
    cudaEventElapsedTime(&tc, S.ev[6], S.ev[7]);
    S.t_schur_ms += ta;
    S.t_chol_ms += tc;
    S.chol_count += 1;
    if (tr) {
      const double fl = (double)cn * cn * cn / 3.0;
      fprintf(stderr, "[bagpu] exact factor: assembly of S (%lld x %lld) %.2f ms, Cholesky (%s) %.2f ms (%.1f TFLOP/s%s)\n",
              (long long)cn, (long long)cn, ta, S.factor64 ? "FP64" : "FP32 storage, 3 x TF32", tc, fl / (tc * 1e-3) / 1e12,
              S.chol.dist_ready ? ", distributed over the ranks" : "");
    }
    if (S.h_scal[S_ERR] != 0.0) {
      h->err = "Schur diagonal block not positive definite";
      return BA_ERR_NUMERIC;
    }
    return BA_OK;
  }
  // the dense matrix from the fixed-point sums (the sums stay intact), in T, and its factorisation; synchronises
  template <typename T>
  int dense_factor(int* info) {
    int rc;
    const int64_t cn = S.cn;
    T* A = reinterpret_cast<T*>(S.d_S);
    k_exact_finish<T><<<dim3(nblk(cn, 256), (unsigned)cn), 256, 0, s>>>(n9, cn, S.d_H, S.d_Cr, S.d_cd, S.d_Sq, A);
    if ((rc = check())) return rc;
    cudaEventRecord(S.ev[6], s);  // the factorisation proper starts here
    if constexpr (sizeof(T) == 8) rc = S.chol.dist_ready ? chol_factor_dist(h, S.chol, A, s) : chol_factor(h, S.chol, A, s, nullptr);
    else rc = chol_factor32(h, S.chol, A, s, nullptr);  // replicated on every rank: the tcgen05 factorisation of a single GPU
                                                         // is faster than the distributed one's chain of diagonal blocks
    if (rc) return rc;
    BA_CUDA(cudaMemcpyAsync(info, S.chol.d_info, sizeof(int), cudaMemcpyDeviceToHost, s));
    BA_CUDA(cudaStreamSynchronize(s));
    return BA_OK;
  }
  // Mixed-precision solve: FP64 CG on S dc = b with the matrix-free FP64 product (the operator PCG applies),
  // preconditioned by the FP32 factor: M^-1 r = D^-1 (L32 L32')^-1 D^-1 r.  M^-1 S = I + O(cond * 1e-7): a handful of
  // iterations.  Stops at ||r|| / ||b|| <= tol, or when the residual stagnates below 1e-10 (its rounding floor); the
  // TRUE residual ||b - S xc|| / ||b|| is then measured and reported like the exact solve's.  Returns 1 when the
  // factor is no usable preconditioner (not positive definite in CG, or no convergence in mixed_max_cg iterations):
  // the caller factorises in FP64.
  int mixed_solve(double tol, int* iters) {
    int rc;
    const int64_t cn = S.cn;
    double* rs = S.d_ex;        // r / cd, padded: right-hand side of the sweeps
    double* xs = S.d_ex + cn;   // their result
    const float* L32 = reinterpret_cast<const float*>(S.d_S);
    const bool p2p = h->nranks > 1 && h->p2p.ready;
    BA_CUDA(cudaMemsetAsync(S.d_scal + S_MBAD, 0, sizeof(double), s));
    k_mixed_init<<<1, RED_THREADS, 0, s>>>(n9, cn, b, S.d_cd, xc, r, rs, S.d_scal);
    double rel = 1.0, prev = 1.0;
    int it = 0;
    bool ok = false;
    const int maxit = std::max(1, h->mixed_max_cg);
    while (it < maxit) {
      if ((rc = chol_solve32(h, S.chol, L32, rs, xs, s))) return rc;
      k_mixed_dir<<<1, RED_THREADS, 0, s>>>(n9, xs, S.d_cd, r, p, S.d_scal, it == 0 ? 1 : 0);
      BA_CUDA(cudaMemsetAsync(S.d_scal + S_DONE, 0, sizeof(double), s));  // the product kernels honour S_DONE
      if ((rc = s_product(false))) return rc;
      if (p2p) k_seq_inc<<<1, 1, 0, s>>>(h->p2p.d_seq);
      k_mixed_update<<<1, RED_THREADS, 0, s>>>(n9, p, q, S.d_cd, xc, r, rs, S.d_scal);
      ++it;
      if ((rc = check())) return rc;
      if (it < 2) continue;  // (never done after one iteration: save the round trip)
      if ((rc = read_scalars())) return rc;
      if (S.h_scal[S_ERR] != 0.0) break;
      if (S.h_scal[S_MBAD] != 0.0) break;
      prev = rel;
      rel = S.h_scal[S_REL];
      if (!std::isfinite(rel)) break;
      if (rel <= tol || (rel <= 1e-10 && rel > 0.25 * prev)) {
        ok = true;
        break;
      }
    }
    if (S.h_scal[S_ERR] == 3.0) {
      h->err = "peer-memory exchange timed out (a rank is missing or stalled)";
      return BA_ERR_COMM;
    }
    if (!ok) {
      if (trace_on())
        fprintf(stderr, "[bagpu] mixed solve: no convergence with the FP32 factor (%d iterations, residual %.2e%s)\n", it, rel,
                S.h_scal[S_MBAD] != 0.0 ? ", non-positive curvature" : "");
      BA_CUDA(cudaMemsetAsync(S.d_scal + S_MBAD, 0, sizeof(double), s));
      return 1;
    }
    // the true residual of xc, matrix-free (the recursive one above drifts by rounding)
    BA_CUDA(cudaMemcpyAsync(p, xc, sizeof(double) * (size_t)n9, cudaMemcpyDeviceToDevice, s));
    BA_CUDA(cudaMemsetAsync(S.d_scal + S_DONE, 0, sizeof(double), s));
    if ((rc = s_product(false))) return rc;
    if (p2p) k_seq_inc<<<1, 1, 0, s>>>(h->p2p.d_seq);
    k_exact_resid<<<1, RED_THREADS, 0, s>>>(n9, cn, b, q, S.d_cd, rs, S.d_scal, 1);
    *iters = it;
    S.last_solver = BA_SOLVER_MIXED;
    S.last_converged = 1;
    S.last_iters = it;
    S.last_rel = -1.0;  // filled from S_REL at the next read of the scalars
    return check();
  }
  // Exact solve, solve phase: xc = D^-1 (L L')^-1 D^-1 b, then `exact_refine` refinement steps with the FP64
  // residual b - S xc of the matrix-free product (the same operator PCG applies).  No host round trip: the
  // residual norm of the direct solve is left in S_REL for the next read of the scalars.
  int exact_solve(int* iters) {
    int rc;
    const int64_t cn = S.cn;
    double* rs = S.d_ex;        // scaled right-hand side / residual
    double* xs = S.d_ex + cn;   // scaled solution / correction
    const bool p2p = h->nranks > 1 && h->p2p.ready;
    k_exact_scale<<<nblk(cn, 256), 256, 0, s>>>(n9, cn, b, S.d_cd, rs);
    if ((rc = chol_solve(h, S.chol, S.d_S, rs, xs, s))) return rc;
    k_exact_unscale<<<nblk(n9, 256), 256, 0, s>>>(n9, xs, S.d_cd, xc, 0);
    const int steps = std::max(0, h->exact_refine);
    for (int it = 0; it < steps; ++it) {
      BA_CUDA(cudaMemcpyAsync(p, xc, sizeof(double) * (size_t)n9, cudaMemcpyDeviceToDevice, s));
      BA_CUDA(cudaMemsetAsync(S.d_scal + S_DONE, 0, sizeof(double), s));  // the product kernels honour S_DONE
      if ((rc = s_product(false))) return rc;
      if (p2p) k_seq_inc<<<1, 1, 0, s>>>(h->p2p.d_seq);
      k_exact_resid<<<1, RED_THREADS, 0, s>>>(n9, cn, b, q, S.d_cd, rs, S.d_scal, it == 0 ? 1 : 0);
      if ((rc = chol_solve(h, S.chol, S.d_S, rs, xs, s))) return rc;
      k_exact_unscale<<<nblk(n9, 256), 256, 0, s>>>(n9, xs, S.d_cd, xc, 1);
    }
    *iters = steps;
    S.last_solver = BA_SOLVER_EXACT;
    S.last_converged = 1;
    S.last_iters = steps;
    S.last_rel = -1.0;  // filled from S_REL at the next read of the scalars
    return check();
  }
  int solve(double tol, int maxit, int* iters) {
    if (!S.exact) return pcg(tol, maxit, iters);
    if (!S.factor64) {
      int rc = mixed_solve(tol, iters);
      if (rc != 1) return rc;
      // the FP32 factor did not do: the FP64 factorisation of the same sums, then the exact solve
      S.mixed_fallbacks += 1;
      S.factor64 = true;
      int info = 0;
      if ((rc = dense_factor<double>(&info))) return rc;
      cudaEventRecord(S.ev[7], s);
      BA_CUDA(cudaEventSynchronize(S.ev[7]));
      float tc = 0;
      cudaEventElapsedTime(&tc, S.ev[6], S.ev[7]);
      S.t_chol_ms += tc;
      S.chol_count += 1;
      if (info == -2) {
        h->err = "distributed Cholesky: a peer's flag never arrived (a rank is missing or stalled)";
        return BA_ERR_COMM;
      }
      if (info != 0) {
        h->err = "Cholesky of the reduced camera system: non-positive pivot";
        return BA_ERR_NUMERIC;
      }
    }
    return exact_solve(iters);
  }
  // Ac = [P Z]' S [P Z], then Ac^-1.  P block: direct assembly in one pass over the points (k_coarse_assemble;
  // BAGPU_COARSE_PRODUCTS=1: column by column with CDOF ncl applications of S, the cross-check).  Z block
  // (deflation vectors, dense): one application of S per vector.
  int coarse_setup() {
    const int m = S.mc + S.kz;
    S.coarse_gen = S.z_gen;
    if (m == 0) return BA_OK;
    int rc;
    static const bool by_products = getenv("BAGPU_COARSE_PRODUCTS") != nullptr;
    const int cpc = 28 * S.ctas_per_cluster;
    const bool p2p = h->nranks > 1 && h->p2p.ready;
    if (S.ncl > 0 && !by_products) {
      const int mcl = S.mc;
      k_coarse_diag<<<nblk(mcl, 64), 64, 0, s>>>(ncams, cpc, mcl, S.d_H, S.d_cdiag);
      BA_CUDA(cudaMemsetAsync(S.d_Acq, 0, sizeof(long long) * (size_t)(mcl * mcl), s));
      const size_t smem = sizeof(unsigned long long) * (size_t)(mcl * mcl);
      const unsigned grid = (unsigned)std::min<int64_t>(nblk(npl, 256), 148 * 2);
      k_coarse_assemble<<<grid, 256, smem, s>>>(S.d_pstart, npl, nl, h->d_cam, S.d_Jp, S.d_Vinv, cpc, mcl, S.d_cdiag,
                                                reinterpret_cast<unsigned long long*>(S.d_Acq));
      if ((rc = check())) return rc;
      if ((rc = allreduce_sum_i64(h, S.d_Acq, (size_t)(mcl * mcl)))) return rc;
      k_coarse_finish<<<nblk((int64_t)mcl * mcl, 256), 256, 0, s>>>(ncams, cpc, mcl, S.d_H, S.d_cdiag, S.d_Acq, S.d_Ac,
                                                                    m);
      if ((rc = check())) return rc;
    }
    if ((S.ncl > 0 && by_products) || S.kz > 0) {
      BA_CUDA(cudaMemsetAsync(S.d_scal + S_DONE, 0, sizeof(double), s));  // the product kernels honour S_DONE
      if (S.ncl > 0 && by_products)
        for (int col = 0; col < S.mc; ++col) {
          k_coarse_basis<<<nblk(n9, 256), 256, 0, s>>>(n9, 9 * cpc, col, p);
          if ((rc = s_product(false))) return rc;
          k_coarse_restrict<<<S.ncl, 288, 0, s>>>(ncams, cpc, m, col, q, S.d_Ac);
          if (p2p) k_seq_inc<<<1, 1, 0, s>>>(h->p2p.d_seq);
        }
      // columns S Z_j: P rows by restriction, Z rows by dot products.  DEFL_NR vectors share one pass over the
      // J blocks (k_point_solve_multi / k_cam_pass_multi); a remainder goes through the single-vector product.
      auto z_columns = [&](bool multi) -> int {
        int j = 0;
        for (; multi && j + DEFL_NR <= S.kz; j += DEFL_NR) {
          const double* zj = S.d_Z + (int64_t)j * n9;
          if (S.ntasks)
            k_point_solve_multi<DEFL_NR><<<nblk(S.ntasks, PT_THREADS / 32), PT_THREADS, 0, s>>>(
                S.d_tstart, S.ntasks, h->d_cam, h->d_pnt, h->pnt0, nl, S.d_Jp, zj, n9, S.d_Vinv, S.d_w4);
          if (S.nctasks)
            k_cam_pass_multi<DEFL_NR><<<nblk(S.nctasks, PT_THREADS / 32), PT_THREADS, 0, s>>>(
                S.d_ctask_beg, S.d_ctask_end, S.nctasks, S.d_ctask_cam, S.d_cam_t0, S.d_cam_cnt, S.d_cperm, S.d_pntc, nl,
                h->d_camtab, S.d_x4, S.d_w4, S.d_taskpart, S.d_q4, n9);
          if (S.nempty)
            for (int r = 0; r < DEFL_NR; ++r)
              k_zero_cams<<<nblk((int64_t)S.nempty * 9, 256), 256, 0, s>>>(S.d_empty_cams, (int)S.nempty, 9,
                                                                          S.d_q4 + (int64_t)r * n9, nullptr, nullptr, n9);
          if ((rc = check())) return rc;
          if ((rc = allreduce_sum(h, S.d_q4, (size_t)(DEFL_NR * n9)))) return rc;
          k_defl_hq<<<nblk(DEFL_NR * n9, 256), 256, 0, s>>>(n9, DEFL_NR, S.d_H, zj, n9, S.d_q4);
          for (int r = 0; r < DEFL_NR; ++r) {
            const double* qr = S.d_q4 + (int64_t)r * n9;
            if (S.ncl > 0) k_coarse_restrict<<<S.ncl, 288, 0, s>>>(ncams, cpc, m, S.mc + j + r, qr, S.d_Ac);
            k_defl_zdot<<<S.kz, 256, 0, s>>>(n9, S.d_Z, qr, S.d_Ac, m, S.mc, S.mc + j + r);
          }
        }
        for (; j < S.kz; ++j) {
          BA_CUDA(cudaMemcpyAsync(p, S.d_Z + (int64_t)j * n9, sizeof(double) * (size_t)n9, cudaMemcpyDeviceToDevice, s));
          if ((rc = s_product(false))) return rc;
          if (S.ncl > 0) k_coarse_restrict<<<S.ncl, 288, 0, s>>>(ncams, cpc, m, S.mc + j, q, S.d_Ac);
          k_defl_zdot<<<S.kz, 256, 0, s>>>(n9, S.d_Z, q, S.d_Ac, m, S.mc, S.mc + j);
          if (p2p) k_seq_inc<<<1, 1, 0, s>>>(h->p2p.d_seq);
        }
        return check();
      };
      const bool single_only = getenv("BAGPU_DEFL_SINGLE") != nullptr;
      if ((rc = z_columns(!single_only))) return rc;
      if (S.kz > 0 && getenv("BAGPU_DEFL_CHECK") && !single_only) {
        // cross-check: the same columns vector by vector must agree to rounding
        std::vector<double> a((size_t)m * m), bref((size_t)m * m);
        BA_CUDA(cudaMemcpyAsync(a.data(), S.d_Ac, a.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
        BA_CUDA(cudaStreamSynchronize(s));
        if ((rc = z_columns(false))) return rc;
        BA_CUDA(cudaMemcpyAsync(bref.data(), S.d_Ac, bref.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
        BA_CUDA(cudaStreamSynchronize(s));
        double worst = 0.0, big = 0.0;
        for (int row = 0; row < m; ++row)
          for (int col = S.mc; col < m; ++col) {
            worst = std::max(worst, std::fabs(a[(size_t)row * m + col] - bref[(size_t)row * m + col]));
            big = std::max(big, std::fabs(bref[(size_t)row * m + col]));
          }
        fprintf(stderr, "[bagpu] coarse setup check: multi-vector vs single-vector columns differ by %.3e (max entry %.3e)\n",
                worst, big);
        if (!(worst <= 1e-11 * big)) {
          h->err = "multi-vector Schur product disagrees with the single-vector one";
          return BA_ERR_NUMERIC;
        }
      }
      if (S.kz > 0 && S.mc > 0) k_defl_symfill<<<nblk((int64_t)S.kz * S.mc, 256), 256, 0, s>>>(m, S.mc, S.d_Ac);
    }
    k_coarse_invert<<<1, INV_THREADS, sizeof(double) * (size_t)(m * m), s>>>(m, S.d_Ac, S.d_Aci, S.d_scal);
    S.coarse_gen = S.z_gen;
    return check();
  }
  // After a solve: rebuild the deflation vectors from the Lanczos vectors it harvested (z_j / sqrt(r_j.z_j) and
  // the CG coefficients).  First long solve: the kz_base_max smallest Ritz vectors become the base.  Every later
  // solve: its smallest Ritz vectors -- the slow modes the current space misses -- orthogonalised against the
  // base, replace the DEFL_ROLL refreshed columns.  The small dense algebra runs on the host (ba_ritz.h): the
  // inputs are bit-identical on all ranks, so the vectors are too.  A failed or degenerate harvest keeps the
  // vectors as they are: nothing here can change the solution, only the iteration count.
  int defl_update(int iters) {
    const int mh = std::min(iters, S.hcap);
    const bool first = S.kz == 0;
    if (first ? (mh < 48) : (S.kz_max <= S.kz_base || mh < 16)) return BA_OK;
    std::vector<double> hc((size_t)(2 * S.hcap));
    BA_CUDA(cudaMemcpyAsync(hc.data(), S.d_hcoef, hc.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
    BA_CUDA(cudaStreamSynchronize(s));
    const double *alpha = hc.data(), *beta = hc.data() + S.hcap;
    for (int j = 0; j < mh; ++j)
      if (!(alpha[j] > 0.0) || !(beta[j] >= 0.0) || !std::isfinite(alpha[j]) || !std::isfinite(beta[j])) return BA_OK;
    std::vector<double> d, e, w, V;
    lanczos_tridiagonal(alpha, beta, mh, d, e);
    const int base = first ? 0 : S.kz_base;
    const int ncand = std::min(first ? DEFL_CAND : std::min(DEFL_CAND - base, 2 * DEFL_ROLL), mh);
    if (!tridiag_smallest(d, e, mh, ncand, w, V)) return BA_OK;
    std::vector<double> cf((size_t)mh * ncand);  // row-major mh x ncand
    for (int j = 0; j < mh; ++j)
      for (int c = 0; c < ncand; ++c) cf[(size_t)j * ncand + c] = V[(size_t)c * mh + j];
    for (double v : cf)
      if (!std::isfinite(v)) return BA_OK;
    BA_CUDA(cudaMemcpyAsync(S.d_dsmall, cf.data(), cf.size() * sizeof(double), cudaMemcpyHostToDevice, s));
    if (base)
      BA_CUDA(cudaMemcpyAsync(S.d_Zcand, S.d_Z, sizeof(double) * (size_t)(n9 * base), cudaMemcpyDeviceToDevice, s));
    k_defl_gemm<<<nblk(n9, 128), 128, 0, s>>>(n9, mh, ncand, S.d_harv, S.d_dsmall, S.d_Zcand + (int64_t)base * n9);
    const int n = base + ncand;
    double* d_G = S.d_dsmall + (size_t)S.hcap * DEFL_CAND;
    k_defl_gram<<<(unsigned)(n * n), 256, 0, s>>>(n9, n, S.d_Zcand, d_G);
    int rc = check();
    if (rc) return rc;
    std::vector<double> G((size_t)n * n), Cm;
    BA_CUDA(cudaMemcpyAsync(G.data(), d_G, G.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
    BA_CUDA(cudaStreamSynchronize(s));  // (also: cf is read by the copy above before it goes out of scope)
    for (double v : G)
      if (!std::isfinite(v)) return BA_OK;
    const int kept = select_orthonormal(G, n, first ? S.kz_base_max : S.kz_max, 1e-3, Cm);
    if (kept == 0 || (!first && kept < S.kz_base)) return BA_OK;
    if (!first)  // the base columns must have survived as the first kept ones
      for (int j = 0; j < S.kz_base; ++j)
        if (!(std::fabs(Cm[(size_t)j * kept + j]) > 0.5)) return BA_OK;
    BA_CUDA(cudaMemcpyAsync(S.d_dsmall, Cm.data(), sizeof(double) * (size_t)n * kept, cudaMemcpyHostToDevice, s));
    k_defl_gemm<<<nblk(n9, 128), 128, 0, s>>>(n9, n, kept, S.d_Zcand, S.d_dsmall, S.d_Z);
    BA_CUDA(cudaStreamSynchronize(s));  // Cm goes out of scope
    S.kz = kept;
    S.z_gen += 1;
    if (first) {
      S.kz_base = kept;
      S.defl_iters_first = iters;
    }
    return check();
  }
  int read_scalars() {
    BA_CUDA(cudaMemcpyAsync(S.h_scal, S.d_scal, S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, s));
    BA_CUDA(cudaStreamSynchronize(s));
    return BA_OK;
  }
  // q = S p: point-major pass, camera-major pass, sum over ranks (fused over peer memory, or NCCL), finalise
  int s_product(bool checks, bool finalize = true) {
    int rc;
    if (S.ntasks)
      k_point_solve<0><<<nblk(S.ntasks, PT_THREADS / 32), PT_THREADS, 0, s>>>(
          S.d_tstart, S.ntasks, h->d_cam, h->d_pnt, h->pnt0, nl, S.d_Jp, S.d_F, p, S.d_Vinv, S.d_gp, S.d_w,
          nullptr, nullptr, S.d_part, S.d_scal);
    if (checks && (rc = check())) return rc;
    const ba_p2p_state& P = h->p2p;
    const bool p2p = h->nranks > 1 && P.ready;
    double* mail = p2p ? reinterpret_cast<double*>(static_cast<char*>(P.block) + P.mail_off) : nullptr;
    if (S.nctasks)
      k_cam_pass<2><<<nblk(S.nctasks, PT_THREADS / 32), PT_THREADS, 0, s>>>(
          S.d_ctask_beg, S.d_ctask_end, S.nctasks, S.d_ctask_cam, S.d_cam_t0, S.d_cam_cnt, S.d_cperm, S.d_pntc, nl,
          h->d_camtab, S.d_x4, S.d_F, S.d_w, S.d_T, S.d_taskpart, q, S.d_scal, mail, P.d_seq, n9);
    if (S.nempty)
      k_zero_cams<<<nblk((int64_t)S.nempty * 9, 256), 256, 0, s>>>(S.d_empty_cams, (int)S.nempty, 9, q, mail, P.d_seq,
                                                                   n9);
    if (checks && (rc = check())) return rc;
    const int nvb = (int)nblk(n9, VEC_ROWS);
    if (!p2p && (rc = allreduce_sum(h, q, (size_t)n9))) return rc;
    if (!finalize) return BA_OK;  // the fused small-system kernel finishes the product itself
    if (p2p) {
      // the sum over ranks is fused into the kernel that consumes it (peer loads over NVLink)
      k_pcg_q<true><<<nvb, VEC_THREADS, 0, s>>>(n9, S.d_H, p, q, S.d_pcgpart, S.d_scal, P.d_mail, P.d_flags, P.d_seq,
                                                 h->nranks, h->rank);
    } else {
      k_pcg_q<false><<<nvb, VEC_THREADS, 0, s>>>(n9, S.d_H, p, q, S.d_pcgpart, S.d_scal, nullptr, nullptr, nullptr, 1,
                                                  0);
    }
    return checks ? check() : BA_OK;
  }
  // the vector half of a PCG iteration (or of its initialisation)
  template <bool INIT>
  void pcg_vectors(double tol) {
    const int nvb = (int)nblk(n9, VEC_ROWS);
    double* ppq = S.d_pcgpart;
    double* prz = S.d_pcgpart + nvb;
    const int m = S.mc + S.kz;  // coarse unknowns: cluster components, then deflation vectors
    const bool coarse = m > 0;
    const bool p2p = h->nranks > 1 && h->p2p.ready;
    k_pcg_xr<INIT><<<nvb, VEC_THREADS, 0, s>>>(n9, nvb, b, S.d_Minv, p, q, xc, r, z, ppq, prz, S.d_scal,
                                               S.ncl > 0 ? S.d_cpart : nullptr, S.d_Z, S.kz, S.d_zpart);
    if (coarse)
      k_pcg_coarse<<<1, COARSE_THREADS, 0, s>>>(nvb, S.ctas_per_cluster, m, S.d_cpart, S.d_Aci, S.d_yc, S.d_scal, INIT ? 1 : 0,
                                     S.kz, S.d_zpart);
    k_pcg_p<INIT><<<nvb, VEC_THREADS, 0, s>>>(n9, nvb, z, p, ppq, prz, S.d_scal, tol,
                                              (!INIT && p2p) ? h->p2p.d_seq : nullptr, coarse ? S.d_yc : nullptr,
                                              std::max(S.ctas_per_cluster, 1), S.d_Z, S.kz, S.mc,
                                              S.kz_max > 0 ? S.d_harv : nullptr, S.hcap, S.d_hcoef);
  }
  // one PCG iteration
  int pcg_iteration(double tol, bool checks) {
    const bool small = n9 <= small_rows_max();
    int rc = s_product(checks, !small);
    if (rc) return rc;
    if (small) {
      const ba_p2p_state& P = h->p2p;
      const int cpc = 28 * std::max(S.ctas_per_cluster, 1);
      if (h->nranks > 1 && P.ready)
        k_pcg_small<true><<<1, SMALL_THREADS, 0, s>>>(n9, ncams, S.d_H, S.d_Minv, p, q, xc, r, z, S.d_scal, tol, P.d_mail,
                                                       P.d_flags, P.d_seq, h->nranks, h->rank, S.d_Aci, S.mc, cpc);
      else
        k_pcg_small<false><<<1, SMALL_THREADS, 0, s>>>(n9, ncams, S.d_H, S.d_Minv, p, q, xc, r, z, S.d_scal, tol, nullptr,
                                                        nullptr, nullptr, 1, 0, S.d_Aci, S.mc, cpc);
    } else {
      pcg_vectors<false>(tol);
    }
    return checks ? check() : BA_OK;
  }
  // block-Jacobi PCG on S dc = b; result in xc.  Convergence is decided on the device (S_DONE): the host
  // only polls every PCG_POLL iterations, and those iterations are ONE CUDA-graph launch (the 3 kernels
  // and the NCCL allreduce of each iteration captured once per handle); launches of a finished solve exit
  // at once.  BAGPU_NO_GRAPH=1 or BAGPU_DEBUG_SYNC=1 fall back to plain launches.
  int pcg(double tol, int maxit, int* iters) {
    constexpr int PCG_POLL = 8;
    static const bool no_graph = getenv("BAGPU_NO_GRAPH") != nullptr || getenv("BAGPU_DEBUG_SYNC") != nullptr;
    int rc;
    if (S.kz_max > 0 && !S.d_harv) {  // harvest buffers, sized once by the first solve's iteration limit
      S.hcap = std::max(1, std::min(DEFL_HCAP, maxit));
      if ((rc = dmalloc(h, &S.d_harv, (size_t)(n9 * S.hcap)))) return rc;
      if ((rc = dmalloc(h, &S.d_hcoef, (size_t)(2 * S.hcap)))) return rc;
      if ((rc = dmalloc(h, &S.d_dsmall, (size_t)(S.hcap + DEFL_CAND) * DEFL_CAND))) return rc;
    }
    if (S.coarse_gen != S.z_gen && (rc = coarse_setup())) return rc;  // vectors changed since the last factor()
    pcg_vectors<true>(tol);
    if ((rc = check())) return rc;
    if (!no_graph && !S.pcg_graph_off && (!S.pcg_graph || S.pcg_graph_tol != tol || S.pcg_graph_kz != S.kz)) {
      if (S.pcg_graph) cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(S.pcg_graph));
      S.pcg_graph = nullptr;
      cudaGraph_t g = nullptr;
      if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();      // e.g. the handle runs on the legacy default stream, which cannot be captured
        S.pcg_graph_off = true;  // plain launches from now on
      } else {
        for (int i = 0; i < PCG_POLL && !rc; ++i) rc = pcg_iteration(tol, false);
        const cudaError_t e = cudaStreamEndCapture(s, &g);
        if (rc) return rc;
        BA_CUDA(e);
        cudaGraphExec_t ge = nullptr;
        BA_CUDA(cudaGraphInstantiate(&ge, g, 0));
        cudaGraphDestroy(g);
        S.pcg_graph = ge;
        S.pcg_graph_tol = tol;
        S.pcg_graph_kz = S.kz;
      }
    }
    int launched = 0;
    for (;;) {
      const int chunk = std::min(PCG_POLL, maxit - launched);
      if (chunk == PCG_POLL && S.pcg_graph) {
        BA_CUDA(cudaGraphLaunch(reinterpret_cast<cudaGraphExec_t>(S.pcg_graph), s));
      } else {
        for (int i = 0; i < chunk; ++i)
          if ((rc = pcg_iteration(tol, true))) return rc;
      }
      launched += chunk;
      if ((rc = check())) return rc;
      if ((rc = read_scalars())) return rc;
      if (S.h_scal[S_DONE] != 0.0 || launched >= maxit) break;
    }
    *iters = (int)S.h_scal[S_ITERS];
    S.last_solver = BA_SOLVER_PCG;
    S.last_converged = S.h_scal[S_DONE] == 1.0;
    S.last_iters = *iters;
    S.last_rel = S.h_scal[S_REL];
    if (S.h_scal[S_ERR] == 3.0) {
      // a peer never published its partial sum (k_pcg_q wait timed out): a communication failure, not a numeric
      // one -- leave the deflation state alone and say so
      h->err = "peer-memory exchange timed out (a rank is missing or stalled)";
      return BA_ERR_COMM;
    }
    if (S.h_scal[S_CBAD] != 0.0) {
      // the coarse matrix was not positive definite: the device zeroed its inverse, so this solve ran (correctly)
      // with plain block-Jacobi.  Nearly dependent deflation vectors are the usual cause: drop them for good.
      if (trace_on()) fprintf(stderr, "[bagpu] coarse matrix not positive definite: solve ran without the coarse level\n");
      BA_CUDA(cudaMemsetAsync(S.d_scal + S_CBAD, 0, sizeof(double), s));
      if (S.kz > 0) {
        S.kz = S.kz_base = S.kz_max = 0;
        S.z_gen += 1;
      }
    }
    if (S.kz > 0 && S.h_scal[S_DONE] == 2.0) {
      // the extended coarse matrix lost definiteness (nearly dependent vectors): drop the deflation vectors
      // for good, rebuild the cluster level and solve again
      S.kz = S.kz_base = S.kz_max = 0;
      S.z_gen += 1;
      BA_CUDA(cudaMemsetAsync(S.d_scal + S_ERR, 0, sizeof(double), s));
      BA_CUDA(cudaMemsetAsync(S.d_scal + S_DONE, 0, sizeof(double), s));
      if ((rc = coarse_setup())) return rc;
      return pcg(tol, maxit, iters);
    }
    // (a solve stopped by the iteration limit has harvested just as valid Lanczos vectors as a converged one --
    //  and is the one that needs the deflation most)
    if (S.kz_max > 0 && S.h_scal[S_DONE] != 2.0 && (rc = defl_update(*iters))) return rc;
    if (S.h_scal[S_DONE] == 2.0) {
      h->err = "PCG breakdown (non-finite or non-positive curvature in the reduced camera system)";
      return BA_ERR_NUMERIC;
    }
    return BA_OK;
  }
  // delta from xc: camera part copied, point part by back-substitution; sum (J delta + r)^2 -> S_DR2
  int backsub(bool keep_dr) {
    if (keep_dr && !S.d_dr) {
      int rc = dmalloc(h, &S.d_dr, (size_t)nl);
      if (rc) return rc;
    }
    k_copy_cam_delta<<<nblk(n9, 256), 256, 0, s>>>(n9, xc, S.d_delta + 3 * h->npnts);
    const unsigned nb = nblk(S.ntasks, PT_THREADS / 32);
    k_point_solve<1><<<nb, PT_THREADS, 0, s>>>(S.d_tstart, S.ntasks, h->d_cam, h->d_pnt, h->pnt0, nl, S.d_Jp, S.d_F,
                                              xc, S.d_Vinv, S.d_gp, nullptr, S.d_delta, keep_dr ? S.d_dr : nullptr,
                                              S.d_part, S.d_scal);
    return reduce_to(nb, 1, S_DR2);
  }
  // xt = x + delta / div, norms, residual norm at xt; then every local sum is allreduced and read back
  int trial(double div) {
    const int64_t n = 3 * npl + n9;
    const unsigned nb = (unsigned)std::min<int64_t>(nblk(n, 256), 1024);
    k_step_update<<<nb, 256, 0, s>>>(S.d_x, S.d_delta, S.d_xt, 3 * h->pnt0, 3 * h->pnt1, 3 * h->npnts, h->nvar(), div,
                                     S.d_part);
    k_reduce_parts<<<4, RED_THREADS, 0, s>>>(S.d_part, nb, 4, S.d_part + S.npart, 0);
    // scatter the four sums to their slots: dp2, xp2 local; dc2, xc2 replicated
    BA_CUDA(cudaMemcpyAsync(S.d_scal + S_DP2, S.d_part + S.npart, 2 * sizeof(double), cudaMemcpyDeviceToDevice, s));
    BA_CUDA(cudaMemcpyAsync(S.d_scal + S_DC2, S.d_part + S.npart + 2, 2 * sizeof(double), cudaMemcpyDeviceToDevice, s));
    launch_cam_precompute(S.d_xt, h->npnts, ncams, S.d_camt, s);
    const unsigned nbo = nblk(nl, PT_THREADS);
    k_lm_trial<<<nbo, PT_THREADS, 0, s>>>(h->d_cam, h->d_pnt, h->d_pt2d, S.d_xt, S.d_camt, S.d_part, nl);
    int rc = reduce_to(nbo, 1, S_TR2);
    if (rc) return rc;
    if ((rc = allreduce_sum(h, S.d_scal + S_DR2, 4))) return rc;  // DR2, DP2, XP2, TR2 are contiguous
    return read_scalars();
  }
  int ls_dr(double dd) {
    const unsigned nb = nblk(nl, PT_THREADS);
    k_ls_dr<<<nb, PT_THREADS, 0, s>>>(S.d_dr, S.d_F, nl, dd, S.d_part);
    return reduce_to(nb, 1, S_DR2);
  }
  // full-length host copy of a device vector laid out like x: point slices are summed over ranks
  int gather_full(double* dvec, double* host) {
    if (h->nranks > 1) {
      // zero what this rank does not own (other ranks' points; cameras except on rank 0), then sum
      if (h->pnt0 > 0) BA_CUDA(cudaMemsetAsync(dvec, 0, sizeof(double) * 3 * (size_t)h->pnt0, s));
      if (h->pnt1 < h->npnts)
        BA_CUDA(cudaMemsetAsync(dvec + 3 * h->pnt1, 0, sizeof(double) * 3 * (size_t)(h->npnts - h->pnt1), s));
      if (h->rank != 0) BA_CUDA(cudaMemsetAsync(dvec + 3 * h->npnts, 0, sizeof(double) * (size_t)n9, s));
      int rc = allreduce_sum(h, dvec, (size_t)h->nvar());
      if (rc) return rc;
    }
    BA_CUDA(cudaMemcpyAsync(host, dvec, sizeof(double) * (size_t)h->nvar(), cudaMemcpyDeviceToHost, s));
    BA_CUDA(cudaStreamSynchronize(s));
    return BA_OK;
  }
};

inline double sq_to_half_norm2(double sumsq) {  // norm(v)^2 / 2 the way the reference forms it
  const double n = sqrt(sumsq);
  return n * n / 2;
}

}  // namespace
}  // namespace ba

extern "C" {

void ba_lm_default_params(ba_lm_params* p) {
  if (!p) return;
  const double eps = 2.220446049250313e-16;
  p->restol = p->ortol = p->rtol = cbrt(eps);  // eps^(1/3), src/lm.jl:21-24
  p->satol = p->srtol = p->oatol = p->atol = sqrt(eps);
  p->nu_d = 3; p->nu_m = 3; p->lambda = 30; p->delta_d = 2;  // src/lm.jl:25
  p->ite_max = 200;                                             // src/lm.jl:26
  p->linesearch = 0;
  p->pcg_max_iter = 1000;
  p->pcg_tol = 1e-13;
  p->solver = BA_SOLVER_AUTO;
  p->reserved = 0;
}

int ba_lm_step(ba_handle* h, const double* x, double lambda, double pcg_tol, int32_t pcg_max_iter, double* delta,
               double* dr2, double* obj, double* jtr, int32_t* pcg_iters) {
  if (!h || !x || !delta) {
    if (h) h->err = "null argument";
    return BA_ERR_ARG;
  }
  if (h->group) return ba::group_lm_step(h, x, lambda, pcg_tol, pcg_max_iter, delta, dr2, obj, jtr, pcg_iters);
  int rc = ba::lm_prepare_all(h);
  if (rc) return rc;
  BA_CUDA(cudaSetDevice(h->device));
  ba::Solver sv(h);
  ba_lm_state& S = h->lm;
  BA_CUDA(cudaMemcpyAsync(S.d_x, x, sizeof(double) * (size_t)h->nvar(), cudaMemcpyHostToDevice, h->stream));
  BA_CUDA(cudaMemsetAsync(S.d_scal, 0, ba::S_COUNT * sizeof(double), h->stream));
  if ((rc = sv.build(S.d_x))) return rc;
  if ((rc = sv.factor(lambda))) return rc;
  int it = 0;
  rc = sv.solve(pcg_tol, pcg_max_iter, &it);
  if (pcg_iters) *pcg_iters = it;
  if (rc) return rc;
  if ((rc = sv.backsub(false))) return rc;
  if ((rc = ba::allreduce_sum(h, S.d_scal + ba::S_DR2, 1))) return rc;
  if ((rc = sv.read_scalars())) return rc;
  if (S.last_rel < 0.0) S.last_rel = S.h_scal[ba::S_REL];
  if (S.h_scal[ba::S_ERR] != 0.0) {
    h->err = "Schur diagonal block not positive definite";
    return BA_ERR_NUMERIC;
  }
  if (dr2) *dr2 = ba::sq_to_half_norm2(S.h_scal[ba::S_DR2]);
  if (obj) *obj = ba::sq_to_half_norm2(S.h_scal[ba::S_F2]);
  if (jtr) {  // J'r = -[g_p; g_c]
    std::vector<double> gp((size_t)(3 * sv.npl)), ug((size_t)(sv.ncams * ba::NV));
    BA_CUDA(cudaMemcpyAsync(gp.data(), S.d_gp, gp.size() * 8, cudaMemcpyDeviceToHost, h->stream));
    BA_CUDA(cudaMemcpyAsync(ug.data(), S.d_Ug, ug.size() * 8, cudaMemcpyDeviceToHost, h->stream));
    BA_CUDA(cudaStreamSynchronize(h->stream));
    // other ranks' point entries are not available here: left zero (documented in bagpu.h)
    std::fill(jtr, jtr + h->nvar(), 0.0);
    for (int64_t i = 0; i < 3 * sv.npl; ++i) jtr[3 * h->pnt0 + i] = -gp[(size_t)i];
    for (int64_t c = 0; c < sv.ncams; ++c)
      for (int j = 0; j < 9; ++j) jtr[3 * h->npnts + 9 * c + j] = -ug[(size_t)(c * ba::NV + 45 + j)];
  }
  return sv.gather_full(S.d_delta, delta);
}

int ba_lm_solve(ba_handle* h, double* x_inout, const ba_lm_params* prm_in, ba_lm_stats* st, ba_iter_cb cb,
                void* user) {
  using namespace ba;
  if (!h || !x_inout) {
    if (h) h->err = "null argument";
    return BA_ERR_ARG;
  }
  if (h->group) return ba::group_lm_solve(h, x_inout, prm_in, st, cb, user);
  ba_lm_params prm;
  if (prm_in) prm = *prm_in; else ba_lm_default_params(&prm);
  static const bool trace = getenv("BAGPU_TRACE") != nullptr;  // per-iteration phase times on stderr
  const auto wall_prep = std::chrono::steady_clock::now();
  int rc;
  if (prm.solver != BA_SOLVER_AUTO && prm.solver != h->solver && (rc = ba_set_solver(h, prm.solver))) return rc;
  if ((rc = lm_prepare_all(h))) return rc;
  BA_CUDA(cudaSetDevice(h->device));
  Solver sv(h);
  ba_lm_state& S = h->lm;
  const auto wall0 = std::chrono::steady_clock::now();
  if (trace)
    fprintf(stderr, "[bagpu] lm_prepare %.1f ms\n", std::chrono::duration<double>(wall0 - wall_prep).count() * 1e3);
  const double t_prepare = std::chrono::duration<double>(wall0 - wall_prep).count() * 1e3;  // ~0 on a prepared handle
  double t_eval = 0, t_asm = 0, t_pcg = 0, t_back = 0, worst_rel = 0;
  S.t_schur_ms = S.t_chol_ms = 0.0;
  S.chol_count = 0;
  S.mixed_fallbacks = 0;
  int64_t pcg_total = 0, capped = 0;
  auto rec = [&](int i) { cudaEventRecord(S.ev[i], h->stream); };
  auto lap = [&](int a, int b) {
    float ms = 0;
    cudaEventElapsedTime(&ms, S.ev[a], S.ev[b]);
    return (double)ms;
  };

  BA_CUDA(cudaMemcpyAsync(S.d_x, x_inout, sizeof(double) * (size_t)h->nvar(), cudaMemcpyHostToDevice, h->stream));
  BA_CUDA(cudaMemsetAsync(S.d_scal, 0, S_COUNT * sizeof(double), h->stream));
  int64_t iter = 0;
  // src/lm.jl:39-59
  rec(0);
  if ((rc = sv.build(S.d_x))) return rc;
  rec(1);
  if ((rc = sv.read_scalars())) return rc;
  t_eval += lap(0, 1);
  double norm_r = sqrt(S.h_scal[S_F2]);
  double obj = norm_r * norm_r / 2;
  double norm_Jtr = sqrt(S.h_scal[S_GP2] + S.h_scal[S_GC2]);
  double lambda = std::max(prm.lambda, 1e10 / norm_Jtr);
  double norm_delta = 0, dr2 = 0;
  const double eps_first_order = prm.atol + prm.rtol * norm_Jtr;  // :107
  double old_obj = obj;
  bool small_step = false, first_order = norm_Jtr < eps_first_order, small_residual = norm_r < prm.restol;
  bool small_obj_change = false, tired = iter > prm.ite_max, fail = false, fail2 = false;
  bool need_factor = true;

  while (!(small_step || first_order || small_residual || small_obj_change || tired || fail || fail2)) {
    iter += 1;
    // damped solve (stands in for ldl_factorize + ldl_solve!, src/lm.jl:175-180,227-229)
    rec(0);
    if (need_factor) {
      rc = sv.factor(lambda);
      if (rc == BA_ERR_NUMERIC) {  // non-positive pivot: the SQDException of src/ldl_aux.jl:199 -> status exception
        fail2 = true;
        continue;
      }
      if (rc) return rc;
      need_factor = false;
    }
    rec(1);
    int pit = 0;
    rc = sv.solve(prm.pcg_tol, prm.pcg_max_iter, &pit);
    pcg_total += pit;
    if (rc == BA_ERR_NUMERIC || S.h_scal[S_ERR] != 0.0) {  // like SQDException / NaN step: status exception
      fail2 = true;
      continue;
    }
    if (rc) return rc;
    rec(2);
    if ((rc = sv.backsub(prm.linesearch != 0))) return rc;
    rec(3);
    // src/lm.jl:251-254
    if ((rc = sv.trial(1.0))) return rc;
    rec(4);
    BA_CUDA(cudaEventSynchronize(S.ev[4]));
    t_asm += lap(0, 1); t_pcg += lap(1, 2); t_back += lap(2, 3); t_eval += lap(3, 4);
    if (S.last_rel < 0.0) S.last_rel = S.h_scal[S_REL];  // exact solve: left on the device, read by trial()
    if (!S.last_converged) capped += 1;
    if (S.last_rel > worst_rel) worst_rel = S.last_rel;
    if (trace) {
      fprintf(stderr, "[bagpu] iter %lld: factor %.2f ms, %s %.2f ms (%d iterations, %d deflation vectors, residual %.1e), "
              "backsub %.2f ms, trial %.2f ms, wall %.1f ms\n", (long long)iter, lap(0, 1), S.exact ? "exact solve" : "pcg",
              lap(1, 2), pit, S.kz, S.last_rel, lap(2, 3), lap(3, 4),
              std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count() * 1e3);
      if (!S.last_converged)
        fprintf(stderr, "[bagpu] iter %lld: PCG stopped at pcg_max_iter = %d with relative residual %.2e > pcg_tol = %.1e: "
                "the step is inexact\n", (long long)iter, (int)prm.pcg_max_iter, S.last_rel, prm.pcg_tol);
    }
    dr2 = sq_to_half_norm2(S.h_scal[S_DR2]);
    double norm_rsuiv = sqrt(S.h_scal[S_TR2]);
    double obj_suiv = norm_rsuiv * norm_rsuiv / 2;
    // :257-260
    double pred = obj - dr2, ared = obj - obj_suiv;
    bool step_accepted = ared >= 1e-4 * pred;
    bool acc_str = step_accepted && dr2 <= obj;
    int ntimes = 0;
    // :264-295
    if (prm.linesearch) {
      while (!step_accepted && ntimes < 4) {
        if ((rc = sv.ls_dr(prm.delta_d))) return rc;
        if ((rc = sv.trial(prm.delta_d))) return rc;
        norm_rsuiv = sqrt(S.h_scal[S_TR2]);
        obj_suiv = norm_rsuiv * norm_rsuiv / 2;
        dr2 = sq_to_half_norm2(S.h_scal[S_DR2]);
        pred = obj - dr2; ared = obj - obj_suiv;
        step_accepted = ared >= 1e-4 * pred;
        acc_str = step_accepted && dr2 <= obj;
        ntimes += 1;
      }
    }
    // :297-302
    norm_delta = sqrt(S.h_scal[S_DP2] + S.h_scal[S_DC2]);
    if (std::isnan(norm_delta)) {
      fail2 = true;
      continue;
    }
    // :304
    if (cb) {
      ba_lm_row row;
      row.iter = iter; row.f = obj; row.df = old_obj - obj; row.dfeas = norm_Jtr; row.lambda = lambda;
      row.delta_norm = norm_delta; row.rho = ared / pred; row.accepted = step_accepted; row.acc_str = acc_str;
      row.pcg_iters = pit; row.ntimes = ntimes;
      row.solver = S.last_solver; row.converged = S.last_converged; row.solve_rel = S.last_rel;
      cb(&row, user);
    }
    if (!step_accepted) {
      // :306-325
      lambda = std::max(lambda, 1 / norm_delta) * pow(prm.nu_m, (double)(ntimes + 1));
      need_factor = true;
    } else {
      // :328-338
      if (ntimes > 0) lambda /= pow(prm.nu_d, (double)(ntimes - 1));
      else lambda /= prm.nu_d;
      if (ared >= 0.9 * pred) lambda /= prm.nu_d;
      lambda = std::max(1.0e-8, lambda);
      std::swap(S.d_x, S.d_xt);
      const double norm_x = sqrt(S.h_scal[S_XP2] + S.h_scal[S_XC2]);
      // :341-371: Jacobian, residual, J'r at the new x
      rec(0);
      if ((rc = sv.build(S.d_x))) return rc;
      rec(1);
      if ((rc = sv.read_scalars())) return rc;
      t_eval += lap(0, 1);
      old_obj = obj;
      norm_r = norm_rsuiv;
      obj = obj_suiv;
      need_factor = true;
      // :374-379
      norm_Jtr = sqrt(S.h_scal[S_GP2] + S.h_scal[S_GC2]);
      small_step = norm_delta < prm.satol + prm.srtol * norm_x;
      first_order = norm_Jtr < eps_first_order;
      small_residual = norm_r < prm.restol;
      small_obj_change = old_obj - obj < prm.oatol + prm.ortol * old_obj;
    }
    tired = iter > prm.ite_max;  // :382 (elapsed_time is never updated inside the loop)
  }
  int status = 0;  // :391-405
  if (small_step) status = 1;
  else if (first_order) status = 2;
  else if (small_residual) status = 3;
  else if (small_obj_change) status = 4;
  else if (fail) status = 5;
  else if (fail2) status = 6;
  else if (tired) status = 7;
  if ((rc = sv.gather_full(S.d_x, x_inout))) return rc;
  if (st) {
    st->status = status; st->pad = 0; st->iter = iter; st->objective = obj; st->dual_feas = norm_Jtr;
    st->lambda_final = lambda;
    st->elapsed_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count();
    st->pcg_iters_total = pcg_total;
    st->t_eval_ms = t_eval; st->t_assemble_ms = t_asm; st->t_pcg_ms = t_pcg; st->t_backsub_ms = t_back;
    st->capped_solves = capped; st->worst_solve_rel = worst_rel; st->t_prepare_ms = t_prepare;
    st->t_schur_ms = S.t_schur_ms; st->t_chol_ms = S.t_chol_ms;
    st->chol_n = S.exact ? S.cn : 0; st->chol_count = S.chol_count;
    st->mixed_fallbacks = S.mixed_fallbacks;
  }
  return BA_OK;
}

}  // extern "C"
