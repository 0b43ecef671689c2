// ba_group.cu -- several GPUs behind ONE handle in ONE process (ba_create_multi).
//
// The reference's callers are a single Julia process (src/main.jl:27-30: one BALNLPModel, one Levenberg_Marquardt
// call); SURVEY.md section 8(b) therefore asks for the multi-GPU mode to live inside the library, invisible to the
// caller.  A group handle owns one observation-sharded sub-handle per device (the same shards, kernels and
// collectives as the one-process-per-GPU mode of ba_create_sharded) and one host thread per sub-handle.  Every
// supported entry point fans out to the sub-handles and presents full-length arrays to the caller: per-observation
// outputs (cx, vals, rows, cols, Jv) are written by each rank straight into its slice of the caller's array, and
// vectors laid out like x (delta, J'v, the LM solution) are returned from rank 0 after the ranks' allreduce.
// NCCL communicators are created in-process (ncclCommInitRank from the rank threads); the peer-memory mailboxes of
// the PCG exchange need no CUDA IPC here: peer access is enabled and the raw device pointers are shared.
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include "ba_internal.h"

struct ba_group {
  std::vector<ba_handle*> subs;
  std::vector<std::thread> workers;
  std::mutex m;
  std::condition_variable cv, cv_done;
  std::function<int(ba_handle*, int)> task;
  uint64_t gen = 0;
  int pending = 0;
  std::vector<int> rcs;
  std::vector<uint64_t> seen;
  bool stop = false;
};

namespace ba {
namespace {

void worker_main(ba_group* G, int r) {
  cudaSetDevice(G->subs[(size_t)r]->device);
  for (;;) {
    std::function<int(ba_handle*, int)> f;
    {
      std::unique_lock<std::mutex> g(G->m);
      G->cv.wait(g, [&] { return G->stop || G->gen != G->seen[(size_t)r]; });
      if (G->stop) return;
      G->seen[(size_t)r] = G->gen;
      f = G->task;
    }
    const int rc = f(G->subs[(size_t)r], r);
    {
      std::lock_guard<std::mutex> g(G->m);
      G->rcs[(size_t)r] = rc;
      if (--G->pending == 0) G->cv_done.notify_all();
    }
  }
}

}  // namespace

// run f(sub, rank) on every rank thread at once; first non-zero status wins and its message moves to the parent
int group_run(ba_handle* h, const std::function<int(ba_handle*, int)>& f) {
  ba_group* G = h->group;
  {
    std::unique_lock<std::mutex> g(G->m);
    G->task = f;
    G->pending = (int)G->subs.size();
    G->gen += 1;
  }
  G->cv.notify_all();
  {
    std::unique_lock<std::mutex> g(G->m);
    G->cv_done.wait(g, [&] { return G->pending == 0; });
  }
  for (size_t r = 0; r < G->subs.size(); ++r)
    if (G->rcs[r]) {
      h->err = "rank " + std::to_string(r) + ": " + G->subs[r]->err;
      return G->rcs[r];
    }
  return BA_OK;
}

void group_release(ba_handle* h) {
  ba_group* G = h->group;
  if (!G) return;
  {
    std::lock_guard<std::mutex> g(G->m);
    G->stop = true;
  }
  G->cv.notify_all();
  for (auto& t : G->workers) t.join();
  for (ba_handle* s : G->subs)
    if (s) ba_destroy(s);
  delete G;
  h->group = nullptr;
}

// peer-memory mailboxes without IPC: every rank allocates its block, enables access to the peers' devices and
// takes their raw pointers
int p2p_attach(ba_handle* h, void* const* blocks);  // ba_comm.cu
int p2p_alloc_block(ba_handle* h);                  // ba_comm.cu

int group_residual(ba_handle* h, const double* x, double* cx, double* vals) {
  return group_run(h, [=](ba_handle* s, int) {
    double* c = cx ? cx + 2 * s->obs0 : nullptr;
    double* v = vals ? vals + 24 * s->obs0 : nullptr;
    if (c && v) return ba_residual_jac(s, x, c, v);
    return c ? ba_residual(s, x, c) : ba_jac_coord(s, x, v);
  });
}

int group_jac_structure(ba_handle* h, int64_t* rows, int64_t* cols) {
  return group_run(h, [=](ba_handle* s, int) { return ba_jac_structure(s, rows + 24 * s->obs0, cols + 24 * s->obs0); });
}

int group_jprod(ba_handle* h, const double* x, const double* v, double* Jv) {
  return group_run(h, [=](ba_handle* s, int) { return ba_jprod(s, x, v, Jv + 2 * s->obs0); });
}

int group_jtprod(ba_handle* h, const double* x, const double* v, double* Jtv) {
  const size_t nvar = (size_t)h->nvar();
  std::vector<std::vector<double>> tmp(h->group->subs.size());
  return group_run(h, [&, x, v, Jtv, nvar](ba_handle* s, int r) {
    double* out = Jtv;
    if (r != 0) {
      tmp[(size_t)r].resize(nvar);
      out = tmp[(size_t)r].data();
    }
    return ba_jtprod(s, x, v + 2 * s->obs0, out);  // summed over the ranks inside (NCCL): every rank holds the total
  });
}

int group_lm_step(ba_handle* h, const double* x, double lambda, double pcg_tol, int32_t pcg_max_iter, double* delta,
                  double* dr2, double* obj, double* jtr, int32_t* pcg_iters) {
  const size_t nvar = (size_t)h->nvar(), nr = h->group->subs.size();
  std::vector<std::vector<double>> dl(nr), jt(nr);
  const int rc = group_run(h, [&, x, lambda, pcg_tol, pcg_max_iter, delta, dr2, obj, jtr, pcg_iters](ba_handle* s, int r) {
    double* d = delta;
    if (r != 0) {
      dl[(size_t)r].resize(nvar);
      d = dl[(size_t)r].data();
    }
    double* j = nullptr;
    if (jtr) {
      jt[(size_t)r].resize(nvar);
      j = jt[(size_t)r].data();
    }
    double a = 0, b = 0;
    int32_t it = 0;
    const int rcs = ba_lm_step(s, x, lambda, pcg_tol, pcg_max_iter, d, &a, &b, j, &it);
    if (r == 0) {
      if (dr2) *dr2 = a;
      if (obj) *obj = b;
      if (pcg_iters) *pcg_iters = it;
    }
    return rcs;
  });
  if (rc) return rc;
  if (jtr) {  // each rank knows J'r on its own points and on the cameras
    memcpy(jtr, jt[0].data(), sizeof(double) * nvar);
    for (size_t r = 1; r < nr; ++r) {
      const ba_handle* s = h->group->subs[r];
      memcpy(jtr + 3 * s->pnt0, jt[r].data() + 3 * s->pnt0, sizeof(double) * 3 * (size_t)s->npnts_l());
    }
  }
  return BA_OK;
}

int group_lm_solve(ba_handle* h, double* x_inout, const ba_lm_params* p, ba_lm_stats* st, ba_iter_cb cb, void* user) {
  const size_t nvar = (size_t)h->nvar(), nr = h->group->subs.size();
  std::vector<std::vector<double>> xs(nr);
  return group_run(h, [&, x_inout, p, st, cb, user, nvar](ba_handle* s, int r) {
    if (r == 0) return ba_lm_solve(s, x_inout, p, st, cb, user);  // rows and statistics are identical on all ranks
    xs[(size_t)r].assign(x_inout, x_inout + nvar);  // (rank 0 only writes x_inout at the very end, after the last
    ba_lm_stats tmp;                                //  collective, which every rank reaches after this copy)
    return ba_lm_solve(s, xs[(size_t)r].data(), p, &tmp, nullptr, nullptr);
  });
}

int group_apply(ba_handle* h, const std::function<int(ba_handle*)>& f) {
  for (ba_handle* s : h->group->subs) {
    const int rc = f(s);
    if (rc) {
      h->err = s->err;
      return rc;
    }
  }
  return BA_OK;
}

}  // namespace ba

extern "C" int ba_create_multi(int64_t ncams, int64_t npnts, int64_t nobs, const int64_t* cam, const int64_t* pnt,
                               const double* pt2d, int ngpus, const int* devices, ba_handle** out) {
  if (!out) return BA_ERR_ARG;
  *out = nullptr;
  int avail = 0;
  if (cudaGetDeviceCount(&avail) != cudaSuccess) return BA_ERR_CUDA;
  if (ngpus <= 0) ngpus = avail;  // "all"
  if (ngpus < 1 || ngpus > 16 || (!devices && ngpus > avail)) return BA_ERR_ARG;
  ba_handle* h = new (std::nothrow) ba_handle();
  if (!h) return BA_ERR_ARG;
  *out = h;
  h->ncams = ncams; h->npnts = npnts; h->nobs = nobs;
  h->obs0 = 0; h->obs1 = nobs; h->pnt0 = 0; h->pnt1 = npnts;
  h->device = devices ? devices[0] : 0;
  ba_group* G = new ba_group();
  h->group = G;
  G->subs.assign((size_t)ngpus, nullptr);
  G->rcs.assign((size_t)ngpus, 0);
  G->seen.assign((size_t)ngpus, 0);
  for (int r = 0; r < ngpus; ++r) {
    const int rc = ba::create_impl(ncams, npnts, nobs, cam, pnt, pt2d, devices ? devices[r] : r, r, ngpus,
                                   &G->subs[(size_t)r]);
    if (rc) {
      h->err = G->subs[(size_t)r] ? G->subs[(size_t)r]->err : "ba_create_sharded failed";
      return rc;  // the caller destroys the handle (ba_destroy), which frees what exists
    }
  }
  h->sorted = G->subs[0]->sorted;
  for (int r = 0; r < ngpus; ++r) G->workers.emplace_back(ba::worker_main, G, r);
  if (ngpus == 1) return BA_OK;
  uint8_t id[128];
  if (ba_comm_unique_id(id) != BA_OK) {
    h->err = "NCCL is not available (libnccl.so.2)";
    return BA_ERR_COMM;
  }
  int rc = ba::group_run(h, [&](ba_handle* s, int) { return ba_comm_init(s, id); });  // concurrent: the ranks rendezvous
  if (rc) return rc;
  // peer-memory mailboxes for the per-PCG-iteration exchange; NCCL stays the fallback when peer access is missing
  bool p2p = getenv("BAGPU_NO_P2P") == nullptr;
  for (int a = 0; a < ngpus && p2p; ++a)
    for (int b = 0; b < ngpus && p2p; ++b)
      if (a != b) {
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, G->subs[(size_t)a]->device, G->subs[(size_t)b]->device) != cudaSuccess || !can)
          p2p = false;
      }
  if (p2p) {
    rc = ba::group_run(h, [&](ba_handle* s, int) { return ba::p2p_alloc_block(s); });
    if (rc) return rc;
    std::vector<void*> blocks((size_t)ngpus);
    for (int r = 0; r < ngpus; ++r) blocks[(size_t)r] = G->subs[(size_t)r]->p2p.block;
    rc = ba::group_run(h, [&](ba_handle* s, int) {
      for (ba_handle* o : G->subs)
        if (o != s) {
          const cudaError_t e = cudaDeviceEnablePeerAccess(o->device, 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
            s->err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e);
            return (int)BA_ERR_CUDA;
          }
          cudaGetLastError();
        }
      return ba::p2p_attach(s, blocks.data());
    });
    if (rc) return rc;
  }
  return BA_OK;
}

namespace ba {
const ba_handle* group_first(const ba_handle* h) { return h->group ? h->group->subs[0] : h; }
}  // namespace ba
