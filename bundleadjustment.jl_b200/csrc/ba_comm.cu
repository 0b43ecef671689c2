// ba_comm.cu -- NCCL plumbing for the observation-sharded (one process per GPU) mode.
//
// The reference has no communication layer at all; this is the only exchange
// the sharded hot path needs: sum-allreduce of camera-sized buffers and a handful of scalars over
// NVLink 5 / NVSwitch.  NCCL is bound lazily with dlopen so that libbagpu.so loads (and every
// single-GPU entry point works) in processes that never touch NCCL; inside a torch process the
// already-loaded libnccl.so.2 is reused.
#include <dlfcn.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "ba_internal.h"

namespace {

// The few NCCL symbols used, declared locally (stable since NCCL 2.0) so that no NCCL header is
// needed at build time.
typedef struct { char internal[128]; } nccl_uid;
typedef int (*fn_get_uid)(nccl_uid*);
typedef int (*fn_comm_init_rank)(ncclComm**, int, nccl_uid, int);
typedef int (*fn_comm_destroy)(ncclComm*);
typedef int (*fn_allreduce)(const void*, void*, size_t, int /*dtype*/, int /*op*/, ncclComm*, cudaStream_t);
typedef int (*fn_allgather)(const void*, void*, size_t, int /*dtype*/, ncclComm*, cudaStream_t);
typedef const char* (*fn_errstr)(int);
constexpr int NCCL_FLOAT64 = 8, NCCL_SUM = 0;

struct nccl_api {
  void* so = nullptr;
  fn_get_uid get_uid = nullptr;
  fn_comm_init_rank init_rank = nullptr;
  fn_comm_destroy destroy = nullptr;
  fn_allreduce allreduce = nullptr;
  fn_allgather allgather = nullptr;
  fn_errstr errstr = nullptr;
  bool ok = false;
};

nccl_api& api() {
  static nccl_api a;
  if (a.ok || a.so) return a;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    a.so = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (a.so) break;
  }
  if (!a.so) return a;
  a.get_uid = (fn_get_uid)dlsym(a.so, "ncclGetUniqueId");
  a.init_rank = (fn_comm_init_rank)dlsym(a.so, "ncclCommInitRank");
  a.destroy = (fn_comm_destroy)dlsym(a.so, "ncclCommDestroy");
  a.allreduce = (fn_allreduce)dlsym(a.so, "ncclAllReduce");
  a.allgather = (fn_allgather)dlsym(a.so, "ncclAllGather");
  a.errstr = (fn_errstr)dlsym(a.so, "ncclGetErrorString");
  a.ok = a.get_uid && a.init_rank && a.destroy && a.allreduce;
  return a;
}

}  // namespace

namespace ba {

int allreduce_sum(ba_handle* h, double* buf, size_t n) {
  if (h->nranks == 1 || n == 0) return BA_OK;
  if (!h->comm) {
    h->err = "sharded handle used before ba_comm_init";
    return BA_ERR_COMM;
  }
  const int rc = api().allreduce(buf, buf, n, NCCL_FLOAT64, NCCL_SUM, h->comm, h->stream);
  if (rc != 0) {
    h->err = std::string("ncclAllReduce: ") + (api().errstr ? api().errstr(rc) : "error");
    return BA_ERR_COMM;
  }
  return BA_OK;
}

// exact (order-independent) sum of 64-bit integers over the ranks
int allreduce_sum_i64(ba_handle* h, long long* buf, size_t n) {
  if (h->nranks == 1 || n == 0) return BA_OK;
  if (!h->comm) {
    h->err = "sharded handle used before ba_comm_init";
    return BA_ERR_COMM;
  }
  const int rc = api().allreduce(buf, buf, n, 4 /* ncclInt64 */, NCCL_SUM, h->comm, h->stream);
  if (rc != 0) {
    h->err = std::string("ncclAllReduce: ") + (api().errstr ? api().errstr(rc) : "error");
    return BA_ERR_COMM;
  }
  return BA_OK;
}

// host-side all-gather of `bytes` per rank (small control records: IPC handles, pointers); synchronises the stream
int allgather_host(ba_handle* h, const void* send, void* recv, size_t bytes) {
  if (h->nranks == 1) {
    memcpy(recv, send, bytes);
    return BA_OK;
  }
  if (!h->comm || !api().allgather) {
    h->err = "sharded handle used before ba_comm_init";
    return BA_ERR_COMM;
  }
  char* d = nullptr;
  BA_CUDA(cudaMalloc(reinterpret_cast<void**>(&d), bytes * (size_t)(h->nranks + 1)));
  BA_CUDA(cudaMemcpyAsync(d, send, bytes, cudaMemcpyHostToDevice, h->stream));
  const int rc = api().allgather(d, d + bytes, bytes, 0 /* ncclInt8 */, h->comm, h->stream);
  if (rc != 0) {
    cudaFree(d);
    h->err = std::string("ncclAllGather: ") + (api().errstr ? api().errstr(rc) : "error");
    return BA_ERR_COMM;
  }
  BA_CUDA(cudaMemcpyAsync(recv, d + bytes, bytes * (size_t)h->nranks, cudaMemcpyDeviceToHost, h->stream));
  BA_CUDA(cudaStreamSynchronize(h->stream));
  cudaFree(d);
  return BA_OK;
}

// own mailbox block: [flags | pad | mail[2][9 ncams]]
int p2p_alloc_block(ba_handle* h) {
  ba_p2p_state& P = h->p2p;
  if (h->nranks > 16) {
    h->err = "peer-memory exchange supports at most 16 ranks";
    return BA_ERR_ARG;
  }
  BA_CUDA(cudaSetDevice(h->device));
  if (!P.block) {
    const size_t bytes = P.mail_off + 2 * 9 * (size_t)h->ncams * sizeof(double);
    BA_CUDA(cudaMalloc(&P.block, bytes));
    BA_CUDA(cudaMemset(P.block, 0, bytes));
  }
  return BA_OK;
}

// blocks[r] = device pointer to rank r's block, usable from this rank's device (IPC-mapped, or a raw peer
// pointer inside one process)
int p2p_attach(ba_handle* h, void* const* blocks) {
  ba_p2p_state& P = h->p2p;
  BA_CUDA(cudaSetDevice(h->device));
  P.nranks = h->nranks;
  P.rank = h->rank;
  std::vector<double*> mail((size_t)h->nranks);
  std::vector<unsigned long long*> flags((size_t)h->nranks);
  for (int r = 0; r < h->nranks; ++r) {
    P.peer_block[r] = blocks[r];
    flags[(size_t)r] = reinterpret_cast<unsigned long long*>(P.peer_block[r]);
    mail[(size_t)r] = reinterpret_cast<double*>(static_cast<char*>(P.peer_block[r]) + P.mail_off);
  }
  if (!P.d_mail) BA_CUDA(cudaMalloc(reinterpret_cast<void**>(&P.d_mail), sizeof(double*) * 16));
  if (!P.d_flags) BA_CUDA(cudaMalloc(reinterpret_cast<void**>(&P.d_flags), sizeof(unsigned long long*) * 16));
  if (!P.d_seq) BA_CUDA(cudaMalloc(reinterpret_cast<void**>(&P.d_seq), sizeof(unsigned long long)));
  BA_CUDA(cudaMemcpy(P.d_mail, mail.data(), sizeof(double*) * (size_t)h->nranks, cudaMemcpyHostToDevice));
  BA_CUDA(cudaMemcpy(P.d_flags, flags.data(), sizeof(unsigned long long*) * (size_t)h->nranks, cudaMemcpyHostToDevice));
  BA_CUDA(cudaMemset(P.d_seq, 0, sizeof(unsigned long long)));
  static const bool off = getenv("BAGPU_NO_P2P") != nullptr;
  P.ready = !off;
  // a captured PCG graph holds the NCCL variant: drop it
  if (h->lm.pcg_graph) cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(h->lm.pcg_graph));
  h->lm.pcg_graph = nullptr;
  return BA_OK;
}

void comm_release(ba_handle* h) {
  if (h->comm && api().ok) api().destroy(h->comm);
  h->comm = nullptr;
  ba_p2p_state& P = h->p2p;
  for (int r = 0; r < P.nranks; ++r)
    if (P.ipc && P.peer_block[r] && P.peer_block[r] != P.block) cudaIpcCloseMemHandle(P.peer_block[r]);
  cudaFree(P.d_mail);
  cudaFree(P.d_flags);
  cudaFree(P.d_seq);
  cudaFree(P.block);
  P = ba_p2p_state();
}

}  // namespace ba

extern "C" {

int ba_comm_unique_id(uint8_t id128[128]) {
  if (!id128 || !api().ok) return BA_ERR_COMM;
  nccl_uid u;
  if (api().get_uid(&u) != 0) return BA_ERR_COMM;
  memcpy(id128, u.internal, 128);
  return BA_OK;
}

int ba_comm_init(ba_handle* h, const uint8_t id128[128]) {
  if (!h || !id128) return BA_ERR_ARG;
  if (h->nranks == 1) return BA_OK;
  if (!api().ok) {
    h->err = "libnccl.so.2 not found";
    return BA_ERR_COMM;
  }
  BA_CUDA(cudaSetDevice(h->device));
  nccl_uid u;
  memcpy(u.internal, id128, 128);
  const int rc = api().init_rank(&h->comm, h->nranks, u, h->rank);
  if (rc != 0) {
    h->err = std::string("ncclCommInitRank: ") + (api().errstr ? api().errstr(rc) : "error");
    h->comm = nullptr;
    return BA_ERR_COMM;
  }
  // NCCL sets up channels / protocols lazily inside the first collective of each size class (tens to hundreds
  // of ms): pay that here, at the message sizes the solver uses, not in the first LM iteration
  // (scalars, one camera vector, the four vectors of the coarse setup, the per-camera accumulators)
  const size_t sizes[4] = {8, 9 * (size_t)h->ncams, 36 * (size_t)h->ncams, 54 * (size_t)h->ncams};
  double* warm = nullptr;
  BA_CUDA(cudaMalloc(reinterpret_cast<void**>(&warm), std::max<size_t>(sizes[3], 8) * sizeof(double)));
  BA_CUDA(cudaMemsetAsync(warm, 0, std::max<size_t>(sizes[3], 8) * sizeof(double), h->stream));
  int wrc = BA_OK;
  for (int rep = 0; rep < 2 && !wrc; ++rep)
    for (size_t n : sizes)
      if (n && !wrc) wrc = ba::allreduce_sum(h, warm, n);
  BA_CUDA(cudaStreamSynchronize(h->stream));
  cudaFree(warm);
  if (wrc) return wrc;
  // the exact solve's workspace and the peer mappings of its distributed factorisation belong to the communicator
  return ba::lm_exact_workspace(h);
}

int ba_comm_ipc_export(ba_handle* h, uint8_t handle64[64]) {
  if (!h || !handle64) return BA_ERR_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  int rc = ba::p2p_alloc_block(h);
  if (rc) return rc;
  ba_p2p_state& P = h->p2p;
  cudaIpcMemHandle_t mh;
  BA_CUDA(cudaIpcGetMemHandle(&mh, P.block));
  memcpy(handle64, &mh, 64);
  return BA_OK;
}

int ba_comm_ipc_import(ba_handle* h, const uint8_t* handles) {
  if (!h || !handles) return BA_ERR_ARG;
  ba_p2p_state& P = h->p2p;
  if (!P.block) {
    h->err = "ba_comm_ipc_import before ba_comm_ipc_export";
    return BA_ERR_ARG;
  }
  BA_CUDA(cudaSetDevice(h->device));
  std::vector<void*> blocks((size_t)h->nranks, nullptr);
  for (int r = 0; r < h->nranks; ++r) {
    if (r == h->rank) {
      blocks[(size_t)r] = P.block;
    } else {
      cudaIpcMemHandle_t mh;
      memcpy(&mh, handles + 64 * (size_t)r, 64);
      BA_CUDA(cudaIpcOpenMemHandle(&blocks[(size_t)r], mh, cudaIpcMemLazyEnablePeerAccess));
    }
  }
  P.ipc = true;
  return ba::p2p_attach(h, blocks.data());
}

// Back to NCCL for the per-iteration exchange (all ranks must agree: call it on every rank or on none).
int ba_comm_ipc_disable(ba_handle* h) {
  if (!h) return BA_ERR_ARG;
  h->p2p.ready = false;
  if (h->lm.pcg_graph) cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(h->lm.pcg_graph));
  h->lm.pcg_graph = nullptr;
  return BA_OK;
}

}  // extern "C"
