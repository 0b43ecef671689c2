// ba_capi.cu -- the extern "C" boundary of libbagpu.so (see include/bagpu.h for the contract).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>
#include "ba_internal.h"
#include "ba_ritz.h"

namespace {

template <class T>
int dev_alloc(ba_handle* h, T** p, size_t n) {
  if (*p) return BA_OK;
  BA_CUDA(cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(n, 1) * sizeof(T)));
  return BA_OK;
}

int fail(ba_handle* h, int code, const char* msg) {
  if (h) h->err = msg;
  return code;
}

// uploads x (host) and refreshes the camera records
int stage_x(ba_handle* h, const double* x) {
  int rc = dev_alloc(h, &h->d_x, (size_t)h->nvar());
  if (rc) return rc;
  if (h->nranks > 1) {
    // observation-sharded: the per-observation kernels of this rank touch its own points and the cameras only
    const size_t p0 = 3 * (size_t)h->pnt0, np = 3 * (size_t)h->npnts_l(), c0 = 3 * (size_t)h->npnts;
    if (np) BA_CUDA(cudaMemcpyAsync(h->d_x + p0, x + p0, sizeof(double) * np, cudaMemcpyHostToDevice, h->stream));
    BA_CUDA(cudaMemcpyAsync(h->d_x + c0, x + c0, sizeof(double) * 9 * (size_t)h->ncams, cudaMemcpyHostToDevice, h->stream));
    return BA_OK;
  }
  BA_CUDA(cudaMemcpyAsync(h->d_x, x, sizeof(double) * (size_t)h->nvar(), cudaMemcpyHostToDevice, h->stream));
  return BA_OK;
}

int refresh_cams(ba_handle* h, const double* x_dev) {
  ba::launch_cam_precompute(x_dev, h->npnts, h->ncams, h->d_camtab, h->stream);
  BA_CUDA(cudaGetLastError());
  return BA_OK;
}

}  // namespace

namespace ba {
int create_impl(int64_t ncams, int64_t npnts, int64_t nobs, const int64_t* cam, const int64_t* pnt,
                const double* pt2d, int device, int rank, int nranks, ba_handle** out) {
  if (!out) return BA_ERR_ARG;
  *out = nullptr;
  if (ncams < 0 || npnts < 0 || nobs < 0 || (nobs > 0 && (!cam || !pnt || !pt2d)) || nranks < 1 || rank < 0 ||
      rank >= nranks || 9 * ncams + 3 * npnts >= (int64_t)1 << 31 || nobs >= (int64_t)1 << 31)
    return BA_ERR_ARG;
  ba_handle* h = new (std::nothrow) ba_handle();
  if (!h) return BA_ERR_ARG;
  *out = h;  // returned even on failure so that ba_last_error can be read; caller destroys it
  h->device = device;
  h->ncams = ncams; h->npnts = npnts; h->nobs = nobs; h->rank = rank; h->nranks = nranks;
  if (const char* e = getenv("BAGPU_COARSE")) h->coarse_clusters = atoi(e);
  if (const char* e = getenv("BAGPU_DEFLATE")) h->deflate = std::max(0, std::min(32, atoi(e)));
  if (const char* e = getenv("BAGPU_SOLVER"))
    h->solver = !strcmp(e, "pcg") ? BA_SOLVER_PCG : (!strcmp(e, "exact") ? BA_SOLVER_EXACT :
                (!strcmp(e, "mixed") ? BA_SOLVER_MIXED : BA_SOLVER_AUTO));
  if (const char* e = getenv("BAGPU_MIXED_MAX_CG")) h->mixed_max_cg = std::max(1, std::min(64, atoi(e)));
  if (const char* e = getenv("BAGPU_EXACT_REFINE")) h->exact_refine = std::max(0, std::min(8, atoi(e)));
  bool sorted = true;
  for (int64_t k = 0; k < nobs; ++k) {
    if (cam[k] < 1 || cam[k] > ncams || pnt[k] < 1 || pnt[k] > npnts)
      return fail(h, BA_ERR_ARG, "camera/point index out of range (indices are 1-based)");
    if (k && pnt[k] < pnt[k - 1]) sorted = false;
  }
  h->sorted = sorted;
  h->obs0 = 0; h->obs1 = nobs; h->pnt0 = 0; h->pnt1 = npnts;
  if (nranks > 1) {
    std::vector<int64_t> cuts((size_t)nranks + 1);
    int rc = ba_partition_observations(nobs, pnt, nranks, cuts.data());
    if (rc) return fail(h, rc, "observation sharding needs point-major order");
    h->obs0 = cuts[rank]; h->obs1 = cuts[rank + 1];
    // owned points: those whose observations fall into the range (points without any observation
    // are attached to the rank that owns the preceding point; rank 0 takes the leading ones)
    h->pnt0 = (rank == 0) ? 0 : (h->obs0 < nobs ? pnt[h->obs0] - 1 : npnts);
    h->pnt1 = (rank == nranks - 1) ? npnts : (h->obs1 < nobs ? pnt[h->obs1] - 1 : npnts);
  }
  const int64_t nl = h->nobs_l();
  h->h_cam.resize((size_t)nl);
  h->h_pnt.resize((size_t)nl);
  for (int64_t k = 0; k < nl; ++k) {
    h->h_cam[(size_t)k] = (int32_t)(cam[h->obs0 + k] - 1);
    h->h_pnt[(size_t)k] = (int32_t)(pnt[h->obs0 + k] - 1);
  }
  BA_CUDA(cudaSetDevice(device));
  BA_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  h->own_stream = true;
  BA_CUDA(cudaEventCreate(&h->ev_eval0));
  BA_CUDA(cudaEventCreate(&h->ev_eval1));
  int rc;
  if ((rc = dev_alloc(h, &h->d_cam, (size_t)nl))) return rc;
  if ((rc = dev_alloc(h, &h->d_pnt, (size_t)nl))) return rc;
  if ((rc = dev_alloc(h, &h->d_pt2d, (size_t)nl))) return rc;
  if ((rc = dev_alloc(h, &h->d_camtab, (size_t)ncams * 16))) return rc;
  if (nl) {
    BA_CUDA(cudaMemcpyAsync(h->d_cam, h->h_cam.data(), sizeof(int32_t) * (size_t)nl, cudaMemcpyHostToDevice, h->stream));
    BA_CUDA(cudaMemcpyAsync(h->d_pnt, h->h_pnt.data(), sizeof(int32_t) * (size_t)nl, cudaMemcpyHostToDevice, h->stream));
    BA_CUDA(cudaMemcpyAsync(h->d_pt2d, pt2d + 2 * h->obs0, sizeof(double) * 2 * (size_t)nl, cudaMemcpyHostToDevice, h->stream));
  }
  BA_CUDA(cudaStreamSynchronize(h->stream));
  return BA_OK;
}
}  // namespace ba

namespace {
using ba::create_impl;
// FP64 FMA throughput probe: 8 independent chains per thread, operands from kernel arguments so that
// nothing folds at compile time
__global__ void __launch_bounds__(256) k_fp64_fma(double* out, int iters, double a, double b) {
  double x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = (double)(threadIdx.x + j) * 1e-3;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = fma(x[j], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += x[j];
  out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = s;
}
// FP64 tensor-core probe: 8 independent m8n8k4 accumulators per warp, operands in registers
__global__ void __launch_bounds__(256) k_fp64_mma(double* out, int iters, double a, double b) {
  double c[8][2];
#pragma unroll
  for (int j = 0; j < 8; ++j) c[j][0] = c[j][1] = (double)(threadIdx.x + j) * 1e-3;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                   : "+d"(c[j][0]), "+d"(c[j][1])
                   : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
  out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = s;
}
// legacy tensor path probes (mma.sync): TF32 m16n8k8 and BF16 m16n8k16, FP32 accumulate, 8 independent accumulators
template <int KIND>
__global__ void __launch_bounds__(256) k_mma_probe(float* out, int iters, unsigned a0, unsigned b0) {
  float c[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) c[j][i] = (float)(threadIdx.x + j + i) * 1e-3f;
  unsigned a[4] = {a0, a0 + 1, a0 + 2, a0 + 3}, b[2] = {b0, b0 + 1};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = s;
}
// FP32 FMA probe
__global__ void __launch_bounds__(256) k_fp32_fma(float* out, int iters, float a, float b) {
  float x[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) x[j] = (float)(threadIdx.x + j) * 1e-3f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] = fmaf(x[j], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) s += x[j];
  out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = s;
}
}  // namespace

extern "C" {

// development probe: kind 0 = TF32 mma.sync m16n8k8, 1 = BF16 mma.sync m16n8k16, 2 = FP32 FMA; TFLOP/s
int ba_dbg_probe_peak(int device, int kind, double* tflops) {
  if (!tflops || kind < 0 || kind > 2) return BA_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return BA_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return BA_ERR_CUDA;
  const int blocks = prop.multiProcessorCount * 4, iters = 1 << 14;
  float* out = nullptr;
  cudaEvent_t e0, e1;
  if (cudaMalloc(reinterpret_cast<void**>(&out), sizeof(float) * 256 * (size_t)blocks) != cudaSuccess) return BA_ERR_CUDA;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 0.f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, 0);
    if (kind == 0) k_mma_probe<0><<<blocks, 256>>>(out, iters, 0x3f800000u, 0x3f000000u);
    else if (kind == 1) k_mma_probe<1><<<blocks, 256>>>(out, iters, 0x3f803f80u, 0x3f003f00u);
    else k_fp32_fma<<<blocks, 256>>>(out, iters, 0.999999f, 1e-9f);
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && (best == 0.f || ms < best)) best = ms;
  }
  const cudaError_t err = cudaGetLastError();
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  if (err != cudaSuccess || !(best > 0.f)) return BA_ERR_CUDA;
  const double per_warp_iter = kind == 0 ? 8.0 * 2 * 16 * 8 * 8 : (kind == 1 ? 8.0 * 2 * 16 * 8 * 16 : 16.0 * 2 * 32);
  *tflops = per_warp_iter * iters * 8.0 * blocks / (best * 1e-3) / 1e12;
  return BA_OK;
}

int ba_measure_fp64_mma_peak(int device, double* tflops) {
  if (!tflops) return BA_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return BA_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return BA_ERR_CUDA;
  const int blocks = prop.multiProcessorCount * 4, iters = 1 << 14;
  double* out = nullptr;
  cudaEvent_t e0, e1;
  if (cudaMalloc(reinterpret_cast<void**>(&out), sizeof(double) * 256 * (size_t)blocks) != cudaSuccess) return BA_ERR_CUDA;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 0.f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, 0);
    k_fp64_mma<<<blocks, 256>>>(out, iters, 1e-3, 1e-3);
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && (best == 0.f || ms < best)) best = ms;
  }
  const cudaError_t err = cudaGetLastError();
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  if (err != cudaSuccess || !(best > 0.f)) return BA_ERR_CUDA;
  // 8 MMAs of 8 x 8 x 4 (512 flops) per warp and iteration
  *tflops = 512.0 * 8.0 * iters * 8.0 * blocks / (best * 1e-3) / 1e12;
  return BA_OK;
}

int ba_measure_fp64_peak(int device, double* tflops) {
  if (!tflops) return BA_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return BA_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return BA_ERR_CUDA;
  const int blocks = prop.multiProcessorCount * 8, iters = 1 << 16;
  double* out = nullptr;
  cudaEvent_t e0, e1;
  if (cudaMalloc(reinterpret_cast<void**>(&out), sizeof(double) * 256 * (size_t)blocks) != cudaSuccess) return BA_ERR_CUDA;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 0.f;
  for (int rep = 0; rep < 4; ++rep) {  // first repetition warms up; keep the fastest of the rest
    cudaEventRecord(e0, 0);
    k_fp64_fma<<<blocks, 256>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && (best == 0.f || ms < best)) best = ms;
  }
  const cudaError_t err = cudaGetLastError();
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  if (err != cudaSuccess || !(best > 0.f)) return BA_ERR_CUDA;
  *tflops = 2.0 * 8.0 * iters * 256.0 * blocks / (best * 1e-3) / 1e12;
  return BA_OK;
}

const char* ba_version(void) { return "bagpu 0.1 (sm_100a, fp64)"; }

const char* ba_last_error(const ba_handle* h) { return h ? h->err.c_str() : "null handle"; }

int ba_partition_observations(int64_t nobs, const int64_t* pnt, int nranks, int64_t* cuts) {
  if (nranks < 1 || !cuts || (nobs > 0 && !pnt)) return BA_ERR_ARG;
  for (int64_t k = 1; k < nobs; ++k)
    if (pnt[k] < pnt[k - 1]) return BA_ERR_UNSORTED;
  cuts[0] = 0;
  for (int r = 1; r < nranks; ++r) {
    int64_t c = (nobs * r) / nranks;  // balanced by observation count ...
    if (c < cuts[r - 1]) c = cuts[r - 1];
    while (c > 0 && c < nobs && pnt[c] == pnt[c - 1]) ++c;  // ... moved up to the next point boundary
    cuts[r] = c;
  }
  cuts[nranks] = nobs;
  return BA_OK;
}

int ba_create(int64_t ncams, int64_t npnts, int64_t nobs, const int64_t* cam, const int64_t* pnt,
              const double* pt2d, int device, ba_handle** out) {
  return create_impl(ncams, npnts, nobs, cam, pnt, pt2d, device, 0, 1, out);
}

int ba_create_sharded(int64_t ncams, int64_t npnts, int64_t nobs, const int64_t* cam, const int64_t* pnt,
                      const double* pt2d, int device, int rank, int nranks, ba_handle** out) {
  return create_impl(ncams, npnts, nobs, cam, pnt, pt2d, device, rank, nranks, out);
}

int ba_destroy(ba_handle* h) {
  if (!h) return BA_OK;
  if (h->group) {
    ba::group_release(h);
    delete h;
    return BA_OK;
  }
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  ba::lm_release(h);
  ba::comm_release(h);
  cudaFree(h->d_cam); cudaFree(h->d_pnt); cudaFree(h->d_pt2d); cudaFree(h->d_x); cudaFree(h->d_camtab);
  cudaFree(h->d_cx); cudaFree(h->d_vals); cudaFree(h->d_v); cudaFree(h->d_w); cudaFree(h->d_rows);
  cudaFree(h->d_cols);
  if (h->ev_eval0) cudaEventDestroy(h->ev_eval0);
  if (h->ev_eval1) cudaEventDestroy(h->ev_eval1);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return BA_OK;
}

int ba_shard_range(const ba_handle* h, int64_t* obs0, int64_t* obs1, int64_t* pnt0, int64_t* pnt1) {
  if (!h) return BA_ERR_ARG;
  if (obs0) *obs0 = h->obs0;
  if (obs1) *obs1 = h->obs1;
  if (pnt0) *pnt0 = h->pnt0;
  if (pnt1) *pnt1 = h->pnt1;
  return BA_OK;
}

int ba_set_stream(ba_handle* h, void* s) {
  if (!h) return BA_ERR_ARG;
  if (h->group) return fail(h, BA_ERR_ARG, "not available on a multi-GPU handle (ba_create_multi)");
  BA_CUDA(cudaSetDevice(h->device));
  if (h->stream) BA_CUDA(cudaStreamSynchronize(h->stream));
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  h->stream = reinterpret_cast<cudaStream_t>(s);
  h->own_stream = false;
  if (h->lm.pcg_graph) cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(h->lm.pcg_graph));
  h->lm.pcg_graph = nullptr;  // captured on the old stream's work; re-captured on first use
  h->lm.pcg_graph_off = false;
  return BA_OK;
}

int ba_alloc_pinned(uint64_t bytes, void** out) {
  if (!out) return BA_ERR_ARG;
  return cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault) == cudaSuccess ? BA_OK : BA_ERR_CUDA;
}

int ba_free_pinned(void* p) { return cudaFreeHost(p) == cudaSuccess ? BA_OK : BA_ERR_CUDA; }

int ba_sync(ba_handle* h) {
  if (!h) return BA_ERR_ARG;
  if (h->group) return ba::group_apply(h, [](ba_handle* s) { return ba_sync(s); });
  BA_CUDA(cudaSetDevice(h->device));
  BA_CUDA(cudaStreamSynchronize(h->stream));
  return BA_OK;
}

int ba_set_coarse_clusters(ba_handle* h, int n) {
  if (!h || n < 0) return BA_ERR_ARG;
  if (h->group) return ba::group_apply(h, [n](ba_handle* s) { return ba_set_coarse_clusters(s, n); });
  if (n != h->coarse_clusters) {
    BA_CUDA(cudaSetDevice(h->device));
    if (h->stream) BA_CUDA(cudaStreamSynchronize(h->stream));
    ba::lm_release(h);  // schedules and the captured PCG graph depend on it; rebuilt on next use
    h->coarse_clusters = n;
  }
  return BA_OK;
}

int ba_set_deflation(ba_handle* h, int k) {
  if (!h || k < 0 || k > 32) return BA_ERR_ARG;
  if (h->group) return ba::group_apply(h, [k](ba_handle* s) { return ba_set_deflation(s, k); });
  if (k != h->deflate) {
    BA_CUDA(cudaSetDevice(h->device));
    if (h->stream) BA_CUDA(cudaStreamSynchronize(h->stream));
    ba::lm_release(h);  // buffers, harvested vectors and the captured PCG graph depend on it
    h->deflate = k;
  }
  return BA_OK;
}

int ba_set_solver(ba_handle* h, int solver) {
  if (!h || solver < BA_SOLVER_AUTO || solver > BA_SOLVER_MIXED) return fail(h, BA_ERR_ARG, "unknown solver");
  if (h->group) {
    h->solver = solver;
    return ba::group_apply(h, [solver](ba_handle* s) { return ba_set_solver(s, solver); });
  }
  if (solver != h->solver) {
    BA_CUDA(cudaSetDevice(h->device));
    if (h->stream) BA_CUDA(cudaStreamSynchronize(h->stream));
    ba::lm_release(h);  // the dense matrix / PCG buffers are allocated per mode; rebuilt on next use
    h->solver = solver;
  }
  return BA_OK;
}

int ba_last_solve_info(const ba_handle* h, int32_t* solver, int32_t* converged, double* rel, int32_t* iters) {
  if (!h) return BA_ERR_ARG;
  h = ba::group_first(h);
  if (solver) *solver = h->lm.last_solver;
  if (converged) *converged = h->lm.last_converged;
  if (rel) *rel = h->lm.last_rel;
  if (iters) *iters = h->lm.last_iters;
  return BA_OK;
}

int ba_set_profiling(ba_handle* h, int on) {
  if (!h) return BA_ERR_ARG;
  if (h->group) return fail(h, BA_ERR_ARG, "not available on a multi-GPU handle (ba_create_multi)");
  h->profile = on != 0;
  return BA_OK;
}

int ba_last_eval_ms(ba_handle* h, float* ms) {
  if (!h || !ms) return BA_ERR_ARG;
  if (!h->profile) {
    h->err = "ba_last_eval_ms: call ba_set_profiling(h, 1) before the evaluation";
    return BA_ERR_ARG;
  }
  BA_CUDA(cudaSetDevice(h->device));
  BA_CUDA(cudaEventSynchronize(h->ev_eval1));
  BA_CUDA(cudaEventElapsedTime(ms, h->ev_eval0, h->ev_eval1));
  return BA_OK;
}

// ---- device-pointer variants ------------------------------------------------------------------
int ba_residual_dev(ba_handle* h, const double* x, double* cx) {
  if (!h || !x || !cx) return fail(h, BA_ERR_ARG, "null argument");
  if (h->group) return fail(h, BA_ERR_ARG, "device-pointer calls are per GPU: not available on a multi-GPU handle");
  BA_CUDA(cudaSetDevice(h->device));
  int rc = refresh_cams(h, x);
  if (rc) return rc;
  ba::launch_eval(h, x, h->d_camtab, cx, nullptr, h->stream);
  BA_CUDA(cudaGetLastError());
  return BA_OK;
}

int ba_jac_coord_dev(ba_handle* h, const double* x, double* vals) {
  if (!h || !x || !vals) return fail(h, BA_ERR_ARG, "null argument");
  if (h->group) return fail(h, BA_ERR_ARG, "device-pointer calls are per GPU: not available on a multi-GPU handle");
  BA_CUDA(cudaSetDevice(h->device));
  int rc = refresh_cams(h, x);
  if (rc) return rc;
  ba::launch_eval(h, x, h->d_camtab, nullptr, vals, h->stream);
  BA_CUDA(cudaGetLastError());
  return BA_OK;
}

int ba_residual_jac_dev(ba_handle* h, const double* x, double* cx, double* vals) {
  if (!h || !x || !cx || !vals) return fail(h, BA_ERR_ARG, "null argument");
  if (h->group) return fail(h, BA_ERR_ARG, "device-pointer calls are per GPU: not available on a multi-GPU handle");
  BA_CUDA(cudaSetDevice(h->device));
  int rc = refresh_cams(h, x);
  if (rc) return rc;
  ba::launch_eval(h, x, h->d_camtab, cx, vals, h->stream);
  BA_CUDA(cudaGetLastError());
  return BA_OK;
}

int ba_jac_structure_dev(ba_handle* h, int64_t* rows, int64_t* cols) {
  if (!h || !rows || !cols) return fail(h, BA_ERR_ARG, "null argument");
  if (h->group) return fail(h, BA_ERR_ARG, "device-pointer calls are per GPU: not available on a multi-GPU handle");
  BA_CUDA(cudaSetDevice(h->device));
  ba::launch_jac_structure(h, rows, cols, h->stream);
  BA_CUDA(cudaGetLastError());
  return BA_OK;
}

int ba_jprod_dev(ba_handle* h, const double* x, const double* v, double* Jv) {
  if (!h || !x || !v || !Jv) return fail(h, BA_ERR_ARG, "null argument");
  if (h->group) return fail(h, BA_ERR_ARG, "device-pointer calls are per GPU: not available on a multi-GPU handle");
  BA_CUDA(cudaSetDevice(h->device));
  int rc = refresh_cams(h, x);
  if (rc) return rc;
  ba::launch_jprod(h, x, h->d_camtab, v, Jv, h->stream);
  BA_CUDA(cudaGetLastError());
  return BA_OK;
}

int ba_jtprod_dev(ba_handle* h, const double* x, const double* v, double* Jtv) {
  if (!h || !x || !v || !Jtv) return fail(h, BA_ERR_ARG, "null argument");
  if (h->group) return fail(h, BA_ERR_ARG, "device-pointer calls are per GPU: not available on a multi-GPU handle");
  BA_CUDA(cudaSetDevice(h->device));
  int rc = refresh_cams(h, x);
  if (rc) return rc;
  // point-major problems: camera sums by the ordered camera-major pass (deterministic, no atomics on hot
  // cameras); otherwise L2 atomics
  ba::launch_jtprod(h, x, h->d_camtab, v, Jtv, !h->sorted, h->stream);
  BA_CUDA(cudaGetLastError());
  if (h->sorted && (rc = ba::lm_jtprod_cams(h, x, v, Jtv + 3 * h->npnts))) return rc;
  if (h->nranks > 1 && h->comm) return ba::allreduce_sum(h, Jtv, (size_t)h->nvar());
  return BA_OK;
}

// ---- host-pointer calls (the ones Julia's ccall binds) ----------------------------------------
int ba_residual(ba_handle* h, const double* x, double* cx) {
  if (!h || !x || !cx) return fail(h, BA_ERR_ARG, "null argument");
  if (h->group) return ba::group_residual(h, x, cx, nullptr);
  BA_CUDA(cudaSetDevice(h->device));
  int rc;
  if ((rc = stage_x(h, x))) return rc;
  if ((rc = dev_alloc(h, &h->d_cx, 2 * (size_t)h->nobs_l()))) return rc;
  if ((rc = ba_residual_dev(h, h->d_x, h->d_cx))) return rc;
  return ba::copy_to_host(h, cx, h->d_cx, sizeof(double) * 2 * (size_t)h->nobs_l());
}

int ba_jac_coord(ba_handle* h, const double* x, double* vals) {
  if (!h || !x || !vals) return fail(h, BA_ERR_ARG, "null argument");
  if (h->group) return ba::group_residual(h, x, nullptr, vals);
  BA_CUDA(cudaSetDevice(h->device));
  int rc;
  if ((rc = stage_x(h, x))) return rc;
  if ((rc = dev_alloc(h, &h->d_vals, 24 * (size_t)h->nobs_l()))) return rc;
  if ((rc = ba_jac_coord_dev(h, h->d_x, h->d_vals))) return rc;
  return ba::copy_to_host(h, vals, h->d_vals, sizeof(double) * 24 * (size_t)h->nobs_l());
}

int ba_residual_jac(ba_handle* h, const double* x, double* cx, double* vals) {
  if (!h || !x || !cx || !vals) return fail(h, BA_ERR_ARG, "null argument");
  if (h->group) return ba::group_residual(h, x, cx, vals);
  BA_CUDA(cudaSetDevice(h->device));
  int rc;
  if ((rc = stage_x(h, x))) return rc;
  if ((rc = dev_alloc(h, &h->d_cx, 2 * (size_t)h->nobs_l()))) return rc;
  if ((rc = dev_alloc(h, &h->d_vals, 24 * (size_t)h->nobs_l()))) return rc;
  if ((rc = ba_residual_jac_dev(h, h->d_x, h->d_cx, h->d_vals))) return rc;
  if ((rc = ba::copy_to_host(h, cx, h->d_cx, sizeof(double) * 2 * (size_t)h->nobs_l()))) return rc;
  return ba::copy_to_host(h, vals, h->d_vals, sizeof(double) * 24 * (size_t)h->nobs_l());
}

int ba_jac_structure(ba_handle* h, int64_t* rows, int64_t* cols) {
  if (!h || !rows || !cols) return fail(h, BA_ERR_ARG, "null argument");
  if (h->group) return ba::group_jac_structure(h, rows, cols);
  BA_CUDA(cudaSetDevice(h->device));
  int rc;
  const size_t n = 24 * (size_t)h->nobs_l();
  if ((rc = dev_alloc(h, &h->d_rows, n))) return rc;
  if ((rc = dev_alloc(h, &h->d_cols, n))) return rc;
  if ((rc = ba_jac_structure_dev(h, h->d_rows, h->d_cols))) return rc;
  if ((rc = ba::copy_to_host(h, rows, h->d_rows, sizeof(int64_t) * n))) return rc;
  if ((rc = ba::copy_to_host(h, cols, h->d_cols, sizeof(int64_t) * n))) return rc;
  // the structure is needed once per solve (src/lm.jl:53): do not keep 2 x 192 B/obs resident
  cudaFree(h->d_rows); cudaFree(h->d_cols);
  h->d_rows = h->d_cols = nullptr;
  return BA_OK;
}

int ba_jprod(ba_handle* h, const double* x, const double* v, double* Jv) {
  if (!h || !x || !v || !Jv) return fail(h, BA_ERR_ARG, "null argument");
  if (h->group) return ba::group_jprod(h, x, v, Jv);
  BA_CUDA(cudaSetDevice(h->device));
  int rc;
  if ((rc = stage_x(h, x))) return rc;
  if ((rc = dev_alloc(h, &h->d_v, (size_t)h->nvar()))) return rc;
  if ((rc = dev_alloc(h, &h->d_w, std::max((size_t)h->nvar(), 2 * (size_t)h->nobs_l())))) return rc;
  BA_CUDA(cudaMemcpyAsync(h->d_v, v, sizeof(double) * (size_t)h->nvar(), cudaMemcpyHostToDevice, h->stream));
  if ((rc = ba_jprod_dev(h, h->d_x, h->d_v, h->d_w))) return rc;
  BA_CUDA(cudaMemcpyAsync(Jv, h->d_w, sizeof(double) * 2 * (size_t)h->nobs_l(), cudaMemcpyDeviceToHost, h->stream));
  BA_CUDA(cudaStreamSynchronize(h->stream));
  return BA_OK;
}

int ba_jtprod(ba_handle* h, const double* x, const double* v, double* Jtv) {
  if (!h || !x || !v || !Jtv) return fail(h, BA_ERR_ARG, "null argument");
  if (h->group) return ba::group_jtprod(h, x, v, Jtv);
  BA_CUDA(cudaSetDevice(h->device));
  int rc;
  if ((rc = stage_x(h, x))) return rc;
  if ((rc = dev_alloc(h, &h->d_v, (size_t)h->nvar()))) return rc;
  if ((rc = dev_alloc(h, &h->d_w, std::max((size_t)h->nvar(), 2 * (size_t)h->nobs_l())))) return rc;
  BA_CUDA(cudaMemcpyAsync(h->d_w, v, sizeof(double) * 2 * (size_t)h->nobs_l(), cudaMemcpyHostToDevice, h->stream));
  if ((rc = ba_jtprod_dev(h, h->d_x, h->d_w, h->d_v))) return rc;
  BA_CUDA(cudaMemcpyAsync(Jtv, h->d_v, sizeof(double) * (size_t)h->nvar(), cudaMemcpyDeviceToHost, h->stream));
  BA_CUDA(cudaStreamSynchronize(h->stream));
  return BA_OK;
}

// ---- host-only numerical helpers of the PCG deflation space, exported for CPU unit tests --------------
int ba_dbg_tridiag_eig(const double* alpha, const double* beta, int32_t m, double* evals, double* evecs_colmajor) {
  if (!alpha || !beta || m < 1 || !evals || !evecs_colmajor) return BA_ERR_ARG;
  std::vector<double> d, e, V;
  ba::lanczos_tridiagonal(alpha, beta, m, d, e);
  if (!ba::tridiag_eig(d, e, m, V)) return BA_ERR_NUMERIC;
  std::copy(d.begin(), d.end(), evals);
  std::copy(V.begin(), V.end(), evecs_colmajor);
  return BA_OK;
}

int ba_dbg_tridiag_smallest(const double* alpha, const double* beta, int32_t m, int32_t k, int32_t full_below,
                            double* evals, double* evecs_colmajor) {
  if (!alpha || !beta || m < 1 || k < 1 || k > m || !evals || !evecs_colmajor) return BA_ERR_ARG;
  std::vector<double> d, e, w, V;
  ba::lanczos_tridiagonal(alpha, beta, m, d, e);
  if (!ba::tridiag_smallest(d, e, m, k, w, V, full_below)) return BA_ERR_NUMERIC;
  std::copy(w.begin(), w.end(), evals);
  std::copy(V.begin(), V.end(), evecs_colmajor);
  return BA_OK;
}

int ba_dbg_select_columns(const double* gram_rowmajor, int32_t n, int32_t k, double tol, double* coeff_rowmajor,
                          int32_t* kept) {
  if (!gram_rowmajor || n < 1 || k < 1 || !coeff_rowmajor || !kept) return BA_ERR_ARG;
  std::vector<double> G(gram_rowmajor, gram_rowmajor + (size_t)n * n), Cm;
  *kept = ba::select_orthonormal(G, n, k, tol, Cm);
  std::copy(Cm.begin(), Cm.begin() + (size_t)n * (*kept), coeff_rowmajor);
  return BA_OK;
}

}  // extern "C"
