// ba_eval.cu -- per-observation operator kernels (K1-K4) for sm_100a, FP64.
//
//   K1 k_cam_precompute : camera parameters -> 128-byte records (once per x)
//   K2 k_eval           : cons! and/or jac_coord!   (src/BALNLPModels.jl:115-122, :161-206)
//   K3 k_jac_structure  : jac_structure!             (src/BALNLPModels.jl:125-158)
//   K4 k_jprod/k_jtprod : J v and J' v, matrix-free  (semantics of src/lma_aux.jl:194-212)
//
// All of these are HBM-bound streaming kernels: one thread per observation, indices and pt2d
// read with coalesced streaming loads, point/camera parameters gathered through L1/L2 (they are
// reused across observations), outputs written with full-sector coalesced streaming stores.  The
// 24 Jacobian values of one observation are contiguous in the reference layout (192 B), i.e.
// strided across the lanes of a warp, so each warp transposes its 32x24 tile through a padded
// shared-memory buffer and writes 6144 contiguous bytes with 16-byte stores.  The same buffer first
// stages the warp's 32 camera records, fetched cooperatively as whole 128-byte lines.
#include "ba_internal.h"
#include "ba_math.cuh"

namespace ba {

#ifndef EVAL_MINBLOCKS
#define EVAL_MINBLOCKS 5  // resident blocks per SM the register allocation targets (A/B-tested)
#endif
constexpr int EVAL_THREADS = 128;           // 4 warps
#ifndef EVAL_SWIZZLE
#define EVAL_SWIZZLE 0  // 0: rows padded to 13 words; 1: dense rows + XOR swizzle (conflict-free but slower: DESIGN.md)
#endif
constexpr int STAGE_ROW = EVAL_SWIZZLE ? 12 : 13;  // double2 per staged observation

__global__ void __launch_bounds__(128) k_cam_precompute(const double* __restrict__ xcam, int64_t ncams,
                                                        double* __restrict__ camtab) {
  // programmatic dependent launch: the evaluation kernel that follows may start its prologue (index and
  // point loads) now; it waits (griddepcontrol.wait) before touching the camera records written here
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= ncams) return;
  double c9[9], rec[CAM_REC];
#pragma unroll
  for (int i = 0; i < 9; ++i) c9[i] = xcam[c * 9 + i];
  cam_precompute(c9, rec);
  double2* out = reinterpret_cast<double2*>(camtab + c * CAM_REC);
#pragma unroll
  for (int i = 0; i < CAM_REC / 2; ++i) out[i] = make_double2(rec[2 * i], rec[2 * i + 1]);
}

__device__ __forceinline__ void load_cam(const double* __restrict__ camtab, int c, double* cam) {
  const double2* src = reinterpret_cast<const double2*>(camtab + (int64_t)c * CAM_REC);
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const double2 t = __ldg(src + i);
    cam[2 * i] = t.x;
    cam[2 * i + 1] = t.y;
  }
}

template <bool WCX, bool WVALS>
__global__ void __launch_bounds__(EVAL_THREADS, EVAL_MINBLOCKS)
k_eval(const int32_t* __restrict__ cam_idx, const int32_t* __restrict__ pnt_idx,
       const double2* __restrict__ pt2d, const double* __restrict__ xpts,
       const double* camtab, double* __restrict__ cx, double* __restrict__ vals,
       int64_t nobs) {
  constexpr int WROW = WVALS ? 32 * STAGE_ROW : 32 * CAM_ROW2;  // double2 per warp
  __shared__ double2 stage[(EVAL_THREADS / 32) * WROW];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t wbase = (blockIdx.x * (int64_t)(EVAL_THREADS / 32) + warp) * 32;
  if (wbase >= nobs) return;
  const int64_t k = wbase + lane;
  const bool valid = k < nobs;
  double2* st = stage + warp * WROW;
  int c = 0, p = 0;
  double2 ob = make_double2(0.0, 0.0);
  if (valid) {
    c = __ldcs(cam_idx + k);
    p = __ldcs(pnt_idx + k);
    ob = __ldcs(pt2d + k);
  }
  double X[3], cam[14];
  const double* xp = xpts + (int64_t)p * 3;
  X[0] = __ldg(xp);
  X[1] = __ldg(xp + 1);
  X[2] = __ldg(xp + 2);
  asm volatile("griddepcontrol.wait;" ::: "memory");  // camera records of k_cam_precompute (no-op without PDL)
  warp_stage_cams<true>(camtab, c, lane, st);
  read_staged_cam(st, lane, cam);
  ObsBlock o;
  if (WVALS) {
    eval_block(X, cam, ob.x, ob.y, o);
  } else {
    eval_residual(X, cam, ob.x, ob.y, o.F);
  }
  if (WCX && valid) __stcs(reinterpret_cast<double2*>(cx) + k, make_double2(o.F[0], o.F[1]));
  if (WVALS) {
    __syncwarp();  // every lane has its camera record in registers: the buffer can be reused
    // Transpose through shared memory with dense 192-byte rows and an XOR swizzle of the 16-byte column
    // index by ((row >> 1) & 3): the per-lane row writes (row stride 12 words) and the transposed reads
    // (32 consecutive words per instruction) are both bank-conflict free, because the swizzle only permutes
    // words inside aligned groups of four.
    double2* row = st + lane * STAGE_ROW;
    const int sw = EVAL_SWIZZLE ? (lane >> 1) & 3 : 0;
    // reference order: row 1 = [A(3) B(9)], row 2 likewise; per-entry NaN -> 0
    row[0 ^ sw] = make_double2(nan0(o.A[0]), nan0(o.A[1]));
    row[1 ^ sw] = make_double2(nan0(o.A[2]), nan0(o.B[0]));
    row[2 ^ sw] = make_double2(nan0(o.B[1]), nan0(o.B[2]));
    row[3 ^ sw] = make_double2(nan0(o.B[3]), nan0(o.B[4]));
    row[4 ^ sw] = make_double2(nan0(o.B[5]), nan0(o.B[6]));
    row[5 ^ sw] = make_double2(nan0(o.B[7]), nan0(o.B[8]));
    row[6 ^ sw] = make_double2(nan0(o.A[3]), nan0(o.A[4]));
    row[7 ^ sw] = make_double2(nan0(o.A[5]), nan0(o.B[9]));
    row[8 ^ sw] = make_double2(nan0(o.B[10]), nan0(o.B[11]));
    row[9 ^ sw] = make_double2(nan0(o.B[12]), nan0(o.B[13]));
    row[10 ^ sw] = make_double2(nan0(o.B[14]), nan0(o.B[15]));
    row[11 ^ sw] = make_double2(nan0(o.B[16]), nan0(o.B[17]));
    __syncwarp();
    const int nval = (int)min((int64_t)32, nobs - wbase);
    double2* dst = reinterpret_cast<double2*>(vals) + wbase * 12;
#pragma unroll
    for (int m = 0; m < 12; ++m) {
      const int q = lane + 32 * m;
      const int r = q / 12, cidx = q - 12 * r;
      if (r < nval) __stcs(dst + q, st[r * STAGE_ROW + (EVAL_SWIZZLE ? cidx ^ ((r >> 1) & 3) : cidx)]);
    }
  }
}

// K3: 12 threads per observation, each writes one (rows, cols) pair of 16 bytes.
__global__ void __launch_bounds__(256)
k_jac_structure(const int32_t* __restrict__ cam_idx, const int32_t* __restrict__ pnt_idx, int64_t nobs,
                int64_t obs0, int64_t npnts, longlong2* __restrict__ rows, longlong2* __restrict__ cols) {
  const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (q >= nobs * 12) return;
  const int64_t k = q / 12;
  const int j = (int)(q - 12 * k);  // pair index inside the observation: entries 2j, 2j+1
  const int64_t c = __ldg(cam_idx + k), p = __ldg(pnt_idx + k);
  const int64_t row = 2 * (obs0 + k) + 1 + (j >= 6 ? 1 : 0);  // 1-based: 2k-1 then 2k
  const int64_t ip = 3 * p, ic = 3 * npnts + 9 * c;
  const int m0 = (2 * j) % 12, m1 = m0 + 1;
  longlong2 cv;
  cv.x = (m0 < 3) ? ip + m0 + 1 : ic + (m0 - 3) + 1;
  cv.y = (m1 < 3) ? ip + m1 + 1 : ic + (m1 - 3) + 1;
  __stcs(rows + q, make_longlong2(row, row));
  __stcs(cols + q, cv);
}

__global__ void __launch_bounds__(128)
k_jprod(const int32_t* __restrict__ cam_idx, const int32_t* __restrict__ pnt_idx,
        const double2* __restrict__ pt2d, const double* __restrict__ x, int64_t npnts,
        const double* __restrict__ camtab, const double* __restrict__ v, double* __restrict__ Jv,
        int64_t nobs) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= nobs) return;
  const int c = __ldcs(cam_idx + k), p = __ldcs(pnt_idx + k);
  const double2 ob = __ldcs(pt2d + k);
  double X[3], cam[14];
  const double* xp = x + (int64_t)p * 3;
  X[0] = __ldg(xp); X[1] = __ldg(xp + 1); X[2] = __ldg(xp + 2);
  load_cam(camtab, c, cam);
  ObsBlock o;
  eval_block(X, cam, ob.x, ob.y, o);
  const double* vp = v + (int64_t)p * 3;
  const double* vc = v + 3 * npnts + (int64_t)c * 9;
  double s0 = 0.0, s1 = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double t = __ldg(vp + i);
    s0 += nan0(o.A[i]) * t;
    s1 += nan0(o.A[3 + i]) * t;
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const double t = __ldg(vc + i);
    s0 += nan0(o.B[i]) * t;
    s1 += nan0(o.B[9 + i]) * t;
  }
  __stcs(reinterpret_cast<double2*>(Jv) + k, make_double2(s0, s1));
}

// J'v: point-side sums are reduced inside the warp over runs of equal point id (observations
// are point-major in BAL order) before one atomic per run.  Camera-side sums: CAMS = true sends them to L2
// atomics (any observation order); CAMS = false leaves them to the ordered camera-major pass of ba_lm.cu
// (point-major problems: deterministic, and no contention on cameras with 10^4 observations).
template <bool CAMS>
__global__ void __launch_bounds__(128)
k_jtprod(const int32_t* __restrict__ cam_idx, const int32_t* __restrict__ pnt_idx,
         const double2* __restrict__ pt2d, const double* __restrict__ x, int64_t npnts,
         const double* __restrict__ camtab, const double2* __restrict__ v, double* __restrict__ Jtv,
         int64_t nobs) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool valid = k < nobs;
  int c = 0, p = -1;
  double gp[3] = {0.0, 0.0, 0.0};
  if (valid) {
    c = __ldcs(cam_idx + k);
    p = __ldcs(pnt_idx + k);
    const double2 ob = __ldcs(pt2d + k);
    const double2 w = __ldcs(v + k);
    double X[3], cam[14];
    const double* xp = x + (int64_t)p * 3;
    X[0] = __ldg(xp); X[1] = __ldg(xp + 1); X[2] = __ldg(xp + 2);
    load_cam(camtab, c, cam);
    ObsBlock o;
    eval_block(X, cam, ob.x, ob.y, o);
#pragma unroll
    for (int i = 0; i < 3; ++i) gp[i] = nan0(o.A[i]) * w.x + nan0(o.A[3 + i]) * w.y;
    if (CAMS) {
      double* jc = Jtv + 3 * npnts + (int64_t)c * 9;
#pragma unroll
      for (int i = 0; i < 9; ++i) atomicAdd(jc + i, nan0(o.B[i]) * w.x + nan0(o.B[9 + i]) * w.y);
    }
  }
  // segmented inclusive scan over runs of equal p
  const int pprev = __shfl_up_sync(0xffffffffu, p, 1);
  const bool head = (lane == 0) || (p != pprev);
  const unsigned hm = __ballot_sync(0xffffffffu, head);
  const int seg0 = 31 - __clz(hm & (0xffffffffu >> (31 - lane)));
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double t = shfl_up_d(gp[i], d);
      if (lane - d >= seg0) gp[i] += t;
    }
  }
  const bool tail = (lane == 31) || ((hm >> (lane + 1)) & 1u);  // next lane starts a new run
  if (valid && tail) {
    double* jp = Jtv + (int64_t)p * 3;
    atomicAdd(jp, gp[0]);
    atomicAdd(jp + 1, gp[1]);
    atomicAdd(jp + 2, gp[2]);
  }
}

void launch_cam_precompute(const double* x, int64_t npnts, int64_t ncams, double* camtab, cudaStream_t s) {
  const int threads = 128;
  const int blocks = (int)((ncams + threads - 1) / threads);
  k_cam_precompute<<<blocks, threads, 0, s>>>(x + 3 * npnts, ncams, camtab);
}

void launch_eval(const ba_handle* h, const double* x, const double* camtab, double* cx, double* vals,
                 cudaStream_t s) {
  const int64_t n = h->nobs_l();
  if (n == 0) return;
  const int per_block = EVAL_THREADS;
  const unsigned blocks = (unsigned)((n + per_block - 1) / per_block);
  // Launched as a programmatic dependent of k_cam_precompute (PDL): its blocks are scheduled and run their
  // prologue while the 4-us precompute kernel drains.  Timing events would sit between the two kernels and
  // break that edge, so they are recorded only when profiling is on (ba_set_profiling).
  if (h->profile && h->ev_eval0) cudaEventRecord(h->ev_eval0, s);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(EVAL_THREADS);
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = h->profile ? 0 : 1;
  if (cx && vals)
    cudaLaunchKernelEx(&cfg, k_eval<true, true>, (const int32_t*)h->d_cam, (const int32_t*)h->d_pnt,
                       (const double2*)h->d_pt2d, x, camtab, cx, vals, n);
  else if (vals)
    cudaLaunchKernelEx(&cfg, k_eval<false, true>, (const int32_t*)h->d_cam, (const int32_t*)h->d_pnt,
                       (const double2*)h->d_pt2d, x, camtab, cx, vals, n);
  else
    cudaLaunchKernelEx(&cfg, k_eval<true, false>, (const int32_t*)h->d_cam, (const int32_t*)h->d_pnt,
                       (const double2*)h->d_pt2d, x, camtab, cx, vals, n);
  if (h->profile && h->ev_eval1) cudaEventRecord(h->ev_eval1, s);
}

void launch_jac_structure(const ba_handle* h, int64_t* rows, int64_t* cols, cudaStream_t s) {
  const int64_t n = h->nobs_l();
  if (n == 0) return;
  const int threads = 256;
  const unsigned blocks = (unsigned)((n * 12 + threads - 1) / threads);
  k_jac_structure<<<blocks, threads, 0, s>>>(h->d_cam, h->d_pnt, n, h->obs0, h->npnts,
                                            reinterpret_cast<longlong2*>(rows),
                                            reinterpret_cast<longlong2*>(cols));
}

void launch_jprod(const ba_handle* h, const double* x, const double* camtab, const double* v, double* Jv,
                  cudaStream_t s) {
  const int64_t n = h->nobs_l();
  if (n == 0) return;
  const unsigned blocks = (unsigned)((n + 127) / 128);
  k_jprod<<<blocks, 128, 0, s>>>(h->d_cam, h->d_pnt, h->d_pt2d, x, h->npnts, camtab, v, Jv, n);
}

void launch_jtprod(const ba_handle* h, const double* x, const double* camtab, const double* v, double* Jtv,
                   bool with_cameras, cudaStream_t s) {
  const int64_t n = h->nobs_l();
  cudaMemsetAsync(Jtv, 0, sizeof(double) * (size_t)h->nvar(), s);
  if (n == 0) return;
  const unsigned blocks = (unsigned)((n + 127) / 128);
  if (with_cameras)
    k_jtprod<true><<<blocks, 128, 0, s>>>(h->d_cam, h->d_pnt, h->d_pt2d, x, h->npnts, camtab,
                                         reinterpret_cast<const double2*>(v), Jtv, n);
  else
    k_jtprod<false><<<blocks, 128, 0, s>>>(h->d_cam, h->d_pnt, h->d_pt2d, x, h->npnts, camtab,
                                          reinterpret_cast<const double2*>(v), Jtv, n);
}

}  // namespace ba
