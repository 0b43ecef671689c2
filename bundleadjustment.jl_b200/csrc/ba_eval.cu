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
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <cuda.h>  // CUtensorMap and the encoder's signature only: the entry point is resolved at run time
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

// TMA variant of the Jacobian store (TMA = true): the 32 x 24 values of a full warp are exactly 6144 contiguous
// bytes of `vals`.  Seen as a 2-D tensor of 128-byte rows (16 doubles), they are a 48-row box; every lane drops
// its 12 16-byte pieces into the warp's shared-memory tile at the positions the hardware 128-byte swizzle expects
// (piece c of row R at R * 128 + ((c ^ (R & 7)) << 4): at most 2-way bank conflicts instead of the 4-way ones of a
// dense tile), and ONE bulk tensor store per warp (cp.async.bulk.tensor, UTMASTG in SASS) moves the tile: no
// transposed shared-memory reads and no per-lane global stores.  The last, partial warp keeps the generic path.
template <bool WCX, bool WVALS, bool TMA>
__global__ void __launch_bounds__(EVAL_THREADS, EVAL_MINBLOCKS)
k_eval(const int32_t* __restrict__ cam_idx, const int32_t* __restrict__ pnt_idx,
       const double2* __restrict__ pt2d, const double* __restrict__ xpts,
       const double* camtab, double* __restrict__ cx, double* __restrict__ vals,
       int64_t nobs, const __grid_constant__ CUtensorMap tmap) {
  constexpr int WROW = TMA ? 384 : (WVALS ? 32 * STAGE_ROW : 32 * CAM_ROW2);  // double2 per warp (TMA: 6144 B)
  __shared__ __align__(1024) double2 stage[(EVAL_THREADS / 32) * WROW];
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
    // reference order: row 1 = [A(3) B(9)], row 2 likewise; per-entry NaN -> 0
    double2 v[12];
    v[0] = make_double2(nan0(o.A[0]), nan0(o.A[1]));
    v[1] = make_double2(nan0(o.A[2]), nan0(o.B[0]));
    v[2] = make_double2(nan0(o.B[1]), nan0(o.B[2]));
    v[3] = make_double2(nan0(o.B[3]), nan0(o.B[4]));
    v[4] = make_double2(nan0(o.B[5]), nan0(o.B[6]));
    v[5] = make_double2(nan0(o.B[7]), nan0(o.B[8]));
    v[6] = make_double2(nan0(o.A[3]), nan0(o.A[4]));
    v[7] = make_double2(nan0(o.A[5]), nan0(o.B[9]));
    v[8] = make_double2(nan0(o.B[10]), nan0(o.B[11]));
    v[9] = make_double2(nan0(o.B[12]), nan0(o.B[13]));
    v[10] = make_double2(nan0(o.B[14]), nan0(o.B[15]));
    v[11] = make_double2(nan0(o.B[16]), nan0(o.B[17]));
    const int nval = (int)min((int64_t)32, nobs - wbase);
    if (TMA && nval == 32) {
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        const int q = 12 * lane + j, R = q >> 3, cc = q & 7;
        st[R * 8 + (cc ^ (R & 7))] = v[j];
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the TMA unit
      __syncwarp();
      if (lane == 0) {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(st);
        const int row0 = (int)(wbase / 32) * 48;  // 128-byte row of vals where this warp's 6144 bytes start
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&tmap),
                     "r"(0), "r"(row0), "r"(sa)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the tile has been read: the CTA may retire
      }
      return;
    }
    // generic path: transpose through shared memory (rows padded to 13 words, or dense rows with an XOR swizzle of
    // the 16-byte column index by ((row >> 1) & 3)) and write 6144 contiguous bytes with 16-byte stores
    constexpr int SROW = TMA ? 12 : STAGE_ROW;
    double2* row = st + lane * SROW;
    const int sw = (!TMA && EVAL_SWIZZLE) ? (lane >> 1) & 3 : 0;
#pragma unroll
    for (int j = 0; j < 12; ++j) row[j ^ sw] = v[j];
    __syncwarp();
    double2* dst = reinterpret_cast<double2*>(vals) + wbase * 12;
#pragma unroll
    for (int m = 0; m < 12; ++m) {
      const int q = lane + 32 * m;
      const int r = q / 12, cidx = q - 12 * r;
      if (r < nval) __stcs(dst + q, st[r * SROW + ((!TMA && EVAL_SWIZZLE) ? cidx ^ ((r >> 1) & 3) : cidx)]);
    }
  }
}

// Persistent variant: every warp walks over tiles of 32 observations (grid-stride) and loads the indices and the
// observed pixel of its NEXT tile before it evaluates the current one, so that the DRAM round trip of the index
// stream (the head of the dependent chain index -> gathers -> ~250 FP64 instructions -> stores) overlaps the
// arithmetic instead of being paid once per tile at 28 % occupancy.
template <bool WCX, bool WVALS>
__global__ void __launch_bounds__(EVAL_THREADS, EVAL_MINBLOCKS)
k_eval_persist(const int32_t* __restrict__ cam_idx, const int32_t* __restrict__ pnt_idx,
               const double2* __restrict__ pt2d, const double* __restrict__ xpts, const double* camtab,
               double* __restrict__ cx, double* __restrict__ vals, int64_t nobs) {
  constexpr int WROW = WVALS ? 32 * STAGE_ROW : 32 * CAM_ROW2;  // double2 per warp
  __shared__ double2 stage[(EVAL_THREADS / 32) * WROW];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t ntiles = (nobs + 31) / 32, stride = (int64_t)gridDim.x * (EVAL_THREADS / 32);
  int64_t tile = blockIdx.x * (int64_t)(EVAL_THREADS / 32) + warp;
  if (tile >= ntiles) return;
  double2* st = stage + warp * WROW;
  int cn = 0, pn = 0;
  double2 obn = make_double2(0.0, 0.0);
  {
    const int64_t k = tile * 32 + lane;
    if (k < nobs) {
      cn = __ldcs(cam_idx + k);
      pn = __ldcs(pnt_idx + k);
      obn = __ldcs(pt2d + k);
    }
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");  // camera records of k_cam_precompute (no-op without PDL)
  for (; tile < ntiles; tile += stride) {
    const int64_t wbase = tile * 32, k = wbase + lane;
    const bool valid = k < nobs;
    const int c = cn, p = pn;
    const double2 ob = obn;
    {
      const int64_t k2 = (tile + stride) * 32 + lane;  // next tile of this warp
      cn = 0; pn = 0;
      if (k2 < nobs) {
        cn = __ldcs(cam_idx + k2);
        pn = __ldcs(pnt_idx + k2);
        obn = __ldcs(pt2d + k2);
      }
    }
    double X[3], cam[14];
    const double* xp = xpts + (int64_t)p * 3;
    X[0] = __ldg(xp);
    X[1] = __ldg(xp + 1);
    X[2] = __ldg(xp + 2);
    warp_stage_cams<true>(camtab, c, lane, st);
    read_staged_cam(st, lane, cam);
    ObsBlock o;
    if (WVALS) {
      eval_block(X, cam, ob.x, ob.y, o);
    } else {
      eval_residual(X, cam, ob.x, ob.y, o.F);
    }
    if (WCX && valid) __stcs(reinterpret_cast<double2*>(cx) + k, make_double2(o.F[0], o.F[1]));
    __syncwarp();  // every lane has its camera record in registers: the buffer can be reused
    if (WVALS) {
      double2* row = st + lane * STAGE_ROW;
      row[0] = make_double2(nan0(o.A[0]), nan0(o.A[1]));
      row[1] = make_double2(nan0(o.A[2]), nan0(o.B[0]));
      row[2] = make_double2(nan0(o.B[1]), nan0(o.B[2]));
      row[3] = make_double2(nan0(o.B[3]), nan0(o.B[4]));
      row[4] = make_double2(nan0(o.B[5]), nan0(o.B[6]));
      row[5] = make_double2(nan0(o.B[7]), nan0(o.B[8]));
      row[6] = make_double2(nan0(o.A[3]), nan0(o.A[4]));
      row[7] = make_double2(nan0(o.A[5]), nan0(o.B[9]));
      row[8] = make_double2(nan0(o.B[10]), nan0(o.B[11]));
      row[9] = make_double2(nan0(o.B[12]), nan0(o.B[13]));
      row[10] = make_double2(nan0(o.B[14]), nan0(o.B[15]));
      row[11] = make_double2(nan0(o.B[16]), nan0(o.B[17]));
      __syncwarp();
      const int nval = (int)min((int64_t)32, nobs - wbase);
      double2* dst = reinterpret_cast<double2*>(vals) + wbase * 12;
#pragma unroll
      for (int m = 0; m < 12; ++m) {
        const int q = lane + 32 * m;
        const int r = q / 12, cidx = q - 12 * r;
        if (r < nval) __stcs(dst + q, st[r * STAGE_ROW + cidx]);
      }
      __syncwarp();  // the tile has been read before the next camera staging overwrites it
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

namespace {
typedef CUresult (*fn_tmap_encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// vals (24 nobs doubles) as a 2-D tensor of 128-byte rows; box = 48 rows (one warp's 32 observations), 128-byte
// swizzle.  Only whole rows are described: the rows of every FULL warp lie inside (a partial last warp never
// uses the map).  false: no encoder / unaligned pointer / switched off -> generic stores.
bool eval_tmap(double* vals, int64_t nobs, CUtensorMap* out) {
  // measured slower than the generic stores (profiles/r02_k_eval_ab.md): opt-in with BAGPU_EVAL_TMA=1
  static const bool off = !(getenv("BAGPU_EVAL_TMA") && atoi(getenv("BAGPU_EVAL_TMA")) != 0);
  static fn_tmap_encode enc = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    cudaGetLastError();
    return reinterpret_cast<fn_tmap_encode>(f);
  }();
  const int64_t rows = (nobs / 32) * 48;  // rows covered by full warps
  if (off || !enc || rows == 0 || (reinterpret_cast<uintptr_t>(vals) & 15)) return false;
  const cuuint64_t gdim[2] = {16, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {128};
  const cuuint32_t box[2] = {16, 48}, estride[2] = {1, 1};
  return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, vals, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

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
  // bulk tensor store of the Jacobian tile (one per warp) when the driver provides the encoder and vals is 16-byte
  // aligned; opt-in: BAGPU_EVAL_TMA=1 (A/B)
  CUtensorMap tm;
  memset(&tm, 0, sizeof tm);
  bool tma = false;
  if (vals) tma = eval_tmap(vals, n, &tm);
#define BA_EVAL_LAUNCH(WCX, WVALS, TMA)                                                                        \
  cudaLaunchKernelEx(&cfg, k_eval<WCX, WVALS, TMA>, (const int32_t*)h->d_cam, (const int32_t*)h->d_pnt,       \
                     (const double2*)h->d_pt2d, x, camtab, cx, vals, n, tm)
  cudaError_t le;
  static const int persist = getenv("BAGPU_EVAL_PERSIST") ? atoi(getenv("BAGPU_EVAL_PERSIST")) : 0;
  if (persist > 0 && !tma) {  // grid-stride warps with index prefetch; `persist` resident-block multiples per SM
    cudaLaunchConfig_t pc = cfg;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    pc.gridDim = dim3((unsigned)std::min<int64_t>(blocks, (int64_t)sms * EVAL_MINBLOCKS * persist));
#define BA_EVAL_PLAUNCH(WCX, WVALS)                                                                           \
  cudaLaunchKernelEx(&pc, k_eval_persist<WCX, WVALS>, (const int32_t*)h->d_cam, (const int32_t*)h->d_pnt,    \
                     (const double2*)h->d_pt2d, x, camtab, cx, vals, n)
    if (cx && vals) le = BA_EVAL_PLAUNCH(true, true);
    else if (vals) le = BA_EVAL_PLAUNCH(false, true);
    else le = BA_EVAL_PLAUNCH(true, false);
#undef BA_EVAL_PLAUNCH
  } else if (cx && vals) le = tma ? BA_EVAL_LAUNCH(true, true, true) : BA_EVAL_LAUNCH(true, true, false);
  else if (vals) le = tma ? BA_EVAL_LAUNCH(false, true, true) : BA_EVAL_LAUNCH(false, true, false);
  else le = BA_EVAL_LAUNCH(true, false, false);
#undef BA_EVAL_LAUNCH
  if (le != cudaSuccess) h->err = std::string("k_eval launch: ") + cudaGetErrorString(le);  // also left for cudaGetLastError
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
