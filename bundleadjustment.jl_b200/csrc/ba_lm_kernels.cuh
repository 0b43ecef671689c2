// ba_lm_kernels.cuh -- device kernels of the LM path (K5-K8): evaluation/assembly, the two passes of the
// Schur product, PCG vector kernels (with the fused peer-memory exchange), coarse level, reductions.
// Included once, by ba_lm.cu, which holds the host side (schedules, launch sequences, LM loop, C ABI).
// The algorithm and its reference citations are described at the top of ba_lm.cu.
#pragma once
#include "ba_internal.h"
#include "ba_math.cuh"

namespace ba {

#ifndef PS_MINBLOCKS
#define PS_MINBLOCKS 5  // resident blocks per SM targeted by k_point_solve (A/B-tested)
#endif
constexpr int NV = 54;         // per-camera accumulators: 45 (symmetric 9x9) + 9
constexpr int PT_THREADS = 128;
constexpr int RED_THREADS = 1024;

// device scalar slots (doubles)
enum {
  S_F2 = 0,   // sum F^2 at the current iterate                 } local partial sums: allreduced
  S_GP2,      // sum g_p^2                                      } over ranks in sharded mode
  S_DR2,      // sum (J delta + r)^2                            }
  S_DP2,      // sum delta_p^2                                  }
  S_XP2,      // sum (x + delta)_p^2                            }
  S_TR2,      // sum F^2 at the trial iterate                   }
  S_NLOCAL = 8,
  S_GC2 = 8,  // sum g_c^2 (cameras are replicated)
  S_DC2,      // sum delta_c^2
  S_XC2,      // sum (x + delta)_c^2
  S_RZ = 16, S_RZ0, S_PQ, S_DONE, S_ITERS, S_REL, S_ERR, S_RZN, S_RCY,
  S_ITK,      // S_ITERS as it was when the current PCG iteration started (read by every CTA of k_pcg_p)
  S_CBAD,     // the coarse matrix [P Z]'S[P Z] had a non-positive pivot: its inverse was replaced by zero, i.e. the
              // solves run with plain block-Jacobi (not an error: the preconditioner never changes the solution)
  S_MBAD,     // mixed-precision solve: CG met a non-positive curvature (the FP32 factor is no usable preconditioner:
              // the host falls back to the FP64 factorisation; not an error)
  S_MRZ,      // mixed-precision solve: r.z of the current CG iteration (k_pcg_q, part of the product, rewrites S_RZ)
  S_COUNT = 32
};

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------
template <int NT>
__device__ __forceinline__ double block_sum(double v, double* sh) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int i = 0; i < NT / 32; ++i) t += sh[i];
  return t;
}

// out[slot0 + j] = sum_b part[b * K + j], one block per j, fixed summation order
__global__ void __launch_bounds__(RED_THREADS)
k_reduce_parts(const double* __restrict__ part, int64_t nblocks, int K, double* __restrict__ out, int slot0) {
  __shared__ double sh[RED_THREADS / 32];
  const int j = blockIdx.x;
  double s = 0.0;
  for (int64_t b = threadIdx.x; b < nblocks; b += RED_THREADS) s += part[b * K + j];
  s = block_sum<RED_THREADS>(s, sh);
  if (threadIdx.x == 0) out[slot0 + j] = s;
}

// ---------------------------------------------------------------------------------------------
// K5a: evaluate F and the J blocks at x (thread per observation), store them point-major (12 planes of
// (row1,row2) pairs); the camera-major passes recompute the camera part instead of reading a second copy
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PT_THREADS)
k_lm_build(const int32_t* __restrict__ cam_idx, const int32_t* __restrict__ pnt_idx,
           const double2* __restrict__ pt2d, const double* __restrict__ xpts,
           const double* __restrict__ camtab, double2* __restrict__ Jp, double2* __restrict__ F,
           double* __restrict__ part, int64_t nl) {
  __shared__ double2 stage[(PT_THREADS / 32) * 32 * CAM_ROW2];
  __shared__ double sh[PT_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t k = blockIdx.x * (int64_t)PT_THREADS + threadIdx.x;
  const bool valid = k < nl;
  double f2 = 0.0;
  if (k - lane < nl) {  // warps wholly past the end only take part in the block sum
    int c = 0, p = 0;
    double2 ob = make_double2(0.0, 0.0);
    if (valid) {
      c = __ldg(cam_idx + k);
      p = __ldg(pnt_idx + k);
      ob = __ldg(pt2d + k);
    }
    double X[3], cam[14];
    const double* xp = xpts + (int64_t)p * 3;
    X[0] = __ldg(xp);
    X[1] = __ldg(xp + 1);
    X[2] = __ldg(xp + 2);
    double2* st = stage + warp * 32 * CAM_ROW2;
    warp_stage_cams(camtab, c, lane, st);
    read_staged_cam(st, lane, cam);
    ObsBlock o;
    eval_block(X, cam, ob.x, ob.y, o);
    if (valid) {
#pragma unroll
      for (int j = 0; j < 3; ++j) Jp[(int64_t)j * nl + k] = make_double2(nan0(o.A[j]), nan0(o.A[3 + j]));
#pragma unroll
      for (int j = 0; j < 9; ++j) Jp[(int64_t)(3 + j) * nl + k] = make_double2(nan0(o.B[j]), nan0(o.B[9 + j]));
      F[k] = make_double2(o.F[0], o.F[1]);
      f2 = o.F[0] * o.F[0] + o.F[1] * o.F[1];
    }
  }
  f2 = block_sum<PT_THREADS>(f2, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = f2;
}

// trial point: residual norm only
__global__ void __launch_bounds__(PT_THREADS)
k_lm_trial(const int32_t* __restrict__ cam_idx, const int32_t* __restrict__ pnt_idx,
           const double2* __restrict__ pt2d, const double* __restrict__ xpts,
           const double* __restrict__ camtab, double* __restrict__ part, int64_t nl) {
  __shared__ double2 stage[(PT_THREADS / 32) * 32 * CAM_ROW2];
  __shared__ double sh[PT_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t k = blockIdx.x * (int64_t)PT_THREADS + threadIdx.x;
  const bool valid = k < nl;
  double f2 = 0.0;
  if (k - lane < nl) {  // warps wholly past the end only take part in the block sum
    int c = 0, p = 0;
    double2 ob = make_double2(0.0, 0.0);
    if (valid) {
      c = __ldg(cam_idx + k);
      p = __ldg(pnt_idx + k);
      ob = __ldg(pt2d + k);
    }
    double X[3], cam[14], Fv[2];
    const double* xp = xpts + (int64_t)p * 3;
    X[0] = __ldg(xp);
    X[1] = __ldg(xp + 1);
    X[2] = __ldg(xp + 2);
    double2* st = stage + warp * 32 * CAM_ROW2;
    warp_stage_cams(camtab, c, lane, st);
    read_staged_cam(st, lane, cam);
    eval_residual(X, cam, ob.x, ob.y, Fv);
    if (valid) f2 = Fv[0] * Fv[0] + Fv[1] * Fv[1];
  }
  f2 = block_sum<PT_THREADS>(f2, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = f2;
}

// V_p = sum A'A (6, symmetric: 00 01 02 11 12 22), g_p = -sum A'F; thread per local point
__global__ void __launch_bounds__(PT_THREADS)
k_point_assemble(const int32_t* __restrict__ pstart, int64_t npl, int64_t nl, const double2* __restrict__ Jp,
                 const double2* __restrict__ F, double* __restrict__ V, double* __restrict__ gp,
                 double* __restrict__ part) {
  __shared__ double sh[PT_THREADS / 32];
  const int64_t p = blockIdx.x * (int64_t)PT_THREADS + threadIdx.x;
  double g2 = 0.0;
  if (p < npl) {
    double v00 = 0, v01 = 0, v02 = 0, v11 = 0, v12 = 0, v22 = 0, g0 = 0, g1 = 0, g2v = 0;
    const int k1 = pstart[p + 1];
    for (int k = pstart[p]; k < k1; ++k) {
      const double2 a0 = Jp[k], a1 = Jp[nl + k], a2 = Jp[2 * nl + k], f = F[k];
      v00 += a0.x * a0.x + a0.y * a0.y;
      v01 += a0.x * a1.x + a0.y * a1.y;
      v02 += a0.x * a2.x + a0.y * a2.y;
      v11 += a1.x * a1.x + a1.y * a1.y;
      v12 += a1.x * a2.x + a1.y * a2.y;
      v22 += a2.x * a2.x + a2.y * a2.y;
      g0 -= a0.x * f.x + a0.y * f.y;
      g1 -= a1.x * f.x + a1.y * f.y;
      g2v -= a2.x * f.x + a2.y * f.y;
    }
    double* vo = V + p * 6;
    vo[0] = v00; vo[1] = v01; vo[2] = v02; vo[3] = v11; vo[4] = v12; vo[5] = v22;
    gp[p * 3] = g0; gp[p * 3 + 1] = g1; gp[p * 3 + 2] = g2v;
    g2 = (g0 * g0 + g1 * g1) + g2v * g2v;
  }
  g2 = block_sum<PT_THREADS>(g2, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = g2;
}

// (V + lambda I)^-1 (symmetric 3x3 by cofactors) and wp = Vinv gp; thread per local point
__global__ void __launch_bounds__(PT_THREADS)
k_point_inv(int64_t npl, double lambda, const double* __restrict__ V, const double* __restrict__ gp,
            double* __restrict__ Vinv, double* __restrict__ wp) {
  const int64_t p = blockIdx.x * (int64_t)PT_THREADS + threadIdx.x;
  if (p >= npl) return;
  const double* v = V + p * 6;
  const double a = v[0] + lambda, b = v[1], c = v[2], d = v[3] + lambda, e = v[4], f = v[5] + lambda;
  const double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
  const double det = (a * c00 + b * c01) + c * c02;
  const double id = 1.0 / det;
  const double i00 = c00 * id, i01 = c01 * id, i02 = c02 * id;
  const double i11 = (a * f - c * c) * id, i12 = (b * c - a * e) * id, i22 = (a * d - b * b) * id;
  double* o = Vinv + p * 6;
  o[0] = i00; o[1] = i01; o[2] = i02; o[3] = i11; o[4] = i12; o[5] = i22;
  const double g0 = gp[p * 3], g1 = gp[p * 3 + 1], g2 = gp[p * 3 + 2];
  wp[p * 3] = (i00 * g0 + i01 * g1) + i02 * g2;
  wp[p * 3 + 1] = (i01 * g0 + i11 * g1) + i12 * g2;
  wp[p * 3 + 2] = (i02 * g0 + i12 * g1) + i22 * g2;
}

// per observation: w_k = A_k wp_p (right-hand side) and T_k = A_k Vinv_p A_k' (preconditioner)
__global__ void __launch_bounds__(PT_THREADS)
k_point_prep(const int32_t* __restrict__ pnt_idx, int64_t pnt0, int64_t nl, const double2* __restrict__ Jp,
             const double* __restrict__ Vinv, const double* __restrict__ wp, double2* __restrict__ w,
             double* __restrict__ T) {
  const int64_t k = blockIdx.x * (int64_t)PT_THREADS + threadIdx.x;
  if (k >= nl) return;
  const int64_t p = __ldg(pnt_idx + k) - pnt0;
  const double2 a0 = Jp[k], a1 = Jp[nl + k], a2 = Jp[2 * nl + k];
  const double* vi = Vinv + p * 6;
  const double i00 = vi[0], i01 = vi[1], i02 = vi[2], i11 = vi[3], i12 = vi[4], i22 = vi[5];
  const double w0 = wp[p * 3], w1 = wp[p * 3 + 1], w2 = wp[p * 3 + 2];
  w[k] = make_double2((a0.x * w0 + a1.x * w1) + a2.x * w2, (a0.y * w0 + a1.y * w1) + a2.y * w2);
  // t1 = Vinv a(row 1), t2 = Vinv a(row 2)
  const double t10 = (i00 * a0.x + i01 * a1.x) + i02 * a2.x, t11 = (i01 * a0.x + i11 * a1.x) + i12 * a2.x,
               t12 = (i02 * a0.x + i12 * a1.x) + i22 * a2.x;
  const double t20 = (i00 * a0.y + i01 * a1.y) + i02 * a2.y, t21 = (i01 * a0.y + i11 * a1.y) + i12 * a2.y,
               t22 = (i02 * a0.y + i12 * a1.y) + i22 * a2.y;
  T[k] = (a0.x * t10 + a1.x * t11) + a2.x * t12;
  T[nl + k] = (a0.x * t20 + a1.x * t21) + a2.x * t22;
  T[2 * nl + k] = (a0.y * t20 + a1.y * t21) + a2.y * t22;
}

// ---------------------------------------------------------------------------------------------
// camera-major passes: one warp per task (a slice of one camera's observations)
// The camera record is warp-uniform; each lane recomputes the camera part B of its observation's block.
//   MODE 0: accumulate U = B'B (45) and g_c = -B'F (9)
//   MODE 1: accumulate B' T B (45) and B' w (9)         (Schur diagonal blocks, right-hand side)
//   MODE 2: accumulate B' w (9)                         (Schur product)
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(PT_THREADS, MODE == 2 ? 4 : 1)
k_cam_pass(const int32_t* __restrict__ tbeg, const int32_t* __restrict__ tend, int64_t nctasks,
           const int32_t* __restrict__ task_cam, const int32_t* __restrict__ cam_t0, int32_t* __restrict__ cam_cnt,
           const int32_t* __restrict__ cperm, const int32_t* __restrict__ pntc, int64_t nl,
           const double* __restrict__ camtab, const double2* __restrict__ x4, const double2* __restrict__ F,
           const double2* __restrict__ w, const double* __restrict__ T, double* taskpart, double* __restrict__ out,
           const double* __restrict__ scal, double* mail, const unsigned long long* seqp, int64_t n9) {
  if (MODE == 2 && scal[S_DONE] != 0.0) return;
  constexpr int NACC = (MODE == 2) ? 9 : NV;
  if (MODE == 2 && mail) out = mail + ((*seqp) & 1ull) * n9;  // this exchange's half of the peer-visible mailbox
  const int lane = threadIdx.x & 31;
  const int64_t task = blockIdx.x * (int64_t)(PT_THREADS / 32) + (threadIdx.x >> 5);
  if (task >= nctasks) return;
  const int c = task_cam[task];
  double cam[14];
  {
    const double2* src = reinterpret_cast<const double2*>(camtab + (int64_t)c * CAM_REC);  // warp-uniform
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      const double2 t = __ldg(src + i);
      cam[2 * i] = t.x;
      cam[2 * i + 1] = t.y;
    }
  }
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = 0.0;
  const int e = tend[task];
  // Software pipeline: the gathers of the next observation (index -> point sector / w: two dependent L2 round
  // trips) are issued before the ~300 FP64 operations of the current one, so they overlap.
  struct Obs {
    double2 xa, xb, wk;
    double t00, t01, t11;
  };
  auto fetch = [&](int pos, Obs& ob) {
    const int k = __ldg(cperm + pos);
    const int p = __ldg(pntc + pos);
    ob.xa = __ldg(x4 + 2 * (int64_t)p);  // one 32-byte sector
    ob.xb = __ldg(x4 + 2 * (int64_t)p + 1);
    ob.t00 = 1.0;
    ob.t01 = 0.0;
    ob.t11 = 1.0;
    if (MODE == 0) {
      const double2 f = F[k];
      ob.wk = make_double2(-f.x, -f.y);
    } else {
      ob.wk = w[k];
      if (MODE == 1) {
        ob.t00 = T[k];
        ob.t01 = T[nl + k];
        ob.t11 = T[2 * nl + k];
      }
    }
  };
  int pos = tbeg[task] + lane;
  Obs cur, nxt;
  if (pos < e) fetch(pos, cur);
  for (; pos < e; pos += 32) {
    const bool more = pos + 32 < e;
    if (more) fetch(pos + 32, nxt);
    const double2 xa = cur.xa, xb = cur.xb, wk = cur.wk;
    const double t00 = cur.t00, t01 = cur.t01, t11 = cur.t11;
    if (more) cur = nxt;
    const double X[3] = {xa.x, xa.y, xb.x};
    ObsBlock o;
    eval_block<false>(X, cam, 0.0, 0.0, o);  // camera part only
    double2 B[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) B[j] = make_double2(nan0(o.B[j]), nan0(o.B[9 + j]));
    if (MODE == 2) {
#pragma unroll
      for (int j = 0; j < 9; ++j) acc[j] += B[j].x * wk.x + B[j].y * wk.y;
    } else {
      int q = 0;
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const double ci_x = (MODE == 0) ? B[i].x : t00 * B[i].x + t01 * B[i].y;
        const double ci_y = (MODE == 0) ? B[i].y : t01 * B[i].x + t11 * B[i].y;
#pragma unroll
        for (int j = i; j < 9; ++j) {
          acc[q] += ci_x * B[j].x + ci_y * B[j].y;
          ++q;
        }
      }
#pragma unroll
      for (int j = 0; j < 9; ++j) acc[45 + j] += B[j].x * wk.x + B[j].y * wk.y;
    }
  }
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = warp_sum(acc[i]);
  // ordered two-level sum without a second kernel: the task that finishes last for its camera adds the
  // camera's task partials in task order (threadfence reduction; partials are read past L1)
  const int tb0 = cam_t0[c], nt = cam_t0[c + 1] - tb0;
  if (nt == 1) {
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) out[(int64_t)c * NACC + i] = acc[i];
    }
    return;
  }
  int last = 0;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) taskpart[task * NACC + i] = acc[i];
    __threadfence();
    last = (atomicAdd(cam_cnt + c, 1) == nt - 1);
  }
  last = __shfl_sync(0xffffffffu, last, 0);
  if (last) {
    __threadfence();
    for (int j = lane; j < NACC; j += 32) {
      double s = 0.0;
      for (int t = 0; t < nt; ++t) s += __ldcg(taskpart + (int64_t)(tb0 + t) * NACC + j);
      out[(int64_t)c * NACC + j] = s;
    }
    if (lane == 0) cam_cnt[c] = 0;  // ready for the next pass (kernel boundaries order this)
  }
}

// x4[p] = (X_p, 0): one 32-byte sector per point for the random gathers of the camera-major passes
__global__ void __launch_bounds__(256)
k_pad_points(const double* __restrict__ x, int64_t p_lo, int64_t p_hi, double2* __restrict__ x4) {
  const int64_t p = p_lo + blockIdx.x * (int64_t)256 + threadIdx.x;
  if (p >= p_hi) return;
  x4[2 * p] = make_double2(x[3 * p], x[3 * p + 1]);
  x4[2 * p + 1] = make_double2(x[3 * p + 2], 0.0);
}

// cameras without observations on this rank: their sums are zero
__global__ void k_zero_cams(const int32_t* __restrict__ cams, int n, int nacc, double* __restrict__ out,
                            double* mail, const unsigned long long* seqp, int64_t n9) {
  if (mail) out = mail + ((*seqp) & 1ull) * n9;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n * nacc) out[(int64_t)cams[i / nacc] * nacc + (i % nacc)] = 0.0;
}

__device__ __forceinline__ int sym9(int i, int j) {  // i <= j, row-wise upper triangle
  return i * 9 - (i * (i - 1)) / 2 + (j - i);
}

// H = U + lambda I, Minv = (H - corr)^-1 by Cholesky, b = g_c - rhs; thread per camera
__global__ void __launch_bounds__(64)
k_cam_finish(int64_t ncams, double lambda, const double* __restrict__ Ug, const double* __restrict__ Cr,
             double* __restrict__ H, double* __restrict__ Minv, double* __restrict__ b, double* __restrict__ scal) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= ncams) return;
  const double* u = Ug + c * NV;
  const double* cr = Cr + c * NV;
  double M[9][9], L[9][9], Li[9][9];
  for (int i = 0; i < 9; ++i)
    for (int j = i; j < 9; ++j) {
      const double h = u[sym9(i, j)] + (i == j ? lambda : 0.0);
      H[c * 81 + i * 9 + j] = h;
      H[c * 81 + j * 9 + i] = h;
      M[i][j] = M[j][i] = h - cr[sym9(i, j)];
    }
  bool ok = true;
  for (int j = 0; j < 9; ++j) {
    double d = M[j][j];
    for (int q = 0; q < j; ++q) d -= L[j][q] * L[j][q];
    if (!(d > 0.0)) ok = false;
    const double lj = sqrt(d);
    L[j][j] = lj;
    for (int i = j + 1; i < 9; ++i) {
      double s = M[i][j];
      for (int q = 0; q < j; ++q) s -= L[i][q] * L[j][q];
      L[i][j] = s / lj;
    }
  }
  if (!ok) scal[S_ERR] = 1.0;
  // Li = L^-1 (lower)
  for (int j = 0; j < 9; ++j) {
    Li[j][j] = 1.0 / L[j][j];
    for (int i = j + 1; i < 9; ++i) {
      double s = 0.0;
      for (int q = j; q < i; ++q) s -= L[i][q] * Li[q][j];
      Li[i][j] = s / L[i][i];
    }
  }
  for (int i = 0; i < 9; ++i)
    for (int j = i; j < 9; ++j) {
      double s = 0.0;
      for (int q = j; q < 9; ++q) s += Li[q][i] * Li[q][j];
      Minv[c * 81 + i * 9 + j] = s;
      Minv[c * 81 + j * 9 + i] = s;
    }
  for (int j = 0; j < 9; ++j) b[c * 9 + j] = u[45 + j] - cr[45 + j];
}

// sum of squares of the g_c part of Ug (single block)
__global__ void __launch_bounds__(RED_THREADS)
k_gc_norm(int64_t ncams, const double* __restrict__ Ug, double* __restrict__ scal) {
  __shared__ double sh[RED_THREADS / 32];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < ncams * 9; i += RED_THREADS) {
    const double g = Ug[(i / 9) * NV + 45 + (i % 9)];
    s += g * g;
  }
  s = block_sum<RED_THREADS>(s, sh);
  if (threadIdx.x == 0) scal[S_GC2] = s;
}

// ---------------------------------------------------------------------------------------------
// point-major pass of the Schur product (MODE 0) and the back-substitution (MODE 1)
// ---------------------------------------------------------------------------------------------
// Warp-cooperative gather of the 32 lanes' 9-vectors (72 B each) from a camera-sized vector: each
// load instruction covers ~3.5 whole records instead of 32 scattered 8-byte words.
__device__ __forceinline__ void warp_stage_vec9(const double* __restrict__ v, int cam_of_lane, int lane,
                                                double* __restrict__ rows /* 288 */) {
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const int e = 32 * i + lane;
    const int r = e / 9;
    const int c = __shfl_sync(0xffffffffu, cam_of_lane, r);
    rows[e] = __ldg(v + (int64_t)c * 9 + (e - 9 * r));
  }
  __syncwarp();
}

struct PtLane {
  double2 A0, A1, A2;  // point part of the block (columns 0..2)
  double2 B[9];        // camera part
  double vi[6];        // (V_p + lambda I)^-1, prefetched
  double g[3];         // g_p (back-substitution only)
  double2 f;           // residual (back-substitution only)
  double2 y;
  double t0, t1, t2;
};

// Issue every global load of one observation up front (block planes, the point's inverse, and for the
// back-substitution g_p and F): the only dependent chain left is index -> address.
template <int MODE>
__device__ __forceinline__ void pt_load(PtLane& L, bool valid, int64_t k, int64_t p, int64_t nl,
                                        const double2* __restrict__ Jp, const double2* __restrict__ F,
                                        const double* __restrict__ Vinv, const double* __restrict__ gp) {
  if (valid) {
    L.A0 = __ldcs(Jp + k);
    L.A1 = __ldcs(Jp + nl + k);
    L.A2 = __ldcs(Jp + 2 * nl + k);
#pragma unroll
    for (int j = 0; j < 9; ++j) L.B[j] = __ldcs(Jp + (int64_t)(3 + j) * nl + k);
#pragma unroll
    for (int i = 0; i < 6; ++i) L.vi[i] = __ldg(Vinv + p * 6 + i);
    if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 3; ++i) L.g[i] = __ldg(gp + p * 3 + i);
      L.f = __ldg(F + k);
    }
  } else {
    L.A0 = L.A1 = L.A2 = make_double2(0.0, 0.0);
#pragma unroll
    for (int j = 0; j < 9; ++j) L.B[j] = make_double2(0.0, 0.0);
#pragma unroll
    for (int i = 0; i < 6; ++i) L.vi[i] = 0.0;
    L.g[0] = L.g[1] = L.g[2] = 0.0;
    L.f = make_double2(0.0, 0.0);
  }
}

// y = B v_c (v staged in shared memory), t = A' y
__device__ __forceinline__ void pt_first(PtLane& L, const double* __restrict__ vrow, int lane) {
  double yx = 0.0, yy = 0.0;
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    const double vj = vrow[lane * 9 + j];
    yx += L.B[j].x * vj;
    yy += L.B[j].y * vj;
  }
  L.y = make_double2(yx, yy);
  L.t0 = L.A0.x * yx + L.A0.y * yy;
  L.t1 = L.A1.x * yx + L.A1.y * yy;
  L.t2 = L.A2.x * yx + L.A2.y * yy;
}

template <int MODE>
__global__ void __launch_bounds__(PT_THREADS, PS_MINBLOCKS)
k_point_solve(const int32_t* __restrict__ tstart, int64_t ntasks, const int32_t* __restrict__ cam_idx,
              const int32_t* __restrict__ pnt_idx, int64_t pnt0, int64_t nl, const double2* __restrict__ Jp,
              const double2* __restrict__ F, const double* __restrict__ vcam, const double* __restrict__ Vinv,
              const double* __restrict__ gp, double2* __restrict__ w_out, double* __restrict__ delta,
              double2* __restrict__ dr_out, double* __restrict__ part, const double* __restrict__ scal) {
  __shared__ double vst[(PT_THREADS / 32) * 288];
  __shared__ double sh[PT_THREADS / 32];
  if (MODE == 0 && scal[S_DONE] != 0.0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t task = blockIdx.x * (int64_t)(PT_THREADS / 32) + warp;
  double* vrow = vst + warp * 288;
  double acc_dr2 = 0.0;

  // per-observation second half, given the point's summed t
  auto second_half = [&](int64_t k, int64_t p, const PtLane& L, double T0, double T1, double T2, bool writer) {
    const double i00 = L.vi[0], i01 = L.vi[1], i02 = L.vi[2], i11 = L.vi[3], i12 = L.vi[4], i22 = L.vi[5];
    if (MODE == 0) {
      const double u0 = (i00 * T0 + i01 * T1) + i02 * T2, u1 = (i01 * T0 + i11 * T1) + i12 * T2,
                   u2 = (i02 * T0 + i12 * T1) + i22 * T2;
      w_out[k] = make_double2((L.A0.x * u0 + L.A1.x * u1) + L.A2.x * u2, (L.A0.y * u0 + L.A1.y * u1) + L.A2.y * u2);
    } else {
      const double r0 = L.g[0] - T0, r1 = L.g[1] - T1, r2 = L.g[2] - T2;
      const double d0 = (i00 * r0 + i01 * r1) + i02 * r2, d1 = (i01 * r0 + i11 * r1) + i12 * r2,
                   d2 = (i02 * r0 + i12 * r1) + i22 * r2;
      const double ex = ((L.A0.x * d0 + L.A1.x * d1) + L.A2.x * d2) + L.y.x + L.f.x;
      const double ey = ((L.A0.y * d0 + L.A1.y * d1) + L.A2.y * d2) + L.y.y + L.f.y;
      acc_dr2 += ex * ex + ey * ey;
      if (dr_out) dr_out[k] = make_double2(-ex, -ey);
      if (writer) {
        double* dp = delta + (pnt0 + p) * 3;
        dp[0] = d0;
        dp[1] = d1;
        dp[2] = d2;
      }
    }
  };

  if (task < ntasks) {
    const int64_t t0 = tstart[task], t1 = tstart[task + 1];
    if (t1 - t0 <= 32) {
      // whole points packed into one warp: segmented sums over runs of equal point id
      const int64_t k = t0 + lane;
      const bool valid = k < t1;
      int c = 0, p = -1;
      if (valid) {
        c = __ldg(cam_idx + k);
        p = (int)(__ldg(pnt_idx + k) - pnt0);
      }
      PtLane L;
      pt_load<MODE>(L, valid, k, p, nl, Jp, F, Vinv, gp);
      warp_stage_vec9(vcam, c, lane, vrow);
      pt_first(L, vrow, lane);
      const int pprev = __shfl_up_sync(0xffffffffu, p, 1);
      const bool head = (lane == 0) || (p != pprev);
      const unsigned hm = __ballot_sync(0xffffffffu, head);
      const int seg0 = 31 - __clz(hm & (0xffffffffu >> (31 - lane)));
      double s0 = L.t0, s1 = L.t1, s2 = L.t2;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const double a = shfl_up_d(s0, d), b = shfl_up_d(s1, d), cc = shfl_up_d(s2, d);
        if (lane - d >= seg0) {
          s0 += a;
          s1 += b;
          s2 += cc;
        }
      }
      const unsigned above = (lane == 31) ? 0u : (hm >> (lane + 1));
      const int tail = above ? lane + __ffs(above) - 1 : 31;
      const double T0 = shfl_d(s0, tail), T1 = shfl_d(s1, tail), T2 = shfl_d(s2, tail);
      if (valid) second_half(k, p, L, T0, T1, T2, lane == tail);
    } else {
      // one point with more than 32 observations: sum over chunks, then a second sweep
      const int64_t p = __ldg(pnt_idx + t0) - pnt0;
      double T0 = 0.0, T1 = 0.0, T2 = 0.0;
      for (int pass = 0; pass < 2; ++pass) {
        for (int64_t base = t0; base < t1; base += 32) {
          const int64_t k = base + lane;
          const bool valid = k < t1;
          const int c = valid ? __ldg(cam_idx + k) : 0;
          PtLane L;
          pt_load<MODE>(L, valid, k, p, nl, Jp, F, Vinv, gp);
          __syncwarp();
          warp_stage_vec9(vcam, c, lane, vrow);
          pt_first(L, vrow, lane);
          if (pass == 0) {
            T0 += warp_sum(L.t0);
            T1 += warp_sum(L.t1);
            T2 += warp_sum(L.t2);
          } else if (valid) {
            second_half(k, p, L, T0, T1, T2, k == t0);
          }
        }
      }
    }
  }
  if (MODE == 1) {
    acc_dr2 = block_sum<PT_THREADS>(acc_dr2, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = acc_dr2;
  }
}

// xt = x + delta on this rank's live entries (its point slice and all cameras), with the norms the
// LM tests need; delta is first divided by `div` (back-tracking, src/lm.jl:266).  part: 4 per block.
__global__ void __launch_bounds__(256)
k_step_update(const double* __restrict__ x, double* __restrict__ delta, double* __restrict__ xt, int64_t p_lo,
              int64_t p_hi, int64_t c_lo, int64_t c_hi, double div, double* __restrict__ part) {
  __shared__ double sh[256 / 32];
  const int64_t np = p_hi - p_lo, n = np + (c_hi - c_lo);
  double dp2 = 0, xp2 = 0, dc2 = 0, xc2 = 0;
  for (int64_t i = blockIdx.x * (int64_t)256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const bool is_p = i < np;
    const int64_t j = is_p ? p_lo + i : c_lo + (i - np);
    double d = delta[j];
    if (div != 1.0) {
      d = d / div;
      delta[j] = d;
    }
    const double v = x[j] + d;
    xt[j] = v;
    if (is_p) {
      dp2 += d * d;
      xp2 += v * v;
    } else {
      dc2 += d * d;
      xc2 += v * v;
    }
  }
  dp2 = block_sum<256>(dp2, sh);
  xp2 = block_sum<256>(xp2, sh);
  dc2 = block_sum<256>(dc2, sh);
  xc2 = block_sum<256>(xc2, sh);
  if (threadIdx.x == 0) {
    part[blockIdx.x * 4] = dp2;
    part[blockIdx.x * 4 + 1] = xp2;
    part[blockIdx.x * 4 + 2] = dc2;
    part[blockIdx.x * 4 + 3] = xc2;
  }
}

// back-tracking on the LDL path: dr <- (dr - r) / dd  (src/lm.jl:277-279), with its squared norm
__global__ void __launch_bounds__(PT_THREADS)
k_ls_dr(double2* __restrict__ dr, const double2* __restrict__ F, int64_t nl, double dd, double* __restrict__ part) {
  __shared__ double sh[PT_THREADS / 32];
  const int64_t k = blockIdx.x * (int64_t)PT_THREADS + threadIdx.x;
  double s = 0.0;
  if (k < nl) {
    const double2 a = dr[k], f = F[k];
    const double2 n = make_double2((a.x - f.x) / dd, (a.y - f.y) / dd);
    dr[k] = n;
    s = n.x * n.x + n.y * n.y;
  }
  s = block_sum<PT_THREADS>(s, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}

// ---------------------------------------------------------------------------------------------
// PCG on the reduced camera system: camera-sized vector work.  One iteration needs two global dot
// products (p.Sp and r.z) with vector updates between them; they run as three small multi-CTA
// kernels whose dot products are per-CTA partials summed in fixed order by every CTA of the next
// kernel (reproducible, no atomics, no host round trip; alpha, beta and the convergence flag live on
// the device).  A CTA owns whole cameras (28 cameras = 252 rows), so z = Minv r needs only
// __syncthreads.  [A single 8-CTA thread-block cluster with DSMEM reductions did the same work in
// 27 us on Venice (latency-bound: 4 dependent passes over 2.3 MB from 8 SMs) and does not scale to
// 13682 cameras; this version spreads the matrix reads over all SMs: profiles/r01_pcg_vector_*.]
// ---------------------------------------------------------------------------------------------
constexpr int VEC_THREADS = 256;
constexpr int CDOF = 6;  // coarse unknowns per camera cluster: the pose components (r, t); adding k1, k2, f to the
                       // coarse space does not reduce PCG iterations further (prototype: 165 vs 166)
constexpr int VEC_ROWS = 252;  // 28 cameras x 9 rows per CTA
constexpr int KZ_LIMIT = 48;   // deflation vectors at most (32 base + 16 refreshed)

__device__ __forceinline__ double row9(const double* __restrict__ M, const double* v, int64_t i) {
  const int64_t c = i / 9;
  const double* m = M + c * 81 + (i - c * 9) * 9;
  const double* x = v + c * 9;
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < 9; ++j) s += __ldg(m + j) * x[j];
  return s;
}

// the same total in every thread of every CTA: partials are added in index order
__device__ __forceinline__ double sum_partials(const double* part, int n, double* sh) {
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += VEC_THREADS) s += part[i];
  return block_sum<VEC_THREADS>(s, sh);
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// K7c: q = (U + lambda I) p - sum_k B'w = S p, per-CTA partials of p.q; also publishes the r.z of the
// previous iteration as the current one.
//   P2P = false: on entry q holds sum_k B'w (already allreduced by NCCL in sharded mode).
//   P2P = true : the sum over ranks is FUSED here.  Every rank's camera pass left its partial in its own
//     IPC-exported mailbox half (seq & 1); block 0 tells every peer "my partial #seq+1 is complete" with a
//     release store into the peer's flag slot (over NVLink), every CTA waits until all ranks have said so,
//     then each row adds the R partials in rank order straight from peer memory -- the same order on every
//     rank, so all ranks get bit-identical q.  Two mailbox halves suffice: nobody can be two exchanges ahead
//     of a rank that has not yet passed this wait.  A wait is bounded (~2 s) and flags an error instead of
//     hanging.
template <bool P2P>
__global__ void __launch_bounds__(VEC_THREADS)
k_pcg_q(int64_t n9, const double* __restrict__ H, const double* p, double* q, double* __restrict__ part_pq,
        double* scal, double* const* __restrict__ mails, unsigned long long* const* __restrict__ flags,
        const unsigned long long* seqp, int nranks, int rank) {
  __shared__ double sh[VEC_THREADS / 32];
  __shared__ int timed_out;
  if (scal[S_DONE] != 0.0) return;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    scal[S_RZ] = scal[S_RZN];
    scal[S_ITK] = scal[S_ITERS];
  }
  unsigned long long seq = 0;
  if (P2P) {
    seq = *seqp;
    if (threadIdx.x == 0) timed_out = 0;
    if (blockIdx.x == 0 && threadIdx.x < nranks) {
      __threadfence_system();  // the camera pass of this rank finished before this kernel started
      st_release_sys(flags[threadIdx.x] + rank, seq + 1);
    }
    __syncthreads();
    if (threadIdx.x < nranks) {
      const unsigned long long* f = flags[rank] + threadIdx.x;  // my slots, written by the peers
      const long long t0 = clock64();
      while (ld_acquire_sys(f) < seq + 1) {
        if (clock64() - t0 > 4000000000ll) {
          timed_out = 1;
          break;
        }
      }
    }
    __syncthreads();
    if (timed_out) {
      if (threadIdx.x == 0) {
        scal[S_DONE] = 2.0;
        scal[S_ERR] = 3.0;
      }
      return;
    }
  }
  const int64_t i = blockIdx.x * (int64_t)VEC_ROWS + threadIdx.x;
  double pq = 0.0;
  if (threadIdx.x < VEC_ROWS && i < n9) {
    double sum;
    if (P2P) {
      const int64_t off = (int64_t)(seq & 1ull) * n9 + i;
      sum = 0.0;
      for (int r = 0; r < nranks; ++r) sum += ld_volatile_f64(mails[r] + off);
    } else {
      sum = q[i];
    }
    const double qi = row9(H, p, i) - sum;
    q[i] = qi;
    pq = p[i] * qi;
  }
  pq = block_sum<VEC_THREADS>(pq, sh);
  if (threadIdx.x == 0) part_pq[blockIdx.x] = pq;
}

// Dot products of this warp's NU deflation vectors (j = warp, warp + 8, ...) with the CTA's slice of r (rs, in
// shared memory).  All NU x 8 loads of a lane are independent and unpredicated (row indices are clamped, the
// clamped rows get a zero weight) so that they are in flight together.
template <int NU, bool FULL>
__device__ __forceinline__ void warp_zdot_rows(const double* __restrict__ zb, int64_t n9, int nrow, const double* rs,
                                               double* __restrict__ zp) {
  constexpr int NW = VEC_THREADS / 32, NB = (VEC_ROWS + 31) / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double* zu[NU];
#pragma unroll
  for (int u = 0; u < NU; ++u) zu[u] = zb + (int64_t)(warp + NW * u) * n9;
  double t[NU];
#pragma unroll
  for (int u = 0; u < NU; ++u) t[u] = 0.0;
#pragma unroll
  for (int bt = 0; bt < NB; ++bt) {
    const int a = lane + 32 * bt;
    // a full CTA (252 rows) only clamps in its last batch: the other row offsets are compile-time constants
    const int ac = (FULL && bt < NB - 1) ? a : min(a, nrow - 1);
    const double ra = (a < nrow) ? rs[ac] : 0.0;
#pragma unroll
    for (int u = 0; u < NU; ++u) t[u] += zu[u][ac] * ra;
  }
#pragma unroll
  for (int u = 0; u < NU; ++u) {
    const double tt = warp_sum(t[u]);
    if (lane == 0) zp[warp + NW * u] = tt;
  }
}
template <int NU>
__device__ __forceinline__ void warp_zdot(const double* __restrict__ zb /* Z + first row of the CTA */, int64_t n9,
                                          int kz, int nrow, const double* rs, double* __restrict__ zp) {
  if (nrow == VEC_ROWS) warp_zdot_rows<NU, true>(zb, n9, nrow, rs, zp);
  else warp_zdot_rows<NU, false>(zb, n9, nrow, rs, zp);
}

// K7d: alpha = r.z / p.q; xc += alpha p; r -= alpha q; z = Minv r; per-CTA partials of r.z.
// INIT: r = b, xc = 0 instead of the update.
template <bool INIT>
__global__ void __launch_bounds__(VEC_THREADS)
k_pcg_xr(int64_t n9, int nparts, const double* __restrict__ b, const double* __restrict__ Minv, const double* p,
         const double* q, double* xc, double* r, double* z, const double* part_pq, double* __restrict__ part_rz,
         const double* scal, double* __restrict__ cpart, const double* __restrict__ Zb, int kz,
         double* __restrict__ zpart) {
  __shared__ double sh[VEC_THREADS / 32];
  __shared__ double rs[VEC_ROWS];
  const int64_t i = blockIdx.x * (int64_t)VEC_ROWS + threadIdx.x;
  const bool live = threadIdx.x < VEC_ROWS && i < n9;
  if (INIT) {
    if (live) {
      r[i] = b[i];
      xc[i] = 0.0;
    }
  } else {
    if (scal[S_DONE] != 0.0) return;
    const double pq = sum_partials(part_pq, nparts, sh);
    if (!(pq > 0.0)) return;  // breakdown: flagged by k_pcg_p
    const double alpha = scal[S_RZ] / pq;
    if (live) {
      xc[i] += alpha * p[i];
      r[i] -= alpha * q[i];
    }
  }
  __syncthreads();
  double rz = 0.0;
  if (live) {
    const double zi = row9(Minv, r, i);
    z[i] = zi;
    rz = r[i] * zi;
  }
  rz = block_sum<VEC_THREADS>(rz, sh);
  if (threadIdx.x == 0) part_rz[blockIdx.x] = rz;
  if (cpart || kz > 0) {
    if (threadIdx.x < VEC_ROWS) rs[threadIdx.x] = live ? r[i] : 0.0;
    __syncthreads();
  }
  if (cpart) {  // restriction to the coarse space: this CTA's 28 cameras summed per component (fixed order)
    if (threadIdx.x < 9) {
      double t = 0.0;
      for (int c = 0; c < VEC_ROWS / 9; ++c) t += rs[c * 9 + threadIdx.x];
      cpart[blockIdx.x * 9 + threadIdx.x] = t;
    }
  }
  if (kz > 0) {  // restriction to the deflation vectors: this CTA's slice of Z_j . r
    const int warp = threadIdx.x >> 5;
    const int64_t i0 = blockIdx.x * (int64_t)VEC_ROWS;
    const int nrow = (int)min((long long)VEC_ROWS, (long long)(n9 - i0));
    const int nu = (kz - warp + VEC_THREADS / 32 - 1) / (VEC_THREADS / 32);  // vectors of this warp (uniform)
    double* zp = zpart + (int64_t)blockIdx.x * kz;
    const double* zb = Zb + i0;
    switch (nu) {
      case 6: warp_zdot<6>(zb, n9, kz, nrow, rs, zp); break;
      case 5: warp_zdot<5>(zb, n9, kz, nrow, rs, zp); break;
      case 4: warp_zdot<4>(zb, n9, kz, nrow, rs, zp); break;
      case 3: warp_zdot<3>(zb, n9, kz, nrow, rs, zp); break;
      case 2: warp_zdot<2>(zb, n9, kz, nrow, rs, zp); break;
      case 1: warp_zdot<1>(zb, n9, kz, nrow, rs, zp); break;
      default: break;
    }
  }
}

// Coarse level of the two-level preconditioner  M^-1 = blkdiag(S_cc)^-1 + P Ac^-1 P':  P = piecewise-constant
// interpolation of the pose components from `ncl` clusters of consecutive cameras (CDOF = 6 coarse unknowns per
// cluster), Ac = P' S P.
// rc = P' r from the per-CTA partials (a CTA never straddles clusters), yc = Ac^-1 rc, and r.(P yc) = rc.yc
// for the r.z dot product.  One CTA; m = CDOF ncl <= 144.
constexpr int COARSE_THREADS = 1024, COARSE_GROUPS = COARSE_THREADS / 144;  // 7 groups of 144 threads
__global__ void __launch_bounds__(COARSE_THREADS)
k_pcg_coarse(int nvb, int ctas_per_cluster, int m, const double* __restrict__ cpart, const double* __restrict__ Aci,
             double* __restrict__ yc, double* scal, int init, int kz, const double* __restrict__ zpart) {
  __shared__ double rc[144], yy[144], part[COARSE_GROUPS][144];
  if (!init && scal[S_DONE] != 0.0) return;
  const int t = threadIdx.x, lane = t & 31;
  const int mcl = m - kz;  // cluster unknowns first, then one unknown per deflation vector
  if (t < mcl) {
    const int I = t / CDOF, jj = t - CDOF * I;
    const int b0 = I * ctas_per_cluster, b1 = min(nvb, b0 + ctas_per_cluster);
    double v = 0.0;
    for (int bb = b0; bb < b1; ++bb) v += cpart[bb * 9 + jj];
    rc[t] = v;
  }
  for (int j = t >> 5; j < kz; j += COARSE_THREADS / 32) {  // Z_j . r: a warp sums the per-CTA partials of one vector
    double v = 0.0;
    for (int bb = lane; bb < nvb; bb += 32) v += zpart[(int64_t)bb * kz + j];
    v = warp_sum(v);
    if (lane == 0) rc[mcl + j] = v;
  }
  __syncthreads();
  // yc = Aci rc: the m columns are split over 7 thread groups (Aci is exactly symmetric, so group g reads rows
  // of its column range: coalesced), partial sums added in group order
  const int g = t / 144, a = t - 144 * g;
  if (g < COARSE_GROUPS) {
    const int chunk = (m + COARSE_GROUPS - 1) / COARSE_GROUPS;
    const int q0 = g * chunk, q1 = min(m, q0 + chunk);
    double v = 0.0;
    if (a < m) {
#pragma unroll 8
      for (int bq = q0; bq < q1; ++bq) v += Aci[bq * m + a] * rc[bq];
    }
    part[g][a] = v;
  }
  __syncthreads();
  if (t < m) {
    double v = 0.0;
#pragma unroll
    for (int gg = 0; gg < COARSE_GROUPS; ++gg) v += part[gg][t];
    yy[t] = v;
    yc[t] = v;
  }
  __syncthreads();
  if (t < 32) {  // r.(coarse correction) = rc.yc
    double v = 0.0;
    for (int bq = t; bq < m; bq += 32) v += rc[bq] * yy[bq];
    v = warp_sum(v);
    if (t == 0) scal[S_RCY] = v;
  }
}

// K7e: beta = r.z(new) / r.z(old); p = z + beta p; iteration count and convergence flag.
// INIT: p = z, r0.z0.
template <bool INIT>
__global__ void __launch_bounds__(VEC_THREADS)
k_pcg_p(int64_t n9, int nparts, const double* z, double* p, const double* part_pq, const double* part_rz,
        double* scal, double tol, unsigned long long* seqp, const double* __restrict__ yc, int ctas_per_cluster,
        const double* __restrict__ Zb, int kz, int mcl, double* __restrict__ harv, int hcap,
        double* __restrict__ hcoef) {
  __shared__ double sh[VEC_THREADS / 32];
  const int64_t i = blockIdx.x * (int64_t)VEC_ROWS + threadIdx.x;
  const bool live = threadIdx.x < VEC_ROWS && i < n9;
  const bool lead = blockIdx.x == 0 && threadIdx.x == 0;
  // coarse correction of this row: z_i = (Minv r)_i + (P yc)_i + (Z yz)_i; its share of r.z is scal[S_RCY]
  double zc = 0.0;
  if (yc && live) {
    const int comp = (int)(i % 9);
    if (mcl > 0 && comp < CDOF) zc = yc[(blockIdx.x / ctas_per_cluster) * CDOF + comp];
#pragma unroll 16
    for (int j = 0; j < kz; ++j) zc += Zb[(int64_t)j * n9 + i] * yc[mcl + j];
  }
  if (INIT) {
    const double rz = sum_partials(part_rz, nparts, sh) + (yc ? scal[S_RCY] : 0.0);
    if (live) p[i] = z[i] + zc;
    if (harv && live && hcap > 0) harv[i] = (z[i] + zc) / sqrt(rz);  // Lanczos vector 0
    if (lead) {
      scal[S_RZN] = rz;
      scal[S_RZ0] = rz;
      scal[S_ITERS] = 0.0;
      scal[S_REL] = 1.0;
      // zero right-hand side: the zero step is the solution; NaN: let the caller see it
      scal[S_DONE] = (rz == 0.0) ? 1.0 : ((rz == rz) ? 0.0 : 2.0);
    }
    return;
  }
  if (scal[S_DONE] != 0.0) return;  // (a CTA that starts after the lead thread flagged convergence may skip
                                    //  its slice of p: p is dead once the solve is done)
  const double pq = sum_partials(part_pq, nparts, sh);
  if (!(pq > 0.0)) {  // breakdown: S is SPD in exact arithmetic, so this is NaN/Inf or lost definiteness
    if (lead) {
      scal[S_DONE] = 2.0;
      scal[S_PQ] = pq;
    }
    return;
  }
  const double rzn = sum_partials(part_rz, nparts, sh) + (yc ? scal[S_RCY] : 0.0);
  const double beta = rzn / scal[S_RZ];
  if (live) p[i] = (z[i] + zc) + beta * p[i];
  if (harv) {  // harvest the Lanczos vector z_t / sqrt(r_t.z_t) and the CG coefficients of step t - 1
    const int t = (int)scal[S_ITK] + 1;
    if (t < hcap && live) harv[(int64_t)t * n9 + i] = (z[i] + zc) / sqrt(rzn);
    if (lead && t <= hcap) {
      hcoef[t - 1] = scal[S_RZ] / pq;
      hcoef[hcap + t - 1] = beta;
    }
  }
  if (lead && seqp) *seqp += 1;  // next exchange uses the other mailbox half (no CTA of this kernel reads it)
  if (lead) {
    const double rel = sqrt(rzn / scal[S_RZ0]);
    scal[S_RZN] = rzn;
    scal[S_PQ] = pq;
    scal[S_ITERS] += 1.0;
    scal[S_REL] = rel;
    if (!(rel > tol)) scal[S_DONE] = (rel == rel) ? 1.0 : 2.0;
  }
}

// Small camera systems (9 ncams <= 1152 rows, i.e. up to 128 cameras): the whole vector half of a PCG iteration
// -- k_pcg_q, k_pcg_xr, k_pcg_coarse and k_pcg_p above -- as ONE CTA, the two dot products being block sums.
// Measured: LadyBug-49 (441 rows) 20.9 vs 34.6 us per PCG iteration; from ~2000 rows on the single CTA is the
// slower choice (Trafalgar-257, 2313 rows: 50.7 vs 46.6 us), hence the threshold.
constexpr int SMALL_THREADS = 1024;
template <bool P2P>
__global__ void __launch_bounds__(SMALL_THREADS)
k_pcg_small(int64_t n9, int64_t ncams, const double* __restrict__ H, const double* __restrict__ Minv, double* p,
            double* q, double* xc, double* r, double* z, double* scal, double tol, double* const* __restrict__ mails,
            unsigned long long* const* __restrict__ flags, unsigned long long* seqp, int nranks, int rank,
            const double* __restrict__ Aci, int m, int cams_per_cluster) {
  __shared__ double sh[SMALL_THREADS / 32];
  __shared__ double rc[144], yy[144];
  __shared__ int timed_out;
  if (scal[S_DONE] != 0.0) return;
  const double rz = scal[S_RZN], rz0 = scal[S_RZ0];
  unsigned long long seq = 0;
  if (P2P) {  // same handshake as k_pcg_q<true>
    seq = *seqp;
    if (threadIdx.x == 0) timed_out = 0;
    if (threadIdx.x < nranks) {
      __threadfence_system();
      st_release_sys(flags[threadIdx.x] + rank, seq + 1);
    }
    __syncthreads();
    if (threadIdx.x < nranks) {
      const unsigned long long* f = flags[rank] + threadIdx.x;
      const long long t0 = clock64();
      while (ld_acquire_sys(f) < seq + 1) {
        if (clock64() - t0 > 4000000000ll) {
          timed_out = 1;
          break;
        }
      }
    }
    __syncthreads();
    if (timed_out) {
      if (threadIdx.x == 0) {
        scal[S_DONE] = 2.0;
        scal[S_ERR] = 3.0;
      }
      return;
    }
  }
  double pq = 0.0;
  for (int64_t i = threadIdx.x; i < n9; i += SMALL_THREADS) {
    double sum;
    if (P2P) {
      const int64_t off = (int64_t)(seq & 1ull) * n9 + i;
      sum = 0.0;
      for (int rr = 0; rr < nranks; ++rr) sum += ld_volatile_f64(mails[rr] + off);
    } else {
      sum = q[i];
    }
    const double qi = row9(H, p, i) - sum;
    q[i] = qi;
    pq += p[i] * qi;
  }
  pq = block_sum<SMALL_THREADS>(pq, sh);
  if (!(pq > 0.0)) {
    if (threadIdx.x == 0) {
      scal[S_DONE] = 2.0;
      scal[S_PQ] = pq;
    }
    return;
  }
  const double alpha = rz / pq;
  for (int64_t i = threadIdx.x; i < n9; i += SMALL_THREADS) {
    xc[i] += alpha * p[i];
    r[i] -= alpha * q[i];
  }
  __syncthreads();
  double rzb = 0.0;
  for (int64_t i = threadIdx.x; i < n9; i += SMALL_THREADS) {
    const double zi = row9(Minv, r, i);
    z[i] = zi;
    rzb += r[i] * zi;
  }
  double rcy = 0.0;
  if (m > 0) {  // coarse level: rc = P' r, yc = Ac^-1 rc
    if (threadIdx.x < m) {
      const int I = threadIdx.x / CDOF, jj = threadIdx.x - CDOF * I;
      const int64_t c0 = (int64_t)I * cams_per_cluster, c1 = min((long long)ncams, (long long)(c0 + cams_per_cluster));
      double t = 0.0;
      for (int64_t c = c0; c < c1; ++c) t += r[c * 9 + jj];
      rc[threadIdx.x] = t;
    }
    __syncthreads();
    if (threadIdx.x < m) {
      double t = 0.0;
      for (int bq = 0; bq < m; ++bq) t += Aci[bq * m + threadIdx.x] * rc[bq];  // (symmetric: coalesced)
      yy[threadIdx.x] = t;
    }
    __syncthreads();
    for (int bq = 0; bq < m; ++bq) rcy += rc[bq] * yy[bq];
  }
  const double rzn = block_sum<SMALL_THREADS>(rzb, sh) + rcy;
  const double beta = rzn / rz;
  for (int64_t i = threadIdx.x; i < n9; i += SMALL_THREADS) {
    double zc = 0.0;
    if (m > 0) {
      const int comp = (int)(i % 9);
      if (comp < CDOF) zc = yy[(int)((i / 9) / cams_per_cluster) * CDOF + comp];
    }
    p[i] = (z[i] + zc) + beta * p[i];
  }
  if (threadIdx.x == 0) {
    const double rel = sqrt(rzn / rz0);
    scal[S_RZN] = rzn;
    scal[S_PQ] = pq;
    scal[S_ITERS] += 1.0;
    scal[S_REL] = rel;
    if (!(rel > tol)) scal[S_DONE] = (rel == rel) ? 1.0 : 2.0;
    if (P2P) *seqp += 1;
  }
}

// Direct assembly of the coarse matrix  Ac = P'HP - sum_p G_p Vinv_p G_p',  G_p[I] = sum_{k in p, cam in I} B_k[:, :6]' A_k:
// one pass over the points instead of CDOF*ncl applications of S.  Sums are accumulated in 64-bit fixed
// point (entries normalised by sqrt(diag P'HP), scale 2^40), first per CTA in shared memory, then globally:
// integer addition is associative, so the result does not depend on the order of the atomics (or of the
// ranks) and the preconditioner -- hence the PCG iterates -- stay bit-reproducible.
constexpr int NCL_MAX = 24;
constexpr double CQ_SCALE = 1099511627776.0;  // 2^40

// d[a] = sqrt(sum_{c in I} H_c[jj][jj]) for a = (I, jj): one thread per coarse unknown
__global__ void k_coarse_diag(int64_t ncams, int cams_per_cluster, int m, const double* __restrict__ H,
                              double* __restrict__ d) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= m) return;
  const int I = a / CDOF, jj = a - CDOF * I;
  const int64_t c0 = (int64_t)I * cams_per_cluster, c1 = min((long long)ncams, (long long)(c0 + cams_per_cluster));
  double t = 0.0;
  for (int64_t c = c0; c < c1; ++c) t += H[c * 81 + jj * 10];
  d[a] = sqrt(t);
}

__global__ void __launch_bounds__(256)
k_coarse_assemble(const int32_t* __restrict__ pstart, int64_t npl, int64_t nl, const int32_t* __restrict__ cam_idx,
                  const double2* __restrict__ Jp, const double* __restrict__ Vinv, int cams_per_cluster, int m,
                  const double* __restrict__ d, unsigned long long* __restrict__ Acq) {
  extern __shared__ unsigned long long sacc[];  // m x m
  for (int e = threadIdx.x; e < m * m; e += 256) sacc[e] = 0ull;
  __syncthreads();
  for (int64_t p = blockIdx.x * (int64_t)256 + threadIdx.x; p < npl; p += (int64_t)gridDim.x * 256) {
    const int k0 = pstart[p], k1 = pstart[p + 1];
    if (k0 == k1) continue;
    const double* vi = Vinv + p * 6;
    const double V[3][3] = {{vi[0], vi[1], vi[2]}, {vi[1], vi[3], vi[4]}, {vi[2], vi[4], vi[5]}};
    int n = 0, ids[NCL_MAX];
    double G[NCL_MAX][18];
    for (int k = k0; k < k1; ++k) {
      const int I = __ldg(cam_idx + k) / cams_per_cluster;
      int e = n - 1;
      while (e >= 0 && ids[e] != I) --e;  // cameras ascend within a point in BAL order: normally the last entry
      if (e < 0) {
        e = n++;
        ids[e] = I;
        for (int q = 0; q < 18; ++q) G[e][q] = 0.0;
      }
      const double2 a0 = Jp[k], a1 = Jp[nl + k], a2 = Jp[2 * nl + k];
#pragma unroll
      for (int a = 0; a < CDOF; ++a) {
        const double2 b = Jp[(int64_t)(3 + a) * nl + k];
        G[e][a * 3 + 0] += b.x * a0.x + b.y * a0.y;
        G[e][a * 3 + 1] += b.x * a1.x + b.y * a1.y;
        G[e][a * 3 + 2] += b.x * a2.x + b.y * a2.y;
      }
    }
    for (int e1 = 0; e1 < n; ++e1) {
      double T[CDOF][3];
      for (int a = 0; a < CDOF; ++a)
        for (int t = 0; t < 3; ++t)
          T[a][t] = (G[e1][a * 3] * V[0][t] + G[e1][a * 3 + 1] * V[1][t]) + G[e1][a * 3 + 2] * V[2][t];
      for (int e2 = 0; e2 < n; ++e2)
        for (int a = 0; a < CDOF; ++a) {
          const int row = ids[e1] * CDOF + a;
          for (int b = 0; b < CDOF; ++b) {
            const int col = ids[e2] * CDOF + b;
            const double v = (T[a][0] * G[e2][b * 3] + T[a][1] * G[e2][b * 3 + 1]) + T[a][2] * G[e2][b * 3 + 2];
            const long long qv = __double2ll_rn(v / (d[row] * d[col]) * CQ_SCALE);
            atomicAdd(&sacc[row * m + col], (unsigned long long)qv);
          }
        }
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < m * m; e += 256)
    if (sacc[e]) atomicAdd(&Acq[e], sacc[e]);
}

// Ac = P'HP - dequantised Schur part
__global__ void __launch_bounds__(256)
k_coarse_finish(int64_t ncams, int cams_per_cluster, int m, const double* __restrict__ H, const double* __restrict__ d,
                const long long* __restrict__ Acq, double* __restrict__ Ac, int ld) {
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= m * m) return;
  const int row = e / m, col = e - row * m;
  const int I = row / CDOF, a = row - CDOF * I, J = col / CDOF, b = col - CDOF * J;
  double t = 0.0;
  if (I == J) {
    const int64_t c0 = (int64_t)I * cams_per_cluster, c1 = min((long long)ncams, (long long)(c0 + cams_per_cluster));
    for (int64_t c = c0; c < c1; ++c) t += H[c * 81 + a * 9 + b];
  }
  Ac[row * ld + col] = t - ((double)Acq[e] / CQ_SCALE) * (d[row] * d[col]);  // ld > m: room for the deflation block
}

// coarse-space setup: basis vector (cluster I, component j) of P, restriction of S v, inversion of Ac
__global__ void __launch_bounds__(256)
k_coarse_basis(int64_t n9, int rows_per_cluster, int col, double* __restrict__ v) {
  const int64_t i = blockIdx.x * (int64_t)256 + threadIdx.x;
  if (i >= n9) return;
  const int I = col / CDOF, j = col - CDOF * I;
  v[i] = ((int)(i / rows_per_cluster) == I && (int)(i % 9) == j) ? 1.0 : 0.0;
}

// Ac[:, col] = P' q: one CTA per cluster, 32 x 9 threads, fixed summation order
__global__ void __launch_bounds__(288)
k_coarse_restrict(int64_t ncams, int cams_per_cluster, int m, int col, const double* __restrict__ q,
                  double* __restrict__ Ac) {
  __shared__ double part[288];
  const int I = blockIdx.x, t = threadIdx.x / 9, jj = threadIdx.x - 9 * t;
  const int64_t c0 = (int64_t)I * cams_per_cluster, c1 = min(ncams, c0 + cams_per_cluster);
  double sum = 0.0;
  for (int64_t c = c0 + t; c < c1; c += 32) sum += q[c * 9 + jj];
  part[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x < CDOF) {
    double tot = 0.0;
    for (int u = 0; u < 32; ++u) tot += part[u * 9 + threadIdx.x];
    Ac[(I * CDOF + threadIdx.x) * m + col] = tot;
  }
}

// Aci = Ac^-1 (m <= 144, SPD) by in-place Gauss-Jordan without pivoting on a symmetrised copy held in shared
// memory (m^2 doubles <= 162 KB); one CTA.  Pivot step k: row k becomes row_k / pivot with 1 / pivot in column
// k; every other row i subtracts A[i][k] times it, its column k starting from zero.
constexpr int INV_THREADS = 1024;
__global__ void __launch_bounds__(INV_THREADS)
k_coarse_invert(int m, const double* __restrict__ Ac, double* __restrict__ Aci, double* scal) {
  extern __shared__ double Ash[];  // m x m
  __shared__ double colk[144], rowk[144];
  for (int e = threadIdx.x; e < m * m; e += INV_THREADS) {
    const int a = e / m, bq = e - a * m;
    Ash[e] = 0.5 * (Ac[a * m + bq] + Ac[bq * m + a]);
  }
  __syncthreads();
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  for (int k = 0; k < m; ++k) {
    const double piv = Ash[k * m + k];
    if (!(piv > 0.0) && threadIdx.x == 0) bad = 1;
    if (threadIdx.x < m) {
      colk[threadIdx.x] = Ash[threadIdx.x * m + k];
      rowk[threadIdx.x] = ((int)threadIdx.x == k ? 1.0 : Ash[k * m + threadIdx.x]) / piv;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < m * m; e += INV_THREADS) {
      const int a = e / m, bq = e - a * m;
      Ash[e] = (a == k) ? rowk[bq] : ((bq == k ? 0.0 : Ash[e]) - colk[a] * rowk[bq]);
    }
    __syncthreads();
  }
  // lost definiteness (rounding of the fixed-point sums at tiny lambda, nearly dependent deflation vectors):
  // a zero "inverse" switches the coarse level off for the solves that use it
  const bool zero = bad != 0;
  for (int e = threadIdx.x; e < m * m; e += INV_THREADS) {
    const int a = e / m, bq = e - a * m;
    Aci[e] = zero ? 0.0 : 0.5 * (Ash[a * m + bq] + Ash[bq * m + a]);
  }
  if (zero && threadIdx.x == 0) scal[S_CBAD] = 1.0;
}

// ---------------------------------------------------------------------------------------------
// Schur product for NR vectors at once (setup of the deflation block of the coarse matrix: S Z_j for all j).
// The J blocks -- the bulk of the traffic of the point-major pass -- and the recomputed camera parts of the
// camera-major pass are shared by the NR right-hand sides.  Same arithmetic per vector as k_point_solve<0> /
// k_cam_pass<2>; no done-flag, no mailbox: the sum over ranks goes through NCCL once per NR vectors.
// ---------------------------------------------------------------------------------------------
template <int NR>
__global__ void __launch_bounds__(PT_THREADS, 4)
k_point_solve_multi(const int32_t* __restrict__ tstart, int64_t ntasks, const int32_t* __restrict__ cam_idx,
                    const int32_t* __restrict__ pnt_idx, int64_t pnt0, int64_t nl, const double2* __restrict__ Jp,
                    const double* __restrict__ vcam, int64_t vstride, const double* __restrict__ Vinv,
                    double2* __restrict__ w_out /* NR planes of nl */) {
  __shared__ double vst[(PT_THREADS / 32) * 288];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t task = blockIdx.x * (int64_t)(PT_THREADS / 32) + warp;
  double* vrow = vst + warp * 288;
  if (task >= ntasks) return;  // (no block-wide barrier below)
  auto second_half = [&](int64_t k, const PtLane& L, double T0, double T1, double T2, int r) {
    const double i00 = L.vi[0], i01 = L.vi[1], i02 = L.vi[2], i11 = L.vi[3], i12 = L.vi[4], i22 = L.vi[5];
    const double u0 = (i00 * T0 + i01 * T1) + i02 * T2, u1 = (i01 * T0 + i11 * T1) + i12 * T2,
                 u2 = (i02 * T0 + i12 * T1) + i22 * T2;
    w_out[(int64_t)r * nl + k] =
        make_double2((L.A0.x * u0 + L.A1.x * u1) + L.A2.x * u2, (L.A0.y * u0 + L.A1.y * u1) + L.A2.y * u2);
  };
  const int64_t t0 = tstart[task], t1 = tstart[task + 1];
  if (t1 - t0 <= 32) {
    const int64_t k = t0 + lane;
    const bool valid = k < t1;
    int c = 0, p = -1;
    if (valid) {
      c = __ldg(cam_idx + k);
      p = (int)(__ldg(pnt_idx + k) - pnt0);
    }
    PtLane L;
    pt_load<0>(L, valid, k, p, nl, Jp, nullptr, Vinv, nullptr);
    const int pprev = __shfl_up_sync(0xffffffffu, p, 1);
    const bool head = (lane == 0) || (p != pprev);
    const unsigned hm = __ballot_sync(0xffffffffu, head);
    const int seg0 = 31 - __clz(hm & (0xffffffffu >> (31 - lane)));
    const unsigned above = (lane == 31) ? 0u : (hm >> (lane + 1));
    const int tail = above ? lane + __ffs(above) - 1 : 31;
    for (int r = 0; r < NR; ++r) {
      __syncwarp();  // every lane is done with the previous vector's rows
      warp_stage_vec9(vcam + (int64_t)r * vstride, c, lane, vrow);
      pt_first(L, vrow, lane);
      double s0 = L.t0, s1 = L.t1, s2 = L.t2;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const double a = shfl_up_d(s0, d), b = shfl_up_d(s1, d), cc = shfl_up_d(s2, d);
        if (lane - d >= seg0) {
          s0 += a;
          s1 += b;
          s2 += cc;
        }
      }
      const double T0 = shfl_d(s0, tail), T1 = shfl_d(s1, tail), T2 = shfl_d(s2, tail);
      if (valid) second_half(k, L, T0, T1, T2, r);
    }
  } else {
    // one point with more than 32 observations (rare): vector by vector, as k_point_solve does
    const int64_t p = __ldg(pnt_idx + t0) - pnt0;
    for (int r = 0; r < NR; ++r) {
      double T0 = 0.0, T1 = 0.0, T2 = 0.0;
      for (int pass = 0; pass < 2; ++pass) {
        for (int64_t base = t0; base < t1; base += 32) {
          const int64_t k = base + lane;
          const bool valid = k < t1;
          const int c = valid ? __ldg(cam_idx + k) : 0;
          PtLane L;
          pt_load<0>(L, valid, k, p, nl, Jp, nullptr, Vinv, nullptr);
          __syncwarp();
          warp_stage_vec9(vcam + (int64_t)r * vstride, c, lane, vrow);
          pt_first(L, vrow, lane);
          if (pass == 0) {
            T0 += warp_sum(L.t0);
            T1 += warp_sum(L.t1);
            T2 += warp_sum(L.t2);
          } else if (valid) {
            second_half(k, L, T0, T1, T2, r);
          }
        }
      }
    }
  }
}

template <int NR>
__global__ void __launch_bounds__(PT_THREADS, 3)
k_cam_pass_multi(const int32_t* __restrict__ tbeg, const int32_t* __restrict__ tend, int64_t nctasks,
                 const int32_t* __restrict__ task_cam, const int32_t* __restrict__ cam_t0,
                 int32_t* __restrict__ cam_cnt, const int32_t* __restrict__ cperm, const int32_t* __restrict__ pntc,
                 int64_t nl, const double* __restrict__ camtab, const double2* __restrict__ x4,
                 const double2* __restrict__ w /* NR planes of nl */, double* taskpart /* nctasks x 9 NR */,
                 double* __restrict__ out /* NR x n9 */, int64_t n9) {
  constexpr int NACC = 9 * NR;
  const int lane = threadIdx.x & 31;
  const int64_t task = blockIdx.x * (int64_t)(PT_THREADS / 32) + (threadIdx.x >> 5);
  if (task >= nctasks) return;
  const int c = task_cam[task];
  double cam[14];
  {
    const double2* src = reinterpret_cast<const double2*>(camtab + (int64_t)c * CAM_REC);  // warp-uniform
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      const double2 t = __ldg(src + i);
      cam[2 * i] = t.x;
      cam[2 * i + 1] = t.y;
    }
  }
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = 0.0;
  const int e = tend[task];
  for (int pos = tbeg[task] + lane; pos < e; pos += 32) {
    const int k = __ldg(cperm + pos);
    const int p = __ldg(pntc + pos);
    const double2 xa = __ldg(x4 + 2 * (int64_t)p), xb = __ldg(x4 + 2 * (int64_t)p + 1);
    double2 wk[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) wk[r] = w[(int64_t)r * nl + k];
    const double X[3] = {xa.x, xa.y, xb.x};
    ObsBlock o;
    eval_block<false>(X, cam, 0.0, 0.0, o);  // camera part only
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const double bx = nan0(o.B[j]), by = nan0(o.B[9 + j]);
#pragma unroll
      for (int r = 0; r < NR; ++r) acc[r * 9 + j] += bx * wk[r].x + by * wk[r].y;
    }
  }
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = warp_sum(acc[i]);
  const int tb0 = cam_t0[c], nt = cam_t0[c + 1] - tb0;
  if (nt == 1) {
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) out[(int64_t)(i / 9) * n9 + (int64_t)c * 9 + (i % 9)] = acc[i];
    }
    return;
  }
  int last = 0;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) taskpart[task * NACC + i] = acc[i];
    __threadfence();
    last = (atomicAdd(cam_cnt + c, 1) == nt - 1);
  }
  last = __shfl_sync(0xffffffffu, last, 0);
  if (last) {
    __threadfence();
    for (int j = lane; j < NACC; j += 32) {
      double s = 0.0;
      for (int t = 0; t < nt; ++t) s += __ldcg(taskpart + (int64_t)(tb0 + t) * NACC + j);
      out[(int64_t)(j / 9) * n9 + (int64_t)c * 9 + (j % 9)] = s;
    }
    if (lane == 0) cam_cnt[c] = 0;
  }
}

// q_r = H v_r - q_r for NR vectors (v_r = v + r vstride, q_r = q + r n9): finishes S v_r
__global__ void __launch_bounds__(256)
k_defl_hq(int64_t n9, int nr, const double* __restrict__ H, const double* __restrict__ v, int64_t vstride,
          double* __restrict__ q) {
  const int64_t e = blockIdx.x * (int64_t)256 + threadIdx.x;
  if (e >= n9 * nr) return;
  const int64_t r = e / n9, i = e - r * n9;
  q[e] = row9(H, v + r * vstride, i) - q[e];
}

// ---------------------------------------------------------------------------------------------
// Deflation vectors: the coarse space is [P | Z] with Z (n9 x kz, column-major,
// Euclidean-orthonormal) built from Ritz vectors harvested from the PCG solves themselves.
// ---------------------------------------------------------------------------------------------
// Ac[(mcl + i) * ld + col] = Z_i . q  (column `col` of the Z rows of [P Z]' S [P Z]); one CTA per i
__global__ void __launch_bounds__(256)
k_defl_zdot(int64_t n9, const double* __restrict__ Zb, const double* __restrict__ q, double* __restrict__ Ac, int ld,
            int mcl, int col) {
  __shared__ double sh[8];
  const double* zi = Zb + (int64_t)blockIdx.x * n9;
  double t = 0.0;
  for (int64_t a = threadIdx.x; a < n9; a += 256) t += zi[a] * q[a];
  t = block_sum<256>(t, sh);
  if (threadIdx.x == 0) Ac[(int64_t)(mcl + blockIdx.x) * ld + col] = t;
}

// the Z-rows x P-columns block is the transpose of the P-rows x Z-columns block (filled by k_coarse_restrict)
__global__ void k_defl_symfill(int m, int mcl, double* __restrict__ Ac) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int kz = m - mcl;
  if (e >= kz * mcl) return;
  const int j = e / mcl, a = e - j * mcl;
  Ac[(mcl + j) * m + a] = Ac[a * m + mcl + j];
}

// Out (n9 x N, column-major) = In (n9 x K, column-major) * Cf (K x N, row-major), N <= 64: thread per row
constexpr int DG_N = 64, DG_TJ = 32;
__global__ void __launch_bounds__(128)
k_defl_gemm(int64_t n9, int K, int N, const double* __restrict__ In, const double* __restrict__ Cf,
            double* __restrict__ Out) {
  __shared__ double cs[DG_TJ][DG_N];
  const int64_t i = blockIdx.x * (int64_t)128 + threadIdx.x;
  double acc[DG_N];
#pragma unroll
  for (int c = 0; c < DG_N; ++c) acc[c] = 0.0;
  for (int j0 = 0; j0 < K; j0 += DG_TJ) {
    const int nj = min(DG_TJ, K - j0);
    __syncthreads();
    for (int e = threadIdx.x; e < DG_TJ * DG_N; e += 128) {
      const int jj = e / DG_N, c = e - jj * DG_N;
      cs[jj][c] = (jj < nj && c < N) ? Cf[(int64_t)(j0 + jj) * N + c] : 0.0;
    }
    __syncthreads();
    if (i < n9)
      for (int jj = 0; jj < nj; ++jj) {
        const double v = In[(int64_t)(j0 + jj) * n9 + i];
#pragma unroll
        for (int c = 0; c < DG_N; ++c) acc[c] += v * cs[jj][c];
      }
  }
  if (i < n9) {
#pragma unroll
    for (int c = 0; c < DG_N; ++c)
      if (c < N) Out[(int64_t)c * n9 + i] = acc[c];
  }
}

// G (N x N, row-major) = Y' Y for Y (n9 x N, column-major); one CTA per entry of the upper triangle
__global__ void __launch_bounds__(256)
k_defl_gram(int64_t n9, int N, const double* __restrict__ Y, double* __restrict__ G) {
  __shared__ double sh[8];
  const int a = blockIdx.x / N, b = blockIdx.x - a * N;
  if (a > b) return;
  const double *ya = Y + (int64_t)a * n9, *yb = Y + (int64_t)b * n9;
  double t = 0.0;
  for (int64_t e = threadIdx.x; e < n9; e += 256) t += ya[e] * yb[e];
  t = block_sum<256>(t, sh);
  if (threadIdx.x == 0) {
    G[a * N + b] = t;
    G[b * N + a] = t;
  }
}

// ---------------------------------------------------------------------------------------------
// Exact solve: the reduced camera system assembled explicitly,
//     S~ = D^-1 (U + lambda I - W (V + lambda I)^-1 W') D^-1,   D = sqrt(diag(U + lambda I))
// (the Jacobi scaling is the camera part of the reference's column normalisation, src/lma_aux.jl:143-178), for
// the dense Cholesky of ba_chol.cu.  Off-diagonal blocks: per point, every pair of its observations contributes
// Yh_k1 Yh_k2' with Yh_k = D_c^-1 (B_k'A_k) L_p, (V_p + lambda I)^-1 = L_p L_p'.  |Yh_k1[i] . Yh_k2[j]| <= 1 and so
// is every partial sum (Cauchy-Schwarz against diag(W V^-1 W') <= diag(U)), hence the sums are accumulated in
// 64-bit fixed point (scale 2^61: resolution 4e-19, no overflow): integer addition is associative, so the
// atomics -- and the allreduce over ranks -- give bit-identical results in any order.  Diagonal blocks come from
// the ordered FP64 sums of the camera pass (H - Cr).
// ---------------------------------------------------------------------------------------------
constexpr double EX_SCALE = 2305843009213693952.0;  // 2^61

// The fixed-point sums live in a PACKED lower-triangular layout of 128 x 128 tiles (tile (ti, tj), tj <= ti, at
// (ti (ti + 1) / 2 + tj) * 128 * 128, row-major inside): half the bytes to clear and to sum over the ranks.
__device__ __forceinline__ int64_t ex_packed(int64_t r, int64_t c) {
  const int64_t ti = r >> 7, tj = c >> 7;
  return ((ti * (ti + 1) / 2 + tj) << 14) + ((r & 127) << 7) + (c & 127);
}

// cd[i] = sqrt(U_ii + lambda) for the n9 camera unknowns
__global__ void __launch_bounds__(256)
k_exact_diag(int64_t n9, double lambda, const double* __restrict__ Ug, double* __restrict__ cd) {
  const int64_t i = blockIdx.x * (int64_t)256 + threadIdx.x;
  if (i >= n9) return;
  const int64_t c = i / 9;
  const int j = (int)(i - 9 * c);
  cd[i] = sqrt(Ug[c * NV + sym9(j, j)] + lambda);
}

// Yh_k (9 x 3, row-major, 27 doubles per observation) = D_c^-1 (B_k' A_k) L_p; thread per observation
__global__ void __launch_bounds__(PT_THREADS)
k_exact_y(const int32_t* __restrict__ cam_idx, const int32_t* __restrict__ pnt_idx, int64_t pnt0, int64_t nl,
          const double2* __restrict__ Jp, const double* __restrict__ Vinv, const double* __restrict__ cd,
          double* __restrict__ Yh) {
  __shared__ double st[PT_THREADS * 27];
  const int64_t k = blockIdx.x * (int64_t)PT_THREADS + threadIdx.x;
  if (k < nl) {
    const int64_t p = __ldg(pnt_idx + k) - pnt0;
    const int c = __ldg(cam_idx + k);
    const double* vi = Vinv + p * 6;
    // Cholesky of the symmetric 3 x 3 inverse: Vinv = L L'
    const double l00 = sqrt(vi[0]), l10 = vi[1] / l00, l20 = vi[2] / l00;
    const double l11 = sqrt(vi[3] - l10 * l10), l21 = (vi[4] - l20 * l10) / l11;
    const double l22 = sqrt(vi[5] - (l20 * l20 + l21 * l21));
    const double2 a0 = Jp[k], a1 = Jp[nl + k], a2 = Jp[2 * nl + k];
    double* o = st + threadIdx.x * 27;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const double2 b = Jp[(int64_t)(3 + i) * nl + k];
      const double e0 = b.x * a0.x + b.y * a0.y, e1 = b.x * a1.x + b.y * a1.y, e2 = b.x * a2.x + b.y * a2.y;
      const double di = 1.0 / __ldg(cd + (int64_t)c * 9 + i);
      o[i * 3 + 0] = ((e0 * l00 + e1 * l10) + e2 * l20) * di;
      o[i * 3 + 1] = (e1 * l11 + e2 * l21) * di;
      o[i * 3 + 2] = (e2 * l22) * di;
    }
  }
  __syncthreads();
  // coalesced store of the block's 128 x 27 doubles
  const int64_t base = blockIdx.x * (int64_t)PT_THREADS * 27;
  const int64_t lim = nl * 27;
  for (int e = threadIdx.x; e < PT_THREADS * 27; e += PT_THREADS)
    if (base + e < lim) Yh[base + e] = st[e];
}

// one warp per observation k1: its pairs (k1, k2), k2 earlier in the same point
__global__ void __launch_bounds__(256)
k_exact_assemble(const int32_t* __restrict__ pstart, const int32_t* __restrict__ pnt_idx, int64_t pnt0,
                 const int32_t* __restrict__ cam_idx, int64_t nl, const double* __restrict__ Yh,
                 unsigned long long* __restrict__ Sq) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = blockIdx.x * (int64_t)8 + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * 8;
  const int la = lane < 27 ? lane : 0;
  for (int64_t k1 = warp0; k1 < nl; k1 += nwarps) {
    const int64_t p = __ldg(pnt_idx + k1) - pnt0;
    const int k0 = __ldg(pstart + p);
    if (k0 >= k1) continue;
    const int c1 = __ldg(cam_idx + k1);
    const double y1 = __ldg(Yh + k1 * 27 + la);
    for (int64_t k2 = k0; k2 < k1; ++k2) {
      const int c2 = __ldg(cam_idx + k2);
      const double y2 = __ldg(Yh + k2 * 27 + la);
      const bool swap = c1 < c2;
      const double yr = swap ? y2 : y1, yc = swap ? y1 : y2;  // row / column camera
      const int cr = swap ? c2 : c1, cc = swap ? c1 : c2;
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int e = la + 27 * u, i = e / 9, j = e - 9 * i;
        double v = __shfl_sync(0xffffffffu, yr, i * 3) * __shfl_sync(0xffffffffu, yc, j * 3);
        v += __shfl_sync(0xffffffffu, yr, i * 3 + 1) * __shfl_sync(0xffffffffu, yc, j * 3 + 1);
        v += __shfl_sync(0xffffffffu, yr, i * 3 + 2) * __shfl_sync(0xffffffffu, yc, j * 3 + 2);
        bool live = lane < 27;
        if (cr == cc) {  // the same camera twice on one point (not in BAL files): symmetric part, lower triangle
          double v2 = __shfl_sync(0xffffffffu, yc, i * 3) * __shfl_sync(0xffffffffu, yr, j * 3);
          v2 += __shfl_sync(0xffffffffu, yc, i * 3 + 1) * __shfl_sync(0xffffffffu, yr, j * 3 + 1);
          v2 += __shfl_sync(0xffffffffu, yc, i * 3 + 2) * __shfl_sync(0xffffffffu, yr, j * 3 + 2);
          v += v2;
          live = live && i >= j;
        }
        if (live) {
          const long long q = __double2ll_rn(v * EX_SCALE);
          atomicAdd(Sq + ex_packed((int64_t)cr * 9 + i, (int64_t)cc * 9 + j), (unsigned long long)q);
        }
      }
    }
  }
}

// packed fixed point -> row-major double over the lower triangle, diagonal blocks added, identity on the padding
template <typename T>  // T = double: the exact solve; float: the matrix of the mixed-precision factor
__global__ void __launch_bounds__(256)
k_exact_finish(int64_t n9, int64_t cn, const double* __restrict__ H, const double* __restrict__ Cr,
               const double* __restrict__ cd, const long long* __restrict__ Sq, T* __restrict__ S) {
  const int64_t r = blockIdx.y;
  const int64_t c = blockIdx.x * (int64_t)256 + threadIdx.x;
  if (c > r || c >= cn) return;
  T* sp = S + r * cn + c;
  if (r >= n9) {
    *sp = (r == c) ? (T)1 : (T)0;
    return;
  }
  double v = -((double)Sq[ex_packed(r, c)] / EX_SCALE);
  const int64_t br = r / 9, bc = c / 9;
  if (br == bc) {
    const int i = (int)(r - 9 * br), j = (int)(c - 9 * bc);  // j <= i
    v += (H[br * 81 + i * 9 + j] - Cr[br * NV + sym9(j, i)]) / (cd[r] * cd[c]);
  }
  *sp = (T)v;
}

// out[i] = in[i] / cd[i] (i < n9), 0 on the padding
__global__ void __launch_bounds__(256)
k_exact_scale(int64_t n9, int64_t cn, const double* __restrict__ in, const double* __restrict__ cd,
              double* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)256 + threadIdx.x;
  if (i < cn) out[i] = (i < n9) ? in[i] / cd[i] : 0.0;
}

// x (+)= xs / cd: the unscaled solution (or its correction)
__global__ void __launch_bounds__(256)
k_exact_unscale(int64_t n9, const double* __restrict__ xs, const double* __restrict__ cd, double* __restrict__ x,
                int accumulate) {
  const int64_t i = blockIdx.x * (int64_t)256 + threadIdx.x;
  if (i < n9) x[i] = (accumulate ? x[i] : 0.0) + xs[i] / cd[i];
}

// r = b - q (q = S x from the matrix-free product), rs = r / cd for the next correction;
// scal[S_RZN] = sum r^2, scal[S_RZ0] = sum b^2, first step: scal[S_REL] = ||r|| / ||b|| (one CTA, fixed order)
__global__ void __launch_bounds__(RED_THREADS)
k_exact_resid(int64_t n9, int64_t cn, const double* __restrict__ b, const double* __restrict__ q,
              const double* __restrict__ cd, double* __restrict__ rs, double* __restrict__ scal, int first) {
  __shared__ double sh[RED_THREADS / 32];
  double r2 = 0.0, b2 = 0.0;
  for (int64_t i = threadIdx.x; i < cn; i += RED_THREADS) {
    double rv = 0.0;
    if (i < n9) {
      const double bi = b[i];
      rv = bi - q[i];
      r2 += rv * rv;
      b2 += bi * bi;
      rv /= cd[i];
    }
    rs[i] = rv;
  }
  r2 = block_sum<RED_THREADS>(r2, sh);
  b2 = block_sum<RED_THREADS>(b2, sh);
  if (threadIdx.x == 0) {
    scal[S_RZN] = r2;
    scal[S_RZ0] = b2;
    if (first) scal[S_REL] = sqrt(r2 / b2);  // residual of the direct solve, before any refinement
  }
}

// ---- mixed-precision exact solve: FP64 CG on the matrix-free S, preconditioned by the FP32 factor of the assembled,
// Jacobi-scaled S (M^-1 = D^-1 (L32 L32')^-1 D^-1).  The camera system has 9 ncams <= ~40 K rows here: the vector
// work of an iteration is two single-CTA kernels with fixed-order sums (identical on every rank).
// start: x = 0, r = b, rs = r / cd (zero on the padding); scal[S_RZ0] = b.b
__global__ void __launch_bounds__(RED_THREADS)
k_mixed_init(int64_t n9, int64_t cn, const double* __restrict__ b, const double* __restrict__ cd, double* __restrict__ x,
             double* __restrict__ r, double* __restrict__ rs, double* __restrict__ scal) {
  __shared__ double sh[RED_THREADS / 32];
  double b2 = 0.0;
  for (int64_t i = threadIdx.x; i < cn; i += RED_THREADS) {
    double v = 0.0;
    if (i < n9) {
      v = b[i];
      x[i] = 0.0;
      r[i] = v;
      b2 += v * v;
      v /= cd[i];
    }
    rs[i] = v;
  }
  b2 = block_sum<RED_THREADS>(b2, sh);
  if (threadIdx.x == 0) {
    scal[S_RZ0] = b2;
    scal[S_RZN] = b2;
    scal[S_REL] = 1.0;
    scal[S_ITERS] = 0.0;
  }
}
// after the sweeps (xs = (L L')^-1 rs): z = xs / cd, rz' = r.z, beta = first ? 0 : rz' / rz, p = z + beta p
__global__ void __launch_bounds__(RED_THREADS)
k_mixed_dir(int64_t n9, const double* __restrict__ xs, const double* __restrict__ cd, const double* __restrict__ r,
            double* __restrict__ p, double* __restrict__ scal, int first) {
  __shared__ double sh[RED_THREADS / 32];
  double rz = 0.0;
  for (int64_t i = threadIdx.x; i < n9; i += RED_THREADS) rz += r[i] * (xs[i] / cd[i]);
  rz = block_sum<RED_THREADS>(rz, sh);
  const double beta = first ? 0.0 : rz / scal[S_MRZ];
  for (int64_t i = threadIdx.x; i < n9; i += RED_THREADS) p[i] = xs[i] / cd[i] + (first ? 0.0 : beta * p[i]);
  __syncthreads();  // everybody has read the old r.z
  if (threadIdx.x == 0) {
    scal[S_MRZ] = rz;
    if (!(rz > 0.0) && scal[S_RZN] > 0.0) scal[S_MBAD] = 1.0;  // the preconditioner is not positive definite
  }
}
// after q = S p: alpha = r.z / p.q, x += alpha p, r -= alpha q, rs = r / cd; scal[S_REL] = ||r|| / ||b||
__global__ void __launch_bounds__(RED_THREADS)
k_mixed_update(int64_t n9, const double* __restrict__ p, const double* __restrict__ q, const double* __restrict__ cd,
               double* __restrict__ x, double* __restrict__ r, double* __restrict__ rs, double* __restrict__ scal) {
  __shared__ double sh[RED_THREADS / 32];
  double pq = 0.0;
  for (int64_t i = threadIdx.x; i < n9; i += RED_THREADS) pq += p[i] * q[i];
  pq = block_sum<RED_THREADS>(pq, sh);
  const double alpha = pq > 0.0 ? scal[S_MRZ] / pq : 0.0;
  double r2 = 0.0;
  for (int64_t i = threadIdx.x; i < n9; i += RED_THREADS) {
    x[i] += alpha * p[i];
    const double rv = r[i] - alpha * q[i];
    r[i] = rv;
    r2 += rv * rv;
    rs[i] = rv / cd[i];
  }
  r2 = block_sum<RED_THREADS>(r2, sh);
  if (threadIdx.x == 0) {
    scal[S_PQ] = pq;
    scal[S_RZN] = r2;
    scal[S_REL] = sqrt(r2 / scal[S_RZ0]);
    scal[S_ITERS] += 1.0;
    if (!(pq > 0.0)) scal[S_MBAD] = 1.0;
  }
}

// out[i] = src[idx[i]]
__global__ void __launch_bounds__(256)
k_gather_i32(int64_t n, const int32_t* __restrict__ idx, const int32_t* __restrict__ src, int32_t* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)256 + threadIdx.x;
  if (i < n) out[i] = __ldg(src + __ldg(idx + i));
}

__global__ void k_seq_inc(unsigned long long* seqp) { *seqp += 1; }

__global__ void k_copy_cam_delta(int64_t n9, const double* __restrict__ xc, double* __restrict__ delta_c) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n9) delta_c[i] = xc[i];
}

}  // namespace ba
