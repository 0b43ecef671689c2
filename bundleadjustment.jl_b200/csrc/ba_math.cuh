// ba_math.cuh -- per-camera precompute and per-observation BAL block, FP64, sm_100a.
//
// Reference semantics (file:line in CelestineAngla/BundleAdjustment.jl):
//   residual  : src/BALNLPModels.jl:11-36 (projection!), :115-122 (cons!)
//   Jacobian  : src/BALNLPModels.jl:161-206 (jac_coord!) with src/JacobianByHand.jl:5-101
//               denseJ(2x12) = (JP3*JP2)*JP1, columns [X(3) | r(3) | t(3) | k1 k2 f], NaN -> 0.
//
// B200-first formulation (not a transcription).  Everything that depends only on the camera is
// computed once per camera (K1) into one 128-byte, 128-byte-aligned record (exactly one L2 line,
// four sectors):
//     [ k(3) | c | s | a | g | t(3) | k1 k2 f | 3 pad ]      c = cos th, s = sin th,
//                                                            a = s/th,  g = (1-c)/th,  k = r/th
// With R = c I + (1-c) k k^T + s [k]x (Rodrigues) and N = a I + (1-a) k k^T + g [k]x (the left
// Jacobian of SO(3)) the rotation part of the reference's closed form is
//     dP1/dr = -[R X]x N          (src/JacobianByHand.jl:41-57; verified to <= 1e-13 relative)
// and neither matrix is ever formed: R^T v and N^T v are applied as vector formulas.  Per
// observation this is ~200 FP64 operations and one reciprocal instead of ~500 flops + sincos +
// sqrt + 6 divisions, which keeps the evaluation pass HBM-bound on B200.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ba {

constexpr int CAM_REC = 16;  // doubles per camera record (128 B)
enum { CK0 = 0, CK1 = 1, CK2 = 2, CC = 3, CS = 4, CA = 5, CG = 6, CT0 = 7, CT1 = 8, CT2 = 9, CK1D = 10,
       CK2D = 11, CF = 12 };

// K1: camera parameters c9 = (r, t, k1, k2, f) -> record.  theta == 0 yields NaNs, exactly like
// the reference, which has no small-angle branch (src/BALNLPModels.jl:19-25).
__device__ __forceinline__ void cam_precompute(const double* __restrict__ c9, double* __restrict__ rec) {
  const double r0 = c9[0], r1 = c9[1], r2 = c9[2];
  const double th = sqrt((r0 * r0 + r1 * r1) + r2 * r2);
  double s, c;
  sincos(th, &s, &c);
  rec[CK0] = r0 / th;
  rec[CK1] = r1 / th;
  rec[CK2] = r2 / th;
  rec[CC] = c;
  rec[CS] = s;
  rec[CA] = s / th;
  rec[CG] = (1.0 - c) / th;
  rec[CT0] = c9[3];
  rec[CT1] = c9[4];
  rec[CT2] = c9[5];
  rec[CK1D] = c9[6];
  rec[CK2D] = c9[7];
  rec[CF] = c9[8];
  rec[13] = rec[14] = rec[15] = 0.0;
}

// One observation's block.  F is the projection minus the observed pixel (cons!).
struct ObsBlock {
  double F[2];
  double A[6];   // dF/dX   2x3 row-major
  double B[18];  // dF/dC   2x9 row-major, columns [r(3) t(3) k1 k2 f]
};

__device__ __forceinline__ bool nonfinite(double v) {
  return (__double2hiint(v) & 0x7ff00000) == 0x7ff00000;
}
__device__ __forceinline__ double nan0(double v) { return (v != v) ? 0.0 : v; }

// First half of the projection: Y = R X (Rodrigues, same operation order as projection!,
// src/BALNLPModels.jl:19-27), P1 = Y + t, P2 = -P1.xy / P1.z, distortion (:11-14, :29-30).
struct Proj {
  double Y0, Y1, Y2, iz, u, v, n2, sg, fs;
};
__device__ __forceinline__ void project_core(const double X[3], const double* __restrict__ cam, Proj& q) {
  const double k0 = cam[CK0], k1 = cam[CK1], k2 = cam[CK2], c = cam[CC], s = cam[CS];
  const double w0 = k1 * X[2] - k2 * X[1], w1 = k2 * X[0] - k0 * X[2], w2 = k0 * X[1] - k1 * X[0];  // k x X
  const double d = (k0 * X[0] + k1 * X[1]) + k2 * X[2];
  const double e = (1.0 - c) * d;
  q.Y0 = (c * X[0] + s * w0) + e * k0;
  q.Y1 = (c * X[1] + s * w1) + e * k1;
  q.Y2 = (c * X[2] + s * w2) + e * k2;
  const double px = q.Y0 + cam[CT0], py = q.Y1 + cam[CT1], pz = q.Y2 + cam[CT2];
  q.iz = 1.0 / pz;
  q.u = -px * q.iz;
  q.v = -py * q.iz;
  q.n2 = q.u * q.u + q.v * q.v;
  q.sg = (1.0 + cam[CK1D] * q.n2) + cam[CK2D] * (q.n2 * q.n2);
  q.fs = cam[CF] * q.sg;
}

// Residual only (trial-point and cons! kernels); bit-identical to ObsBlock::F of eval_block on
// finite inputs because both go through project_core.
__device__ __forceinline__ void eval_residual(const double X[3], const double* __restrict__ cam,
                                              double ox, double oy, double F[2]) {
  Proj q;
  project_core(X, cam, q);
  F[0] = q.fs * q.u - ox;
  F[1] = q.fs * q.v - oy;
}

// Faithful dense path: reproduces the reference's NaN/Inf propagation through the dense
// products (JP3*JP2)*JP1 including their structural zeros (src/BALNLPModels.jl:197).  Only taken
// when the fast path produced a non-finite value, so its cost does not matter.
// Arguments and result travel by value (param space): no generic pointers into a caller's stack.
struct DenseIn {
  double X[3];
  double cam[13];
  double ox, oy;
};
static __device__ __noinline__ ObsBlock eval_block_dense(const DenseIn in) {
  const double* X = in.X;
  const double* cam = in.cam;
  const double ox = in.ox, oy = in.oy;
  ObsBlock o;
  const double kx = cam[CK0], ky = cam[CK1], kz = cam[CK2], c = cam[CC], s = cam[CS];
  const double a = cam[CA], g = cam[CG], b = 1.0 - a, omc = 1.0 - c;
  // R: the entries of JP1[1:3,1:3] (src/JacobianByHand.jl:38-40,45-47,52-54); N likewise
  const double R[9] = {c + omc * (kx * kx),        -s * kz + (omc * ky) * kx, s * ky + (omc * kz) * kx,
                       s * kz + (omc * ky) * kx,   c + omc * (ky * ky),       -s * kx + (omc * ky) * kz,
                       -s * ky + (omc * kx) * kz,  s * kx + (omc * ky) * kz,  c + omc * (kz * kz)};
  const double N[9] = {a + b * (kx * kx),        (b * kx) * ky - g * kz, (b * kx) * kz + g * ky,
                       (b * ky) * kx + g * kz,   a + b * (ky * ky),      (b * ky) * kz - g * kx,
                       (b * kz) * kx - g * ky,   (b * kz) * ky + g * kx, a + b * (kz * kz)};
  const double Y[3] = {R[0] * X[0] + R[1] * X[1] + R[2] * X[2], R[3] * X[0] + R[4] * X[1] + R[5] * X[2],
                       R[6] * X[0] + R[7] * X[1] + R[8] * X[2]};
  const double p[3] = {Y[0] + cam[CT0], Y[1] + cam[CT1], Y[2] + cam[CT2]};
  const double k1 = cam[CK1D], k2 = cam[CK2D], f = cam[CF];
  double JP1[6][12], JP2[5][6], JP3[2][5], T[2][6];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 12; ++j) JP1[i][j] = 0.0;
  for (int i = 0; i < 5; ++i)
    for (int j = 0; j < 6; ++j) JP2[i][j] = 0.0;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) JP1[i][j] = R[3 * i + j];
    JP1[i][6 + i] = 1.0;
    JP1[3 + i][9 + i] = 1.0;
  }
  // dP1/dr = -[Y]x N : column j = N_col_j x Y
  for (int j = 0; j < 3; ++j) {
    const double n0 = N[j], n1 = N[3 + j], n2 = N[6 + j];
    JP1[0][3 + j] = n1 * Y[2] - n2 * Y[1];
    JP1[1][3 + j] = n2 * Y[0] - n0 * Y[2];
    JP1[2][3 + j] = n0 * Y[1] - n1 * Y[0];
  }
  JP2[2][3] = JP2[3][4] = JP2[4][5] = 1.0;
  double u, v;
  if (p[2] == 0.0) {  // JacobianByHand.jl:15-18, :67-69: NaN marker poisons the whole block
    JP2[0][0] = __longlong_as_double(0x7ff8000000000000LL);
    u = JP2[0][0] * p[0];
    v = JP2[0][0] * p[1];
  } else {
    JP2[0][0] = -1.0 / p[2];
    JP2[0][2] = p[0] / (p[2] * p[2]);
    JP2[1][1] = JP2[0][0];
    JP2[1][2] = p[1] / (p[2] * p[2]);
    u = (-p[0]) / p[2];
    v = (-p[1]) / p[2];
  }
  const double n2 = u * u + v * v, n4 = n2 * n2;
  const double sg = (1.0 + k1 * n2) + k2 * n4;
  const double aa = (2 * k1) * u + k2 * (4 * ((u * u) * u) + (4 * u) * (v * v));
  const double bb = (2 * k1) * v + k2 * (4 * ((v * v) * v) + (4 * v) * (u * u));
  JP3[0][0] = f * sg + (f * aa) * u; JP3[0][1] = (f * bb) * u; JP3[0][2] = (f * n2) * u;
  JP3[0][3] = (f * n4) * u;          JP3[0][4] = sg * u;
  JP3[1][0] = (f * aa) * v;          JP3[1][1] = f * sg + (f * bb) * v; JP3[1][2] = (f * n2) * v;
  JP3[1][3] = (f * n4) * v;          JP3[1][4] = sg * v;
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 6; ++j) {
      double acc = 0.0;
      for (int q = 0; q < 5; ++q) acc += JP3[i][q] * JP2[q][j];
      T[i][j] = acc;
    }
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 12; ++j) {
      double acc = 0.0;
      for (int q = 0; q < 6; ++q) acc += T[i][q] * JP1[q][j];
      const double val = (acc != acc) ? 0.0 : acc;  // per-entry NaN -> 0 (BALNLPModels.jl:201)
      if (j < 3) o.A[3 * i + j] = val; else o.B[9 * i + (j - 3)] = val;
    }
  // the residual keeps its NaN/Inf (BALNLPModels.jl:120 is commented out)
  const double uu = (-p[0]) / p[2], vv = (-p[1]) / p[2];
  const double m2 = uu * uu + vv * vv;
  const double fs = f * ((1.0 + k1 * m2) + k2 * (m2 * m2));
  o.F[0] = fs * uu - ox;
  o.F[1] = fs * vv - oy;
  return o;
}

// out = M^T v for M = alpha I + beta k k^T + gamma [k]x  (so M^T v = alpha v + beta (k.v) k - gamma k x v)
__device__ __forceinline__ void apply_T(double alpha, double beta, double gamma, double k0, double k1, double k2,
                                        double v0, double v1, double v2, double& o0, double& o1, double& o2) {
  const double d = beta * ((k0 * v0 + k1 * v1) + k2 * v2);
  o0 = (alpha * v0 + d * k0) - gamma * (k1 * v2 - k2 * v1);
  o1 = (alpha * v1 + d * k1) - gamma * (k2 * v0 - k0 * v2);
  o2 = (alpha * v2 + d * k2) - gamma * (k0 * v1 - k1 * v0);
}

// Fast path: residual + 2x12 Jacobian block from the camera record.  WANT_A = false skips the point part
// A = G R (callers that only need the camera part B; B is then bit-identical to the full evaluation, and a
// non-finite A implies a non-finite G, which the t-columns of B carry, so the fallback triggers the same way).
template <bool WANT_A = true>
__device__ __forceinline__ void eval_block(const double X[3], const double* __restrict__ cam,
                                           double ox, double oy, ObsBlock& o) {
  Proj q;
  project_core(X, cam, q);
  const double Y0 = q.Y0, Y1 = q.Y1, Y2 = q.Y2, iz = q.iz, u = q.u, v = q.v, n2 = q.n2, sg = q.sg, fs = q.fs;
  const double kd1 = cam[CK1D], kd2 = cam[CK2D], f = cam[CF];
  const double k0 = cam[CK0], k1 = cam[CK1], k2 = cam[CK2], c = cam[CC], s = cam[CS], a = cam[CA], g = cam[CG];
  const double n4 = n2 * n2;
  o.F[0] = fs * u - ox;
  o.F[1] = fs * v - oy;
  // JP3 (2x5): d(f*sg*P2)/d(P2x,P2y,k1,k2,f); 4u^3+4uv^2 = 4 u n2
  const double e = f * (2.0 * kd1 + 4.0 * kd2 * n2);
  const double j00 = fs + (e * u) * u, j01 = (e * v) * u, j11 = fs + (e * v) * v;
  const double fn2 = f * n2, fn4 = f * n4;
  // G = JP3[:,0:2] * JP2[0:2,0:3] with JP2 = -iz [1 0 u; 0 1 v]
  const double g00 = -iz * j00, g01 = -iz * j01, g02 = -iz * (j00 * u + j01 * v);
  const double g10 = g01, g11 = -iz * j11, g12 = -iz * (j01 * u + j11 * v);
  // A = G R : row i = (R^T g_i)^T
  if (WANT_A) {
    apply_T(c, 1.0 - c, s, k0, k1, k2, g00, g01, g02, o.A[0], o.A[1], o.A[2]);
    apply_T(c, 1.0 - c, s, k0, k1, k2, g10, g11, g12, o.A[3], o.A[4], o.A[5]);
  }
  // dF/dr = G (-[Y]x N) : row i = (N^T (Y x g_i))^T
  apply_T(a, 1.0 - a, g, k0, k1, k2, Y1 * g02 - Y2 * g01, Y2 * g00 - Y0 * g02, Y0 * g01 - Y1 * g00,
          o.B[0], o.B[1], o.B[2]);
  apply_T(a, 1.0 - a, g, k0, k1, k2, Y1 * g12 - Y2 * g11, Y2 * g10 - Y0 * g12, Y0 * g11 - Y1 * g10,
          o.B[9], o.B[10], o.B[11]);
  o.B[3] = g00; o.B[4] = g01; o.B[5] = g02;
  o.B[12] = g10; o.B[13] = g11; o.B[14] = g12;
  o.B[6] = fn2 * u; o.B[7] = fn4 * u; o.B[8] = sg * u;
  o.B[15] = fn2 * v; o.B[16] = fn4 * v; o.B[17] = sg * v;
  // Any NaN/Inf in the reference's dense products shows up in at least one output entry
  // (t-columns carry G, k1/k2/f-columns carry JP3, X/r-columns carry R and N).  Rare: redo
  // the block the reference's way so zeros, NaN -> 0 and surviving Infs land identically.
  bool bad = false;
  if (WANT_A) {
#pragma unroll
    for (int i = 0; i < 6; ++i) bad |= nonfinite(o.A[i]);
  }
#pragma unroll
  for (int i = 0; i < 18; ++i) bad |= nonfinite(o.B[i]);
  if (bad) {
    DenseIn in;
    in.X[0] = X[0]; in.X[1] = X[1]; in.X[2] = X[2];
#pragma unroll
    for (int i = 0; i < 13; ++i) in.cam[i] = cam[i];
    in.ox = ox;
    in.oy = oy;
    o = eval_block_dense(in);
  }
}

// ---- small warp helpers ------------------------------------------------------------------
__device__ __forceinline__ double shfl_up_d(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += shfl_xor_d(v, m);
  return v;
}

// Warp-cooperative gather of the 32 lanes' camera records (128 B each) into shared memory: every
// load instruction moves four whole records as 4 x 128 contiguous bytes (4 L1 wavefronts instead of
// 32 for a per-lane strided read), then each lane reads its own record.  Rows are padded to 9
// double2 (144 B, odd multiple of 16 B) so the per-lane 16-byte reads are bank-conflict free.
constexpr int CAM_ROW2 = 9;  // double2 per staged camera row
// COHERENT: read the records with ld.global.cg instead of the non-coherent path.  Needed when the kernel
// runs as a programmatic dependent launch of the kernel that wrote them (k_eval after k_cam_precompute):
// ld.global.nc may be hoisted above griddepcontrol.wait, because the compiler takes nc data to be constant
// for the whole kernel.
template <bool COHERENT = false>
__device__ __forceinline__ void warp_stage_cams(const double* camtab, int cam_of_lane, int lane,
                                                double2* __restrict__ rows /* 32 * CAM_ROW2 */) {
  const int sub = lane >> 3, ch = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = 4 * i + sub;
    const int c = __shfl_sync(0xffffffffu, cam_of_lane, r);
    const double2* src = reinterpret_cast<const double2*>(camtab + (int64_t)c * CAM_REC) + ch;
    double2 t;
    if (COHERENT) {
#if !defined(EVAL_CAM_LDCA) || EVAL_CAM_LDCA  // default on: +1.5 % on Venice-1778 (profiles/r02_k_eval_ab.md)
      // cached in L1 (popular cameras are re-read by many warps of an SM); a volatile asm with a memory clobber
      // stays below the volatile griddepcontrol.wait, and L1 holds nothing of camtab from before the wait
      asm volatile("ld.global.ca.v2.f64 {%0, %1}, [%2];" : "=d"(t.x), "=d"(t.y) : "l"(src) : "memory");
#else
      t = __ldcg(src);
#endif
    } else {
      t = __ldg(src);
    }
    rows[r * CAM_ROW2 + ch] = t;
  }
  __syncwarp();
}
__device__ __forceinline__ void read_staged_cam(const double2* __restrict__ rows, int lane, double* cam) {
#pragma unroll
  for (int i = 0; i < 7; ++i) {  // 13 doubles used: 7 double2
    const double2 t = rows[lane * CAM_ROW2 + i];
    cam[2 * i] = t.x;
    cam[2 * i + 1] = t.y;
  }
}

}  // namespace ba
