// ba_chol.cu -- blocked FP64 Cholesky A = L L' of the (explicitly assembled, Jacobi-scaled) reduced camera
// system, and the two triangular sweeps; hand-written for sm_100a.
//
// Reference counterpart: the numeric factorisation and the L / D / L' sweeps of src/ldl_aux.jl:122-201,4-42 on
// the camera block (after the ordering has eliminated residual rows and points), SURVEY.md section 8 row f2.
//
// Layout: A is cn x cn row-major (cn a multiple of 128), lower triangle live.  Right-looking by 128-column panels:
//   potrf  (1 CTA)        : the 128 x 128 diagonal block in shared memory (16-column sub-panels), plus its
//                           inverse Linv_kk (in-place triangular inversion), kept for the panel solve and
//                           the substitution sweeps;
//   trsm   (1 CTA / tile) : P_i <- P_i Linv_kk'                       (128 x 128 x 128 product)
//   syrk   (1 CTA / tile) : A_ij <- A_ij - P_i P_j'  for k < j <= i   (the n^3/3 flops)
// The two products are one micro-kernel: 128 x 128 x 128 C = A B' with both operands k-contiguous, staged
// through shared memory by cp.async (3 stages of 16 k), FP64 tensor-core MMAs (mma.sync.m8n8k4.f64 -- DMMA in
// SASS; tcgen05 has no f64 kind), 8 warps x (64 x 32) accumulators in registers.  Look-ahead: the update of
// the next panel's column is launched first, then the next potrf + trsm run on a side stream under the rest
// of the trailing update.
// Roofline: n^3/3 flops against the FP64 peak (ba_measure_fp64_peak; DMMA and DFMA peaks coincide on B200).
#include <algorithm>
#include <cstdlib>
#include "ba_internal.h"
#include "ba_chol.h"

namespace ba {
namespace {

constexpr int CT = CHOL_TILE;     // 128
constexpr int KC = 16;            // k per pipeline stage
constexpr int LDSM = KC + 4;      // padded row stride (doubles): fragment loads are bank-conflict free
constexpr int STAGES = 3;
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_SMEM = STAGES * 2 * CT * LDSM * (int)sizeof(double);  // 122880 B

__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// acc = A (128 x 128, rows lda apart) * B' (B: 128 x 128, rows ldb apart); accumulator fragment layout of
// m8n8k4: warp (wm, wn) of 2 x 4 owns rows wm*64.., columns wn*32..; tile (mi, ni): lane holds row 8 mi + lane/4,
// columns 8 ni + 2 (lane%4) + {0, 1}.
__device__ __forceinline__ void tile_abt(const double* __restrict__ A, int64_t lda, const double* __restrict__ B,
                                         int64_t ldb, double (&acc)[8][4][2], double* sm) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mi = 0; mi < 8; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
  auto load_stage = [&](int stage, int kc) {
    double* As = sm + stage * (2 * CT * LDSM);
    double* Bs = As + CT * LDSM;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = tid + i * GEMM_THREADS;  // 1024 16-byte pieces per operand
      const int row = c >> 3, part = c & 7;
      cp_async16(As + row * LDSM + part * 2, A + (int64_t)row * lda + kc * KC + part * 2);
      cp_async16(Bs + row * LDSM + part * 2, B + (int64_t)row * ldb + kc * KC + part * 2);
    }
  };
  constexpr int NK = CT / KC;  // 8
  load_stage(0, 0);
  cp_async_commit();
  load_stage(1, 1);
  cp_async_commit();
#pragma unroll 1
  for (int kc = 0; kc < NK; ++kc) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();  // stage kc has landed for everybody; everybody is done with stage kc - 1
    if (kc + STAGES - 1 < NK) load_stage((kc + STAGES - 1) % STAGES, kc + STAGES - 1);
    cp_async_commit();
    const double* As = sm + (kc % STAGES) * (2 * CT * LDSM) + (wm * 64 + g) * LDSM + t;
    const double* Bs = sm + (kc % STAGES) * (2 * CT * LDSM) + CT * LDSM + (wn * 32 + g) * LDSM + t;
#pragma unroll
    for (int ks = 0; ks < KC / 4; ++ks) {
      double a[8], b[4];
#pragma unroll
      for (int mi = 0; mi < 8; ++mi) a[mi] = As[mi * 8 * LDSM + ks * 4];
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) b[ni] = Bs[ni * 8 * LDSM + ks * 4];
#pragma unroll
      for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) dmma8x8x4(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();
}

// trailing update: tile (i, j) = (j0 + blockIdx.y + blockIdx.x, j0 + blockIdx.y), i < nb: A_ij -= P_i P_j'
__global__ void __launch_bounds__(GEMM_THREADS, 1)
k_chol_syrk(double* __restrict__ A, int64_t ld, int k, int j0, int nb) {
  extern __shared__ __align__(16) double sm[];
  const int j = j0 + blockIdx.y, i = j + blockIdx.x;
  if (i >= nb) return;
  const double* Pi = A + (int64_t)i * CT * ld + (int64_t)k * CT;
  const double* Pj = A + (int64_t)j * CT * ld + (int64_t)k * CT;
  double acc[8][4][2];
  tile_abt(Pi, ld, Pj, ld, acc, sm);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp >> 2, wn = warp & 3, g = lane >> 2, t = lane & 3;
  double* C = A + ((int64_t)i * CT + wm * 64 + g) * ld + (int64_t)j * CT + wn * 32 + 2 * t;
#pragma unroll
  for (int mi = 0; mi < 8; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      double2* p = reinterpret_cast<double2*>(C + (int64_t)mi * 8 * ld + ni * 8);
      double2 c = *p;
      c.x -= acc[mi][ni][0];
      c.y -= acc[mi][ni][1];
      *p = c;
    }
}

// panel solve: P_i <- P_i Linv_kk' for the row tiles i = k + 1 + blockIdx.x
__global__ void __launch_bounds__(GEMM_THREADS, 1)
k_chol_trsm(double* __restrict__ A, int64_t ld, int k, const double* __restrict__ Dinv) {
  extern __shared__ __align__(16) double sm[];
  const int i = k + 1 + blockIdx.x;
  double* Pi = A + (int64_t)i * CT * ld + (int64_t)k * CT;
  double acc[8][4][2];
  tile_abt(Pi, ld, Dinv + (int64_t)k * CT * CT, CT, acc, sm);  // (every read of P_i is complete on return)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp >> 2, wn = warp & 3, g = lane >> 2, t = lane & 3;
  double* C = Pi + (int64_t)(wm * 64 + g) * ld + wn * 32 + 2 * t;
#pragma unroll
  for (int mi = 0; mi < 8; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
      *reinterpret_cast<double2*>(C + (int64_t)mi * 8 * ld + ni * 8) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
}

// ---- diagonal block -------------------------------------------------------------------------------------
constexpr int PO_THREADS = 512;
constexpr int PO_LD = CT + 1;  // padded row stride in shared memory
constexpr int PO_SB = 16;      // sub-panel width
constexpr int PO_SMEM = (CT * PO_LD + CT) * (int)sizeof(double);

// A_kk (lower) <- L_kk, Dinv[k] <- L_kk^-1.  info: first non-positive pivot (1-based global index), else untouched.
__global__ void __launch_bounds__(PO_THREADS, 1)
k_chol_potrf(double* __restrict__ A, int64_t ld, int k, double* __restrict__ Dinv, int* __restrict__ info) {
  extern __shared__ __align__(16) double smp[];
  double* a = smp;                  // 128 x 129
  double* col = smp + CT * PO_LD;   // 128 scratch
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* Akk = A + (int64_t)k * CT * ld + (int64_t)k * CT;
  for (int e = tid; e < CT * CT; e += PO_THREADS) {
    const int r = e >> 7, c = e & 127;
    a[r * PO_LD + c] = (c <= r) ? Akk[(int64_t)r * ld + c] : 0.0;
  }
  __syncthreads();
  for (int j0 = 0; j0 < CT; j0 += PO_SB) {
    // (a) 16 x 16 diagonal sub-block, one warp, lane = row
    if (warp == 0) {
      const int r = j0 + (lane & 15);
      for (int j = 0; j < PO_SB; ++j) {
        const double djj = a[(j0 + j) * PO_LD + j0 + j];
        if (!(djj > 0.0) && lane == 0) atomicCAS(info, 0, k * CT + j0 + j + 1);
        const double d = sqrt(djj);
        __syncwarp();
        if (lane < 16) {
          if (lane == j) a[r * PO_LD + j0 + j] = d;
          else if (lane > j) a[r * PO_LD + j0 + j] /= d;
        }
        __syncwarp();
        if (lane < 16 && lane > j) {
          const double lrj = a[r * PO_LD + j0 + j];
          for (int c = j + 1; c <= lane; ++c) a[r * PO_LD + j0 + c] -= lrj * a[(j0 + c) * PO_LD + j0 + j];
        }
        __syncwarp();
      }
    }
    __syncthreads();
    // (b) rows below: forward substitution against the 16 x 16 factor, one thread per row
    const int nbelow = CT - j0 - PO_SB;
    if (tid < nbelow) {
      double* row = a + (j0 + PO_SB + tid) * PO_LD + j0;
      double x[PO_SB];
#pragma unroll
      for (int j = 0; j < PO_SB; ++j) {
        double s = row[j];
#pragma unroll
        for (int q = 0; q < j; ++q) s -= x[q] * a[(j0 + j) * PO_LD + j0 + q];
        x[j] = s / a[(j0 + j) * PO_LD + j0 + j];
      }
#pragma unroll
      for (int j = 0; j < PO_SB; ++j) row[j] = x[j];
    }
    __syncthreads();
    // (c) trailing update of the lower triangle: a[r][c] -= sum_q a[r][j0+q] a[c][j0+q]
    const int base = j0 + PO_SB;
    for (int e = tid; e < nbelow * nbelow; e += PO_THREADS) {
      const int rr = e / nbelow, cc = e - rr * nbelow;
      if (cc > rr) continue;
      const double* pr = a + (base + rr) * PO_LD + j0;
      const double* pc = a + (base + cc) * PO_LD + j0;
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < PO_SB; ++q) s += pr[q] * pc[q];
      a[(base + rr) * PO_LD + base + cc] -= s;
    }
    __syncthreads();
  }
  for (int e = tid; e < CT * CT; e += PO_THREADS) {
    const int r = e >> 7, c = e & 127;
    if (c <= r) Akk[(int64_t)r * ld + c] = a[r * PO_LD + c];
  }
  __syncthreads();
  // in-place inverse of the lower-triangular factor, last column first:
  //   x_jj = 1 / l_jj;  x[j+1:, j] = -x_jj * Xinv[j+1:, j+1:] l[j+1:, j]
  for (int j = CT - 1; j >= 0; --j) {
    const double xjj = 1.0 / a[j * PO_LD + j];
    if (tid > j && tid < CT) col[tid] = a[tid * PO_LD + j];
    __syncthreads();
    // row i in (j, 128): sum_{q = j+1..i} X[i][q] col[q]; 4 threads per row
    {
      const int i = j + 1 + (tid >> 2), part = tid & 3;
      double s = 0.0;
      if (i < CT)
        for (int q = j + 1 + part; q <= i; q += 4) s += a[i * PO_LD + q] * col[q];
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (i < CT && part == 0) a[i * PO_LD + j] = -xjj * s;
    }
    if (tid == 0) a[j * PO_LD + j] = xjj;
    __syncthreads();
  }
  double* D = Dinv + (int64_t)k * CT * CT;
  for (int e = tid; e < CT * CT; e += PO_THREADS) {
    const int r = e >> 7, c = e & 127;
    D[e] = (c <= r) ? a[r * PO_LD + c] : 0.0;
  }
}

// ---- substitution sweeps ----------------------------------------------------------------------------------
constexpr int SV_THREADS = 256;

// step k of L y = b: every CTA forms y_k = Linv_kk w_k; CTA 0 stores it, CTA c > 0 updates row tile i = k + c:
// w_i -= L_ik y_k
__global__ void __launch_bounds__(SV_THREADS)
k_chol_fwd(const double* __restrict__ L, int64_t ld, const double* __restrict__ Dinv, double* __restrict__ w,
           double* __restrict__ y, int k) {
  __shared__ double wk[CT], yk[CT];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < CT) wk[tid] = w[(int64_t)k * CT + tid];
  __syncthreads();
  const double* D = Dinv + (int64_t)k * CT * CT;
  for (int r = warp; r < CT; r += SV_THREADS / 32) {  // warp per row, lanes across the row
    double s = 0.0;
#pragma unroll
    for (int u = 0; u < 4; ++u) s += D[r * CT + lane + 32 * u] * wk[lane + 32 * u];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) yk[r] = s;
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    if (tid < CT) y[(int64_t)k * CT + tid] = yk[tid];
    return;
  }
  const int i = k + blockIdx.x;
  const double* T = L + (int64_t)i * CT * ld + (int64_t)k * CT;
  for (int r = warp; r < CT; r += SV_THREADS / 32) {
    double s = 0.0;
#pragma unroll
    for (int u = 0; u < 4; ++u) s += T[(int64_t)r * ld + lane + 32 * u] * yk[lane + 32 * u];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) w[(int64_t)i * CT + r] -= s;
  }
}

// step i of L' x = y (i descending): every CTA forms x_i = Linv_ii' y_i; CTA i stores it, CTA kk < i updates
// y_kk -= L_{i,kk}' x_i
__global__ void __launch_bounds__(SV_THREADS)
k_chol_bwd(const double* __restrict__ L, int64_t ld, const double* __restrict__ Dinv, double* __restrict__ y,
           double* __restrict__ x, int i) {
  __shared__ double yi[CT], xi[CT], half[CT];
  const int tid = threadIdx.x;
  if (tid < CT) yi[tid] = y[(int64_t)i * CT + tid];
  __syncthreads();
  const double* D = Dinv + (int64_t)i * CT * CT;
  const int c = tid & (CT - 1), h = tid >> 7;  // column, half of the rows
  {
    double s = 0.0;
    for (int r = h * 64; r < h * 64 + 64; ++r) s += D[r * CT + c] * yi[r];
    if (h == 1) half[c] = s;
    __syncthreads();
    if (h == 0) xi[c] = s + half[c];
    __syncthreads();
  }
  const int kk = blockIdx.x;
  if (kk == i) {
    if (tid < CT) x[(int64_t)i * CT + tid] = xi[tid];
    return;
  }
  const double* T = L + (int64_t)i * CT * ld + (int64_t)kk * CT;
  double s = 0.0;
  for (int r = h * 64; r < h * 64 + 64; ++r) s += T[(int64_t)r * ld + c] * xi[r];
  if (h == 1) half[c] = s;
  __syncthreads();
  if (h == 0) y[(int64_t)kk * CT + c] -= s + half[c];
}

}  // namespace

int chol_plan_init(ba_handle* h, chol_plan& P, int64_t cn) {
  if (P.cn == cn && P.d_Dinv) return BA_OK;
  chol_plan_release(P);
  P.cn = cn;
  const int64_t nb = cn / CT;
  BA_CUDA(cudaMalloc(reinterpret_cast<void**>(&P.d_Dinv), sizeof(double) * (size_t)(nb * CT * CT)));
  BA_CUDA(cudaMalloc(reinterpret_cast<void**>(&P.d_y), sizeof(double) * (size_t)cn));
  BA_CUDA(cudaMalloc(reinterpret_cast<void**>(&P.d_w), sizeof(double) * (size_t)cn));
  BA_CUDA(cudaMalloc(reinterpret_cast<void**>(&P.d_info), sizeof(int)));
  BA_CUDA(cudaStreamCreateWithFlags(&P.side, cudaStreamNonBlocking));
  BA_CUDA(cudaEventCreateWithFlags(&P.ev_col, cudaEventDisableTiming));
  BA_CUDA(cudaEventCreateWithFlags(&P.ev_panel, cudaEventDisableTiming));
  BA_CUDA(cudaEventCreateWithFlags(&P.ev_join, cudaEventDisableTiming));
  // per device (a process may hold handles on several devices): set once per plan
  BA_CUDA(cudaFuncSetAttribute(k_chol_syrk, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
  BA_CUDA(cudaFuncSetAttribute(k_chol_trsm, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
  BA_CUDA(cudaFuncSetAttribute(k_chol_potrf, cudaFuncAttributeMaxDynamicSharedMemorySize, PO_SMEM));
  P.attrs_set = true;
  return BA_OK;
}

void chol_plan_release(chol_plan& P) {
  cudaFree(P.d_Dinv);
  cudaFree(P.d_Dinv32);
  cudaFree(P.d_y);
  cudaFree(P.d_w);
  cudaFree(P.d_info);
  if (P.solve_graph) cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(P.solve_graph));
  if (P.side) cudaStreamDestroy(P.side);
  if (P.ev_col) cudaEventDestroy(P.ev_col);
  if (P.ev_panel) cudaEventDestroy(P.ev_panel);
  if (P.ev_join) cudaEventDestroy(P.ev_join);
  P = chol_plan();
}

int chol_factor(ba_handle* h, chol_plan& P, double* A, cudaStream_t s, int* info_host) {
  const int64_t cn = P.cn, ld = cn;
  const int nb = (int)(cn / CT);
  static const bool no_lookahead = getenv("BAGPU_CHOL_NO_LOOKAHEAD") != nullptr;
  BA_CUDA(cudaMemsetAsync(P.d_info, 0, sizeof(int), s));
  k_chol_potrf<<<1, PO_THREADS, PO_SMEM, s>>>(A, ld, 0, P.d_Dinv, P.d_info);
  if (nb > 1) k_chol_trsm<<<nb - 1, GEMM_THREADS, GEMM_SMEM, s>>>(A, ld, 0, P.d_Dinv);
  for (int k = 0; k + 1 < nb; ++k) {
    // panel k is final in A[k+1:, k].  Column k+1 of the trailing matrix first ...
    k_chol_syrk<<<dim3(nb - k - 1, 1), GEMM_THREADS, GEMM_SMEM, s>>>(A, ld, k, k + 1, nb);
    const bool rest = k + 2 < nb;
    if (rest && !no_lookahead) {
      // ... then panel k+1 (potrf + trsm) on the side stream, under the rest of the update
      BA_CUDA(cudaEventRecord(P.ev_col, s));
      BA_CUDA(cudaStreamWaitEvent(P.side, P.ev_col, 0));
      k_chol_potrf<<<1, PO_THREADS, PO_SMEM, P.side>>>(A, ld, k + 1, P.d_Dinv, P.d_info);
      k_chol_trsm<<<nb - k - 2, GEMM_THREADS, GEMM_SMEM, P.side>>>(A, ld, k + 1, P.d_Dinv);
      BA_CUDA(cudaEventRecord(P.ev_panel, P.side));
      k_chol_syrk<<<dim3(nb - k - 2, nb - k - 2), GEMM_THREADS, GEMM_SMEM, s>>>(A, ld, k, k + 2, nb);
      BA_CUDA(cudaStreamWaitEvent(s, P.ev_panel, 0));
    } else {
      if (rest) k_chol_syrk<<<dim3(nb - k - 2, nb - k - 2), GEMM_THREADS, GEMM_SMEM, s>>>(A, ld, k, k + 2, nb);
      k_chol_potrf<<<1, PO_THREADS, PO_SMEM, s>>>(A, ld, k + 1, P.d_Dinv, P.d_info);
      if (rest) k_chol_trsm<<<nb - k - 2, GEMM_THREADS, GEMM_SMEM, s>>>(A, ld, k + 1, P.d_Dinv);
    }
  }
  BA_CUDA(cudaGetLastError());
  if (info_host) {
    BA_CUDA(cudaMemcpyAsync(info_host, P.d_info, sizeof(int), cudaMemcpyDeviceToHost, s));
    BA_CUDA(cudaStreamSynchronize(s));
    if (*info_host != 0) {
      h->err = "Cholesky of the reduced camera system: non-positive pivot";
      return BA_ERR_NUMERIC;
    }
  }
  return BA_OK;
}

int chol_solve(ba_handle* h, chol_plan& P, const double* L, const double* b, double* x, cudaStream_t s) {
  const int64_t cn = P.cn;
  const int nb = (int)(cn / CT);
  BA_CUDA(cudaMemcpyAsync(P.d_w, b, sizeof(double) * (size_t)cn, cudaMemcpyDeviceToDevice, s));
  for (int k = 0; k < nb; ++k) k_chol_fwd<<<nb - k, SV_THREADS, 0, s>>>(L, cn, P.d_Dinv, P.d_w, P.d_y, k);
  for (int i = nb - 1; i >= 0; --i) k_chol_bwd<<<i + 1, SV_THREADS, 0, s>>>(L, cn, P.d_Dinv, P.d_y, x, i);
  BA_CUDA(cudaGetLastError());
  return BA_OK;
}

}  // namespace ba

// ---- debug / benchmark entry: factor and solve a caller-supplied SPD matrix (tests, roofline of the factorisation) --
extern "C" int ba_dbg_chol(int device, int64_t n, const double* A_rowmajor, const double* b, double* x, double* L_out,
                           float* factor_ms, float* solve_ms) {
  if (n < 1 || !A_rowmajor || !b || !x) return BA_ERR_ARG;
  ba_handle hh;
  ba_handle* h = &hh;
  if (cudaSetDevice(device) != cudaSuccess) return BA_ERR_CUDA;
  const int64_t cn = ba::chol_padded(n);
  ba::chol_plan P;
  double *dA = nullptr, *db = nullptr;
  cudaStream_t s = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
  int rc = BA_OK, info = 0;
  auto done = [&](int code) {
    cudaFree(dA); cudaFree(db);
    ba::chol_plan_release(P);
    if (s) cudaStreamDestroy(s);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (e2) cudaEventDestroy(e2);
    return code;
  };
  if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) return done(BA_ERR_CUDA);
  cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
  if (cudaMalloc(reinterpret_cast<void**>(&dA), sizeof(double) * (size_t)(cn * cn)) != cudaSuccess) return done(BA_ERR_CUDA);
  if (cudaMalloc(reinterpret_cast<void**>(&db), sizeof(double) * (size_t)cn) != cudaSuccess) return done(BA_ERR_CUDA);
  if ((rc = ba::chol_plan_init(h, P, cn))) return done(rc);
  // pad with the identity
  std::vector<double> pad((size_t)cn, 0.0);
  cudaMemsetAsync(dA, 0, sizeof(double) * (size_t)(cn * cn), s);
  cudaMemcpy2DAsync(dA, sizeof(double) * (size_t)cn, A_rowmajor, sizeof(double) * (size_t)n, sizeof(double) * (size_t)n,
                    (size_t)n, cudaMemcpyHostToDevice, s);
  for (int64_t r = n; r < cn; ++r) {
    const double one = 1.0;
    cudaMemcpyAsync(dA + r * cn + r, &one, sizeof(double), cudaMemcpyHostToDevice, s);
  }
  cudaMemsetAsync(db, 0, sizeof(double) * (size_t)cn, s);
  cudaMemcpyAsync(db, b, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s);
  cudaStreamSynchronize(s);
  cudaEventRecord(e0, s);
  rc = ba::chol_factor(h, P, dA, s, nullptr);
  cudaEventRecord(e1, s);
  if (!rc) rc = ba::chol_solve(h, P, dA, db, db, s);
  cudaEventRecord(e2, s);
  if (rc) return done(rc);
  if (cudaStreamSynchronize(s) != cudaSuccess) return done(BA_ERR_CUDA);
  cudaMemcpy(&info, P.d_info, sizeof(int), cudaMemcpyDeviceToHost);
  if (factor_ms) cudaEventElapsedTime(factor_ms, e0, e1);
  if (solve_ms) cudaEventElapsedTime(solve_ms, e1, e2);
  cudaMemcpy(x, db, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost);
  if (L_out)
    cudaMemcpy2D(L_out, sizeof(double) * (size_t)n, dA, sizeof(double) * (size_t)cn, sizeof(double) * (size_t)n, (size_t)n,
                 cudaMemcpyDeviceToHost);
  if (cudaGetLastError() != cudaSuccess) return done(BA_ERR_CUDA);
  return done(info ? BA_ERR_NUMERIC : BA_OK);
}
