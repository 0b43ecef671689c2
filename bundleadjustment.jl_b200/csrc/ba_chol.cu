// ba_chol.cu -- blocked Cholesky A = L L' of the (explicitly assembled, Jacobi-scaled) reduced camera system, in FP64
// (the exact solver) and in FP32 on the tcgen05 tensor cores (the mixed-precision solver), and the two triangular
// sweeps; hand-written for sm_100a.
//
// Reference counterpart: the numeric factorisation and the L / D / L' sweeps of src/ldl_aux.jl:122-201,4-42 on
// the camera block (after the ordering has eliminated residual rows and points); the
// FP32 instantiation is the reference's facto_type < T mode (src/lm.jl:92-98,165-173), row f4.
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
// FP32 (templates instantiated for float): the same panels, look-ahead and diagonal blocks (still factorised in FP64:
// a latency chain); the panel solve on the legacy tensor path (mma.sync TF32, three-term split); the trailing update --
// the n^3/3 flops -- in k_chol_syrk_tc: persistent, warp-specialised, tcgen05.mma kind::tf32 with the accumulators in
// tensor memory (see there).  Sweeps: k_chol_sweep, both of them as one persistent cooperative kernel.
#include <unistd.h>
#include <algorithm>
#include <cstdlib>
#include "ba_internal.h"
#include "ba_chol.h"

namespace ba {
namespace {

constexpr int CT = CHOL_TILE;     // 128
constexpr int TM = 64;            // rows of a CTA tile of the two products (columns: CT)
constexpr int STAGES = 3;
constexpr int GEMM_THREADS = 256;
// Element type of the factorisation.  double: the exact solve (DMMA).  float: the mixed-precision factor
// (src/lm.jl:92-98 facto_type below the model type) -- FP32 storage, products on the tensor cores as three TF32
// MMAs per product (a = a_hi + a_lo split in registers: a_lo b_hi + a_hi b_lo + a_hi b_hi, FP32 accumulate),
// i.e. FP32-level accuracy at the TF32 rate / 3; the factor then preconditions FP64 CG on the FP64 operator.
template <typename T> struct mkt;
template <> struct mkt<double> {
  static constexpr int KC = 16;        // k per pipeline stage (128 bytes of a row)
  static constexpr int LDSM = KC + 4;  // padded row stride: fragment loads are bank-conflict free
  using T2 = double2;
};
template <> struct mkt<float> {
  static constexpr int KC = 32;
  static constexpr int LDSM = KC + 4;
  using T2 = float2;
};
template <typename T>
constexpr int gemm_smem() { return STAGES * (TM + CT) * mkt<T>::LDSM * (int)sizeof(T); }  // 92160 / 82944 B: two CTAs per SM

__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// x = hi + lo + O(2^-23 |x|): hi = x rounded to TF32 (low 13 mantissa bits zero), lo = the remainder rounded to TF32
// (the MMA would truncate it otherwise)
__device__ __forceinline__ void split_tf32(float x, unsigned& hi, unsigned& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(x - __uint_as_float(hi)));
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

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_volatile_16(const void* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
// bounded wait for a flag written by another rank (~2 s): false on time-out
__device__ __forceinline__ bool wait_epoch(const unsigned long long* f, unsigned long long epoch) {
  const long long t0 = clock64();
  while (ld_acquire_sys_u64(f) < epoch)
    if (clock64() - t0 > 4000000000ll) return false;
  return true;
}
// two adjacent elements as doubles (the sweeps compute in FP64 whatever the factor's storage type)
__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ double2 ld2(const float* p) {
  const float2 v = *reinterpret_cast<const float2*>(p);
  return make_double2((double)v.x, (double)v.y);
}

// Accumulators of one 64 x 128 CTA tile, 8 warps as 2 x 4, warp (wm, wn) owns rows wm*32.., columns wn*32...
// each(f) visits them as pairs of adjacent columns: f(row, column, v0, v1), coordinates relative to the CTA tile.
template <typename T> struct acc_frag;
template <> struct acc_frag<double> {  // m8n8k4: tile (mi, ni): lane holds row 8 mi + lane/4, columns 8 ni + 2 (lane%4) + {0, 1}
  double v[4][4][2];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) v[mi][ni][0] = v[mi][ni][1] = 0.0;
  }
  template <class F>
  __device__ __forceinline__ void each(F f) const {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r0 = (warp >> 2) * 32 + (lane >> 2), c0 = (warp & 3) * 32 + 2 * (lane & 3);
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) f(r0 + 8 * mi, c0 + 8 * ni, v[mi][ni][0], v[mi][ni][1]);
  }
};
template <> struct acc_frag<float> {  // m16n8k8: tile (mi, ni): c0 c1 = row 16 mi + lane/4, c2 c3 = that row + 8
  float v[2][4][4];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
#pragma unroll
        for (int e = 0; e < 4; ++e) v[mi][ni][e] = 0.f;
  }
  template <class F>
  __device__ __forceinline__ void each(F f) const {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r0 = (warp >> 2) * 32 + (lane >> 2), c0 = (warp & 3) * 32 + 2 * (lane & 3);
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        f(r0 + 16 * mi, c0 + 8 * ni, v[mi][ni][0], v[mi][ni][1]);
        f(r0 + 16 * mi + 8, c0 + 8 * ni, v[mi][ni][2], v[mi][ni][3]);
      }
  }
};

// one pipeline stage of MMAs from shared memory (As, Bs: this warp's / lane's fragment origin)
__device__ __forceinline__ void mma_stage(const double* As, const double* Bs, acc_frag<double>& acc) {
  constexpr int LD = mkt<double>::LDSM;
#pragma unroll
  for (int ks = 0; ks < mkt<double>::KC / 4; ++ks) {
    double a[4], b[4];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) a[mi] = As[mi * 8 * LD + ks * 4];
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) b[ni] = Bs[ni * 8 * LD + ks * 4];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) dmma8x8x4(acc.v[mi][ni][0], acc.v[mi][ni][1], a[mi], b[ni]);
  }
}
__device__ __forceinline__ void mma_stage(const float* As, const float* Bs, acc_frag<float>& acc) {
  constexpr int LD = mkt<float>::LDSM;
#pragma unroll
  for (int ks = 0; ks < mkt<float>::KC / 8; ++ks) {
    unsigned ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      const float* p = As + mi * 16 * LD + ks * 8;  // a0 (g, t)  a1 (g + 8, t)  a2 (g, t + 4)  a3 (g + 8, t + 4)
      split_tf32(p[0], ah[mi][0], al[mi][0]);
      split_tf32(p[8 * LD], ah[mi][1], al[mi][1]);
      split_tf32(p[4], ah[mi][2], al[mi][2]);
      split_tf32(p[8 * LD + 4], ah[mi][3], al[mi][3]);
    }
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      const float* p = Bs + ni * 8 * LD + ks * 8;   // b0 (k = t, n = g)  b1 (k = t + 4, n = g)
      split_tf32(p[0], bh[ni][0], bl[ni][0]);
      split_tf32(p[4], bh[ni][1], bl[ni][1]);
    }
    // small terms first; term by term over the eight tiles, so that dependent MMAs are eight issues apart
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) mma_tf32(acc.v[mi][ni], al[mi], bh[ni]);
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) mma_tf32(acc.v[mi][ni], ah[mi], bl[ni]);
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) mma_tf32(acc.v[mi][ni], ah[mi], bh[ni]);
  }
}

// acc = A (64 x 128 k, rows lda apart) * B' (B: 128 x 128 k, rows ldb apart), both operands k-contiguous.
// Two such CTAs share an SM (8 + 8 warps): while one waits (prologue loads, epilogue, the fixed issue distance of
// dependent MMAs) the other keeps the tensor pipe busy, and the half-height tiles halve the tail of the last wave.
// NK = number of k-chunks of KC: CT / KC for one 128-column panel, twice that for two adjacent panels.
template <typename T>
__device__ __forceinline__ void tile_abt(const T* __restrict__ A, int64_t lda, const T* __restrict__ B, int64_t ldb,
                                         acc_frag<T>& acc, T* sm, int NK) {
  constexpr int KC = mkt<T>::KC, LDSM = mkt<T>::LDSM, EPP = 16 / (int)sizeof(T);  // elements per 16-byte piece
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3, g = lane >> 2, t = lane & 3;
  acc.clear();
  auto load_stage = [&](int stage, int kc) {
    T* As = sm + stage * ((TM + CT) * LDSM);
    T* Bs = As + TM * LDSM;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = tid + i * GEMM_THREADS;  // 512 16-byte pieces of A (8 per row)
      const int row = c >> 3, part = c & 7;
      cp_async16(As + row * LDSM + part * EPP, A + (int64_t)row * lda + kc * KC + part * EPP);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = tid + i * GEMM_THREADS;  // 1024 pieces of B
      const int row = c >> 3, part = c & 7;
      cp_async16(Bs + row * LDSM + part * EPP, B + (int64_t)row * ldb + kc * KC + part * EPP);
    }
  };
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
    const T* As = sm + (kc % STAGES) * ((TM + CT) * LDSM) + (wm * 32 + g) * LDSM + t;
    const T* Bs = sm + (kc % STAGES) * ((TM + CT) * LDSM) + TM * LDSM + (wn * 32 + g) * LDSM + t;
    mma_stage(As, Bs, acc);
  }
  cp_async_wait<0>();
  __syncthreads();
}

// trailing update, half tiles: A_ij[half] -= P_i[half] P_j', j = j0 + blockIdx.y.
//   single GPU : i = j + blockIdx.x / 2
//   DIST       : i = i0 + R (blockIdx.x / 2), the rank's own tile rows from i0 on; the CTA first waits until every
//                rank's tiles of panel k have landed in this rank's copy of the matrix (flags over NVLink)
template <typename T, bool DIST>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
k_chol_syrk(T* __restrict__ A, int64_t ld, int k, int j0, int nb, int i0, chol_peers P, int* __restrict__ info,
            int nk, int kwait) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  T* sm = reinterpret_cast<T*>(sm_raw);
  const int j = j0 + blockIdx.y, half = blockIdx.x & 1;
  const int i = DIST ? i0 + P.R * (int)(blockIdx.x >> 1) : j + (int)(blockIdx.x >> 1);
  if (i >= nb || i < j) return;
  if (DIST) {
    __shared__ int ok;
    if (threadIdx.x == 0) ok = 1;
    __syncthreads();
    // (a rank reports its panels in order, so the flag of the last panel used implies the earlier ones)
    if ((int)threadIdx.x < P.R && !wait_epoch(P.ctl[P.q] + CHOL_NBMAX + 16 * kwait + threadIdx.x, P.epoch)) ok = 0;
    __syncthreads();
    if (!ok) {
      if (threadIdx.x == 0) atomicCAS(info, 0, -2);
      return;
    }
  }
  const T* Pi = A + ((int64_t)i * CT + half * TM) * ld + (int64_t)k * CT;
  const T* Pj = A + (int64_t)j * CT * ld + (int64_t)k * CT;
  acc_frag<T> acc;
  tile_abt<T>(Pi, ld, Pj, ld, acc, sm, nk);  // nk chunks: one panel, or the panels k and k + 1 (adjacent columns) in one pass
  T* C = A + ((int64_t)i * CT + half * TM) * ld + (int64_t)j * CT;
  using T2 = typename mkt<T>::T2;
  acc.each([&](int r, int c, T v0, T v1) {
    T2* p = reinterpret_cast<T2*>(C + (int64_t)r * ld + c);
    T2 cv = *p;
    cv.x -= v0;
    cv.y -= v1;
    *p = cv;
  });
}

// ---- trailing update of the FP32 factorisation on the 5th-generation tensor cores (tcgen05, accumulators in TMEM) ----
// A_ij (128 x 128, FP32) -= P_i P_j' over nk chunks of 32 k, P = the panels k (and k + 1), for every tile (i >= j) of
// the columns [j0, j0 + ncol): a PERSISTENT, warp-specialised kernel, one CTA per SM, tiles dealt round-robin.
// FP32-level accuracy from TF32 MMAs by the three-term split p = hi + lo (both TF32).  Roles:
//   * 8 producer warps: load the FP32 operand chunks (two chunks ahead, in registers), split them and store hi and lo
//     into a 3-stage ring in shared memory, in the canonical K-major 128-byte-swizzled layout (row r at 128 r, 16-byte
//     piece c at position c ^ (r & 7)); they run ahead across tile boundaries, so there is no pipeline drain per tile;
//   * 1 MMA thread: per 8 k, lo*hi' + hi*lo' into one TMEM accumulator and hi*hi' into another (tcgen05.mma kind::tf32,
//     M = N = 128) -- two accumulators, so that the small cross terms are not rounded at the magnitude of the large one;
//     tcgen05.commit hands the stage back to the producers and, after the last chunk, the accumulators to the epilogue;
//   * 4 epilogue warps: tcgen05.ld of both accumulators, sum, subtract from the tile in global memory, while the next
//     tile is already being accumulated in the other half of the tensor memory (2 x 256 columns).
// The MMA's M side (TMEM lanes) is P_j and its N side (TMEM columns) is P_i, i.e. the accumulator is the TRANSPOSED
// update: an epilogue warp's 32 lanes then hold 32 consecutive columns of one row of A_ij -- coalesced 128-byte accesses.
constexpr int TC_PROD_WARPS = 8, TC_EPI_WARPS = 4;
constexpr int TC_PROD_THREADS = 32 * TC_PROD_WARPS;
constexpr int TC_THREADS = 32 * (TC_PROD_WARPS + 1 + TC_EPI_WARPS);  // 416
constexpr int TC_KC = 32;                            // floats per row of a stage = 128 bytes = one swizzle atom
constexpr int TC_OPER = CT * 128;                    // bytes of one operand tile of a stage (128 rows x 128 B)
constexpr int TC_STAGE = 4 * TC_OPER;                // P_i hi, P_i lo, P_j hi, P_j lo
constexpr int TC_STAGES = 3;
constexpr int TC_SMEM = TC_STAGES * TC_STAGE + 1024; // + slack to align the tiles to 1024 B (swizzle atom)
constexpr unsigned TC_COLS = 512;                    // TMEM columns: two buffers x (accumulator of hi*hi' + of the cross terms)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100: version 1): start address, LBO 1, SBO 1024 B
__device__ __forceinline__ unsigned long long umma_desc_sw128(unsigned saddr) {
  return (unsigned long long)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((unsigned long long)(1024 >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
// D[tmem] (+)= A[smem] B[smem]', TF32 inputs, FP32 accumulate; issued by one thread for the CTA
__device__ __forceinline__ void umma_tf32(unsigned d_tmem, unsigned long long adesc, unsigned long long bdesc,
                                          unsigned idesc, unsigned accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait (~1 s) on an mbarrier phase: false on time-out (a wrong descriptor must not hang the GPU)
__device__ __forceinline__ bool mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned a = smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
    unsigned done;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (done) return true;
    if (clock64() - t0 > 2000000000ll) return false;
  }
}
__device__ __forceinline__ void tmem_ld16(unsigned taddr, float (&v)[16]) {
  unsigned r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// x = hi + lo + O(2^-23 |x|) with hi and lo TF32 (13 low mantissa bits zero), round to nearest (ties away) by integer
// arithmetic: five full-rate instructions per element (cvt.rna.tf32.f32 costs about four each, twice)
__device__ __forceinline__ void split_tf32_fast(float x, unsigned& hi, unsigned& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
  lo = (__float_as_uint(x - __uint_as_float(hi)) + 0x1000u) & 0xFFFFE000u;
}

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __maxnreg__(128)  // 13 warps are allocated as 16 (granularity 4): 16 x 32 x 128 registers = the SM's file
k_chol_syrk_tc(float* __restrict__ A, int64_t ld, int k, int j0, int ncol, int nb, int* __restrict__ info, int nk) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  __shared__ __align__(8) unsigned long long full[TC_STAGES], empty[TC_STAGES], tfull[2], tempty[2];
  __shared__ unsigned tmem_base_sh;
  __shared__ int ok_sh;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) ok_sh = *reinterpret_cast<volatile int*>(info) == 0 ? 1 : 0;  // an earlier failure (time-out, pivot): nothing to do
  __syncthreads();
  if (!ok_sh) return;
  unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~(uintptr_t)1023);
  if (warp == TC_PROD_WARPS) {  // the MMA warp allocates the tensor memory of the CTA (and frees it at the end)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)), "r"(TC_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int st = 0; st < TC_STAGES; ++st) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[st])), "r"(TC_PROD_THREADS) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[st])) : "memory");
    }
    for (int bf = 0; bf < 2; ++bf) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&tfull[bf])) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&tempty[bf])), "r"(32 * TC_EPI_WARPS) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = tmem_base_sh;
  // tiles of this launch: columns j = j0 .. j0 + ncol - 1 (< nb), rows i = j .. nb - 1; tile t -> (j, i)
  const int jend = min(j0 + ncol, nb);
  int ntiles = 0;
  for (int j = j0; j < jend; ++j) ntiles += nb - j;
  // this CTA's tiles, t = blockIdx.x, + gridDim.x, ...: (column j, row j + rem), advanced incrementally
  struct tile_iter {
    int j, rem, nb;
    __device__ __forceinline__ void advance(int by) {
      rem += by;
      while (j < nb && rem >= nb - j) {  // (j == nb: past the last tile)
        rem -= nb - j;
        ++j;
      }
    }
    __device__ __forceinline__ int row() const { return j + rem; }
  };
  auto first_tile = [&]() {
    tile_iter it = {j0, 0, nb};
    it.advance((int)blockIdx.x);
    return it;
  };
  // instruction descriptor: D FP32 (bits 4-5 = 1), A and B TF32 (bits 7-9, 10-12 = 2), both K-major, N / 8 at bit 17,
  // M / 16 at bit 24
  constexpr unsigned IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(CT >> 3) << 17) | ((unsigned)(CT >> 4) << 24);
  bool failed = false;
  if (warp < TC_PROD_WARPS) {
    // ---- producers: chunk sequence g = 0, 1, ... over (tile, chunk); this thread's four 16-byte pieces of an operand
    // tile: piece e = tid + 256 q -> row e / 8, piece e % 8 of the row
    const int row0 = tid >> 3, pc = tid & 7;
    const int soff = row0 * 128 + ((pc ^ (row0 & 7)) << 4);  // (row + 32 q) & 7 == row & 7
    struct chunk_regs { float4 a[4], b[4]; };
    const int nchunks = ((ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * nk;  // of this CTA
    tile_iter it = first_tile();
    int cnext = 0;  // chunk of the tile `it` the next fetch brings
    auto fetch = [&](chunk_regs& v) {
      const int c = cnext, i = it.row(), j = it.j;
      if (++cnext == nk) {
        cnext = 0;
        it.advance((int)gridDim.x);
      }
      const float* Pi = A + ((int64_t)i * CT + row0) * ld + (int64_t)k * CT + c * TC_KC + pc * 4;
      const float* Pj = A + ((int64_t)j * CT + row0) * ld + (int64_t)k * CT + c * TC_KC + pc * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) v.a[q] = *reinterpret_cast<const float4*>(Pi + (int64_t)(32 * q) * ld);
      if (i != j) {
#pragma unroll
        for (int q = 0; q < 4; ++q) v.b[q] = *reinterpret_cast<const float4*>(Pj + (int64_t)(32 * q) * ld);
      }
      return i != j;
    };
    auto put1 = [&](const float4& v, unsigned char* hi, int off) {
      uint4 h, l;
      split_tf32_fast(v.x, h.x, l.x);
      split_tf32_fast(v.y, h.y, l.y);
      split_tf32_fast(v.z, h.z, l.z);
      split_tf32_fast(v.w, h.w, l.w);
      *reinterpret_cast<uint4*>(hi + off) = h;
      *reinterpret_cast<uint4*>(hi + TC_OPER + off) = l;
    };
    auto step = [&](int g, const chunk_regs& v, bool both) {
      const int st = g % TC_STAGES;
      unsigned char* base = tiles + st * TC_STAGE;
      // the MMAs that read this stage TC_STAGES chunks ago are done
      if (g >= TC_STAGES && !mbar_wait(&empty[st], (unsigned)((g / TC_STAGES - 1) & 1))) failed = true;
#pragma unroll
      for (int q = 0; q < 4; ++q) put1(v.a[q], base, soff + q * 32 * 128);
      if (both) {
#pragma unroll
        for (int q = 0; q < 4; ++q) put1(v.b[q], base + 2 * TC_OPER, soff + q * 32 * 128);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
      mbar_arrive(&full[st]);
    };
    chunk_regs v0, v1;
    bool b0 = false, b1 = false;
    if (nchunks > 0) b0 = fetch(v0);
    if (nchunks > 1) b1 = fetch(v1);
    for (int g = 0; g < nchunks && !failed; g += 2) {  // (nk is even, so nchunks is)
      step(g, v0, b0);
      if (g + 2 < nchunks) b0 = fetch(v0);
      step(g + 1, v1, b1);
      if (g + 3 < nchunks) b1 = fetch(v1);
    }
  } else if (warp == TC_PROD_WARPS) {
    // ---- MMA issuer (one lane)
    if (lane == 0) {
      int g = 0, tc = 0;
      tile_iter it = first_tile();
      for (int t = blockIdx.x; t < ntiles && !failed; t += gridDim.x, ++tc, it.advance((int)gridDim.x)) {
        const int i = it.row(), j = it.j;
        const int bf = tc & 1;
        // the epilogue has drained this accumulator buffer (two tiles ago)
        if (tc >= 2 && !mbar_wait(&tempty[bf], (unsigned)((tc / 2 - 1) & 1))) failed = true;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned d = tmem + bf * 2 * CT;
        for (int c = 0; c < nk && !failed; ++c, ++g) {
          const int st = g % TC_STAGES;
          if (!mbar_wait(&full[st], (unsigned)((g / TC_STAGES) & 1))) failed = true;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const unsigned sa = smem_u32(tiles + st * TC_STAGE);
          // M side (TMEM lanes) = P_j, N side (TMEM columns) = P_i; on the diagonal P_j = P_i
          const unsigned long long nhi = umma_desc_sw128(sa), nlo = umma_desc_sw128(sa + TC_OPER);
          const unsigned long long mhi = (i == j) ? nhi : umma_desc_sw128(sa + 2 * TC_OPER);
          const unsigned long long mlo = (i == j) ? nlo : umma_desc_sw128(sa + 3 * TC_OPER);
#pragma unroll
          for (int ks = 0; ks < TC_KC / 8; ++ks) {  // 8 k = 32 bytes further along the rows: + 2 in the address field
            const unsigned acc = (c > 0 || ks > 0) ? 1u : 0u;
            umma_tf32(d + CT, mlo + 2 * ks, nhi + 2 * ks, IDESC, acc);
            umma_tf32(d + CT, mhi + 2 * ks, nlo + 2 * ks, IDESC, 1u);
            umma_tf32(d, mhi + 2 * ks, nhi + 2 * ks, IDESC, acc);
          }
          umma_commit(&empty[st]);  // arrives once these (and all earlier) MMAs have completed
        }
        umma_commit(&tfull[bf]);
      }
    }
  } else {
    // ---- epilogue: warp e reads TMEM lanes 32 (warp % 4) .. = columns of the tile; TMEM column n = row n of the tile
    const int q4 = warp & 3;
    int tc = 0;
    tile_iter it = first_tile();
    for (int t = blockIdx.x; t < ntiles && !failed; t += gridDim.x, ++tc, it.advance((int)gridDim.x)) {
      const int i = it.row(), j = it.j;
      const int bf = tc & 1;
      if (!__all_sync(0xffffffffu, mbar_wait(&tfull[bf], (unsigned)((tc / 2) & 1)))) {  // (warp-uniform: tcgen05.ld is .aligned)
        failed = true;
        break;
      }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float* C = A + (int64_t)i * CT * ld + (int64_t)j * CT + 32 * q4 + lane;
      const unsigned ta = tmem + ((unsigned)(32 * q4) << 16) + (unsigned)(bf * 2 * CT);
#pragma unroll 1
      for (int cc = 0; cc < 8; ++cc) {  // 16 rows of the tile per pass: 16 coalesced loads in flight per lane
        float d0[16], d1[16], cv[16];
#pragma unroll
        for (int n = 0; n < 16; ++n) cv[n] = C[(int64_t)(cc * 16 + n) * ld];
        tmem_ld16(ta + cc * 16, d0);
        tmem_ld16(ta + CT + cc * 16, d1);
#pragma unroll
        for (int n = 0; n < 16; ++n) C[(int64_t)(cc * 16 + n) * ld] = cv[n] - (d0[n] + d1[n]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&tempty[bf]);
    }
  }
  if (failed) atomicCAS(info, 0, -3);  // an mbarrier phase never completed
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == TC_PROD_WARPS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TC_COLS) : "memory");
}

// panel solve: P_i[half] <- P_i[half] Linv_kk'.
//   single GPU : row tiles i = k + 1 + blockIdx.x / 2
//   DIST       : the rank's own row tiles i = i0 + R (blockIdx.x / 2); the result is stored into EVERY rank's copy of
//                the matrix (peer-memory stores over NVLink fused into the epilogue: the all-gather of the panel),
//                and the CTA that finishes last tells every rank that this rank's part of panel k is complete
template <typename T, bool DIST>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
k_chol_trsm(T* __restrict__ A, int64_t ld, int k, const T* __restrict__ Dinv, int i0, chol_peers P,
            int* __restrict__ cnt) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  T* sm = reinterpret_cast<T*>(sm_raw);
  const int half = blockIdx.x & 1;
  const int i = DIST ? i0 + P.R * (int)(blockIdx.x >> 1) : k + 1 + (int)(blockIdx.x >> 1);
  const int64_t off = ((int64_t)i * CT + half * TM) * ld + (int64_t)k * CT;
  acc_frag<T> acc;
  tile_abt<T>(A + off, ld, Dinv + (int64_t)k * CT * CT, CT, acc, sm, CT / mkt<T>::KC);  // (every read of these rows is complete on return)
  using T2 = typename mkt<T>::T2;
  if constexpr (!DIST) {
    acc.each([&](int r, int c, T v0, T v1) {
      T2 o;
      o.x = v0;
      o.y = v1;
      *reinterpret_cast<T2*>(A + off + (int64_t)r * ld + c) = o;
    });
  } else {
    for (int q = 0; q < P.R; ++q) {
      T* C = static_cast<T*>(P.S[q]) + off;
      acc.each([&](int r, int c, T v0, T v1) {
        T2 o;
        o.x = v0;
        o.y = v1;
        *reinterpret_cast<T2*>(C + (int64_t)r * ld + c) = o;
      });
    }
    __threadfence_system();
    __syncthreads();
    __shared__ int last;
    if (threadIdx.x == 0) last = (atomicAdd(cnt, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (last) {
      __threadfence_system();
      if ((int)threadIdx.x < P.R) st_release_sys_u64(P.ctl[threadIdx.x] + CHOL_NBMAX + 16 * k + P.q, P.epoch);
      if (threadIdx.x == 0) *cnt = 0;  // ready for the next panel (kernel boundaries order this)
    }
  }
}

// a rank without tile rows below k still reports "done with panel k" (it has fetched the diagonal block)
__global__ void k_chol_signal(int k, chol_peers P) {
  if ((int)threadIdx.x < P.R) st_release_sys_u64(P.ctl[threadIdx.x] + CHOL_NBMAX + 16 * k + P.q, P.epoch);
}

// end of a distributed factorisation: every rank has reported every panel, i.e. all peer stores into this rank's
// matrix have landed and nobody reads this rank's diagonal blocks any more (the matrix may be reused)
__global__ void __launch_bounds__(256)
k_chol_wait_all(int nb, chol_peers P, int* __restrict__ info) {
  for (int e = threadIdx.x; e < nb * P.R; e += 256) {
    const int k = e / P.R, r = e - k * P.R;
    if (!wait_epoch(P.ctl[P.q] + CHOL_NBMAX + 16 * k + r, P.epoch)) atomicCAS(info, 0, -2);
  }
}

// non-owners pull L_kk and Linv_kk from the owner of tile row k once its flag says they are complete
// (16-byte pieces: PR of them per row of a block)
template <typename T>
__global__ void __launch_bounds__(256)
k_chol_fetch_diag(T* __restrict__ A, int64_t ld, int k, T* __restrict__ Dinv, chol_peers P, int owner,
                  int* __restrict__ info) {
  __shared__ int ok;
  if (threadIdx.x == 0) ok = wait_epoch(P.ctl[P.q] + k, P.epoch) ? 1 : 0;
  __syncthreads();
  if (!ok) {
    if (threadIdx.x == 0) atomicCAS(info, 0, -2);
    return;
  }
  constexpr int EPP = 16 / (int)sizeof(T), PR = CT / EPP;
  const int64_t base = (int64_t)k * CT * ld + (int64_t)k * CT;
  const T* sL = static_cast<const T*>(P.S[owner]) + base;
  const T* sD = static_cast<const T*>(P.D[owner]) + (int64_t)k * CT * CT;
  T* dD = Dinv + (int64_t)k * CT * CT;
  const int n2 = CT * PR;  // pieces per block
  for (int e = blockIdx.x * 256 + threadIdx.x; e < 2 * n2; e += gridDim.x * 256) {
    if (e < n2) {
      const int r = e / PR, c2 = e - r * PR;
      *reinterpret_cast<uint4*>(A + base + (int64_t)r * ld + EPP * c2) = ld_volatile_16(sL + (int64_t)r * ld + EPP * c2);
    } else {
      const int f = e - n2;
      *reinterpret_cast<uint4*>(dD + EPP * f) = ld_volatile_16(sD + EPP * f);
    }
  }
}

// ---- diagonal block -------------------------------------------------------------------------------------
// One CTA factorises the 128 x 128 diagonal block in shared memory and inverts the factor.  It sits on the critical
// path of every panel step, so it is organised for latency (ncu: the first version spent 45 % of its time in the
// one-warp 16 x 16 steps and 30 % in the scalar loops of the inversion):
//   * 16-column sub-panels; the 16 x 16 diagonal block is factorised by ONE warp in registers (rows in lanes,
//     columns by shuffles, rsqrt instead of sqrt + division on the dependent chain);
//   * the rows below are solved against it by forward substitution, one thread per row (independent rows);
//   * the trailing update uses 4 x 4 register patches;
//   * after the loop the eight 16 x 16 diagonal factors are inverted by eight warps AT ONCE, and the inverse of the
//     whole factor is assembled by recursive doubling ([L11 0; L21 L22]^-1 = [X11 0; -X22 L21 X11, X22]: three
//     levels of two small products, again in 4 x 4 register patches).
constexpr int PO_THREADS = 512;
constexpr int PO_LD = CT + 4;  // padded row stride in shared memory (= 4 mod 16: conflict-free DMMA fragment loads)
constexpr int PO_SB = 16;      // sub-panel width
constexpr int PO_SMEM = (CT * PO_LD + CT + 64 * 68) * (int)sizeof(double);

// One 8 x 8 tile of a product on the FP64 tensor pipe, operands in shared memory: acc += A[0:8, 0:K] * op(B),
// A rows lda apart (k contiguous); BT: B given as rows of n with k contiguous (B[n][k]), else B[k][n].
// Lane (g, t) = (lane / 4, lane % 4) ends up with acc rows g, columns 2 t and 2 t + 1.
template <bool BT>
__device__ __forceinline__ void po_mma(const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
                                       int K, int g, int t, double& c0, double& c1) {
#pragma unroll 4
  for (int k0 = 0; k0 < K; k0 += 4) {
    const double av = A[g * lda + k0 + t];
    const double bv = BT ? B[g * ldb + k0 + t] : B[(k0 + t) * ldb + g];
    dmma8x8x4(c0, c1, av, bv);
  }
}

// A_kk (lower) <- L_kk, Dinv[k] <- L_kk^-1.  info: first non-positive pivot (1-based global index), else untouched.
// T: storage type of the matrix, of L and of Linv; the block itself is factorised and inverted in FP64 either way
// (it is a latency-bound chain, not a throughput problem).
template <typename T>
__global__ void __launch_bounds__(PO_THREADS, 1)
k_chol_potrf(T* __restrict__ A, int64_t ld, int k, T* __restrict__ Dinv, int* __restrict__ info,
             long long* __restrict__ prof, chol_peers P) {
  extern __shared__ __align__(16) double smp[];
  long long tk[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tl = 0;  // phase cycle counts (BAGPU_POTRF_PROF; thread 0 only)
#define PO_LAP(slot)                          \
  if (prof && threadIdx.x == 0) {             \
    const long long now_ = clock64();         \
    tk[slot] += now_ - tl;                    \
    tl = now_;                                \
  }
  if (prof && threadIdx.x == 0) tl = clock64();
  double* a = smp;                       // 128 x 129: the block, then its inverse
  double* rinv = smp + CT * PO_LD;       // 128: reciprocals of the diagonal of L
  double* tmp = rinv + CT;               // 64 x 65 scratch of the doubling levels
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  T* Akk = A + (int64_t)k * CT * ld + (int64_t)k * CT;
  for (int e = tid; e < CT * CT; e += PO_THREADS) {
    const int r = e >> 7, c = e & 127;
    a[r * PO_LD + c] = (c <= r) ? (double)Akk[(int64_t)r * ld + c] : 0.0;
  }
  __syncthreads();
  PO_LAP(0)
  for (int j0 = 0; j0 < CT; j0 += PO_SB) {
    // (a) 16 x 16 diagonal sub-block: Cholesky by 16 lanes of warp 0; lane r keeps row r in registers, a column
    // travels through shared memory (one store, one warp barrier, broadcast loads); no division on the chain:
    // l_jj = d_jj * rsqrt(d_jj), l_rj = a_rj * rsqrt(d_jj), every lane forms rsqrt(d_jj) itself
    if (warp == 0) {
      const int r = lane & 15;
      double* blk = a + j0 * PO_LD + j0;
      double row[PO_SB];
#pragma unroll
      for (int c = 0; c < PO_SB; ++c) row[c] = blk[r * PO_LD + c];  // zeros above the diagonal
#pragma unroll
      for (int j = 0; j < PO_SB; ++j) {
        if (lane == j) blk[j * PO_LD + j] = row[j];  // the updated pivot
        __syncwarp();
        const double djj = blk[j * PO_LD + j];
        if (!(djj > 0.0) && lane == 0) atomicCAS(info, 0, k * CT + j0 + j + 1);
        const double rs = rsqrt(djj);
        const double l = row[j] * rs;  // (rows above j hold zeros in this column)
        row[j] = l;
        __syncwarp();                  // everybody has read the pivot before it is overwritten by l_jj
        if (lane < PO_SB) blk[r * PO_LD + j] = l;
        if (lane == j) rinv[j0 + j] = rs;
        __syncwarp();
#pragma unroll
        for (int c = j + 1; c < PO_SB; ++c) {
          const double lc = blk[c * PO_LD + j];
          if (c <= r) row[c] -= l * lc;
        }
      }
    }
    __syncthreads();
    PO_LAP(1)
    // (b) rows below: x L16' = row by forward substitution, one thread per row
    const int nbelow = CT - j0 - PO_SB;
    if (tid < nbelow) {
      double* rowp = a + (j0 + PO_SB + tid) * PO_LD + j0;
      double x[PO_SB];
#pragma unroll
      for (int j = 0; j < PO_SB; ++j) {
        double sacc = rowp[j];
#pragma unroll
        for (int q = 0; q < j; ++q) sacc -= x[q] * a[(j0 + j) * PO_LD + j0 + q];
        x[j] = sacc * rinv[j0 + j];
      }
#pragma unroll
      for (int j = 0; j < PO_SB; ++j) rowp[j] = x[j];
    }
    __syncthreads();
    PO_LAP(2)
    // (c) trailing update of the lower triangle, a[r][c] -= sum_q a[r][j0+q] a[c][j0+q], in 8 x 8 tiles on the
    // FP64 tensor pipe (4 DMMAs per tile): every lane of every warp works even when few tiles are left, and a tile
    // needs 8 shared-memory loads per 256 multiply-adds instead of the 128 of scalar 4 x 4 patches
    {
      const int nt8 = nbelow >> 3, ntile = nt8 * (nt8 + 1) / 2, base = j0 + PO_SB;
      const int g = lane >> 2, t = lane & 3;
      for (int tile = warp; tile < ntile; tile += PO_THREADS / 32) {
        int tr = (int)((sqrtf(8.0f * (float)tile + 1.0f) - 1.0f) * 0.5f);
        while (tr * (tr + 1) / 2 > tile) --tr;
        while ((tr + 1) * (tr + 2) / 2 <= tile) ++tr;
        const int tc = tile - tr * (tr + 1) / 2;
        const int r0 = base + 8 * tr, c0 = base + 8 * tc;
        double s0 = 0.0, s1 = 0.0;
        po_mma<true>(a + r0 * PO_LD + j0, PO_LD, a + c0 * PO_LD + j0, PO_LD, PO_SB, g, t, s0, s1);
        double* cp = a + (r0 + g) * PO_LD + c0 + 2 * t;
        if (tr != tc) {
          double2 cv = *reinterpret_cast<double2*>(cp);
          cv.x -= s0;
          cv.y -= s1;
          *reinterpret_cast<double2*>(cp) = cv;
        } else {  // diagonal tile: the zeros above the diagonal stay
          if (2 * t <= g) cp[0] -= s0;
          if (2 * t + 1 <= g) cp[1] -= s1;
        }
      }
    }
    __syncthreads();
    PO_LAP(3)
  }
  for (int e = tid; e < CT * CT; e += PO_THREADS) {
    const int r = e >> 7, c = e & 127;
    if (c <= r) Akk[(int64_t)r * ld + c] = (T)a[r * PO_LD + c];
  }
  __syncthreads();  // the factor has been read out before its storage is reused
  PO_LAP(4)
  // inverse of the factor, in place.  Diagonal 16 x 16 blocks first: thread (block b, column c) solves L_bb x = e_c
  // by forward substitution, L read from shared memory (the 16 threads of a block read the same word: broadcast)
  {
    const int b0 = (tid >> 4) * PO_SB, c = tid & 15;
    double x[PO_SB];
    if (tid < CT) {
      const double* blk = a + b0 * PO_LD + b0;
#pragma unroll
      for (int i = 0; i < PO_SB; ++i) {
        double sacc = (i == c) ? 1.0 : 0.0;
#pragma unroll
        for (int q = 0; q < i; ++q) sacc -= blk[i * PO_LD + q] * x[q];  // x[q] == 0 for q < c
        x[i] = (i >= c) ? sacc * rinv[b0 + i] : 0.0;
      }
    }
    __syncthreads();  // every thread has read L before the blocks are overwritten
    if (tid < CT) {
#pragma unroll
      for (int i = 0; i < PO_SB; ++i) a[(b0 + i) * PO_LD + b0 + c] = x[i];  // X[i][c] (zeros above the diagonal)
    }
  }
  __syncthreads();
  PO_LAP(5)
  // X21 = -X22 (L21 X11) for block sizes 16, 32, 64; both products in 8 x 8 tiles on the FP64 tensor pipe
  {
    const int g = lane >> 2, t = lane & 3;
    for (int m = PO_SB; m < CT; m <<= 1) {
      const int npair = CT / (2 * m), m8 = m >> 3, per = m8 * m8, tld = m + 4;
      // T = L21 X11
      for (int item = warp; item < npair * per; item += PO_THREADS / 32) {
        const int pi = item / per, rem = item - pi * per, tr = rem / m8, tc = rem - tr * m8, o = pi * 2 * m;
        double s0 = 0.0, s1 = 0.0;
        po_mma<false>(a + (o + m + 8 * tr) * PO_LD + o, PO_LD, a + o * PO_LD + o + 8 * tc, PO_LD, m, g, t, s0, s1);
        *reinterpret_cast<double2*>(tmp + (pi * m + 8 * tr + g) * tld + 8 * tc + 2 * t) = make_double2(s0, s1);
      }
      __syncthreads();
      // X21 = -X22 T
      for (int item = warp; item < npair * per; item += PO_THREADS / 32) {
        const int pi = item / per, rem = item - pi * per, tr = rem / m8, tc = rem - tr * m8, o = pi * 2 * m;
        double s0 = 0.0, s1 = 0.0;
        po_mma<false>(a + (o + m + 8 * tr) * PO_LD + o + m, PO_LD, tmp + pi * m * tld + 8 * tc, tld, m, g, t, s0, s1);
        *reinterpret_cast<double2*>(a + (o + m + 8 * tr + g) * PO_LD + o + 8 * tc + 2 * t) = make_double2(-s0, -s1);
      }
      __syncthreads();
    }
  }
  PO_LAP(6)
  T* D = Dinv + (int64_t)k * CT * CT;
  for (int e = tid; e < CT * CT; e += PO_THREADS) {
    const int r = e >> 7, c = e & 127;
    D[e] = (c <= r) ? (T)a[r * PO_LD + c] : (T)0;
  }
  PO_LAP(7)
  if (prof && threadIdx.x == 0)
    for (int i = 0; i < 8; ++i) prof[i] = tk[i];
#undef PO_LAP
  if (P.R > 1) {  // distributed: L_kk and Linv_kk are complete in this rank's memory -- every rank may fetch them
    __threadfence_system();
    __syncthreads();
    if (tid < P.R) st_release_sys_u64(P.ctl[tid] + k, P.epoch);
  }
}

// ---- substitution sweeps ----------------------------------------------------------------------------------
// 2 cn/128 dependent steps, each a tiny kernel: organised for latency (every global load of a step is issued up
// front -- the tile loads do not depend on the small solve with the diagonal block), launched as one CUDA graph.
constexpr int SV_THREADS = 512;

// step k of L y = b: every CTA forms y_k = Linv_kk w_k; CTA 0 stores it, CTA c > 0 updates row tile i = k + c:
// w_i -= L_ik y_k.  Warp w owns rows 8 w .. 8 w + 7; a lane holds columns 2 lane, 2 lane + 1, 64 + 2 lane, 65 + 2 lane.
template <typename T>
__global__ void __launch_bounds__(SV_THREADS)
k_chol_fwd(const T* __restrict__ L, int64_t ld, const T* __restrict__ Dinv, double* __restrict__ w,
           double* __restrict__ y, int k) {
  __shared__ double yk[CT];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i = k + blockIdx.x;
  const T* D = Dinv + (int64_t)k * CT * CT + (int64_t)(warp * 8) * CT + 2 * lane;
  const T* Tl = L + ((int64_t)i * CT + warp * 8) * ld + (int64_t)k * CT + 2 * lane;
  double2 d0[8], d1[8], t0[8], t1[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    d0[r] = ld2(D + r * CT);
    d1[r] = ld2(D + r * CT + 64);
  }
  if (blockIdx.x > 0) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      t0[r] = ld2(Tl + (int64_t)r * ld);
      t1[r] = ld2(Tl + (int64_t)r * ld + 64);
    }
  }
  const double2 w0 = *reinterpret_cast<const double2*>(w + (int64_t)k * CT + 2 * lane);
  const double2 w1 = *reinterpret_cast<const double2*>(w + (int64_t)k * CT + 64 + 2 * lane);
  double s[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) s[r] = (d0[r].x * w0.x + d0[r].y * w0.y) + (d1[r].x * w1.x + d1[r].y * w1.y);
#pragma unroll
  for (int o = 16; o; o >>= 1)
#pragma unroll
    for (int r = 0; r < 8; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < 8; ++r) yk[warp * 8 + r] = s[r];
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    if (tid < CT) y[(int64_t)k * CT + tid] = yk[tid];
    return;
  }
  const double y00 = yk[2 * lane], y01 = yk[2 * lane + 1], y10 = yk[64 + 2 * lane], y11 = yk[65 + 2 * lane];
#pragma unroll
  for (int r = 0; r < 8; ++r) s[r] = (t0[r].x * y00 + t0[r].y * y01) + (t1[r].x * y10 + t1[r].y * y11);
#pragma unroll
  for (int o = 16; o; o >>= 1)
#pragma unroll
    for (int r = 0; r < 8; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
  if (lane == 0) {
    double* wi = w + (int64_t)i * CT + warp * 8;
#pragma unroll
    for (int r = 0; r < 8; ++r) wi[r] -= s[r];
  }
}

// step i of L' x = y (i descending): every CTA forms x_i = Linv_ii' y_i; CTA i stores it, CTA kk < i updates
// y_kk -= L_{i,kk}' x_i.  Thread (c, part): column c, rows 32 part .. 32 part + 31 (coalesced across c).
template <typename T>
__global__ void __launch_bounds__(SV_THREADS)
k_chol_bwd(const T* __restrict__ L, int64_t ld, const T* __restrict__ Dinv, double* __restrict__ y,
           double* __restrict__ x, int i) {
  __shared__ double yi[CT], xi[CT], part[3][CT];
  const int tid = threadIdx.x;
  const int c = tid & (CT - 1), h = tid >> 7;  // column, quarter of the rows
  const int kk = blockIdx.x;
  const T* D = Dinv + (int64_t)i * CT * CT + (int64_t)(h * 32) * CT + c;
  const T* Tl = L + ((int64_t)i * CT + h * 32) * ld + (int64_t)kk * CT + c;
  double dv[32], tv[32];
#pragma unroll
  for (int r = 0; r < 32; ++r) dv[r] = (double)D[r * CT];
  if (kk != i) {
#pragma unroll
    for (int r = 0; r < 32; ++r) tv[r] = (double)Tl[(int64_t)r * ld];
  }
  if (tid < CT) yi[tid] = y[(int64_t)i * CT + tid];
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int r = 0; r < 32; ++r) s += dv[r] * yi[h * 32 + r];
  if (h > 0) part[h - 1][c] = s;
  __syncthreads();
  if (h == 0) xi[c] = ((s + part[0][c]) + part[1][c]) + part[2][c];
  __syncthreads();
  if (kk == i) {
    if (tid < CT) x[(int64_t)i * CT + tid] = xi[tid];
    return;
  }
  s = 0.0;
#pragma unroll
  for (int r = 0; r < 32; ++r) s += tv[r] * xi[h * 32 + r];
  if (h > 0) part[h - 1][c] = s;
  __syncthreads();
  if (h == 0) y[(int64_t)kk * CT + c] -= ((s + part[0][c]) + part[1][c]) + part[2][c];
}

// ---- both sweeps as ONE persistent kernel ---------------------------------------------------------------------
// CTA i owns tile row i of L (forward: y_i = Linv_ii (w_i - sum_{k<i} L_ik y_k)) and tile column i (backward:
// x_i = Linv_ii' (y_i - sum_{k>i} L_ki' x_k)).  The CTAs are co-resident (cooperative launch, one per SM) and hand the
// vectors over through global memory with the flag INSIDE the data: an element travels as one 16-byte store
// {low word, flag, high word, flag} (flag = the launch's epoch, so nothing is ever cleared) and a consumer polls the
// element itself until both flags match -- no fence, no separate flag store, no extra barrier on the dependent chain,
// which is one such hand-over + two 128 x 128 products per step instead of one kernel launch per step;
// everything off the chain (CTA i consuming y_k for k < i - 1) runs as early as its input exists.  The tile a CTA
// will need next is in registers before it waits for the vector that goes with it; Linv_ii sits in shared memory.
// Sums in FP64 whatever the storage type T of the factor.
template <typename T> struct swt;
template <> struct swt<float> { static constexpr int TPR = 2; };    // threads per row (forward) / per column (backward)
template <> struct swt<double> { static constexpr int TPR = 4; };
constexpr int SW_PAD = 4;
template <typename T> constexpr int sweep_threads() { return CT * swt<T>::TPR; }
template <typename T> constexpr int sweep_smem() { return CT * (CT + SW_PAD) * (int)sizeof(T); }

// a double and its flag as ONE 16-byte store / load; each 8-byte half carries the flag, so a reader that sees both
// flags has both words whatever the granularity at which the store became visible
__device__ __forceinline__ void ll_store(uint4* p, double v, unsigned flag) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned)__double2loint(v)), "r"(flag),
               "r"((unsigned)__double2hiint(v)), "r"(flag)
               : "memory");
}
__device__ __forceinline__ bool ll_load(const uint4* p, unsigned flag, double& v) {  // bounded (~1 s)
  const long long t0 = clock64();
  for (;;) {
    unsigned a, fa, b, fb;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(fa), "=r"(b), "=r"(fb) : "l"(p) : "memory");
    if (fa == flag && fb == flag) {
      v = __hiloint2double((int)b, (int)a);
      return true;
    }
    if (clock64() - t0 > 2000000000ll) {
      v = 0.0;
      return false;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(CT * swt<T>::TPR, 1)
k_chol_sweep(const T* __restrict__ L, int64_t ld, const T* __restrict__ Dinv, const double* __restrict__ w,
             uint4* yq, uint4* xq, double* x, int nb, unsigned flag, int* __restrict__ info) {
  constexpr int TPR = swt<T>::TPR, EPT = CT / TPR, NT = CT * TPR, LDS = CT + SW_PAD;
  extern __shared__ __align__(16) unsigned char sm_raw[];
  T* Ds = reinterpret_cast<T*>(sm_raw);  // Linv_ii, rows LDS apart
  __shared__ double vec[CT], mine[CT], part[TPR][CT];
  __shared__ int ok_sh;
  const int tid = threadIdx.x, i = blockIdx.x;
  if (tid == 0) ok_sh = 1;
  {  // Linv_ii -> shared memory (off the dependent chain)
    constexpr int EPP = 16 / (int)sizeof(T), PR = CT / EPP;
    const T* D = Dinv + (int64_t)i * CT * CT;
    for (int e = tid; e < CT * PR; e += NT) {
      const int r = e / PR, c = (e - r * PR) * EPP;
      cp_async16(Ds + r * LDS + c, D + r * CT + c);
    }
    cp_async_commit();
  }
  // wait for vector k of a sweep (each of 128 threads polls its own element), bring it into vec[]
  auto get_vec = [&](const uint4* vq, int k) {
    __syncthreads();  // everybody is done with the previous contents of vec
    if (tid < CT) {
      double v;
      if (!ll_load(vq + (int64_t)k * CT + tid, flag, v)) ok_sh = 0;
      vec[tid] = v;
    }
    __syncthreads();
  };
  // ---- forward: thread (r, pt) holds row r, columns pt * EPT ... of the current tile
  {
    const int r = tid / TPR, pt = tid - r * TPR;
    T t[EPT];
    auto load_tile = [&](int k) {
      const T* src = L + ((int64_t)i * CT + r) * ld + (int64_t)k * CT + pt * EPT;
#pragma unroll
      for (int q = 0; q < EPT * (int)sizeof(T) / 16; ++q)
        *reinterpret_cast<uint4*>(&t[q * (16 / (int)sizeof(T))]) = __ldg(reinterpret_cast<const uint4*>(src) + q);
    };
    double s = 0.0;
    if (i > 0) load_tile(0);
    for (int k = 0; k < i; ++k) {
      get_vec(yq, k);
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
      for (int c = 0; c < EPT; c += 4) {
        a0 += (double)t[c] * vec[pt * EPT + c];
        a1 += (double)t[c + 1] * vec[pt * EPT + c + 1];
        a2 += (double)t[c + 2] * vec[pt * EPT + c + 2];
        a3 += (double)t[c + 3] * vec[pt * EPT + c + 3];
      }
      s += (a0 + a1) + (a2 + a3);
      if (k + 1 < i) load_tile(k + 1);  // in flight while the next vector is awaited
    }
#pragma unroll
    for (int o = 1; o < TPR; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (pt == 0) mine[r] = w[(int64_t)i * CT + r] - s;
    cp_async_wait<0>();
    __syncthreads();
    double a0 = 0.0, a1 = 0.0;
#pragma unroll
    for (int c = 0; c < EPT; c += 2) {
      a0 += (double)Ds[r * LDS + pt * EPT + c] * mine[pt * EPT + c];
      a1 += (double)Ds[r * LDS + pt * EPT + c + 1] * mine[pt * EPT + c + 1];
    }
    double yv = a0 + a1;
#pragma unroll
    for (int o = 1; o < TPR; o <<= 1) yv += __shfl_xor_sync(0xffffffffu, yv, o);
    __syncthreads();  // everybody has read mine[] (w_i') before it becomes y_i
    if (pt == 0) {
      mine[r] = yv;
      ll_store(yq + (int64_t)i * CT + r, yv, flag);
    }
    __syncthreads();
  }
  // ---- backward: thread (c, g) holds column c, rows g * EPT ... of the current tile L_ki
  {
    const int c = tid % CT, g = tid / CT;
    T t[EPT];
    auto load_tile = [&](int k) {
      const T* src = L + ((int64_t)k * CT + g * EPT) * ld + (int64_t)i * CT + c;
#pragma unroll
      for (int q = 0; q < EPT; ++q) t[q] = __ldg(src + (int64_t)q * ld);
    };
    double s = 0.0;
    if (i + 1 < nb) load_tile(nb - 1);
    for (int k = nb - 1; k > i; --k) {
      get_vec(xq, k);
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
      for (int q = 0; q < EPT; q += 4) {
        a0 += (double)t[q] * vec[g * EPT + q];
        a1 += (double)t[q + 1] * vec[g * EPT + q + 1];
        a2 += (double)t[q + 2] * vec[g * EPT + q + 2];
        a3 += (double)t[q + 3] * vec[g * EPT + q + 3];
      }
      s += (a0 + a1) + (a2 + a3);
      if (k - 1 > i) load_tile(k - 1);
    }
    part[g][c] = s;
    __syncthreads();
    if (tid < CT) {
      double tot = part[0][tid];
#pragma unroll
      for (int q = 1; q < TPR; ++q) tot += part[q][tid];
      vec[tid] = mine[tid] - tot;  // y_i - sum_k L_ki' x_k
    }
    __syncthreads();
    double a0 = 0.0, a1 = 0.0;
#pragma unroll
    for (int q = 0; q < EPT; q += 2) {
      a0 += (double)Ds[(g * EPT + q) * LDS + c] * vec[g * EPT + q];
      a1 += (double)Ds[(g * EPT + q + 1) * LDS + c] * vec[g * EPT + q + 1];
    }
    part[g][c] = a0 + a1;
    __syncthreads();
    if (tid < CT) {
      double tot = part[0][tid];
#pragma unroll
      for (int q = 1; q < TPR; ++q) tot += part[q][tid];
      ll_store(xq + (int64_t)i * CT + tid, tot, flag);
      x[(int64_t)i * CT + tid] = tot;
    }
    __syncthreads();
    if (tid == 0 && !ok_sh) atomicCAS(info, 0, -4);  // a vector never arrived
  }
}

// BAGPU_CHOL_NO_TC=1: the FP32 trailing update on the legacy tensor path (mma.sync, three TF32 terms) -- the A/B switch
inline bool tc_enabled() {
  static const bool v = getenv("BAGPU_CHOL_NO_TC") == nullptr;
  return v;
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
  BA_CUDA(cudaMalloc(reinterpret_cast<void**>(&P.d_x), sizeof(double) * (size_t)cn));
  BA_CUDA(cudaMalloc(reinterpret_cast<void**>(&P.d_info), sizeof(int)));
  {  // the panel stream must get SMs while the trailing update fills the machine: highest priority
    int lo = 0, hi = 0;
    BA_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    BA_CUDA(cudaStreamCreateWithPriority(&P.side, cudaStreamNonBlocking, hi));
  }
  BA_CUDA(cudaEventCreateWithFlags(&P.ev_col, cudaEventDisableTiming));
  BA_CUDA(cudaEventCreateWithFlags(&P.ev_panel, cudaEventDisableTiming));
  BA_CUDA(cudaEventCreateWithFlags(&P.ev_join, cudaEventDisableTiming));
  // per device (a process may hold handles on several devices): set once per plan
  BA_CUDA((cudaFuncSetAttribute(k_chol_syrk<double, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem<double>())));
  BA_CUDA((cudaFuncSetAttribute(k_chol_syrk<double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem<double>())));
  BA_CUDA((cudaFuncSetAttribute(k_chol_trsm<double, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem<double>())));
  BA_CUDA((cudaFuncSetAttribute(k_chol_trsm<double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem<double>())));
  BA_CUDA((cudaFuncSetAttribute(k_chol_syrk<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem<float>())));
  BA_CUDA((cudaFuncSetAttribute(k_chol_trsm<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem<float>())));
  BA_CUDA(cudaFuncSetAttribute(k_chol_sweep<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, sweep_smem<float>()));
  BA_CUDA(cudaFuncSetAttribute(k_chol_sweep<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, sweep_smem<double>()));
  BA_CUDA(cudaMalloc(&P.d_vq, 2 * sizeof(uint4) * (size_t)cn));
  BA_CUDA(cudaMemset(P.d_vq, 0, 2 * sizeof(uint4) * (size_t)cn));
  {
    int dev = 0;
    BA_CUDA(cudaGetDevice(&dev));
    BA_CUDA(cudaDeviceGetAttribute(&P.sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  BA_CUDA(cudaFuncSetAttribute(k_chol_syrk_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
  BA_CUDA(cudaFuncSetAttribute(k_chol_potrf<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, PO_SMEM));
  BA_CUDA(cudaFuncSetAttribute(k_chol_potrf<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, PO_SMEM));
  P.attrs_set = true;
  return BA_OK;
}

void chol_plan_release(chol_plan& P) {
  cudaFree(P.d_Dinv);
  cudaFree(P.d_y);
  cudaFree(P.d_w);
  cudaFree(P.d_x);
  cudaFree(P.d_info);
  cudaFree(P.d_prof);
  cudaFree(P.d_ctl);
  cudaFree(P.d_cnt);
  cudaFree(P.d_vq);
  for (int r = 0; r < CHOL_RMAX; ++r)
    if (P.peer_ipc[r])
      for (int j = 0; j < 3; ++j)
        if (P.peer_open[3 * r + j]) cudaIpcCloseMemHandle(P.peer_open[3 * r + j]);
  if (P.solve_graph) cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(P.solve_graph));
  if (P.side) cudaStreamDestroy(P.side);
  if (P.ev_col) cudaEventDestroy(P.ev_col);
  if (P.ev_panel) cudaEventDestroy(P.ev_panel);
  if (P.ev_join) cudaEventDestroy(P.ev_join);
  P = chol_plan();
}

namespace {
template <typename T>
int chol_factor_t(ba_handle* h, chol_plan& P, T* A, cudaStream_t s, int* info_host) {
  const int64_t cn = P.cn, ld = cn;
  T* const Dinv = reinterpret_cast<T*>(P.d_Dinv);
  constexpr int NKP = CT / mkt<T>::KC, GEMM_SMEM = gemm_smem<T>();
  const int nb = (int)(cn / CT);
  static const bool no_lookahead = getenv("BAGPU_CHOL_NO_LOOKAHEAD") != nullptr;
  chol_peers solo = {};
  solo.R = 1;
  BA_CUDA(cudaMemsetAsync(P.d_info, 0, sizeof(int), s));
  static const bool want_prof = getenv("BAGPU_POTRF_PROF") != nullptr;
  long long* prof = nullptr;
  if (want_prof) {  // development aid: phase cycle counts of the first diagonal block
    if (!P.d_prof) BA_CUDA(cudaMalloc(reinterpret_cast<void**>(&P.d_prof), 8 * sizeof(long long)));
    prof = P.d_prof;
  }
  // Two panels per trailing update: the rank-128 updates of panels k and k + 1 are applied together as ONE
  // rank-256 update (the panels are adjacent columns), which halves the read-modify-write traffic of the trailing
  // matrix, the per-tile prologue/epilogue and the number of partial last waves.  Per pair (k even):
  //   panel k final -> column k + 1 updated with panel k (rank 128) -> potrf / trsm of panel k + 1
  //   -> columns k + 2, k + 3 updated with both panels (rank 256) -> [side stream: the next pair's panel work]
  //   || [main stream: the rest of the trailing matrix, rank 256]
  static const bool single = getenv("BAGPU_CHOL_SINGLE_PANEL") != nullptr;  // A/B: one panel per update
  const int step = (single || nb < 24) ? 1 : 2;  // small matrices are bound by the chain of diagonal blocks: no gain
  auto potrf = [&](int k, cudaStream_t st, long long* pf) {
    k_chol_potrf<T><<<1, PO_THREADS, PO_SMEM, st>>>(A, ld, k, Dinv, P.d_info, pf, solo);
  };
  auto trsm = [&](int k, cudaStream_t st) {
    if (k + 1 < nb) k_chol_trsm<T, false><<<2 * (nb - k - 1), GEMM_THREADS, GEMM_SMEM, st>>>(A, ld, k, Dinv, 0, solo, nullptr);
  };
  // columns [j0, j0 + ncol) of the trailing matrix -= (panels k .. k + npan - 1) (...)'
  auto update = [&](int k, int npan, int j0, int ncol, cudaStream_t st) {
    if (j0 >= nb || ncol <= 0) return;
    if constexpr (sizeof(T) == 4) {
      if (tc_enabled()) {  // FP32: 128 x 128 tiles on the tcgen05 tensor cores, persistent CTAs (one per SM)
        const int nc = std::min(ncol, nb - j0);
        int64_t ntiles = 0;
        for (int j = j0; j < j0 + nc; ++j) ntiles += nb - j;
        // a few tiles per CTA (not the whole update): the high-priority panel kernels of the look-ahead get SMs as
        // CTAs retire, while the operand ring still runs across the tile boundaries inside a CTA
        static const int per_cta = getenv("BAGPU_TC_TILES_PER_CTA") ? std::max(1, atoi(getenv("BAGPU_TC_TILES_PER_CTA"))) : 8;
        const int64_t grid = std::max<int64_t>(std::min<int64_t>(ntiles, P.sm_count), (ntiles + per_cta - 1) / per_cta);
        k_chol_syrk_tc<<<(unsigned)grid, TC_THREADS, TC_SMEM, st>>>(
            A, ld, k, j0, nc, nb, P.d_info, (CT / TC_KC) * npan);
        return;
      }
    }
    k_chol_syrk<T, false><<<dim3(2 * (nb - j0), std::min(ncol, nb - j0)), GEMM_THREADS, GEMM_SMEM, st>>>(
        A, ld, k, j0, nb, 0, solo, nullptr, NKP * npan, 0);
  };
  // the panel work of a pair starting at k (its first column is already up to date), on stream st
  auto pair_panels = [&](int k, cudaStream_t st, long long* pf) {
    potrf(k, st, pf);
    trsm(k, st);
    if (step == 2 && k + 1 < nb) {
      update(k, 1, k + 1, 1, st);
      potrf(k + 1, st, nullptr);
      trsm(k + 1, st);
    }
  };
  pair_panels(0, s, prof);
  for (int k = 0; k + step < nb; k += step) {
    const int npan = std::min(step, nb - k);
    update(k, npan, k + step, step, s);  // the next pair's columns first
    if (k + 2 * step < nb && !no_lookahead) {
      BA_CUDA(cudaEventRecord(P.ev_col, s));
      BA_CUDA(cudaStreamWaitEvent(P.side, P.ev_col, 0));
      pair_panels(k + step, P.side, nullptr);
      BA_CUDA(cudaEventRecord(P.ev_panel, P.side));
      update(k, npan, k + 2 * step, nb, s);  // the rest of the trailing matrix under the panel work
      BA_CUDA(cudaStreamWaitEvent(s, P.ev_panel, 0));
    } else {
      update(k, npan, k + 2 * step, nb, s);
      pair_panels(k + step, s, nullptr);
    }
  }
  BA_CUDA(cudaGetLastError());
  if (prof) {
    long long hp[8];
    BA_CUDA(cudaMemcpyAsync(hp, prof, sizeof hp, cudaMemcpyDeviceToHost, s));
    BA_CUDA(cudaStreamSynchronize(s));
    fprintf(stderr, "[bagpu] potrf cycles: load %lld | (a) %lld (b) %lld (c) %lld | store L %lld | inv16 %lld | doubling %lld "
            "| store Linv %lld\n", hp[0], hp[1], hp[2], hp[3], hp[4], hp[5], hp[6], hp[7]);
  }
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
}  // namespace

int chol_factor(ba_handle* h, chol_plan& P, double* A, cudaStream_t s, int* info_host) {
  return chol_factor_t<double>(h, P, A, s, info_host);
}
int chol_factor32(ba_handle* h, chol_plan& P, float* A32, cudaStream_t s, int* info_host) {
  return chol_factor_t<float>(h, P, A32, s, info_host);
}

// ---- distributed factorisation ------------------------------------------------------------------------------
namespace {
struct peer_record {  // what the ranks exchange (NCCL all-gather of the raw bytes)
  long long pid;
  int device, ok;
  void *S, *D, *ctl;
  cudaIpcMemHandle_t hS, hD, hC;
};
}  // namespace

int chol_dist_setup(ba_handle* h, chol_plan& P, void* A) {
  P.dist_ready = false;
  static const bool off = getenv("BAGPU_CHOL_REPLICATED") != nullptr || getenv("BAGPU_NO_P2P") != nullptr;
  const int R = h->nranks;
  if (R < 2 || R > CHOL_RMAX || !h->comm || P.cn / CT > CHOL_NBMAX) return BA_OK;
  BA_CUDA(cudaSetDevice(h->device));
  const size_t ctl_n = (size_t)CHOL_NBMAX * (1 + 16);
  if (!P.d_ctl) {
    BA_CUDA(cudaMalloc(reinterpret_cast<void**>(&P.d_ctl), ctl_n * sizeof(unsigned long long)));
    BA_CUDA(cudaMemset(P.d_ctl, 0, ctl_n * sizeof(unsigned long long)));
    BA_CUDA(cudaMalloc(reinterpret_cast<void**>(&P.d_cnt), sizeof(int)));
    BA_CUDA(cudaMemset(P.d_cnt, 0, sizeof(int)));
  }
  peer_record mine = {};
  mine.pid = (long long)getpid();
  mine.device = h->device;
  mine.S = A; mine.D = P.d_Dinv; mine.ctl = P.d_ctl;
  mine.ok = off ? 0 : 1;
  if (mine.ok && (cudaIpcGetMemHandle(&mine.hS, A) != cudaSuccess || cudaIpcGetMemHandle(&mine.hD, P.d_Dinv) != cudaSuccess ||
                  cudaIpcGetMemHandle(&mine.hC, P.d_ctl) != cudaSuccess)) {
    cudaGetLastError();
    mine.ok = 0;
  }
  std::vector<peer_record> all((size_t)R);
  int rc = allgather_host(h, &mine, all.data(), sizeof(peer_record));
  if (rc) return rc;
  int ok = 1;
  for (const peer_record& p : all) ok &= p.ok;
  chol_peers V = {};
  V.R = R; V.q = h->rank; V.epoch = 0;
  if (ok) {
    for (int r = 0; r < R && ok; ++r) {
      const peer_record& p = all[(size_t)r];
      if (r == h->rank) {
        V.S[r] = A; V.D[r] = P.d_Dinv; V.ctl[r] = P.d_ctl;
      } else if (p.pid == mine.pid) {  // same process (ba_create_multi): raw pointers + peer access
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, h->device, p.device) != cudaSuccess || !can) { ok = 0; break; }
        const cudaError_t e = cudaDeviceEnablePeerAccess(p.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { ok = 0; break; }
        cudaGetLastError();
        V.S[r] = p.S; V.D[r] = p.D;
        V.ctl[r] = static_cast<unsigned long long*>(p.ctl);
      } else {
        void *a = nullptr, *b = nullptr, *c = nullptr;
        if (cudaIpcOpenMemHandle(&a, p.hS, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
            cudaIpcOpenMemHandle(&b, p.hD, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
            cudaIpcOpenMemHandle(&c, p.hC, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          cudaGetLastError();
          ok = 0;
          break;
        }
        P.peer_ipc[r] = true;
        P.peer_open[3 * r] = a; P.peer_open[3 * r + 1] = b; P.peer_open[3 * r + 2] = c;
        V.S[r] = a; V.D[r] = b;
        V.ctl[r] = static_cast<unsigned long long*>(c);
      }
    }
  }
  // every rank must take the same path: agree on the outcome
  int mine_ok = ok;
  std::vector<int> oks((size_t)R);
  if ((rc = allgather_host(h, &mine_ok, oks.data(), sizeof(int)))) return rc;
  for (int v : oks) ok &= v;
  if (!ok) return BA_OK;  // replicated factorisation
  P.peers = V;
  P.dist_ready = true;
  return BA_OK;
}

namespace {
template <typename T>
int chol_factor_dist_t(ba_handle* h, chol_plan& P, T* A, cudaStream_t s) {
  const int64_t cn = P.cn, ld = cn;
  T* const Dinv = reinterpret_cast<T*>(P.d_Dinv);
  constexpr int NKP = CT / mkt<T>::KC, GEMM_SMEM = gemm_smem<T>();
  const int nb = (int)(cn / CT);
  chol_peers& V = P.peers;
  const int R = V.R, q = V.q;
  V.epoch += 1;  // the same count on every rank: all of them factorise the same sequence of matrices
  auto first_own = [&](int from) { return from + (((q - from) % R) + R) % R; };          // first own tile row >= from
  auto count_own = [&](int from) { const int i0 = first_own(from); return i0 < nb ? (nb - 1 - i0) / R + 1 : 0; };
  BA_CUDA(cudaMemsetAsync(P.d_info, 0, sizeof(int), s));
  // diagonal block k and the panel below it, on stream st
  auto panel = [&](int k, cudaStream_t st) {
    if (k % R == q) k_chol_potrf<T><<<1, PO_THREADS, PO_SMEM, st>>>(A, ld, k, Dinv, P.d_info, nullptr, V);
    else k_chol_fetch_diag<T><<<16, 256, 0, st>>>(A, ld, k, Dinv, V, k % R, P.d_info);
    const int n = count_own(k + 1);
    if (n > 0) k_chol_trsm<T, true><<<2 * n, GEMM_THREADS, GEMM_SMEM, st>>>(A, ld, k, Dinv, first_own(k + 1), V, P.d_cnt);
    else k_chol_signal<<<1, 32, 0, st>>>(k, V);
  };
  // as in the single-GPU factorisation, two panels per trailing update (rank 256) above 24 tile rows
  static const bool single = getenv("BAGPU_CHOL_SINGLE_PANEL") != nullptr;
  const int step = (single || nb < 24) ? 1 : 2;
  // this rank's tiles of the columns [j0, j0 + ncol) -= (panels k .. k + npan - 1) (...)'
  auto update = [&](int k, int npan, int j0, int ncol, cudaStream_t st) {
    if (j0 >= nb || ncol <= 0) return;
    const int n = count_own(j0);
    if (n <= 0) return;
    k_chol_syrk<T, true><<<dim3(2 * n, std::min(ncol, nb - j0)), GEMM_THREADS, GEMM_SMEM, st>>>(
        A, ld, k, j0, nb, first_own(j0), V, P.d_info, NKP * npan, k + npan - 1);
  };
  auto pair_panels = [&](int k, cudaStream_t st) {
    panel(k, st);
    if (step == 2 && k + 1 < nb) {
      update(k, 1, k + 1, 1, st);
      panel(k + 1, st);
    }
  };
  pair_panels(0, s);
  for (int k = 0; k + step < nb; k += step) {
    const int npan = std::min(step, nb - k);
    update(k, npan, k + step, step, s);  // the next pair's columns first (they gate its panel work)
    BA_CUDA(cudaEventRecord(P.ev_col, s));
    BA_CUDA(cudaStreamWaitEvent(P.side, P.ev_col, 0));
    pair_panels(k + step, P.side);
    BA_CUDA(cudaEventRecord(P.ev_panel, P.side));
    update(k, npan, k + 2 * step, nb, s);  // the rest of this rank's trailing matrix under the panel work
    BA_CUDA(cudaStreamWaitEvent(s, P.ev_panel, 0));
  }
  k_chol_wait_all<<<1, 256, 0, s>>>(nb, V, P.d_info);
  BA_CUDA(cudaGetLastError());
  return BA_OK;
}

template <typename T>
int chol_solve_t(ba_handle* h, chol_plan& P, const T* L, const double* b, double* x, cudaStream_t s) {
  const int64_t cn = P.cn;
  const T* const Dinv = reinterpret_cast<const T*>(P.d_Dinv);
  constexpr bool is32 = sizeof(T) == 4;
  // one persistent kernel for both sweeps when its cn / 128 CTAs can be co-resident (one per SM); BAGPU_SWEEP_STEPS=1:
  // the per-step kernels below (one launch per step, as a CUDA graph) -- the A/B switch and the route for larger systems
  const bool by_steps = getenv("BAGPU_SWEEP_STEPS") != nullptr;  // (read per call: the tests switch it)
  if (!by_steps && !P.sweep_off && (int)(cn / CT) <= P.sm_count) {
    int nbi = (int)(cn / CT);
    int64_t ldv = cn;
    if (++P.sweep_epoch == 0) ++P.sweep_epoch;  // (0 is the flag of the freshly cleared buffers)
    unsigned ep = P.sweep_epoch;
    uint4* yq = reinterpret_cast<uint4*>(P.d_vq);
    uint4* xq = yq + cn;
    void* args[] = {(void*)&L, (void*)&ldv, (void*)&Dinv, (void*)&b, (void*)&yq, (void*)&xq, (void*)&x, (void*)&nbi,
                    (void*)&ep, (void*)&P.d_info};
    const cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(&k_chol_sweep<T>), dim3((unsigned)nbi),
                                                      dim3((unsigned)sweep_threads<T>()), args, (size_t)sweep_smem<T>(), s);
    if (e == cudaSuccess) return BA_OK;
    cudaGetLastError();   // e.g. the grid cannot be co-resident on this device / in this context: per-step kernels
    P.sweep_off = true;
  }
  const int nb = (int)(cn / CT);
  static const bool no_graph = getenv("BAGPU_NO_GRAPH") != nullptr || getenv("BAGPU_DEBUG_SYNC") != nullptr;
  // the sweeps always run from P.d_w into P.d_x: fixed pointers, so the 2 nb launches are captured once per
  // (matrix, plan) and replayed as one graph launch
  BA_CUDA(cudaMemcpyAsync(P.d_w, b, sizeof(double) * (size_t)cn, cudaMemcpyDeviceToDevice, s));
  auto sweeps = [&]() {
    for (int k = 0; k < nb; ++k) k_chol_fwd<T><<<nb - k, SV_THREADS, 0, s>>>(L, cn, Dinv, P.d_w, P.d_y, k);
    for (int i = nb - 1; i >= 0; --i) k_chol_bwd<T><<<i + 1, SV_THREADS, 0, s>>>(L, cn, Dinv, P.d_y, P.d_x, i);
  };
  if (!no_graph && !P.graph_off && (!P.solve_graph || P.solve_graph_A != L || P.solve_graph_32 != is32)) {
    if (P.solve_graph) cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(P.solve_graph));
    P.solve_graph = nullptr;
    cudaGraph_t g = nullptr;
    if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();  // e.g. the legacy default stream: plain launches from now on
      P.graph_off = true;
    } else {
      sweeps();
      BA_CUDA(cudaStreamEndCapture(s, &g));
      cudaGraphExec_t ge = nullptr;
      BA_CUDA(cudaGraphInstantiate(&ge, g, 0));
      cudaGraphDestroy(g);
      P.solve_graph = ge;
      P.solve_graph_A = L;
      P.solve_graph_32 = is32;
    }
  }
  if (P.solve_graph && !no_graph) BA_CUDA(cudaGraphLaunch(reinterpret_cast<cudaGraphExec_t>(P.solve_graph), s));
  else sweeps();
  BA_CUDA(cudaMemcpyAsync(x, P.d_x, sizeof(double) * (size_t)cn, cudaMemcpyDeviceToDevice, s));
  BA_CUDA(cudaGetLastError());
  return BA_OK;
}
}  // namespace

int chol_factor_dist(ba_handle* h, chol_plan& P, double* A, cudaStream_t s) { return chol_factor_dist_t<double>(h, P, A, s); }
int chol_solve(ba_handle* h, chol_plan& P, const double* L, const double* b, double* x, cudaStream_t s) {
  return chol_solve_t<double>(h, P, L, b, x, s);
}
int chol_solve32(ba_handle* h, chol_plan& P, const float* L32, const double* b, double* x, cudaStream_t s) {
  return chol_solve_t<float>(h, P, L32, b, x, s);
}

}  // namespace ba

// ---- debug / benchmark entry: factor and solve a caller-supplied SPD matrix (tests, roofline of the factorisation) --
namespace {
template <typename T>
int dbg_chol_t(int device, int64_t n, const double* A_rowmajor, const double* b, double* x, double* L_out,
               float* factor_ms, float* solve_ms) {
  if (n < 1 || !A_rowmajor || !b || !x) return BA_ERR_ARG;
  ba_handle hh;
  ba_handle* h = &hh;
  if (cudaSetDevice(device) != cudaSuccess) return BA_ERR_CUDA;
  const int64_t cn = ba::chol_padded(n);
  ba::chol_plan P;
  T* dA = nullptr;
  double* db = nullptr;
  cudaStream_t s = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
  int rc = BA_OK, info = 0;
  auto done = [&](int code) {
    if (code == BA_ERR_CUDA && !hh.err.empty()) fprintf(stderr, "[bagpu] ba_dbg_chol: %s\n", hh.err.c_str());
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
  if (cudaMalloc(reinterpret_cast<void**>(&dA), sizeof(T) * (size_t)(cn * cn)) != cudaSuccess) return done(BA_ERR_CUDA);
  if (cudaMalloc(reinterpret_cast<void**>(&db), sizeof(double) * (size_t)cn) != cudaSuccess) return done(BA_ERR_CUDA);
  if ((rc = ba::chol_plan_init(h, P, cn))) return done(rc);
  // padded with the identity, in the element type of the factorisation
  std::vector<T> host((size_t)(cn * cn), (T)0);
  for (int64_t r = 0; r < n; ++r)
    for (int64_t c = 0; c < n; ++c) host[(size_t)(r * cn + c)] = (T)A_rowmajor[r * n + c];
  for (int64_t r = n; r < cn; ++r) host[(size_t)(r * cn + r)] = (T)1;
  cudaMemcpyAsync(dA, host.data(), sizeof(T) * host.size(), cudaMemcpyHostToDevice, s);
  cudaMemsetAsync(db, 0, sizeof(double) * (size_t)cn, s);
  cudaMemcpyAsync(db, b, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s);
  cudaStreamSynchronize(s);
  cudaEventRecord(e0, s);
  if (sizeof(T) == 8) rc = ba::chol_factor(h, P, reinterpret_cast<double*>(dA), s, nullptr);
  else rc = ba::chol_factor32(h, P, reinterpret_cast<float*>(dA), s, nullptr);
  cudaEventRecord(e1, s);
  if (!rc) {
    if (sizeof(T) == 8) rc = ba::chol_solve(h, P, reinterpret_cast<double*>(dA), db, db, s);
    else rc = ba::chol_solve32(h, P, reinterpret_cast<float*>(dA), db, db, s);
  }
  cudaEventRecord(e2, s);
  if (rc) return done(rc);
  {
    const cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
      fprintf(stderr, "[bagpu] ba_dbg_chol: %s\n", cudaGetErrorString(e));
      return done(BA_ERR_CUDA);
    }
  }
  cudaMemcpy(&info, P.d_info, sizeof(int), cudaMemcpyDeviceToHost);
  if (factor_ms) cudaEventElapsedTime(factor_ms, e0, e1);
  if (solve_ms) cudaEventElapsedTime(solve_ms, e1, e2);
  cudaMemcpy(x, db, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost);
  if (L_out) {
    cudaMemcpy(host.data(), dA, sizeof(T) * host.size(), cudaMemcpyDeviceToHost);
    for (int64_t r = 0; r < n; ++r)
      for (int64_t c = 0; c < n; ++c) L_out[r * n + c] = (double)host[(size_t)(r * cn + c)];
  }
  if (cudaGetLastError() != cudaSuccess) return done(BA_ERR_CUDA);
  return done(info ? BA_ERR_NUMERIC : BA_OK);
}
}  // namespace

extern "C" int ba_dbg_chol(int device, int64_t n, const double* A_rowmajor, const double* b, double* x, double* L_out,
                           float* factor_ms, float* solve_ms) {
  return dbg_chol_t<double>(device, n, A_rowmajor, b, x, L_out, factor_ms, solve_ms);
}
// the mixed-precision factor: A rounded to FP32, FP32 factor (returned widened), x = (L32 L32')^-1 b with FP64 sweeps
extern "C" int ba_dbg_chol32(int device, int64_t n, const double* A_rowmajor, const double* b, double* x, double* L_out,
                             float* factor_ms, float* solve_ms) {
  return dbg_chol_t<float>(device, n, A_rowmajor, b, x, L_out, factor_ms, solve_ms);
}
