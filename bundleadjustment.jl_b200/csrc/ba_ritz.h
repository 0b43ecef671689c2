// ba_ritz.h -- host-side small dense algebra for the PCG deflation space: eigenpairs of the Lanczos
// tridiagonal that a PCG solve produces for free, and the selection of a well-conditioned subset of the
// Ritz vectors (converged Ritz values come with "ghost" copies).  Header-only, no CUDA: unit-tested on
// the CPU through ba_dbg_tridiag_eig / ba_dbg_tridiag_smallest / ba_dbg_select_columns (tests/test_host.py).
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

namespace ba {

// Lanczos tridiagonal of preconditioned CG (Z~' S Z~ with Z~_j = z_j / sqrt(r_j.z_j)):
//   T_jj = 1/a_j + b_{j-1}/a_{j-1},   T_{j,j+1} = -sqrt(b_j)/a_j      (a = alpha, b = beta of CG)
inline void lanczos_tridiagonal(const double* alpha, const double* beta, int m, std::vector<double>& d,
                                std::vector<double>& e) {
  d.assign((size_t)m, 0.0);
  e.assign((size_t)m, 0.0);  // e[j] couples j and j+1; e[m-1] unused
  for (int j = 0; j < m; ++j) {
    d[(size_t)j] = 1.0 / alpha[j] + (j ? beta[j - 1] / alpha[j - 1] : 0.0);
    if (j + 1 < m) e[(size_t)j] = -std::sqrt(beta[j]) / alpha[j];
  }
}

// Eigen-decomposition of a symmetric tridiagonal matrix (implicit QL with Wilkinson shifts, the classic
// tql2 scheme).  d: diagonal (m) -> eigenvalues ascending; e: off-diagonal (e[j] couples j, j+1);
// V (m x m, column-major, V[i + m*j]) -> eigenvectors in columns.  Returns false if an eigenvalue fails to
// converge in 60 sweeps.  want_vectors = false: eigenvalues only (O(m^2) instead of O(m^3)), V left empty.
inline bool tridiag_eig(std::vector<double>& d, std::vector<double> e, int m, std::vector<double>& V,
                        bool want_vectors = true) {
  V.assign(want_vectors ? (size_t)m * m : 0, 0.0);
  if (want_vectors)
    for (int i = 0; i < m; ++i) V[(size_t)i + (size_t)m * i] = 1.0;
  if (m == 0) return true;
  e[(size_t)m - 1] = 0.0;
  for (int l = 0; l < m; ++l) {
    int iter = 0, mm;
    do {
      for (mm = l; mm < m - 1; ++mm) {
        const double dd = std::fabs(d[(size_t)mm]) + std::fabs(d[(size_t)mm + 1]);
        if (std::fabs(e[(size_t)mm]) <= 2.220446049250313e-16 * dd) break;
      }
      if (mm != l) {
        if (iter++ == 60) return false;
        double g = (d[(size_t)l + 1] - d[(size_t)l]) / (2.0 * e[(size_t)l]);
        double r = std::hypot(g, 1.0);
        g = d[(size_t)mm] - d[(size_t)l] + e[(size_t)l] / (g + (g >= 0 ? std::fabs(r) : -std::fabs(r)));
        double s = 1.0, c = 1.0, p = 0.0;
        int i;
        for (i = mm - 1; i >= l; --i) {
          double f = s * e[(size_t)i], b = c * e[(size_t)i];
          r = std::hypot(f, g);
          e[(size_t)i + 1] = r;
          if (r == 0.0) {
            d[(size_t)i + 1] -= p;
            e[(size_t)mm] = 0.0;
            break;
          }
          s = f / r;
          c = g / r;
          g = d[(size_t)i + 1] - p;
          r = (d[(size_t)i] - g) * s + 2.0 * c * b;
          p = s * r;
          d[(size_t)i + 1] = g + p;
          g = c * r - b;
          for (int k = 0; want_vectors && k < m; ++k) {
            double* vk = &V[(size_t)k];
            f = vk[(size_t)m * (i + 1)];
            vk[(size_t)m * (i + 1)] = s * vk[(size_t)m * i] + c * f;
            vk[(size_t)m * i] = c * vk[(size_t)m * i] - s * f;
          }
        }
        if (r == 0.0 && i >= l) continue;
        d[(size_t)l] -= p;
        e[(size_t)l] = g;
        e[(size_t)mm] = 0.0;
      }
    } while (mm != l);
  }
  // sort ascending, permuting the eigenvector columns
  std::vector<int> idx((size_t)m);
  for (int i = 0; i < m; ++i) idx[(size_t)i] = i;
  std::sort(idx.begin(), idx.end(), [&](int a, int b) { return d[(size_t)a] < d[(size_t)b]; });
  std::vector<double> d2((size_t)m), V2(V.size());
  for (int j = 0; j < m; ++j) {
    d2[(size_t)j] = d[(size_t)idx[(size_t)j]];
    if (want_vectors)
      std::copy(V.begin() + (size_t)m * idx[(size_t)j], V.begin() + (size_t)m * (idx[(size_t)j] + 1),
                V2.begin() + (size_t)m * j);
  }
  d.swap(d2);
  V.swap(V2);
  return true;
}

// One eigenvector of the symmetric tridiagonal (d, e) for the (computed) eigenvalue theta by inverse iteration:
// LU of T - theta I with partial pivoting (the pivot row keeps at most three entries), three solves from a
// fixed pseudo-random start.  x (m) is returned with unit 2-norm.
inline void tridiag_inverse_iteration(const std::vector<double>& d, const std::vector<double>& e, int m, double theta,
                                      double* x) {
  std::vector<double> u0((size_t)m), u1((size_t)m, 0.0), u2((size_t)m, 0.0), mult((size_t)m, 0.0);
  std::vector<char> swapped((size_t)m, 0);
  double tnorm = 0.0;
  for (int i = 0; i < m; ++i)
    tnorm = std::max(tnorm, std::fabs(d[(size_t)i]) + (i + 1 < m ? std::fabs(e[(size_t)i]) : 0.0) +
                                (i ? std::fabs(e[(size_t)i - 1]) : 0.0));
  const double tiny = 2.220446049250313e-16 * std::max(tnorm, 1e-300);
  double w0 = d[0] - theta, w1 = m > 1 ? e[0] : 0.0;  // working row: entries at columns i, i+1
  for (int i = 0; i + 1 < m; ++i) {
    const double sub = e[(size_t)i], dia = d[(size_t)i + 1] - theta, sup = (i + 2 < m) ? e[(size_t)i + 1] : 0.0;
    if (std::fabs(sub) > std::fabs(w0)) {  // the next row becomes the pivot row
      swapped[(size_t)i] = 1;
      u0[(size_t)i] = sub; u1[(size_t)i] = dia; u2[(size_t)i] = sup;
      const double mu = w0 / sub;
      mult[(size_t)i] = mu;
      w0 = w1 - mu * dia;
      w1 = -mu * sup;
    } else {
      if (std::fabs(w0) < tiny) w0 = tiny;
      u0[(size_t)i] = w0; u1[(size_t)i] = w1; u2[(size_t)i] = 0.0;
      const double mu = sub / w0;
      mult[(size_t)i] = mu;
      w0 = dia - mu * w1;
      w1 = sup;
    }
  }
  if (std::fabs(w0) < tiny) w0 = tiny;
  u0[(size_t)m - 1] = w0;
  unsigned long long lcg = 0x9E3779B97F4A7C15ull;
  for (int i = 0; i < m; ++i) {
    lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
    x[i] = 0.5 + (double)(lcg >> 11) * (1.0 / 9007199254740992.0);  // in [0.5, 1.5): no accidental orthogonality
  }
  for (int it = 0; it < 3; ++it) {
    for (int i = 0; i + 1 < m; ++i) {
      if (swapped[(size_t)i]) std::swap(x[i], x[i + 1]);
      x[i + 1] -= mult[(size_t)i] * x[i];
    }
    for (int i = m - 1; i >= 0; --i) {
      double t = x[i];
      if (i + 1 < m) t -= u1[(size_t)i] * x[i + 1];
      if (i + 2 < m) t -= u2[(size_t)i] * x[i + 2];
      x[i] = t / u0[(size_t)i];
    }
    double big = 0.0;
    for (int i = 0; i < m; ++i) big = std::max(big, std::fabs(x[i]));
    if (!(big > 0.0) || big != big) break;
    double n2 = 0.0;
    for (int i = 0; i < m; ++i) {
      x[i] /= big;
      n2 += x[i] * x[i];
    }
    const double inv = 1.0 / std::sqrt(n2);
    for (int i = 0; i < m; ++i) x[i] *= inv;
  }
}

// Sturm counts: cnt[s] = number of eigenvalues of the tridiagonal (d, e2 = squared off-diagonal) smaller than
// x[s], by the LDL' pivot recurrence q_i = (d_i - x) - e2_{i-1} / q_{i-1}.  NS shifts advance together so that
// the divisions -- the whole cost -- are independent and pipeline.
template <int NS>
inline void sturm_counts(const double* d, const double* e2, int m, const double* x, int* cnt, double pivmin) {
  double q[NS];
  for (int s = 0; s < NS; ++s) {
    q[s] = d[0] - x[s];
    cnt[s] = q[s] < 0.0;
  }
  for (int i = 1; i < m; ++i) {
    const double di = d[i], ei = e2[i - 1];
    for (int s = 0; s < NS; ++s) {
      double qq = q[s];
      if (std::fabs(qq) < pivmin) qq = -pivmin;
      qq = (di - x[s]) - ei / qq;
      cnt[s] += qq < 0.0;
      q[s] = qq;
    }
  }
}

// The k smallest eigenvalues (ascending, with multiplicity) by multisection on the Sturm count: each pass cuts
// the bracket of eigenvalue j into NS + 1 parts.  O(k m log(1/eps)) instead of the O(m^2) of a full QL sweep,
// and indifferent to the ghost copies a long CG run leaves in T.
inline void tridiag_smallest_values(const std::vector<double>& d, const std::vector<double>& e, int m, int k,
                                    std::vector<double>& evals) {
  constexpr int NS = 8;
  k = std::min(k, m);
  evals.assign((size_t)k, 0.0);
  if (m == 0) return;
  std::vector<double> e2((size_t)std::max(m - 1, 1), 0.0);
  double gl = d[0], gu = d[0], emax = 0.0;
  for (int i = 0; i < m; ++i) {
    const double r = (i ? std::fabs(e[(size_t)i - 1]) : 0.0) + (i + 1 < m ? std::fabs(e[(size_t)i]) : 0.0);
    gl = std::min(gl, d[(size_t)i] - r);
    gu = std::max(gu, d[(size_t)i] + r);
    if (i + 1 < m) {
      e2[(size_t)i] = e[(size_t)i] * e[(size_t)i];
      emax = std::max(emax, e2[(size_t)i]);
    }
  }
  const double tnorm = std::max(std::fabs(gl), std::fabs(gu));
  const double eps = 2.220446049250313e-16;
  const double pivmin = std::max(2.2250738585072014e-308 * std::max(emax, 1.0), 1e-300);
  gl -= 2.0 * eps * tnorm * m + 2.0 * pivmin;  // Gershgorin bounds, widened by the rounding of the recurrence
  gu += 2.0 * eps * tnorm * m + 2.0 * pivmin;
  // The recurrence resolves eigenvalues to about eps |T| in absolute terms: stop there.  Upper ends found while
  // bracketing eigenvalue j are kept for the later ones (ub), so that the common part of the search
  // -- from |T| down to the scale of the small cluster -- is done once.  (Single-threaded on purpose: every rank
  // of a sharded run must get the same bits, whatever cores it may use.)
  const double atol = 4.0 * eps * tnorm + 2.0 * pivmin;
  auto chunk = [&](int j0, int j1) {
    std::vector<double> ub((size_t)(j1 - j0), gu);
    double lo_prev = gl;
    for (int j = j0; j < j1; ++j) {
      double lo = lo_prev, hi = ub[(size_t)(j - j0)];  // count(lo) <= j < count(hi)
      for (int pass = 0; pass < 64 && hi - lo > atol; ++pass) {
        double x[NS];
        int cnt[NS];
        const double h = (hi - lo) / (NS + 1);
        for (int s = 0; s < NS; ++s) x[s] = lo + h * (s + 1);
        sturm_counts<NS>(d.data(), e2.data(), m, x, cnt, pivmin);
        double nlo = lo, nhi = hi;
        for (int s = NS - 1; s >= 0; --s)  // count(x) > j' makes x an upper end for every eigenvalue j' < count(x)
          for (int jj = std::min(cnt[s], j1) - 1; jj > j && ub[(size_t)(jj - j0)] > x[s]; --jj) ub[(size_t)(jj - j0)] = x[s];
        for (int s = 0; s < NS; ++s) {
          if (cnt[s] <= j) nlo = x[s];  // counts are monotone in x: the last such shift is the new lower end
          else {
            nhi = x[s];
            break;
          }
        }
        if (nlo == lo && nhi == hi) break;  // no progress: the bracket is at rounding level
        lo = nlo;
        hi = nhi;
      }
      evals[(size_t)j] = 0.5 * (lo + hi);
      lo_prev = lo;
    }
  };
  chunk(0, k);
}

// The k smallest eigenpairs of the tridiagonal (d, e): evals (k) ascending, vecs (m x k, column-major).
// Small matrices: full QL with vectors; larger ones: eigenvalues by multisection, vectors by inverse iteration
// (ghost copies of a converged Ritz value then yield nearly parallel vectors, which select_orthonormal drops).
inline bool tridiag_smallest(const std::vector<double>& d, const std::vector<double>& e, int m, int k,
                             std::vector<double>& evals, std::vector<double>& vecs, int full_below = 64) {
  k = std::min(k, m);
  if (m <= full_below) {
    std::vector<double> w(d), V;
    if (!tridiag_eig(w, e, m, V)) return false;
    evals.assign(w.begin(), w.begin() + k);
    vecs.assign(V.begin(), V.begin() + (size_t)m * k);
    return true;
  }
  tridiag_smallest_values(d, e, m, k, evals);
  vecs.assign((size_t)m * k, 0.0);
  for (int j = 0; j < k; ++j) tridiag_inverse_iteration(d, e, m, evals[(size_t)j], &vecs[(size_t)m * j]);
  return true;
}

// Given the Gram matrix G (n x n, row-major) of n candidate vectors y_1..y_n (in preference order), pick up to
// k of them whose orthogonalised remainder is at least `tol` of their norm (Gram-Schmidt in coefficient
// space, i.e. Cholesky with a fixed pivot order and rejection), and return C (n x kept, row-major) such that
// the columns of Y C are orthonormal.  Returns the number kept.
inline int select_orthonormal(const std::vector<double>& G, int n, int k, double tol, std::vector<double>& C) {
  std::vector<std::vector<double>> cols;  // coefficient vectors (length n) of the kept orthonormal vectors
  for (int j = 0; j < n && (int)cols.size() < k; ++j) {
    std::vector<double> c((size_t)n, 0.0);
    c[(size_t)j] = 1.0;
    const double njj = G[(size_t)j * n + j];
    if (!(njj > 0.0)) continue;
    for (int pass = 0; pass < 2; ++pass)  // twice is enough
      for (const auto& u : cols) {
        // <u, c>_G
        double dot = 0.0;
        for (int a = 0; a < n; ++a) {
          if (u[(size_t)a] == 0.0) continue;
          double t = 0.0;
          for (int b = 0; b < n; ++b) t += G[(size_t)a * n + b] * c[(size_t)b];
          dot += u[(size_t)a] * t;
        }
        for (int a = 0; a < n; ++a) c[(size_t)a] -= dot * u[(size_t)a];
      }
    double nrm2 = 0.0;
    for (int a = 0; a < n; ++a) {
      if (c[(size_t)a] == 0.0) continue;
      double t = 0.0;
      for (int b = 0; b < n; ++b) t += G[(size_t)a * n + b] * c[(size_t)b];
      nrm2 += c[(size_t)a] * t;
    }
    if (!(nrm2 > tol * tol * njj)) continue;  // a ghost copy (or numerically dependent): drop it
    const double inv = 1.0 / std::sqrt(nrm2);
    for (auto& v : c) v *= inv;
    cols.push_back(c);
  }
  const int kept = (int)cols.size();
  C.assign((size_t)n * std::max(kept, 1), 0.0);
  for (int j = 0; j < kept; ++j)
    for (int a = 0; a < n; ++a) C[(size_t)a * kept + j] = cols[(size_t)j][(size_t)a];
  return kept;
}

}  // namespace ba
