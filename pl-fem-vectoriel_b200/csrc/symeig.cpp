// Dense symmetric eigensolver for the projected Lanczos matrix (<= ~200 x 200), host side.
// Householder tridiagonalisation followed by implicit QL — the classic EISPACK tred2/tql2 pair.
// ARPACK does the same job with LAPACK dsteqr inside dseigt (scipy eigsh, solver_fem.py:197).
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.h"

namespace plfem {

namespace {

// sum a[i] * b[i] with four independent partial sums (a single running sum is a chain of dependent additions the compiler
// may not reorder: 4 cycles per element); y += alpha * a fused in where asked
inline double dot4(const double* __restrict a, const double* __restrict b, int n) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int i = 0;
  for (; i + 3 < n; i += 4) { s0 += a[i] * b[i]; s1 += a[i + 1] * b[i + 1]; s2 += a[i + 2] * b[i + 2]; s3 += a[i + 3] * b[i + 3]; }
  for (; i < n; ++i) s0 += a[i] * b[i];
  return (s0 + s1) + (s2 + s3);
}
// returns sum col[i] * x[i] and adds col[i] * vj to y[i] in the same pass over col
inline double dot4_axpy(const double* __restrict col, const double* __restrict x, double* __restrict y, double vj, int n) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int i = 0;
  for (; i + 3 < n; i += 4) {
    const double c0 = col[i], c1 = col[i + 1], c2 = col[i + 2], c3 = col[i + 3];
    y[i] += c0 * vj; y[i + 1] += c1 * vj; y[i + 2] += c2 * vj; y[i + 3] += c3 * vj;
    s0 += c0 * x[i]; s1 += c1 * x[i + 1]; s2 += c2 * x[i + 2]; s3 += c3 * x[i + 3];
  }
  for (; i < n; ++i) { y[i] += col[i] * vj; s0 += col[i] * x[i]; }
  return (s0 + s1) + (s2 + s3);
}

inline double hyp(double a, double b) {        // sqrt(a^2 + b^2); the library call only where the plain form could over/underflow
  const double r = std::sqrt(a * a + b * b);
  return (r > 1e-140 && r < 1e140) ? r : std::hypot(a, b);
}

// Implicit QL on the tridiagonal (d, e): e[i] couples i and i+1, e[n-1] = 0.  Every rotation is applied to the columns of Z
// (nrows x n, leading dimension ldz); on exit d holds the eigenvalues in ascending order and Z Z_T.
void ql_implicit(int n, double* d, double* e, double* Z, int nrows, int ldz) {
  auto V = [&](int i, int j) -> double& { return Z[(size_t)j * ldz + i]; };
  double f = 0.0, tst1 = 0.0;
  const double eps = std::pow(2.0, -52.0);
  for (int l = 0; l < n; ++l) {
    tst1 = std::max(tst1, std::fabs(d[l]) + std::fabs(e[l]));
    int m = l;
    while (m < n) {
      if (std::fabs(e[m]) <= eps * tst1) break;
      ++m;
    }
    if (m > l) {
      int iter = 0;
      do {
        ++iter;
        double g = d[l];
        double p = (d[l + 1] - g) / (2.0 * e[l]);
        double r = hyp(p, 1.0);
        if (p < 0) r = -r;
        d[l] = e[l] / (p + r);
        d[l + 1] = e[l] * (p + r);
        const double dl1 = d[l + 1];
        double h = g - d[l];
        for (int i = l + 2; i < n; ++i) d[i] -= h;
        f += h;
        p = d[m];
        double c = 1.0, c2 = c, c3 = c;
        const double el1 = e[l + 1];
        double s = 0.0, s2 = 0.0;
        for (int i = m - 1; i >= l; --i) {
          c3 = c2; c2 = c; s2 = s;
          g = c * e[i];
          h = c * p;
          r = hyp(p, e[i]);
          e[i + 1] = s * r;
          s = e[i] / r;
          c = p / r;
          p = c * d[i] - s * g;
          d[i + 1] = h + s * (c * g + s * d[i]);
          double* __restrict z0 = &V(0, i);
          double* __restrict z1 = &V(0, i + 1);
          for (int k = 0; k < nrows; ++k) {
            const double hk = z1[k];
            z1[k] = s * z0[k] + c * hk;
            z0[k] = c * z0[k] - s * hk;
          }
        }
        p = -s * s2 * c3 * el1 * e[l] / dl1;
        e[l] = s * p;
        d[l] = c * p;
      } while (std::fabs(e[l]) > eps * tst1 && iter < 200);
    }
    d[l] = d[l] + f;
    e[l] = 0.0;
  }
  // sort ascending
  for (int i = 0; i < n - 1; ++i) {
    int k = i; double p = d[i];
    for (int j = i + 1; j < n; ++j) if (d[j] < p) { k = j; p = d[j]; }
    if (k != i) {
      d[k] = d[i]; d[i] = p;
      for (int j = 0; j < nrows; ++j) std::swap(V(j, i), V(j, k));
    }
  }
}

}  // namespace

// a: n*n column-major symmetric matrix on entry; on exit its columns are the eigenvectors.
// w: eigenvalues, ascending.
void symmetric_eigen(int n, std::vector<double>& a, std::vector<double>& w) {
  auto V = [&](int i, int j) -> double& { return a[(size_t)j * n + i]; };
  std::vector<double> d(n), e(n);
  for (int j = 0; j < n; ++j) d[j] = V(n - 1, j);

  // Householder reduction to tridiagonal form
  for (int i = n - 1; i > 0; --i) {
    double scale = 0.0, h = 0.0;
    for (int k = 0; k < i; ++k) scale += std::fabs(d[k]);
    if (scale == 0.0) {
      e[i] = d[i - 1];
      for (int j = 0; j < i; ++j) { d[j] = V(i - 1, j); V(i, j) = 0.0; V(j, i) = 0.0; }
    } else {
      for (int k = 0; k < i; ++k) { d[k] /= scale; h += d[k] * d[k]; }
      double f = d[i - 1];
      double g = std::sqrt(h);
      if (f > 0) g = -g;
      e[i] = scale * g;
      h -= f * g;
      d[i - 1] = f - g;
      for (int j = 0; j < i; ++j) e[j] = 0.0;
      for (int j = 0; j < i; ++j) {
        f = d[j];
        V(j, i) = f;
        g = e[j] + V(j, j) * f;
        g += dot4_axpy(&V(j + 1, j), d.data() + j + 1, e.data() + j + 1, f, i - 1 - j);
        e[j] = g;
      }
      f = 0.0;
      for (int j = 0; j < i; ++j) { e[j] /= h; f += e[j] * d[j]; }
      const double hh = f / (h + h);
      for (int j = 0; j < i; ++j) e[j] -= hh * d[j];
      for (int j = 0; j < i; ++j) {
        f = d[j]; g = e[j];
        for (int k = j; k <= i - 1; ++k) V(k, j) -= (f * e[k] + g * d[k]);
        d[j] = V(i - 1, j);
        V(i, j) = 0.0;
      }
    }
    d[i] = h;
  }
  // accumulate transformations
  for (int i = 0; i < n - 1; ++i) {
    V(n - 1, i) = V(i, i);
    V(i, i) = 1.0;
    const double h = d[i + 1];
    if (h != 0.0) {
      for (int k = 0; k <= i; ++k) d[k] = V(k, i + 1) / h;
      for (int j = 0; j <= i; ++j) {
        const double g = dot4(&V(0, i + 1), &V(0, j), i + 1);
        for (int k = 0; k <= i; ++k) V(k, j) -= g * d[k];
      }
    }
    for (int k = 0; k <= i; ++k) V(k, i + 1) = 0.0;
  }
  for (int j = 0; j < n; ++j) { d[j] = V(n - 1, j); V(n - 1, j) = 0.0; }
  V(n - 1, n - 1) = 1.0;
  e[0] = 0.0;

  for (int i = 1; i < n; ++i) e[i - 1] = e[i];
  e[n - 1] = 0.0;
  ql_implicit(n, d.data(), e.data(), a.data(), n, n);
  w = d;
}

// Eigenvalues (ascending) and only the LAST p ROWS of the eigenvector matrix: tail[j * p + r] = Z(n - p + r, j).
// That is all a Lanczos convergence check reads — the residual bound of a Ritz pair is the norm of the last block of its
// eigenvector times the coupling block — at a fraction of the cost: the tridiagonalisation without accumulating Q
// (4/3 n^3 flops), the reflectors and the QL rotations applied to p rows (O(p n^2)) instead of n.  `a` is destroyed.
void symmetric_eigen_tail(int n, std::vector<double>& a, std::vector<double>& w, int p, std::vector<double>& tail) {
  auto A = [&](int i, int j) -> double& { return a[(size_t)j * n + i]; };     // lower triangle, i >= j
  std::vector<double> d(n), e(n, 0.0), tau(n, 0.0), pv(n), wv(n);
  tail.assign((size_t)p * n, 0.0);
  for (int r = 0; r < p; ++r) tail[(size_t)(n - p + r) * p + r] = 1.0;
  for (int k = 0; k + 1 < n; ++k) {
    const int m = n - k - 1;
    double* x = &A(k + 1, k);                    // becomes the reflector v (v[0] = 1 kept explicitly)
    const double alpha = x[0];
    const double xn2 = dot4(x + 1, x + 1, m - 1);
    d[k] = A(k, k);
    if (xn2 == 0.0) { e[k] = alpha; tau[k] = 0.0; continue; }
    double beta = std::hypot(alpha, std::sqrt(xn2));
    if (alpha > 0) beta = -beta;
    const double tk = (beta - alpha) / beta, sc = 1.0 / (alpha - beta);
    for (int i = 1; i < m; ++i) x[i] *= sc;
    x[0] = 1.0;
    e[k] = beta; tau[k] = tk;
    // pv = tau * A22 v  (A22 = trailing m x m block, lower triangle stored)
    for (int i = 0; i < m; ++i) pv[i] = 0.0;
    for (int j = 0; j < m; ++j) {
      const double* col = &A(k + 1 + j, k + 1 + j);        // A22(j.., j)
      const double vj = x[j];
      pv[j] += col[0] * vj + dot4_axpy(col + 1, x + j + 1, pv.data() + j + 1, vj, m - j - 1);
    }
    for (int i = 0; i < m; ++i) pv[i] *= tk;
    const double dot = dot4(pv.data(), x, m);
    const double al = -0.5 * tk * dot;
    for (int i = 0; i < m; ++i) wv[i] = pv[i] + al * x[i];
    for (int j = 0; j < m; ++j) {
      double* col = &A(k + 1 + j, k + 1 + j);
      const double vj = x[j], wj = wv[j];
      for (int i = 0; i < m - j; ++i) col[i] -= x[j + i] * wj + wv[j + i] * vj;
    }
    // tail rows of Q = H_0 H_1 ...: R <- R H_k on the columns k+1 .. n-1
    for (int r = 0; r < p; ++r) {
      double t = 0.0;
      for (int i = 0; i < m; ++i) t += tail[(size_t)(k + 1 + i) * p + r] * x[i];
      t *= tk;
      for (int i = 0; i < m; ++i) tail[(size_t)(k + 1 + i) * p + r] -= t * x[i];
    }
  }
  d[n - 1] = A(n - 1, n - 1);
  e[n - 1] = 0.0;
  ql_implicit(n, d.data(), e.data(), tail.data(), p, p);
  w = d;
}

}  // namespace plfem
