// Tiny dense linear-algebra helpers in FP64 (__host__ __device__): cyclic Jacobi eigen-decomposition of
// small symmetric matrices, 3x3 SVD built on it.  Used by the pose-recovery kernels.
#pragma once
#include "common.cuh"

namespace dfvo {
namespace sm {

// Symmetric eigen-decomposition A = V diag(w) V^T, n <= 4.  A is destroyed; V columns are eigenvectors.
template <int n>
DFVO_HD void jacobi_eig(double A[n][n], double V[n][n], double w[n]) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < n; ++i) {
      diag += A[i][i] * A[i][i];
      for (int j = i + 1; j < n; ++j) off += A[i][j] * A[i][j];
    }
    if (off <= 1e-32 * diag || off == 0.0) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[p][q];
        if (apq == 0.0) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < n; ++i) w[i] = A[i][i];
}

DFVO_HD double det3(const double M[3][3]) {
  return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
         M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
}

// SVD of a (near) rank-2 3x3 matrix E = U diag(s) Vt with s sorted descending; U, Vt orthogonal.
DFVO_HD void svd3_rank2(const double E[3][3], double U[3][3], double s[3], double Vt[3][3]) {
  double A[3][3], V[3][3], w[3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) { double a = 0; for (int k = 0; k < 3; ++k) a += E[k][i] * E[k][j]; A[i][j] = a; }
  jacobi_eig<3>(A, V, w);
  int idx[3] = {0, 1, 2};
  for (int a = 0; a < 2; ++a)
    for (int b = a + 1; b < 3; ++b)
      if (w[idx[b]] > w[idx[a]]) { int t = idx[a]; idx[a] = idx[b]; idx[b] = t; }
  double v[3][3];   // columns sorted
  for (int c = 0; c < 3; ++c) { s[c] = sqrt(w[idx[c]] > 0 ? w[idx[c]] : 0.0); for (int r = 0; r < 3; ++r) v[r][c] = V[r][idx[c]]; }
  // u_c = E v_c / s_c for the two dominant directions
  double u[3][3];
  for (int c = 0; c < 2; ++c) {
    double nn = 0;
    for (int r = 0; r < 3; ++r) { double a = 0; for (int k = 0; k < 3; ++k) a += E[r][k] * v[k][c]; u[r][c] = a; nn += a * a; }
    nn = sqrt(nn);
    for (int r = 0; r < 3; ++r) u[r][c] = nn > 0 ? u[r][c] / nn : (r == c ? 1.0 : 0.0);
  }
  // re-orthogonalise u1 against u0, third = cross
  double d = u[0][0] * u[0][1] + u[1][0] * u[1][1] + u[2][0] * u[2][1];
  double nn = 0;
  for (int r = 0; r < 3; ++r) { u[r][1] -= d * u[r][0]; nn += u[r][1] * u[r][1]; }
  nn = sqrt(nn);
  for (int r = 0; r < 3; ++r) u[r][1] /= nn;
  u[0][2] = u[1][0] * u[2][1] - u[2][0] * u[1][1];
  u[1][2] = u[2][0] * u[0][1] - u[0][0] * u[2][1];
  u[2][2] = u[0][0] * u[1][1] - u[1][0] * u[0][1];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) { U[r][c] = u[r][c]; Vt[c][r] = v[r][c]; }
}

}  // namespace sm
}  // namespace dfvo
