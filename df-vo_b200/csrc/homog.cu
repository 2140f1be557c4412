// Homography model of the GRIC validity check on the device:
//   H, _ = cv2.findHomography(kp_cur, kp_ref, method=RANSAC, confidence=0.99, ransacReprojThreshold=1)   (E_tracker.py:199-205)
//   H_gric = calc_GRIC(compute_homography_residual(H, kp_cur, kp_ref), 0.8, n, 'HMat')                    (E_tracker.py:206-215, gric.py:40-132)
// Structure of OpenCV's routine (calib3d fundam.cpp / ptsetreg.cpp / levmarq.cpp; restated from the published algorithms,
// pinned against cv2 4.13 in tests/):
//   * both point sets are converted to float32;
//   * RANSACPointSetRegistrator(modelPoints = 4, threshold, confidence, maxIters = 2000): cv::RNG(-1); a subset is four
//     distinct uniform indices, REDRAWN while HomographyEstimatorCallback::checkSubset rejects it (last point collinear
//     with a previous pair in either image, or the four correspondences not orientation-consistent) -- so unlike the
//     essential-matrix / PnP streams the subset sequence depends on the data and is generated here, sequentially;
//   * minimal solver: normalised DLT, the 9x9 L^T L accumulated in double, H = eigenvector of the smallest eigenvalue,
//     de-normalised and scaled to H[2][2] = 1;
//   * error: forward reprojection distance^2 evaluated in float32 with a float32 copy of H; inlier iff err <= thr^2;
//     accept iff good > max(best, 3); niters = RANSACUpdateNumIters(confidence, (N - good) / N, 4, niters);
//   * final model: the same DLT over all inliers, then LMSolver (<= 10 iterations) on the 8 free parameters minimising
//     the reprojection residuals -- restated step for step (Nielsen damping as in cv::LMSolverImpl::run).
// Then the GRIC score of that H (Torr's GRIC with the reference's residual) is reduced on the device, so the host only
// reads one double per frame.
#include "ransac.h"
#include "smallmat.cuh"

namespace dfvo {

struct HState { int niters, best_good, best_iter, it, done, n_subsets; };

namespace hmg {

// Eigenvector of the smallest eigenvalue of the symmetric positive semi-definite 9x9 normal matrix, by inverse iteration on
// A + mu I (mu = 1e-13 trace: keeps the LDL^T factorisation regular when the matrix is exactly singular, as it is for a
// minimal 4-point sample).  The smallest eigenvalue is separated from the next by many orders of magnitude in both uses
// (minimal sample: exact null space; inlier refit: noise level vs signal), so a handful of iterations reach round-off; a
// single GPU thread does this in ~2k dependent flops instead of the ~40k of a cyclic Jacobi sweep sequence.
DFVO_HD_NOINLINE void smallest_eigvec9(double A[9][9], double out[9]) {
  double tr = 0;
  for (int i = 0; i < 9; ++i) tr += A[i][i];
  const double mu = 1e-13 * tr;
  // LDL^T (no pivoting; SPD + shift): L unit lower in-place, D on the diagonal
  double L[9][9], D[9];
  for (int j = 0; j < 9; ++j) {
    double d = A[j][j] + mu;
    for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k] * D[k];
    D[j] = d;
    for (int i = j + 1; i < 9; ++i) {
      double v = A[i][j];
      for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k] * D[k];
      L[i][j] = v / d;
    }
  }
  double x[9], y[9];
  for (int i = 0; i < 9; ++i) x[i] = 1.0 / 3.0 + 0.01 * i;           // generic start (not orthogonal to anything special)
  for (int it = 0; it < 8; ++it) {
    for (int i = 0; i < 9; ++i) { double v = x[i]; for (int k = 0; k < i; ++k) v -= L[i][k] * y[k]; y[i] = v; }
    for (int i = 0; i < 9; ++i) y[i] /= D[i];
    for (int i = 8; i >= 0; --i) { double v = y[i]; for (int k = i + 1; k < 9; ++k) v -= L[k][i] * y[k]; y[i] = v; }
    double nn = 0;
    for (int i = 0; i < 9; ++i) nn += y[i] * y[i];
    nn = 1.0 / sqrt(nn);
    double diff = 0, diffm = 0;
    for (int i = 0; i < 9; ++i) { const double v = y[i] * nn; diff += (v - x[i]) * (v - x[i]); diffm += (v + x[i]) * (v + x[i]); x[i] = v; }
    if ((diff < 1e-30 || diffm < 1e-30) && it > 0) break;
  }
  for (int i = 0; i < 9; ++i) out[i] = x[i];
}

// H from the accumulated normal matrix and the two normalisations (HomographyEstimatorCallback::runKernel tail)
DFVO_HD bool finish_dlt(double LtL[9][9], double cMx, double cMy, double sMx, double sMy, double cmx, double cmy, double smx, double smy,
                        double H[9]) {
  for (int j = 0; j < 9; ++j) for (int k = 0; k < j; ++k) LtL[j][k] = LtL[k][j];          // completeSymm
  double h0[9];
  smallest_eigvec9(LtL, h0);
  const double invHnorm[9] = {1.0 / smx, 0, cmx, 0, 1.0 / smy, cmy, 0, 0, 1};
  const double Hnorm2[9] = {sMx, 0, -cMx * sMx, 0, sMy, -cMy * sMy, 0, 0, 1};
  double T[9], Hh[9];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double a = 0; for (int k = 0; k < 3; ++k) a += invHnorm[3 * i + k] * h0[3 * k + j]; T[3 * i + j] = a; }
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double a = 0; for (int k = 0; k < 3; ++k) a += T[3 * i + k] * Hnorm2[3 * k + j]; Hh[3 * i + j] = a; }
  if (!(fabs(Hh[8]) > 0)) return false;
  const double s = 1.0 / Hh[8];
  for (int i = 0; i < 9; ++i) H[i] = Hh[i] * s;
  return true;
}

// haveCollinearPoints(m, count): only the LAST point against pairs of earlier ones
DFVO_HD bool last_collinear(const float* p, int count) {
  const int i = count - 1;
  for (int j = 0; j < i; ++j) {
    const double dx1 = (double)p[2 * j] - (double)p[2 * i], dy1 = (double)p[2 * j + 1] - (double)p[2 * i + 1];
    for (int k = 0; k < j; ++k) {
      const double dx2 = (double)p[2 * k] - (double)p[2 * i], dy2 = (double)p[2 * k + 1] - (double)p[2 * i + 1];
      if (fabs(dx2 * dy1 - dy2 * dx1) <= 1.1920928955078125e-07 * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2))) return true;
    }
  }
  return false;
}

DFVO_HD double det3pts(const float* p, int a, int b, int c) {
  const double x0 = p[2 * a], y0 = p[2 * a + 1], x1 = p[2 * b], y1 = p[2 * b + 1], x2 = p[2 * c], y2 = p[2 * c + 1];
  // determinant of [[x0,y0,1],[x1,y1,1],[x2,y2,1]] (cv::determinant of a Matx33d)
  return x0 * (y1 * 1.0 - 1.0 * y2) - y0 * (x1 * 1.0 - 1.0 * x2) + 1.0 * (x1 * y2 - y1 * x2);
}

DFVO_HD bool check_subset(const float* s, const float* d) {
  if (last_collinear(s, 4) || last_collinear(d, 4)) return false;
  const int tt[4][3] = {{0, 1, 2}, {1, 2, 3}, {0, 2, 3}, {0, 1, 3}};
  int negative = 0;
  for (int i = 0; i < 4; ++i) negative += (det3pts(s, tt[i][0], tt[i][1], tt[i][2]) * det3pts(d, tt[i][0], tt[i][1], tt[i][2]) < 0) ? 1 : 0;
  return negative == 0 || negative == 4;
}

}  // namespace hmg

// float32 copies of the points (cv::findHomography converts both sets to CV_32FC2)
__global__ void k_h_prepare(const double* __restrict__ p1, const double* __restrict__ p2, int N, float* __restrict__ src, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  src[2 * i] = (float)p1[2 * i]; src[2 * i + 1] = (float)p1[2 * i + 1];
  dst[2 * i] = (float)p2[2 * i]; dst[2 * i + 1] = (float)p2[2 * i + 1];
}

// thread 0 of the block: subsets [i0, i1) of the RANSAC loop (RANSACPointSetRegistrator::getSubset with the homography
// checkSubset); the generator state is carried between rounds.  Returns the number of subsets available so far.
DFVO_D int h_make_subsets(const float* __restrict__ src, const float* __restrict__ dst, int N, int i0, int i1, uint64_t* state_io,
                          int32_t* __restrict__ subsets) {
  uint64_t state = *state_io;
  int made = i0;
  for (int it = i0; it < i1; ++it) {
    int idx[4];
    float s[8], d[8];
    bool found = false;
    for (int attempt = 0; attempt < 10000 && !found; ++attempt) {
      for (int i = 0; i < 4;) {
        state = (uint64_t)(uint32_t)state * 4164903690u + (uint32_t)(state >> 32);
        const int v = (int)((uint32_t)state % (uint32_t)N);
        bool dup = false;
        for (int j = 0; j < i; ++j) dup = dup || idx[j] == v;
        if (dup) continue;
        idx[i] = v;
        s[2 * i] = src[2 * v]; s[2 * i + 1] = src[2 * v + 1]; d[2 * i] = dst[2 * v]; d[2 * i + 1] = dst[2 * v + 1];
        ++i;
      }
      found = hmg::check_subset(s, d);
    }
    if (!found) break;                                          // OpenCV stops the loop here
    for (int i = 0; i < 4; ++i) subsets[it * 4 + i] = idx[i];
    made = it + 1;
  }
  *state_io = state;
  return made;
}

// one thread per iteration: 4-point normalised DLT -> hyp [max_iters][9], ok [max_iters]
DFVO_D void h_hypothesis(const float* __restrict__ src, const float* __restrict__ dst, const int32_t* __restrict__ subsets, int it,
                         double* __restrict__ hyp, int32_t* __restrict__ ok) {
  double Mx[4], My[4], mx[4], my[4];
  double cMx = 0, cMy = 0, cmx = 0, cmy = 0;
  for (int i = 0; i < 4; ++i) {
    const int v = subsets[it * 4 + i];
    Mx[i] = src[2 * v]; My[i] = src[2 * v + 1]; mx[i] = dst[2 * v]; my[i] = dst[2 * v + 1];
    cMx += Mx[i]; cMy += My[i]; cmx += mx[i]; cmy += my[i];
  }
  cMx /= 4; cMy /= 4; cmx /= 4; cmy /= 4;
  double sMx = 0, sMy = 0, smx = 0, smy = 0;
  for (int i = 0; i < 4; ++i) { sMx += fabs(Mx[i] - cMx); sMy += fabs(My[i] - cMy); smx += fabs(mx[i] - cmx); smy += fabs(my[i] - cmy); }
  const double eps = 2.220446049250313e-16;
  if (fabs(smx) < eps || fabs(smy) < eps || fabs(sMx) < eps || fabs(sMy) < eps) { ok[it] = 0; return; }
  smx = 4 / smx; smy = 4 / smy; sMx = 4 / sMx; sMy = 4 / sMy;
  double LtL[9][9];
  for (int j = 0; j < 9; ++j) for (int k = 0; k < 9; ++k) LtL[j][k] = 0;
  for (int i = 0; i < 4; ++i) {
    const double x = (mx[i] - cmx) * smx, y = (my[i] - cmy) * smy, X = (Mx[i] - cMx) * sMx, Y = (My[i] - cMy) * sMy;
    const double Lx[9] = {X, Y, 1, 0, 0, 0, -x * X, -x * Y, -x}, Ly[9] = {0, 0, 0, X, Y, 1, -y * X, -y * Y, -y};
    for (int j = 0; j < 9; ++j) for (int k = j; k < 9; ++k) LtL[j][k] += Lx[j] * Lx[k] + Ly[j] * Ly[k];
  }
  double H[9];
  const bool good = hmg::finish_dlt(LtL, cMx, cMy, sMx, sMy, cmx, cmy, smx, smy, H);
  ok[it] = good ? 1 : 0;
  if (good) for (int i = 0; i < 9; ++i) hyp[(size_t)it * 9 + i] = H[i];
}

// HomographyEstimatorCallback::computeError, float32 throughout
DFVO_D float h_err(const float* Hf, float Mx, float My, float mx, float my) {
  const float ww = 1.f / (Hf[6] * Mx + Hf[7] * My + 1.f);
  const float dx = (Hf[0] * Mx + Hf[1] * My + Hf[2]) * ww - mx;
  const float dy = (Hf[3] * Mx + Hf[4] * My + Hf[5]) * ww - my;
  return dx * dx + dy * dy;
}

DFVO_HD int h_update_num_iters(double p, double ep, int model_points, int max_iters) {
  p = p < 0 ? 0 : (p > 1 ? 1 : p);
  ep = ep < 0 ? 0 : (ep > 1 ? 1 : ep);
  double num = 1.0 - p;
  if (num < 2.2250738585072014e-308) num = 2.2250738585072014e-308;
  double denom = 1.0 - pow(1.0 - ep, (double)model_points);
  if (denom < 2.2250738585072014e-308) return 0;
  num = log(num);
  denom = log(denom);
  if (denom >= 0 || -num >= max_iters * (-denom)) return max_iters;
  return (int)rint(num / denom);
}

// ONE block runs the whole estimation: RANSAC rounds (thread 0 draws the subsets, one thread per minimal solve, one warp
// per hypothesis score, thread 0 replays the acceptance rule; later rounds only if the adaptive iteration count asks
// for them), then the inlier mask of the winner, the DLT over the inliers, the LM refinement and GRIC-H.  A single
// launch matters: the tracker's kernels run beside the next frame's persistent convolution kernels, and every
// dependent launch waits for an SM to take it.
// out: H_out [9], mask [N] (inliers of the refined H), info [4] = {found, inliers, iterations, winning iteration}, gric [1]
__global__ void __launch_bounds__(256)
k_h_ransac(const float* __restrict__ src, const float* __restrict__ dst, const double* __restrict__ p1, const double* __restrict__ p2, int N,
           int max_iters, double prob, float thr2, double* __restrict__ hyp, int32_t* __restrict__ subsets, int32_t* __restrict__ ok,
           int32_t* __restrict__ counts, double* __restrict__ H_out, uint8_t* __restrict__ mask, int32_t* __restrict__ info,
           double* __restrict__ gric) {
  __shared__ double part[8][48];
  __shared__ double acc[48];
  __shared__ double Hs[9], xs[8], xd[8], dvec[8], Amat[8][8], vvec[8];
  __shared__ double S_cur, S_new, lambda, lc;
  __shared__ int stop, ninl;
  __shared__ HState sst;
  __shared__ uint64_t rng_state;
  const int t = threadIdx.x;
  if (t == 0) {
    sst.niters = max_iters; sst.best_good = -1; sst.best_iter = -1; sst.it = 0; sst.done = 0; sst.n_subsets = 0;
    rng_state = 0xFFFFFFFFFFFFFFFFull;                          // cv::RNG((uint64)-1)
  }
  __syncthreads();
  {
    const int bounds[5] = {0, 64, 320, 1088, max_iters};
    for (int rd = 0; rd < 4; ++rd) {
      const int i0 = bounds[rd] < max_iters ? bounds[rd] : max_iters, i1 = bounds[rd + 1] < max_iters ? bounds[rd + 1] : max_iters;
      if (i1 <= i0) continue;
      if (t == 0) sst.n_subsets = h_make_subsets(src, dst, N, i0, i1, &rng_state, subsets);
      __syncthreads();
      for (int it = i0 + t; it < i1; it += 256) {
        if (it < sst.n_subsets) h_hypothesis(src, dst, subsets, it, hyp, ok);
        else ok[it] = 0;
      }
      __syncthreads();
      for (int it = i0 + (t >> 5); it < i1; it += 8) {          // one warp per hypothesis
        int c = 0;
        if (ok[it]) {
          float Hf[8];
          for (int q = 0; q < 8; ++q) Hf[q] = (float)hyp[(size_t)it * 9 + q];
          for (int j = (t & 31); j < N; j += 32) c += (h_err(Hf, src[2 * j], src[2 * j + 1], dst[2 * j], dst[2 * j + 1]) <= thr2) ? 1 : 0;
        }
        for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
        if ((t & 31) == 0) counts[it] = c;
      }
      __syncthreads();
      if (t == 0) {                                             // acceptance rule of RANSACPointSetRegistrator::run
        HState s2 = sst;
        int it = s2.it;
        while (it < s2.niters && it < i1) {
          if (it >= s2.n_subsets) { s2.done = 1; break; }       // getSubset failed: the loop ends
          if (ok[it]) {
            const int good = counts[it];
            const int lim = s2.best_good > 3 ? s2.best_good : 3;
            if (good > lim) {
              s2.best_good = good; s2.best_iter = it;
              s2.niters = h_update_num_iters(prob, (double)(N - good) / (double)N, 4, s2.niters);
            }
          }
          ++it;
        }
        s2.it = it;
        if (it >= s2.niters) s2.done = 1;
        sst = s2;
      }
      __syncthreads();
      if (sst.done) break;
    }
  }
  const HState s = sst;
  auto block_sum_vec = [&](const double* a, int n) {            // sums of n <= 48 per-thread values -> acc[0..n)
    for (int k = 0; k < n; ++k) {
      double v = a[k];
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if ((t & 31) == 0) part[t >> 5][k] = v;
    }
    __syncthreads();
    if (t < n) { double v = 0; for (int w8 = 0; w8 < 8; ++w8) v += part[w8][t]; acc[t] = v; }
    __syncthreads();
  };
  if (s.best_iter < 0) {
    for (int j = t; j < N; j += 256) mask[j] = 0;
    if (t == 0) {
      for (int k = 0; k < 9; ++k) H_out[k] = 0.0;
      info[0] = 0; info[1] = 0; info[2] = s.it; info[3] = -1;
      gric[0] = 0.0;
    }
    return;
  }
  if (t < 9) Hs[t] = hyp[(size_t)s.best_iter * 9 + t];
  __syncthreads();
  {
    float Hf[8];
    for (int q = 0; q < 8; ++q) Hf[q] = (float)Hs[q];
    double cnt[1] = {0};
    for (int j = t; j < N; j += 256) {
      const bool in = h_err(Hf, src[2 * j], src[2 * j + 1], dst[2 * j], dst[2 * j + 1]) <= thr2;
      mask[j] = in ? 1 : 0;
      cnt[0] += in ? 1.0 : 0.0;
    }
    block_sum_vec(cnt, 1);
    if (t == 0) ninl = (int)acc[0];
    __syncthreads();
  }
  const int n = ninl;
  // ---- DLT over the inliers: centroids, mean absolute deviations, L^T L (45 unique sums)
  {
    double a4[4] = {0, 0, 0, 0};
    for (int j = t; j < N; j += 256) if (mask[j]) { a4[0] += src[2 * j]; a4[1] += src[2 * j + 1]; a4[2] += dst[2 * j]; a4[3] += dst[2 * j + 1]; }
    block_sum_vec(a4, 4);
  }
  const double cMx = acc[0] / n, cMy = acc[1] / n, cmx = acc[2] / n, cmy = acc[3] / n;
  __syncthreads();
  {
    double a4[4] = {0, 0, 0, 0};
    for (int j = t; j < N; j += 256) if (mask[j]) {
      a4[0] += fabs((double)src[2 * j] - cMx); a4[1] += fabs((double)src[2 * j + 1] - cMy);
      a4[2] += fabs((double)dst[2 * j] - cmx); a4[3] += fabs((double)dst[2 * j + 1] - cmy);
    }
    block_sum_vec(a4, 4);
  }
  const double sMx = n / acc[0], sMy = n / acc[1], smx = n / acc[2], smy = n / acc[3];
  __syncthreads();
  {
    double a45[45];
    for (int k = 0; k < 45; ++k) a45[k] = 0;
    for (int j = t; j < N; j += 256) if (mask[j]) {
      const double x = ((double)dst[2 * j] - cmx) * smx, y = ((double)dst[2 * j + 1] - cmy) * smy;
      const double X = ((double)src[2 * j] - cMx) * sMx, Y = ((double)src[2 * j + 1] - cMy) * sMy;
      const double Lx[9] = {X, Y, 1, 0, 0, 0, -x * X, -x * Y, -x}, Ly[9] = {0, 0, 0, X, Y, 1, -y * X, -y * Y, -y};
      int q = 0;
      for (int a = 0; a < 9; ++a) for (int b = a; b < 9; ++b) a45[q++] += Lx[a] * Lx[b] + Ly[a] * Ly[b];
    }
    block_sum_vec(a45, 45);
  }
  if (t == 0) {
    double LtL[9][9];
    int q = 0;
    for (int a = 0; a < 9; ++a) for (int b = a; b < 9; ++b) LtL[a][b] = acc[q++];
    double H[9];
    if (hmg::finish_dlt(LtL, cMx, cMy, sMx, sMy, cmx, cmy, smx, smy, H)) for (int k = 0; k < 9; ++k) Hs[k] = H[k];
    for (int k = 0; k < 8; ++k) xs[k] = Hs[k];
    lambda = 1.0; lc = 0.75; stop = 0;
  }
  __syncthreads();
  // ---- LMSolver (<= 10 iterations) on h[0..7]:  r = (proj - m) per inlier;  A = J^T J, v = J^T r
  auto residuals = [&](const double* h, double* Sout, bool with_jac) {
    double a[45];
    const int nn = with_jac ? 45 : 1;
    for (int k = 0; k < nn; ++k) a[k] = 0;
    for (int j = t; j < N; j += 256) if (mask[j]) {
      const double Mx = src[2 * j], My = src[2 * j + 1];
      double ww = h[6] * Mx + h[7] * My + 1.0;
      ww = fabs(ww) > 2.220446049250313e-16 ? 1.0 / ww : 0.0;
      const double xi = (h[0] * Mx + h[1] * My + h[2]) * ww, yi = (h[3] * Mx + h[4] * My + h[5]) * ww;
      const double ex = xi - (double)dst[2 * j], ey = yi - (double)dst[2 * j + 1];
      a[0] += ex * ex + ey * ey;
      if (with_jac) {
        const double Jx[8] = {Mx * ww, My * ww, ww, 0, 0, 0, -Mx * ww * xi, -My * ww * xi};
        const double Jy[8] = {0, 0, 0, Mx * ww, My * ww, ww, -Mx * ww * yi, -My * ww * yi};
        int q = 1;
        for (int u = 0; u < 8; ++u) for (int w2 = u; w2 < 8; ++w2) a[q++] += Jx[u] * Jx[w2] + Jy[u] * Jy[w2];       // 36 entries
        for (int u = 0; u < 8; ++u) a[37 + u] += Jx[u] * ex + Jy[u] * ey;                                              // 8 entries
      }
    }
    block_sum_vec(a, nn);
    if (t == 0) {
      *Sout = acc[0];
      if (with_jac) {
        int q = 1;
        for (int u = 0; u < 8; ++u) for (int w2 = u; w2 < 8; ++w2) { Amat[u][w2] = acc[q]; Amat[w2][u] = acc[q]; ++q; }
        for (int u = 0; u < 8; ++u) vvec[u] = acc[37 + u];
      }
    }
    __syncthreads();
  };
  if (n > 4) {
    residuals(xs, &S_cur, true);
    for (int iter = 0; iter < 10; ++iter) {
      if (t == 0) {
        // solve (A + lambda diag(A)) d = v   (symmetric positive definite: Cholesky; OpenCV uses an eigen-solve)
        double Ap[8][8], L[8][8];
        for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) Ap[i][j] = Amat[i][j];
        for (int i = 0; i < 8; ++i) Ap[i][i] += lambda * Amat[i][i];
        bool pd = true;
        for (int i = 0; i < 8 && pd; ++i)
          for (int k = 0; k <= i; ++k) {
            double sum = Ap[i][k];
            for (int m = 0; m < k; ++m) sum -= L[i][m] * L[k][m];
            if (i == k) { if (sum <= 0) { pd = false; break; } L[i][i] = sqrt(sum); }
            else L[i][k] = sum / L[k][k];
          }
        for (int i = 0; i < 8; ++i) dvec[i] = 0;
        if (pd) {
          double yv[8];
          for (int i = 0; i < 8; ++i) { double sum = vvec[i]; for (int m = 0; m < i; ++m) sum -= L[i][m] * yv[m]; yv[i] = sum / L[i][i]; }
          for (int i = 7; i >= 0; --i) { double sum = yv[i]; for (int m = i + 1; m < 8; ++m) sum -= L[m][i] * dvec[m]; dvec[i] = sum / L[i][i]; }
        }
        for (int i = 0; i < 8; ++i) xd[i] = xs[i] - dvec[i];
      }
      __syncthreads();
      residuals(xd, &S_new, false);
      if (t == 0) {
        // gain ratio R = (S - Sd) / (d . (2 v - A d))
        double dS = 0;
        for (int i = 0; i < 8; ++i) { double ad = 0; for (int j = 0; j < 8; ++j) ad += Amat[i][j] * dvec[j]; dS += dvec[i] * (2.0 * vvec[i] - ad); }
        const double Rg = (S_cur - S_new) / (fabs(dS) > 2.220446049250313e-16 ? dS : 1.0);
        if (Rg > 0.75) { lambda *= 0.5; if (lambda < lc) lambda = 0; }
        else if (Rg < 0.25) {
          double tt = 0;
          for (int i = 0; i < 8; ++i) tt += dvec[i] * vvec[i];
          double nu = (S_new - S_cur) / (fabs(tt) > 2.220446049250313e-16 ? tt : 1.0) + 2.0;
          nu = nu < 2.0 ? 2.0 : (nu > 10.0 ? 10.0 : nu);
          if (lambda == 0) {
            // lambda = lc = 1 / max |diag(A^-1)|: invert through the Cholesky factor of A
            double L2[8][8];
            bool pd2 = true;
            for (int i = 0; i < 8 && pd2; ++i)
              for (int k = 0; k <= i; ++k) {
                double sum = Amat[i][k];
                for (int m = 0; m < k; ++m) sum -= L2[i][m] * L2[k][m];
                if (i == k) { if (sum <= 0) { pd2 = false; break; } L2[i][i] = sqrt(sum); }
                else L2[i][k] = sum / L2[k][k];
              }
            double maxval = 2.220446049250313e-16;
            if (pd2) {
              for (int c = 0; c < 8; ++c) {                     // column c of A^-1: solve A z = e_c, take z[c]
                double yv[8], z[8];
                for (int i = 0; i < 8; ++i) { double sum = (i == c) ? 1.0 : 0.0; for (int m = 0; m < i; ++m) sum -= L2[i][m] * yv[m]; yv[i] = sum / L2[i][i]; }
                for (int i = 7; i >= 0; --i) { double sum = yv[i]; for (int m = i + 1; m < 8; ++m) sum -= L2[m][i] * z[m]; z[i] = sum / L2[i][i]; }
                if (fabs(z[c]) > maxval) maxval = fabs(z[c]);
              }
            }
            lambda = lc = 1.0 / maxval;
            nu *= 0.5;
          }
          lambda *= nu;
        }
      }
      __syncthreads();
      const bool better = S_new < S_cur;
      __syncthreads();
      if (better) {
        if (t < 8) xs[t] = xd[t];
        __syncthreads();
        residuals(xs, &S_cur, true);
      }
      if (t == 0) {
        double dinf = 0;
        for (int i = 0; i < 8; ++i) dinf = fabs(dvec[i]) > dinf ? fabs(dvec[i]) : dinf;
        // OpenCV also stops on |r|_inf < FLT_EPSILON (never true with noisy points); the step test is the live one
        if (!(iter + 1 < 10 && dinf >= 1.1920928955078125e-07)) stop = 1;
      }
      __syncthreads();
      if (stop) break;
    }
    if (t < 8) Hs[t] = xs[t];
    __syncthreads();
  }
  // ---- the returned mask is re-evaluated with the refined H (cv::findHomography does the same before it hands the mask out)
  {
    float Hf[8];
    for (int q = 0; q < 8; ++q) Hf[q] = (float)Hs[q];
    double cnt[1] = {0};
    for (int j = t; j < N; j += 256) {
      const bool in = h_err(Hf, src[2 * j], src[2 * j + 1], dst[2 * j], dst[2 * j + 1]) <= thr2;
      mask[j] = in ? 1 : 0;
      cnt[0] += in ? 1.0 : 0.0;
    }
    block_sum_vec(cnt, 1);
    if (t == 0) ninl = (int)acc[0];
    __syncthreads();
  }
  // ---- GRIC-H over ALL points with pixel coordinates (gric.py:40-132): kp1 = p1 (cur), kp2 = p2 (ref)
  {
    const double* H = Hs;
    double g[1] = {0};
    const double sigmasq1 = 1.0 / (0.8 * 0.8), lam3RD = 2.0 * (4 - 2);
    for (int j = t; j < N; j += 256) {
      const double x0 = p1[2 * j], y0 = p1[2 * j + 1], x1 = p2[2 * j], y1 = p2[2 * j + 1];
      const double w = x0 * H[6] + y0 * H[7] + H[8];
      const double G0[3] = {H[0] - x1 * H[6], H[1] - x1 * H[7], -w}, G1[3] = {H[3] - y1 * H[6], H[4] - y1 * H[7], -w};
      const double magG0 = sqrt(G0[0] * G0[0] + G0[1] * G0[1] + G0[2] * G0[2]), magG1 = sqrt(G1[0] * G1[0] + G1[1] * G1[1] + G1[2] * G1[2]);
      const double alpha = acos((G0[0] * G1[0] + G0[1] * G1[1]) / (magG0 * magG1));
      const double alg0 = x0 * H[0] + y0 * H[1] + H[2] - x1 * w, alg1 = x0 * H[3] + y0 * H[4] + H[5] - y1 * w;
      const double D1 = alg0 / magG0, D2 = alg1 / magG1;
      const double res = (D1 * D1 + D2 * D2 - 2.0 * D1 * D2 * cos(alpha)) / sin(alpha);
      const double tmp = res * sigmasq1;
      g[0] += (tmp <= lam3RD) ? tmp : lam3RD;
    }
    block_sum_vec(g, 1);
    if (t == 0) {
      gric[0] = acc[0] + (double)N * 2.0 * log(4.0) + 8.0 * log(4.0 * (double)N);
      for (int k = 0; k < 9; ++k) H_out[k] = Hs[k];
      info[0] = 1; info[1] = ninl; info[2] = s.it; info[3] = s.best_iter;
    }
  }
}

size_t homography_workspace_bytes(int N, int max_iters) {
  return (size_t)N * 4 * 4 + (size_t)max_iters * (9 * 8 + 4 * 4 + 4 + 4) + sizeof(HState) + 8 + 2048;
}

int homography_ransac(const double* p1, const double* p2, int N, int max_iters, double threshold, double prob, void* workspace, size_t ws_bytes,
                      double* H_out, uint8_t* mask_out, int32_t* info, double* gric, cudaStream_t s) {
  DFVO_REQUIRE(p1 && p2 && H_out && mask_out && info && gric && N >= 4 && max_iters >= 1, DFVO_EINVAL, "homography_ransac args (N=%d)", N);
  DFVO_REQUIRE(ws_bytes >= homography_workspace_bytes(N, max_iters), DFVO_EINVAL, "homography_ransac workspace too small");
  uint8_t* w = reinterpret_cast<uint8_t*>(workspace);
  auto take = [&](size_t bytes) { uint8_t* p = w; w += (bytes + 127) & ~(size_t)127; return p; };
  float* src = (float*)take((size_t)N * 2 * 4);
  float* dst = (float*)take((size_t)N * 2 * 4);
  double* hyp = (double*)take((size_t)max_iters * 9 * 8);
  int32_t* subsets = (int32_t*)take((size_t)max_iters * 4 * 4);
  int32_t* ok = (int32_t*)take((size_t)max_iters * 4);
  int32_t* counts = (int32_t*)take((size_t)max_iters * 4);
  const float thr2 = (float)(threshold * threshold);
  DFVO_LAUNCH(k_h_prepare, dim3(cdiv(N, 128)), dim3(128), 0, s, p1, p2, N, src, dst);
  DFVO_LAUNCH(k_h_ransac, dim3(1), dim3(256), 0, s, src, dst, p1, p2, N, max_iters, prob, thr2, hyp, subsets, ok, counts, H_out, mask_out,
              info, gric);
  DFVO_CHECK_LAUNCH();
  return DFVO_OK;
}

}  // namespace dfvo
