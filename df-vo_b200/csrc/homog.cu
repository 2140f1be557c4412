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

struct HState {
  int niters, best_good, best_iter, it, done, n_subsets;
  int pos, ndraws, failed, pad;      // stream offset of the next attempt, raw draws generated so far
  uint64_t rng;                      // generator state after `ndraws` draws
};

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

// ---- subset stream: "parallel-evaluate / sequential-replay" --------------------------------------------------------------------
// OpenCV draws its 4-point subsets with rejection (checkSubset), so the stream depends on the data and is sequential -- but only
// through HOW MANY raw draws each attempt consumes.  So: (1) one thread extends the raw uniform stream U[k] = next() % N (a few
// clocks per draw); (2) every stream offset k is evaluated IN PARALLEL as if an attempt started there: the four distinct indices,
// the number of draws consumed, checkSubset's verdict; (3) one thread walks the offsets from the current position -- a table
// look-up per attempt from shared memory -- and emits the subsets OpenCV would have produced.  A sequential generator reading
// the points from global memory took 2.5 us per subset (5 ms for the 2000 iterations a non-planar scene needs).
struct HAttempt { int32_t idx[4]; };

DFVO_D uint32_t h_rng_next(uint64_t* state) {
  *state = (uint64_t)(uint32_t)*state * 4164903690u + (uint32_t)(*state >> 32);      // cv::RNG::next (multiply-with-carry)
  return (uint32_t)*state;
}

// Jump-ahead.  cv::RNG is a multiply-with-carry generator: with b = 2^32, S = c*b + x  ->  S' = a*x + c, hence S'*b = S (mod m),
// m = a*b - 1: the state sequence is S_n = S_2 * r^(n-2) mod m for n >= 2 (r = b^-1 mod m; S_1 of the seed 2^64-1 is not yet a
// canonical residue, S_2 is).  Any thread can therefore start the stream at any draw index, and the raw stream of a whole RANSAC
// round is generated in parallel instead of by one thread (64 000 sequential draws were 0.4 ms).
#define H_RNG_M 0xf83f6309ffffffffull          // 4164903690 * 2^32 - 1
#define H_RNG_R 0x00000000f83f630aull          // (2^32)^-1 mod m
#define H_RNG_S1 0xf83f630a07c09cf5ull         // state after the first draw from the seed (uint64)-1
#define H_RNG_S2 0x07848374bac3439cull         // ... after the second
DFVO_D uint64_t h_addmod(uint64_t x, uint64_t y) {            // x, y < m
  const uint64_t s = x + y;
  return (s < x || s >= H_RNG_M) ? s - H_RNG_M : s;
}
DFVO_D uint64_t h_mulmod(uint64_t x, uint64_t y) {            // binary double-and-add (m is within 3 % of 2^64)
  uint64_t r = 0;
  for (int i = 63; i >= 0; --i) {
    r = h_addmod(r, r);
    if ((y >> i) & 1ull) r = h_addmod(r, x);
  }
  return r;
}
// generator state after n draws (n >= 0) from the seed cv::RNG((uint64)-1)
DFVO_D uint64_t h_rng_state_after(uint32_t n) {
  if (n == 0) return 0xFFFFFFFFFFFFFFFFull;
  if (n == 1) return H_RNG_S1;
  uint64_t acc = H_RNG_S2, base = H_RNG_R;
  for (uint32_t e = n - 2; e; e >>= 1) {
    if (e & 1u) acc = h_mulmod(acc, base);
    base = h_mulmod(base, base);
  }
  return acc;
}

// attempt starting at stream offset k: consumed draws (0 if the stream is too short), accepted flag, indices
DFVO_D uint16_t h_eval_attempt(const uint32_t* __restrict__ U, int k, int ndraws, const float* __restrict__ src, const float* __restrict__ dst,
                               HAttempt* __restrict__ out) {
  int idx[4], c = 0;
  float sp[8], dp[8];
  for (int i = 0; i < 4;) {
    if (k + c >= ndraws) return 0;
    const int v = (int)U[k + c];
    ++c;
    bool dup = false;
    for (int j = 0; j < i; ++j) dup = dup || idx[j] == v;
    if (dup) continue;
    idx[i] = v;
    sp[2 * i] = src[2 * v]; sp[2 * i + 1] = src[2 * v + 1]; dp[2 * i] = dst[2 * v]; dp[2 * i + 1] = dst[2 * v + 1];
    ++i;
  }
  const bool found = hmg::check_subset(sp, dp);
  for (int i = 0; i < 4; ++i) out->idx[i] = idx[i];
  return (uint16_t)(c | (found ? 0x8000 : 0));
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


// acceptance rule of RANSACPointSetRegistrator::run over iterations [st.it, min(niters, i1))
DFVO_D void h_replay(HState* st, const int32_t* __restrict__ ok, const int32_t* __restrict__ counts, int N, int i1, double prob) {
  HState s2 = *st;
  if (s2.done) return;
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
  *st = s2;
}

__global__ void k_h_init(HState* st, int max_iters) {
  if (threadIdx.x == 0) {
    st->niters = max_iters; st->best_good = -1; st->best_iter = -1; st->it = 0; st->done = 0; st->n_subsets = 0;
    st->pos = 0; st->ndraws = 0; st->failed = 0; st->rng = 0;
  }
}

#define H_DRAWS_PER_SUBSET 40     // window of raw draws per subset: ~4 when nothing is rejected, ~32 with 60 % gross outliers
#define H_DRAW_CHUNK 32

// round start (one thread): replay of the previous round, then the window of raw draws the coming round may need
__global__ void k_h_round_begin(HState* st, const int32_t* __restrict__ ok, const int32_t* __restrict__ counts, int N, int prev_i1, int i0, int i1,
                                double prob, int max_draws) {
  if (threadIdx.x != 0) return;
  if (prev_i1 > 0) h_replay(st, ok, counts, N, prev_i1, prob);
  if (st->done) return;
  int target = st->pos + (i1 - i0) * H_DRAWS_PER_SUBSET + 1024;
  if (target > max_draws) target = max_draws;
  if (target > st->ndraws) st->ndraws = target;
}

// raw stream U[k] = draw k % N for the window [pos, ndraws): each thread jumps to its chunk and runs the generator from there
__global__ void __launch_bounds__(128)
k_h_draws(const HState* __restrict__ st, int N, uint32_t* __restrict__ U) {
  if (st->done) return;
  const int k0 = st->pos + (blockIdx.x * blockDim.x + threadIdx.x) * H_DRAW_CHUNK;
  if (k0 >= st->ndraws) return;
  uint64_t state = h_rng_state_after((uint32_t)k0);
  const int k1 = k0 + H_DRAW_CHUNK < st->ndraws ? k0 + H_DRAW_CHUNK : st->ndraws;
  for (int k = k0; k < k1; ++k) U[k] = h_rng_next(&state) % (uint32_t)N;
}

// every stream offset of the window as a potential attempt start
__global__ void __launch_bounds__(128)
k_h_attempts(const HState* __restrict__ st, const uint32_t* __restrict__ U, const float* __restrict__ src, const float* __restrict__ dst,
             uint16_t* __restrict__ att, HAttempt* __restrict__ atti) {
  if (st->done) return;
  const int k = st->pos + blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= st->ndraws) return;
  att[k] = h_eval_attempt(U, k, st->ndraws, src, dst, &atti[k]);
}

// the sequential walk: subsets [i0, i1) from the attempt table (thread 0, table staged in shared memory), then the copy-out
__global__ void __launch_bounds__(256)
k_h_walk(HState* st, uint32_t* __restrict__ U, int max_draws, const float* __restrict__ src, const float* __restrict__ dst, int N,
         uint16_t* __restrict__ att, HAttempt* __restrict__ atti, int i0, int i1, int32_t* __restrict__ acc_off, int32_t* __restrict__ subsets, int cap) {
  DFVO_DYN_SMEM(uint16_t, tab);
  __shared__ int made_s;
  if (st->done) return;
  const int t = threadIdx.x, p0 = st->pos;
  int span = st->ndraws - p0;
  if (span > cap) span = cap;
  for (int k = t; k < span; k += 256) tab[k] = att[p0 + k];
  __syncthreads();
  if (t == 0) {
    int p = p0, made = i0, ndraws = st->ndraws;
    bool failed = false;
    for (int it = i0; it < i1 && !failed; ++it) {
      bool found = false;
      for (int attempt = 0; attempt < 10000 && !found; ++attempt) {
        uint16_t a = (p - p0 < span) ? tab[p - p0] : (p < ndraws ? att[p] : (uint16_t)0);
        if ((a & 0x7fff) == 0) {
          // the pre-evaluated window is exhausted (more rejections than H_DRAWS_PER_SUBSET allows for): extend the stream from
          // this offset and evaluate the attempt here, sequentially
          if (ndraws < p + 64) {
            int to = p + 64 < max_draws ? p + 64 : max_draws;
            uint64_t state = h_rng_state_after((uint32_t)ndraws);
            for (int k = ndraws; k < to; ++k) U[k] = h_rng_next(&state) % (uint32_t)N;
            ndraws = to > ndraws ? to : ndraws;
          }
          a = h_eval_attempt(U, p, ndraws, src, dst, &atti[p]);
          if ((a & 0x7fff) == 0) { failed = true; break; }            // stream capacity exhausted
        }
        if (a & 0x8000) { acc_off[it] = p; found = true; }
        p += a & 0x7fff;
      }
      if (!found) break;                                          // OpenCV stops the loop here
      made = it + 1;
    }
    st->pos = p; st->ndraws = ndraws; st->n_subsets = made;
    made_s = made;
  }
  __syncthreads();
  for (int it = i0 + t; it < made_s; it += 256) {
    const HAttempt a = atti[acc_off[it]];
    for (int i = 0; i < 4; ++i) subsets[it * 4 + i] = a.idx[i];
  }
}

__global__ void __launch_bounds__(64)
k_h_hyp(const HState* __restrict__ st, const float* __restrict__ src, const float* __restrict__ dst, const int32_t* __restrict__ subsets, int i0,
        int i1, double* __restrict__ hyp, int32_t* __restrict__ ok) {
  if (st->done) return;
  const int it = i0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (it >= i1) return;
  if (it < st->n_subsets) h_hypothesis(src, dst, subsets, it, hyp, ok);
  else ok[it] = 0;
}

__global__ void __launch_bounds__(256)
k_h_score(const HState* __restrict__ st, const float* __restrict__ src, const float* __restrict__ dst, int N, int i0, int i1, float thr2,
          const double* __restrict__ hyp, const int32_t* __restrict__ ok, int32_t* __restrict__ counts) {
  if (st->done) return;
  const int it = i0 + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (it >= i1) return;
  int c = 0;
  if (ok[it]) {
    float Hf[8];
    for (int q = 0; q < 8; ++q) Hf[q] = (float)hyp[(size_t)it * 9 + q];
    for (int j = lane; j < N; j += 32) c += (h_err(Hf, src[2 * j], src[2 * j + 1], dst[2 * j], dst[2 * j + 1]) <= thr2) ? 1 : 0;
  }
  for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
  if (lane == 0) counts[it] = c;
}

// ONE block finishes the estimation: replay of the last round, the inlier mask of the winner, the DLT over the inliers, the LM
// refinement and GRIC-H.
// out: H_out [9], mask [N] (inliers of the refined H), info [4] = {found, inliers, iterations, winning iteration}, gric [1]
__global__ void __launch_bounds__(256)
k_h_finalize(HState* st, const float* __restrict__ src, const float* __restrict__ dst, const double* __restrict__ p1, const double* __restrict__ p2,
             int N, int max_iters, double prob, float thr2, const double* __restrict__ hyp, const int32_t* __restrict__ ok,
             const int32_t* __restrict__ counts, double* __restrict__ H_out, uint8_t* __restrict__ mask, int32_t* __restrict__ info,
             double* __restrict__ gric) {
  __shared__ double part[8][48];
  __shared__ double acc[48];
  __shared__ double Hs[9], xs[8], xd[8], dvec[8], Amat[8][8], vvec[8];
  __shared__ double S_cur, S_new, lambda, lc;
  __shared__ int stop, ninl;
  __shared__ HState sst;
  const int t = threadIdx.x;
  if (t == 0) {
    h_replay(st, ok, counts, N, max_iters, prob);
    sst = *st;
  }
  __syncthreads();
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

static int h_max_draws(int max_iters) { return max_iters * (H_DRAWS_PER_SUBSET + 8) + 8192; }

size_t homography_workspace_bytes(int N, int max_iters) {
  const size_t nd = (size_t)h_max_draws(max_iters);
  return (size_t)N * 4 * 4 + (size_t)max_iters * (9 * 8 + 4 * 4 + 4 + 4 + 4) + nd * (4 + 2 + sizeof(HAttempt)) + sizeof(HState) + 4096;
}

int homography_ransac(const double* p1, const double* p2, int N, int max_iters, double threshold, double prob, void* workspace, size_t ws_bytes,
                      double* H_out, uint8_t* mask_out, int32_t* info, double* gric, cudaStream_t s) {
  DFVO_REQUIRE(p1 && p2 && H_out && mask_out && info && gric && N >= 4 && max_iters >= 1, DFVO_EINVAL, "homography_ransac args (N=%d)", N);
  DFVO_REQUIRE(ws_bytes >= homography_workspace_bytes(N, max_iters), DFVO_EINVAL, "homography_ransac workspace too small");
  uint8_t* w = reinterpret_cast<uint8_t*>(workspace);
  auto take = [&](size_t bytes) { uint8_t* p = w; w += (bytes + 127) & ~(size_t)127; return p; };
  const int nd = h_max_draws(max_iters);
  float* src = (float*)take((size_t)N * 2 * 4);
  float* dst = (float*)take((size_t)N * 2 * 4);
  double* hyp = (double*)take((size_t)max_iters * 9 * 8);
  int32_t* subsets = (int32_t*)take((size_t)max_iters * 4 * 4);
  int32_t* ok = (int32_t*)take((size_t)max_iters * 4);
  int32_t* counts = (int32_t*)take((size_t)max_iters * 4);
  int32_t* acc_off = (int32_t*)take((size_t)max_iters * 4);
  uint32_t* U = (uint32_t*)take((size_t)nd * 4);
  uint16_t* att = (uint16_t*)take((size_t)nd * 2);
  HAttempt* atti = (HAttempt*)take((size_t)nd * sizeof(HAttempt));
  HState* st = (HState*)take(sizeof(HState));
  const float thr2 = (float)(threshold * threshold);
  DFVO_LAUNCH(k_h_prepare, dim3(cdiv(N, 128)), dim3(128), 0, s, p1, p2, N, src, dst);
  DFVO_LAUNCH(k_h_init, dim3(1), dim3(32), 0, s, st, max_iters);
  // rounds: a scene with a dominant plane (or no outliers) stops inside the first (measured 96 .. 124 iterations at 0 % outliers), a
  // general scene at 30 % outliers inside the second (338 .. 669), anything worse runs all max_iters.  The sequential walk of a round
  // costs ~0.16 us per subset, so the middle round saves ~0.2 ms of the homography chain exactly where that chain is the
  // tracker's critical path (30 % outliers: essential-matrix chain 0.45 ms, homography chain 0.81 -> 0.6 ms)
  const int bounds[4] = {0, 256, 768, max_iters};
  int prev_i1 = 0;
#ifndef DFVO_HOSTSIM
  static bool attr_set = false;
  if (!attr_set) { DFVO_CUDA(cudaFuncSetAttribute(k_h_walk, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); attr_set = true; }
#endif
  for (int rd = 0; rd < 3; ++rd) {
    const int i0 = bounds[rd] < max_iters ? bounds[rd] : max_iters, i1 = bounds[rd + 1] < max_iters ? bounds[rd + 1] : max_iters;
    if (i1 <= i0) continue;
    DFVO_LAUNCH(k_h_round_begin, dim3(1), dim3(32), 0, s, st, ok, counts, N, prev_i1, i0, i1, prob, nd);
    const int window = (i1 - i0) * H_DRAWS_PER_SUBSET + 1024;
    DFVO_LAUNCH(k_h_draws, dim3(cdiv(cdiv(window, H_DRAW_CHUNK), 128)), dim3(128), 0, s, st, N, U);
    DFVO_LAUNCH(k_h_attempts, dim3(cdiv(window, 128)), dim3(128), 0, s, st, U, src, dst, att, atti);
    const int cap = window < 80 * 1024 ? window : 80 * 1024;         // table entries staged in shared memory (2 B each)
    DFVO_LAUNCH(k_h_walk, dim3(1), dim3(256), (size_t)cap * 2, s, st, U, nd, src, dst, N, att, atti, i0, i1, acc_off, subsets, cap);
    DFVO_LAUNCH(k_h_hyp, dim3(cdiv(i1 - i0, 64)), dim3(64), 0, s, st, src, dst, subsets, i0, i1, hyp, ok);
    DFVO_LAUNCH(k_h_score, dim3(cdiv((i1 - i0) * 32, 256)), dim3(256), 0, s, st, src, dst, N, i0, i1, thr2, hyp, ok, counts);
    prev_i1 = i1;
  }
  DFVO_LAUNCH(k_h_finalize, dim3(1), dim3(256), 0, s, st, src, dst, p1, p2, N, max_iters, prob, thr2, hyp, ok, counts, H_out, mask_out, info, gric);
  DFVO_CHECK_LAUNCH();
  return DFVO_OK;
}

}  // namespace dfvo
