// Essential-matrix RANSAC on the device, reproducing cv2.findEssentialMat + cv2.recoverPose as the
// reference calls them (E_tracker.py:223-296; OpenCV calib3d ptsetreg.cpp / five-point.cpp, SURVEY
// Appendix C): "parallel-evaluate / sequential-replay".
//   1. the 5-point subsets OpenCV would draw depend only on N (its RNG is re-seeded per call); the
//      host supplies that table (b200/cvrng.py) and the permutations drawn from the host np.random;
//   2. k_hypotheses: one thread per (repeat, iteration) solves the 5-point problem (fivept.cuh, FP64);
//   3. k_score: one warp per candidate counts Sampson inliers over all N correspondences (FP64);
//   4. k_replay: one thread per repeat walks the iterations in order applying OpenCV's acceptance rule
//      (strict >, first-found wins) and adaptive iteration count, i.e. finds the candidate OpenCV
//      would return and where it would stop;
//   5. k_finalize: inlier mask + GRIC-E residual sum (gric.py:14-37,94-132) of each repeat's winner;
//   6. k_recover_pose_vote / _pick: decomposeEssentialMat + 4 x triangulation + cheirality vote.
// All arithmetic is FP64 (inlier decisions are threshold tests).
#include <stdlib.h>

#include "fivept.cuh"
#include "ops.h"
#include "ransac.h"
#include "smallmat.cuh"

namespace dfvo {

// ---------------------------------------------------------------------------------------------
// stage kernels
// ---------------------------------------------------------------------------------------------
__global__ void k_five_point(const double* __restrict__ x1, const double* __restrict__ x2, int M, double* __restrict__ E,
                             int32_t* __restrict__ n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  double a[10], b[10], e[90];
  for (int k = 0; k < 10; ++k) { a[k] = x1[i * 10 + k]; b[k] = x2[i * 10 + k]; }
  int c = fivept::solve(a, b, e);
  n[i] = c;
  for (int k = 0; k < 90; ++k) E[(size_t)i * 90 + k] = k < 9 * c ? e[k] : 0.0;
}

int five_point(const double* x1, const double* x2, int M, double* E, int32_t* n, cudaStream_t s) {
  DFVO_LAUNCH(k_five_point, dim3(cdiv(M, 64)), dim3(64), 0, s, x1, x2, M, E, n);
  DFVO_CHECK_LAUNCH();
  return DFVO_OK;
}

DFVO_D double sampson(const double* E, double u1, double v1, double u2, double v2) {
  // EMEstimatorCallback::computeError: x2^T E x1 squared over the four gradient terms
  const double a0 = E[0] * u1 + E[1] * v1 + E[2];
  const double a1 = E[3] * u1 + E[4] * v1 + E[5];
  const double a2 = E[6] * u1 + E[7] * v1 + E[8];
  const double b0 = E[0] * u2 + E[3] * v2 + E[6];
  const double b1 = E[1] * u2 + E[4] * v2 + E[7];
  const double x2tEx1 = u2 * a0 + v2 * a1 + a2;
  return x2tEx1 * x2tEx1 / (a0 * a0 + a1 * a1 + b0 * b0 + b1 * b1);
}

// one warp per candidate; counts[m] = #{ i : sampson(E_m, x1_i, x2_i) <= thr2 }.  perm (optional) maps
// slot -> original point index per repeat (repeat = m / cand_per_repeat).
__global__ void __launch_bounds__(256)
k_score(const double* __restrict__ E, const int32_t* __restrict__ valid, int M, const double* __restrict__ x1,
        const double* __restrict__ x2, int N, double thr2, int32_t* __restrict__ counts) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  if (valid && !valid[warp]) { if (lane == 0) counts[warp] = 0; return; }
  double e[9];
  for (int k = 0; k < 9; ++k) e[k] = E[(size_t)warp * 9 + k];
  int c = 0;
  for (int i = lane; i < N; i += 32) {
    double err = sampson(e, x1[2 * i], x1[2 * i + 1], x2[2 * i], x2[2 * i + 1]);
    c += (err <= thr2) ? 1 : 0;
  }
  for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
  if (lane == 0) counts[warp] = c;
}

int score_hypotheses(const double* E, int M, const double* x1, const double* x2, int N, double thr2, int32_t* counts,
                     cudaStream_t s) {
  DFVO_LAUNCH(k_score, dim3(cdiv(M * 32, 256)), dim3(256), 0, s, E, (const int32_t*)nullptr, M, x1, x2, N, thr2, counts);
  DFVO_CHECK_LAUNCH();
  return DFVO_OK;
}

// ---------------------------------------------------------------------------------------------
// fused pipeline
// ---------------------------------------------------------------------------------------------
struct EssState {          // one per repeat, device memory
  int32_t niters, best_good, best_iter, best_cand, it, done, evaluated, pad;
};

// x1n/x2n [R][N][2]: normalised AND permuted points of each repeat
__global__ void k_normalize_perm(const double* __restrict__ p1, const double* __restrict__ p2, const int32_t* __restrict__ perm,
                                 int N, double focal, double cx, double cy, double* __restrict__ x1n, double* __restrict__ x2n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (i >= N) return;
  int src = perm ? perm[(size_t)r * N + i] : i;
  size_t o = ((size_t)r * N + i) * 2;
  x1n[o] = (p1[2 * src] - cx) / focal; x1n[o + 1] = (p1[2 * src + 1] - cy) / focal;
  x2n[o] = (p2[2 * src] - cx) / focal; x2n[o + 1] = (p2[2 * src + 1] - cy) / focal;
}

__global__ void k_ess_init(EssState* st, int R, int max_iters) {
  int r = threadIdx.x;
  if (r < R) { st[r].niters = max_iters; st[r].best_good = -1; st[r].best_iter = -1; st[r].best_cand = -1; st[r].it = 0; st[r].done = 0; st[r].evaluated = 0; }
}

__global__ void k_hypotheses(const double* __restrict__ x1n, const double* __restrict__ x2n, const int32_t* __restrict__ subsets,
                             int N, int i0, int i1, const EssState* __restrict__ st, double* __restrict__ Ecand,
                             int32_t* __restrict__ ncand, int max_iters) {
  int i = i0 + blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (i >= i1) return;
  if (st[r].done) return;
  double a[10], b[10], e[90];
  for (int k = 0; k < 5; ++k) {
    int s = subsets[i * 5 + k];
    size_t o = ((size_t)r * N + s) * 2;
    a[2 * k] = x1n[o]; a[2 * k + 1] = x1n[o + 1];
    b[2 * k] = x2n[o]; b[2 * k + 1] = x2n[o + 1];
  }
  int c = fivept::solve(a, b, e);
  size_t h = (size_t)r * max_iters + i;
  ncand[h] = c;
  for (int k = 0; k < 9 * c; ++k) Ecand[h * 90 + k] = e[k];
}

// Warp-cooperative hypothesis generation (fivept::solve_coop): ten lanes per minimal sample, three samples per warp.
#ifdef DFVO_HOSTSIM
#define HYP_WARPS 1          // the CPU emulation pays per shuffle and per thread of the block
#else
#define HYP_WARPS 4
#endif
__global__ void __launch_bounds__(HYP_WARPS * 32)
k_hypotheses_coop(const double* __restrict__ x1n, const double* __restrict__ x2n, const int32_t* __restrict__ subsets, int N, int i0, int i1,
                  const EssState* __restrict__ st, double* __restrict__ Ecand, int32_t* __restrict__ ncand, int max_iters) {
  __shared__ fivept::CoopShared sm[HYP_WARPS * 3];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, r = blockIdx.y;
  if (st[r].done) return;                                  // uniform per block
  int g = lane / 10, l = lane - 10 * g;
  if (g == 3) { g = 2; l += 10; }                           // lanes 30, 31 ride along with the third group
  int i = i0 + (blockIdx.x * HYP_WARPS + warp) * 3 + g;
  const bool live = i < i1;
  if (!live) i = i1 - 1;                                    // duplicate work on a valid sample, results discarded
  double a[10], b[10];
  for (int k = 0; k < 5; ++k) {
    const int s = subsets[i * 5 + k];
    const size_t o = ((size_t)r * N + s) * 2;
    a[2 * k] = x1n[o]; a[2 * k + 1] = x1n[o + 1];
    b[2 * k] = x2n[o]; b[2 * k + 1] = x2n[o + 1];
  }
  const size_t h = (size_t)r * max_iters + i;
  // dead groups write nowhere: E_out is only dereferenced by lanes whose candidate exists (live is folded into `ok`)
  const int c = fivept::solve_coop(a, b, &sm[warp * 3 + g], l, 10 * g, live, Ecand + h * 90);
  if (live && l == 0) ncand[h] = c;
}

__global__ void __launch_bounds__(256)
k_score_round(const double* __restrict__ Ecand, const int32_t* __restrict__ ncand, const double* __restrict__ x1n,
              const double* __restrict__ x2n, int N, int i0, int i1, double thr2, const EssState* __restrict__ st,
              int32_t* __restrict__ counts, int max_iters) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, r = blockIdx.y;
  const int i = i0 + w / 10, k = w % 10;
  if (i >= i1) return;
  if (st[r].done) return;
  size_t h = (size_t)r * max_iters + i;
  if (k >= ncand[h]) return;
  double e[9];
  for (int q = 0; q < 9; ++q) e[q] = Ecand[h * 90 + 9 * k + q];
  const double* a = x1n + (size_t)r * N * 2;
  const double* b = x2n + (size_t)r * N * 2;
  int c = 0;
  for (int j = lane; j < N; j += 32) c += (sampson(e, a[2 * j], a[2 * j + 1], b[2 * j], b[2 * j + 1]) <= thr2) ? 1 : 0;
  for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
  if (lane == 0) counts[h * 10 + k] = c;
}

DFVO_HD int ransac_update_num_iters(double p, double ep, int model_points, int max_iters) {
  // cv::RANSACUpdateNumIters (ptsetreg.cpp)
  p = p < 0 ? 0 : (p > 1 ? 1 : p);
  ep = ep < 0 ? 0 : (ep > 1 ? 1 : ep);
  double num = 1.0 - p;
  if (num < 2.2250738585072014e-308) num = 2.2250738585072014e-308;
  double denom = 1.0 - pow(1.0 - ep, (double)model_points);
  if (denom < 2.2250738585072014e-308) return 0;
  num = log(num);
  denom = log(denom);
  if (denom >= 0 || -num >= max_iters * (-denom)) return max_iters;
  double q = num / denom;
  return (int)rint(q);           // cvRound: round half to even
}

__global__ void k_replay(const int32_t* __restrict__ ncand, const int32_t* __restrict__ counts, int N, int i1, double prob,
                         EssState* st, int max_iters, int R) {
  int r = threadIdx.x + blockIdx.x * blockDim.x;
  if (r >= R) return;
  EssState s = st[r];
  if (s.done) return;
  int it = s.it;
  while (it < s.niters && it < i1) {
    size_t h = (size_t)r * max_iters + it;
    int nc = ncand[h];
    for (int k = 0; k < nc; ++k) {
      int good = counts[h * 10 + k];
      int lim = s.best_good > 4 ? s.best_good : 4;
      if (good > lim) {
        s.best_good = good; s.best_iter = it; s.best_cand = k;
        s.niters = ransac_update_num_iters(prob, (double)(N - good) / (double)N, 5, s.niters);
      }
    }
    ++it;
  }
  s.it = it;
  s.evaluated = i1;
  if (it >= s.niters) s.done = 1;
  st[r] = s;
}

// Warp version of k_replay (one warp per repeat): the acceptance rule is sequential only through the running best, so each lane
// loads one iteration's candidate counts (max and its first index) -- 32 iterations per round trip to memory instead of one --
// and the warp then applies, in iteration order, only the iterations that beat the running best (a handful).  Same state
// transitions as k_replay: within an iteration the first candidate with the largest count wins, and niters after several
// improvements equals RANSACUpdateNumIters of the last one (it only ever shrinks with the inlier ratio).
__global__ void __launch_bounds__(32)
k_replay_warp(const int32_t* __restrict__ ncand, const int32_t* __restrict__ counts, int N, int i1, double prob, EssState* st, int max_iters) {
  const int r = blockIdx.x, lane = threadIdx.x;
  EssState s = st[r];
  if (s.done) return;
  int it = s.it;
  while (it < s.niters && it < i1) {
    const int mine = it + lane;
    int m = -1, mk = -1;
    if (mine < i1) {
      const size_t h = (size_t)r * max_iters + mine;
      const int nc = ncand[h];
      for (int k = 0; k < nc; ++k) { const int g = counts[h * 10 + k]; if (g > m) { m = g; mk = k; } }
    }
    int from = 0;                                             // lanes below `from` are settled
    while (true) {
      const int lim = s.best_good > 4 ? s.best_good : 4;
      const int upto = (s.niters < i1 ? s.niters : i1) - it;   // iterations of this batch that OpenCV would still run
      const unsigned cand = __ballot_sync(0xffffffffu, lane >= from && lane < upto && m > lim);
      if (!cand) break;
      const int j = __ffs(cand) - 1;
      const int good = __shfl_sync(0xffffffffu, m, j), k = __shfl_sync(0xffffffffu, mk, j);
      // within the winning iteration candidates are visited in order: the first one that beats the running best may be an earlier,
      // smaller one -- but every later strictly larger one replaces it, so the iteration ends on its maximum (first occurrence)
      s.best_good = good; s.best_iter = it + j; s.best_cand = k;
      s.niters = ransac_update_num_iters(prob, (double)(N - good) / (double)N, 5, s.niters);
      from = j + 1;
    }
    const int upto = (s.niters < i1 ? s.niters : i1) - it;
    it += upto < 32 ? (upto > 0 ? upto : 0) : 32;
    if (upto <= 0) break;
  }
  if (lane == 0) {
    s.it = it;
    s.evaluated = i1;
    if (it >= s.niters) s.done = 1;
    st[r] = s;
  }
}

// mask of the winner in ORIGINAL point order (E_tracker.py:278-285 un-permutes), GRIC-E (gric.py), counts.
// out per repeat: E[9], info[4] = {inlier count, iterations, best_iter, best_cand}, gric
__global__ void __launch_bounds__(256)
k_finalize(const double* __restrict__ Ecand, const EssState* __restrict__ st, const double* __restrict__ x1n,
           const double* __restrict__ x2n, const int32_t* __restrict__ perm, const double* __restrict__ p1,
           const double* __restrict__ p2, int N, double thr2, double fx, double fy, double cx, double cy, int max_iters,
           double* __restrict__ E_out, uint8_t* __restrict__ mask_out, int32_t* __restrict__ info, double* __restrict__ gric) {
  __shared__ double red[256];
  const int r = blockIdx.x, t = threadIdx.x;
  const EssState s = st[r];
  double e[9];
  const bool have = s.best_iter >= 0;
  if (have) for (int q = 0; q < 9; ++q) e[q] = Ecand[((size_t)r * max_iters + s.best_iter) * 90 + 9 * s.best_cand + q];
  else for (int q = 0; q < 9; ++q) e[q] = 0.0;
  // F = K^-T E K^-1 (E_tracker.py:261-262); K = [[fx,0,cx],[0,fy,cy],[0,0,1]]
  double F[9];
  {
    // K^-1 = [[1/fx,0,-cx/fx],[0,1/fy,-cy/fy],[0,0,1]]
    const double ki[9] = {1.0 / fx, 0.0, -cx / fx, 0.0, 1.0 / fy, -cy / fy, 0.0, 0.0, 1.0};
    double T[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double a = 0; for (int k = 0; k < 3; ++k) a += e[3 * i + k] * ki[3 * k + j]; T[3 * i + j] = a; }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double a = 0; for (int k = 0; k < 3; ++k) a += ki[3 * k + i] * T[3 * k + j]; F[3 * i + j] = a; }
  }
  const double* a = x1n + (size_t)r * N * 2;
  const double* b = x2n + (size_t)r * N * 2;
  double gsum = 0.0;
  const double sigmasq1 = 1.0 / (0.8 * 0.8), lam3RD = 2.0 * (4 - 3);
  for (int j = t; j < N; j += 256) {
    int src = perm ? perm[(size_t)r * N + j] : j;
    bool inl = have && (sampson(e, a[2 * j], a[2 * j + 1], b[2 * j], b[2 * j + 1]) <= thr2);
    mask_out[(size_t)r * N + src] = inl ? 1 : 0;
    // compute_fundamental_residual(F, kp1=points1, kp2=points2) on pixel coordinates (gric.py:14-37)
    const double u1 = p1[2 * src], v1 = p1[2 * src + 1], u2 = p2[2 * src], v2 = p2[2 * src + 1];
    const double f0 = F[0] * u1 + F[1] * v1 + F[2], f1 = F[3] * u1 + F[4] * v1 + F[5], f2 = F[6] * u1 + F[7] * v1 + F[8];
    const double g0 = F[0] * u2 + F[3] * v2 + F[6], g1 = F[1] * u2 + F[4] * v2 + F[7];
    const double m = u2 * f0 + v2 * f1 + f2;
    const double res = m * m / (f0 * f0 + f1 * f1 + g0 * g0 + g1 * g1);
    const double tmp = res * sigmasq1;
    gsum += (tmp <= lam3RD) ? tmp : lam3RD;
  }
  red[t] = gsum;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) { if (t < off) red[t] += red[t + off]; __syncthreads(); }
  if (t == 0) {
    // calc_GRIC(res, 0.8, n, 'EMat'): sum + n*D*log(R) + K*log(R*n), R=4, D=3, K=5
    gric[r] = red[0] + (double)N * 3.0 * log(4.0) + 5.0 * log(4.0 * (double)N);
    for (int q = 0; q < 9; ++q) E_out[r * 9 + q] = e[q];
    info[r * 4 + 0] = have ? s.best_good : 0; info[r * 4 + 1] = s.it; info[r * 4 + 2] = s.best_iter; info[r * 4 + 3] = s.best_cand;
  }
}

size_t essential_workspace_bytes(int N, int R, int max_iters) {
  size_t b = 0;
  b += (size_t)R * N * 2 * 8 * 2;                 // x1n, x2n
  b += (size_t)R * max_iters * 90 * 8;            // candidates
  b += (size_t)R * max_iters * 4;                 // ncand
  b += (size_t)R * max_iters * 10 * 4;            // counts
  b += (size_t)R * sizeof(EssState);
  return b + 1024;
}

int essential_ransac(const double* p1, const double* p2, int N, const int32_t* perm, int R, const int32_t* subsets, int max_iters,
                     double fx, double fy, double cx, double cy, double threshold, double prob, void* workspace, size_t ws_bytes,
                     double* E_out, uint8_t* mask_out, int32_t* info, double* gric, cudaStream_t s) {
  DFVO_REQUIRE(N >= 5 && R >= 1 && R <= 32 && max_iters >= 1, DFVO_EINVAL, "essential_ransac args (N=%d R=%d)", N, R);
  DFVO_REQUIRE(ws_bytes >= essential_workspace_bytes(N, R, max_iters), DFVO_EINVAL, "essential_ransac workspace too small");
  uint8_t* w = reinterpret_cast<uint8_t*>(workspace);
  auto take = [&](size_t bytes) { uint8_t* p = w; w += (bytes + 127) & ~(size_t)127; return p; };
  double* x1n = (double*)take((size_t)R * N * 2 * 8);
  double* x2n = (double*)take((size_t)R * N * 2 * 8);
  double* Ecand = (double*)take((size_t)R * max_iters * 90 * 8);
  int32_t* ncand = (int32_t*)take((size_t)R * max_iters * 4);
  int32_t* counts = (int32_t*)take((size_t)R * max_iters * 10 * 4);
  EssState* st = (EssState*)take((size_t)R * sizeof(EssState));
  const double focal = fx;                                    // findEssentialMat(focal=fx, pp) (E_tracker.py:231-239)
  const double thr = threshold / focal, thr2 = thr * thr;
  DFVO_LAUNCH(k_normalize_perm, dim3(cdiv(N, 128), R), dim3(128), 0, s, p1, p2, perm, N, focal, cx, cy, x1n, x2n);
  DFVO_LAUNCH(k_ess_init, dim3(1), dim3(32), 0, s, st, R, max_iters);
  // rounds: most scenes stop within the first few dozen iterations; later rounds early-exit on st.done
  const int bounds[4] = {0, 48 < max_iters ? 48 : max_iters, 256 < max_iters ? 256 : max_iters, max_iters};
  for (int rd = 0; rd < 3; ++rd) {
    const int i0 = bounds[rd], i1 = bounds[rd + 1];
    if (i1 <= i0) continue;
    static int coop = -1;
    if (coop < 0) { const char* e = getenv("DFVO_HYP_COOP"); coop = !(e && atoi(e) == 0); }
    if (coop)
      DFVO_LAUNCH(k_hypotheses_coop, dim3(cdiv(i1 - i0, HYP_WARPS * 3), R), dim3(HYP_WARPS * 32), 0, s, x1n, x2n, subsets, N, i0, i1, st, Ecand,
                  ncand, max_iters);
    else
      DFVO_LAUNCH(k_hypotheses, dim3(cdiv(i1 - i0, 32), R), dim3(32), 0, s, x1n, x2n, subsets, N, i0, i1, st, Ecand, ncand, max_iters);
    DFVO_LAUNCH(k_score_round, dim3(cdiv((i1 - i0) * 10 * 32, 256), R), dim3(256), 0, s, Ecand, ncand, x1n, x2n, N, i0, i1, thr2, st,
                counts, max_iters);
    DFVO_LAUNCH(k_replay_warp, dim3(R), dim3(32), 0, s, ncand, counts, N, i1, prob, st, max_iters);
  }
  DFVO_LAUNCH(k_finalize, dim3(R), dim3(256), 0, s, Ecand, st, x1n, x2n, perm, p1, p2, N, thr2, fx, fy, cx, cy, max_iters, E_out,
              mask_out, info, gric);
  DFVO_CHECK_LAUNCH();
  return DFVO_OK;
}

// ---------------------------------------------------------------------------------------------
// cv2.recoverPose(E, points1, points2, focal, pp)  (E_tracker.py:292-295; five-point.cpp)
// decomposeEssentialMat, then every point is triangulated (4x4 DLT, smallest singular vector as in
// cv::triangulatePoints) against the four (R,t) candidates; cheirality masks in OpenCV's order; first
// maximum wins.  out: Rt[12] (R row-major then t), info[5] = {best count, c0..c3}.
// ---------------------------------------------------------------------------------------------
DFVO_D void triangulate_dlt(const double P1[3][4], double u0, double v0, double u1, double v1, double X[4]) {
  // rows of A (cvTriangulatePoints): x*P[2] - P[0], y*P[2] - P[1] for view 0 = [I|0] and view 1 = P1
  double A[4][4] = {{-1.0, 0.0, u0, 0.0}, {0.0, -1.0, v0, 0.0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
  for (int k = 0; k < 4; ++k) { A[2][k] = u1 * P1[2][k] - P1[0][k]; A[3][k] = v1 * P1[2][k] - P1[1][k]; }
  double AtA[4][4], V[4][4], w[4];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) { double a = 0; for (int k = 0; k < 4; ++k) a += A[k][i] * A[k][j]; AtA[i][j] = a; }
  sm::jacobi_eig<4>(AtA, V, w);
  int m = 0;
  for (int i = 1; i < 4; ++i) if (w[i] < w[m]) m = i;
  for (int i = 0; i < 4; ++i) X[i] = V[i][m];
}

// decomposeEssentialMat: R candidates U W Vt / U W^T Vt and t = U[:,2]
DFVO_D void decompose_essential(const double* __restrict__ Eptr, double R[2][3][3], double tv[3]) {
  double E[3][3], U[3][3], s[3], Vt[3][3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) E[i][j] = Eptr[3 * i + j];
  sm::svd3_rank2(E, U, s, Vt);
  if (sm::det3(U) < 0) for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) U[i][j] = -U[i][j];
  if (sm::det3(Vt) < 0) for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Vt[i][j] = -Vt[i][j];
  const double W[3][3] = {{0, 1, 0}, {-1, 0, 0}, {0, 0, 1}};
  for (int which = 0; which < 2; ++which) {
    double T[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
      double a = 0;
      for (int k = 0; k < 3; ++k) a += U[i][k] * (which == 0 ? W[k][j] : W[j][k]);
      T[i][j] = a;
    }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
      double a = 0;
      for (int k = 0; k < 3; ++k) a += T[i][k] * Vt[k][j];
      R[which][i][j] = a;
    }
  }
  for (int i = 0; i < 3; ++i) tv[i] = U[i][2];
}

// Pass 1: one thread per (point, candidate) -- the 4 candidates of a point sit in adjacent lanes.  Every block
// repeats the (tiny) decomposition instead of waiting for a producer kernel.  mask_out[j] receives the 4-bit
// candidate field, info[1+k] the cheirality count of candidate k (atomics; info zeroed by the caller).
__global__ void __launch_bounds__(256)
k_recover_pose_vote(const double* __restrict__ Eptr, const double* __restrict__ p1, const double* __restrict__ p2, int N, double focal,
                    double cx, double cy, double dist, uint8_t* __restrict__ mask_out, int32_t* __restrict__ info) {
  __shared__ double sR[2][3][3], st[3];
  __shared__ int cnt[4];
  const int t = threadIdx.x;
  if (t == 0) decompose_essential(Eptr, sR, st);
  if (t < 4) cnt[t] = 0;
  __syncthreads();
  const int g = blockIdx.x * 256 + t, j = g >> 2, k = g & 3;
  bool m = false;
  if (j < N) {
    const double u0 = (p1[2 * j] - cx) / focal, v0 = (p1[2 * j + 1] - cy) / focal;
    const double u1 = (p2[2 * j] - cx) / focal, v1 = (p2[2 * j + 1] - cy) / focal;
    double P[3][4];
    const double sg = k < 2 ? 1.0 : -1.0;
    for (int a = 0; a < 3; ++a) { for (int b = 0; b < 3; ++b) P[a][b] = sR[k & 1][a][b]; P[a][3] = sg * st[a]; }
    double X[4];
    triangulate_dlt(P, u0, v0, u1, v1, X);
    m = (X[2] * X[3]) > 0;
    const double x = X[0] / X[3], y = X[1] / X[3], z = X[2] / X[3];
    m = m && (z < dist);
    const double z2 = P[2][0] * x + P[2][1] * y + P[2][2] * z + P[2][3];
    m = m && (z2 > 0) && (z2 < dist);
  }
  const unsigned ballot = __ballot_sync(0xffffffffu, m);
  const int lane = t & 31;
  if (j < N && k == 0) mask_out[j] = (uint8_t)((ballot >> lane) & 0xfu);
  if (lane < 4) {                                  // lane k sums candidate k over the warp's 8 points
    const int c = __popc(ballot & (0x11111111u << lane));
    if (c) atomicAdd(&cnt[lane], c);
  }
  __syncthreads();
  if (t < 4 && cnt[t]) atomicAdd(&info[1 + t], cnt[t]);
}

// Pass 2: first maximum wins (OpenCV's order); resolve the bit-field to the winner's mask; block 0 writes R|t.
__global__ void __launch_bounds__(256)
k_recover_pose_pick(const double* __restrict__ Eptr, int N, double* __restrict__ Rt_out, uint8_t* __restrict__ mask_out,
                    int32_t* __restrict__ info) {
  int b = 0;
  for (int k = 1; k < 4; ++k) if (info[1 + k] > info[1 + b]) b = k;
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j < N) mask_out[j] = (mask_out[j] >> b) & 1u;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double R[2][3][3], tv[3];
    decompose_essential(Eptr, R, tv);
    for (int a = 0; a < 3; ++a) for (int q = 0; q < 3; ++q) Rt_out[3 * a + q] = R[b & 1][a][q];
    for (int a = 0; a < 3; ++a) Rt_out[9 + a] = (b < 2 ? 1.0 : -1.0) * tv[a];
    info[0] = info[1 + b];
  }
}

__global__ void k_triangulate_depth(const double* __restrict__ x1, const double* __restrict__ x2, int N, const double* __restrict__ T21,
                                    double* __restrict__ depth2) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double P[3][4];
  for (int a = 0; a < 3; ++a) for (int b = 0; b < 4; ++b) P[a][b] = T21[4 * a + b];
  double X[4];
  triangulate_dlt(P, x1[2 * i], x1[2 * i + 1], x2[2 * i], x2[2 * i + 1], X);
  const double x = X[0] / X[3], y = X[1] / X[3], z = X[2] / X[3];           // X /= X[3]  (ops_3d.py:64)
  depth2[i] = P[2][0] * x + P[2][1] * y + P[2][2] * z + P[2][3];             // X2 = T_2w[:3] @ X
}

// ops_3d.triangulation (ops_3d.py:44-67) for two general views: cv2.triangulatePoints(T_1w[:3], T_2w[:3], kp1, kp2) (the same 4x4 DLT,
// smallest singular vector), X /= X[3], X1 = T_1w[:3] @ X, X2 = T_2w[:3] @ X.  Outputs are [3][N] (any may be null).
__global__ void k_triangulate_points(const double* __restrict__ x1, const double* __restrict__ x2, int N, const double* __restrict__ T1w,
                                     const double* __restrict__ T2w, double* __restrict__ Xw, double* __restrict__ X1, double* __restrict__ X2) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double P0[3][4], P1[3][4];
  for (int a = 0; a < 3; ++a) for (int b = 0; b < 4; ++b) { P0[a][b] = T1w[4 * a + b]; P1[a][b] = T2w[4 * a + b]; }
  const double u0 = x1[2 * i], v0 = x1[2 * i + 1], u1 = x2[2 * i], v1 = x2[2 * i + 1];
  double A[4][4];
  for (int k = 0; k < 4; ++k) {
    A[0][k] = u0 * P0[2][k] - P0[0][k]; A[1][k] = v0 * P0[2][k] - P0[1][k];
    A[2][k] = u1 * P1[2][k] - P1[0][k]; A[3][k] = v1 * P1[2][k] - P1[1][k];
  }
  double AtA[4][4], V[4][4], w[4];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) { double a = 0; for (int k = 0; k < 4; ++k) a += A[k][r] * A[k][c]; AtA[r][c] = a; }
  sm::jacobi_eig<4>(AtA, V, w);
  int m = 0;
  for (int r = 1; r < 4; ++r) if (w[r] < w[m]) m = r;
  const double X[4] = {V[0][m] / V[3][m], V[1][m] / V[3][m], V[2][m] / V[3][m], 1.0};
  for (int a = 0; a < 3; ++a) {
    if (Xw) Xw[(size_t)a * N + i] = X[a];
    if (X1) X1[(size_t)a * N + i] = P0[a][0] * X[0] + P0[a][1] * X[1] + P0[a][2] * X[2] + P0[a][3];
    if (X2) X2[(size_t)a * N + i] = P1[a][0] * X[0] + P1[a][1] * X[1] + P1[a][2] * X[2] + P1[a][3];
  }
}

int triangulate_points(const double* x1, const double* x2, int N, const double* T1w, const double* T2w, double* Xw, double* X1, double* X2,
                       cudaStream_t s) {
  DFVO_LAUNCH(k_triangulate_points, dim3(cdiv(N, 128)), dim3(128), 0, s, x1, x2, N, T1w, T2w, Xw, X1, X2);
  DFVO_CHECK_LAUNCH();
  return DFVO_OK;
}

int triangulate_depth(const double* x1, const double* x2, int N, const double* T21, double* depth2, cudaStream_t s) {
  DFVO_LAUNCH(k_triangulate_depth, dim3(cdiv(N, 128)), dim3(128), 0, s, x1, x2, N, T21, depth2);
  DFVO_CHECK_LAUNCH();
  return DFVO_OK;
}

int recover_pose(const double* E, const double* p1, const double* p2, int N, double focal, double cx, double cy, double* Rt_out,
                 uint8_t* mask_out, int32_t* info, cudaStream_t s) {
  DFVO_CUDA(cudaMemsetAsync(info, 0, 5 * sizeof(int32_t), s));
  DFVO_LAUNCH(k_recover_pose_vote, dim3(cdiv(4 * N, 256)), dim3(256), 0, s, E, p1, p2, N, focal, cx, cy, 50.0, mask_out, info);
  DFVO_LAUNCH(k_recover_pose_pick, dim3(cdiv(N, 256)), dim3(256), 0, s, E, N, Rt_out, mask_out, info);
  DFVO_CHECK_LAUNCH();
  return DFVO_OK;
}

}  // namespace dfvo
