// ============================================================================
// multih_oracle.cpp — CPU FP64 restatement of Multi-H's data-parallel hot path.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke check
// in __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs may
// load it.  The product (libmultih_b200.so) never links, calls or falls back
// to anything in this directory.
//
// It restates, OpenCV-free and in double precision, the reference functions
// listed in SURVEY.md §8(a); every function cites the reference file:line it
// follows (paths relative to /root/reference/MultiH/MultiH/).
//
// Parity status: PINNED ON THE REFERENCE SOURCE.  oracle/Makefile compiles the reference's own MultiH.cpp,
// MeanShiftClustering.h, Homography_RefineHAFCallback.h, Homography_Refine3PTCallback.h, the LMSolverImpl of Utilities.hpp and
// GCO, unmodified and where they lie under /root/reference, against oracle/cvshim/mini_cv.hpp (a stand-in for the OpenCV 3.1
// surface those files use) into oracle/_ref/libmultih_ref.so / libgco_ref.so (wrappers: ref_multih_wrapper.cpp,
// gco_ref_wrapper.cpp; the reference's MSVC project / PPL build is not run).  tests/test_oracle.py checks this restatement
// against that library function by function (GetHomographyHAF, GetHomography3PT + NormalizePoints, HomographyHAFNonminimal,
// dataEnergy, smoothnessEnergy, MeanShiftClustering<double>::Cluster, the LM solver + 3PT callback) and as a whole
// (MultiH::Process() on the bundled pair: same survivors, labels and homographies, with and without the LM polish).
// Second, independent pins kept from round 1: (1) tests/golden/*.npz, produced by tests/golden/make_golden.py — line-by-line
// transliterations of the same reference functions on the real OpenCV numerical routines (cv2.eigen, cv2.invert(DECOMP_SVD),
// cv2.solvePoly; cv2 4.13, the reference pins 3.1.0), including numpy / Python-list transliterations of
// MeanShiftClustering::Cluster (golden_meanshift.npz) and HomographyCompatibilityCheck (golden_compat.npz) with the MSVC rand(),
// (2) analytic known-answer tests (noise-free plane => generating H), (3) the integer cost constants 4901 / 9802 / 0..200 at
// default parameters.  What stays unpinned is only the output FILE of the shipped multih.exe (another build, OpenCV's RANSAC for
// F, FLANN's randomised trees, unseeded rand(): SURVEY.md §4) — parity is against the reference source run on identical inputs.
//
// Third-party arithmetic restated here (absent from /root/reference): OpenCV
// 3.1.0 cv::eigen on symmetric input (Jacobi; eigenvalues descending,
// eigenvectors in rows), cv::Mat::inv(DECOMP_SVD) (Jacobi SVD + backSubst with
// threshold 2*DBL_EPSILON*sum(w)), cv::Mat::inv() on 3x3.
// ============================================================================
#include <algorithm>
#include <array>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

namespace {

// ---------------------------------------------------------------------------
// cv::eigen restatement for symmetric n x n (n <= 8): cyclic Jacobi, then sort
// eigenvalues descending, eigenvectors returned in ROWS (MH.cpp:893-895 takes
// EVec.row(3) == smallest eigenvalue).
// ---------------------------------------------------------------------------
void sym_eigen(int n, const double* Ain, double* evals, double* evecs_rows) {
  double A[64], V[64];
  for (int i = 0; i < n * n; ++i) A[i] = Ain[i];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) V[i * n + j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < n; ++i) {
      diag += A[i * n + i] * A[i * n + i];
      for (int j = i + 1; j < n; ++j) off += A[i * n + j] * A[i * n + j];
    }
    if (off <= 1e-300 || off <= 1e-34 * diag) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        double apq = A[p * n + q];
        if (apq == 0.0) continue;
        double app = A[p * n + p], aqq = A[q * n + q];
        double theta = (aqq - app) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {  // A <- A J
          double akp = A[k * n + p], akq = A[k * n + q];
          A[k * n + p] = c * akp - s * akq;
          A[k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {  // A <- J^T A
          double apk = A[p * n + k], aqk = A[q * n + k];
          A[p * n + k] = c * apk - s * aqk;
          A[q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {  // V <- V J (columns are eigenvectors)
          double vkp = V[k * n + p], vkq = V[k * n + q];
          V[k * n + p] = c * vkp - s * vkq;
          V[k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  int order[8];
  for (int i = 0; i < n; ++i) order[i] = i;
  std::sort(order, order + n, [&](int a, int b) { return A[a * n + a] > A[b * n + b]; });
  for (int r = 0; r < n; ++r) {
    int c = order[r];
    evals[r] = A[c * n + c];
    for (int k = 0; k < n; ++k) evecs_rows[r * n + k] = V[k * n + c];
  }
}

void mat3_mul(const double* A, const double* B, double* C) {
  double T[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
  std::memcpy(C, T, sizeof(T));
}

// Epipole in image 2: last row of eigen(F F^T), divided by z (MH.cpp:789-793;
// same construction on the normalised F at MH.cpp:1013-1017).
void epipole2(const double* F, double* e /*3*/) {
  double FFt[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) FFt[i * 3 + j] = F[i * 3] * F[j * 3] + F[i * 3 + 1] * F[j * 3 + 1] + F[i * 3 + 2] * F[j * 3 + 2];
  double ev[3], evec[9];
  sym_eigen(3, FFt, ev, evec);
  e[0] = evec[6] / evec[8];
  e[1] = evec[7] / evec[8];
  e[2] = 1.0;
}

// The six HAF rows of one affine correspondence (MH.cpp:859-887 == :938-966).
inline void haf_rows(const double* p /*x1 y1 x2 y2*/, const double* a /*a11 a12 a21 a22*/, const double* F, double ex,
                     double ey, double R[6][4]) {
  const double x1 = p[0], y1 = p[1], x2 = p[2], y2 = p[3];
  const double a11 = a[0], a12 = a[1], a21 = a[2], a22 = a[3];
  R[0][0] = a11 * x1 + x2 - ex; R[0][1] = a11 * y1;           R[0][2] = a11; R[0][3] = -F[3];
  R[1][0] = a12 * x1;           R[1][1] = a12 * y1 + x2 - ex; R[1][2] = a12; R[1][3] = -F[4];
  R[2][0] = a21 * x1 + y2 - ey; R[2][1] = a21 * y1;           R[2][2] = a21; R[2][3] = F[0];
  R[3][0] = a22 * x1;           R[3][1] = a22 * y1 + y2 - ey; R[3][2] = a22; R[3][3] = F[1];
  R[4][0] = ex * x1 - x2 * x1;  R[4][1] = ex * y1 - x2 * y1;  R[4][2] = ex - x2;
  R[4][3] = x1 * F[3] + y1 * F[4] + F[5];
  R[5][0] = ey * x1 - y2 * x1;  R[5][1] = ey * y1 - y2 * y1;  R[5][2] = ey - y2;
  R[5][3] = -(x1 * F[0] + y1 * F[1] + F[2]);
}

// H from v = (h31,h32,h33,lambda) (MH.cpp:899-909 == :979-989).
inline void haf_assemble(const double* v, const double* F, double ex, double ey, double* H) {
  H[6] = v[0]; H[7] = v[1]; H[8] = v[2];
  const double lam = v[3];
  H[3] = ey * H[6] - lam * F[0];
  H[4] = ey * H[7] - lam * F[1];
  H[5] = ey * H[8] - lam * F[2];
  H[0] = ex * H[6] + lam * F[3];
  H[1] = ex * H[7] + lam * F[4];
  H[2] = ex * H[8] + lam * F[5];
}

void parallel_for(int64_t n, int threads, const std::function<void(int64_t, int64_t)>& body) {
  if (threads <= 1 || n < 2) { body(0, n); return; }
  std::vector<std::thread> pool;
  int64_t chunk = (n + threads - 1) / threads;
  for (int t = 0; t < threads; ++t) {
    int64_t b = t * chunk, e = std::min<int64_t>(n, b + chunk);
    if (b >= e) break;
    pool.emplace_back([=, &body] { body(b, e); });
  }
  for (auto& th : pool) th.join();
}

// C round(): half away from zero, then int conversion (MH.cpp:479,502,503,510).
inline int c_round(double v) { return (int)std::round(v); }

// Hartley normalisation of n 2-D points (3PTcb.h:161-195, CV_64F branch):
// centroid -> 0, mean distance -> sqrt(2);  T = [s 0 -mx*s; 0 s -my*s; 0 0 1].
void normalize_points(const double* pts, int n, double* out, double* T) {
  double mx = 0, my = 0;
  for (int i = 0; i < n; ++i) { mx += pts[2 * i]; my += pts[2 * i + 1]; }
  mx *= 1.0 / n; my *= 1.0 / n;
  double avg = 0;
  for (int i = 0; i < n; ++i) {
    out[2 * i] = pts[2 * i] - mx; out[2 * i + 1] = pts[2 * i + 1] - my;
    avg += std::sqrt(out[2 * i] * out[2 * i] + out[2 * i + 1] * out[2 * i + 1]);
  }
  avg /= n;
  const double ratio = std::sqrt(2.0) / avg;
  for (int i = 0; i < 2 * n; ++i) out[i] *= ratio;
  T[0] = ratio; T[1] = 0; T[2] = -mx * ratio;
  T[3] = 0; T[4] = ratio; T[5] = -my * ratio;
  T[6] = 0; T[7] = 0; T[8] = 1;
}

// inverse of the similarity T above (exact; cv::Mat::inv() on 3x3, MH.cpp:1009,1054)
void inv_similarity(const double* T, double* Ti) {
  const double s = T[0];
  Ti[0] = 1.0 / s; Ti[1] = 0; Ti[2] = -T[2] / s;
  Ti[3] = 0; Ti[4] = 1.0 / s; Ti[5] = -T[5] / s;
  Ti[6] = 0; Ti[7] = 0; Ti[8] = 1;
}

// x = pinv(A) b for A (m x 3), as cv::Mat::inv(DECOMP_SVD)*b (MH.cpp:1038):
// one-sided (Hestenes) Jacobi SVD, singular values <= 2*DBL_EPSILON*sum(w)
// dropped (OpenCV SVBkSb threshold).
void pinv3_solve(const std::vector<double>& Ain, const std::vector<double>& b, int m, double* x) {
  std::vector<double> U(Ain);  // m x 3, columns orthogonalised in place
  double V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool changed = false;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double a = 0, bb = 0, g = 0;
        for (int k = 0; k < m; ++k) {
          a += U[k * 3 + p] * U[k * 3 + p];
          bb += U[k * 3 + q] * U[k * 3 + q];
          g += U[k * 3 + p] * U[k * 3 + q];
        }
        if (std::fabs(g) <= DBL_EPSILON * std::sqrt(a * bb) || g == 0.0) continue;
        changed = true;
        double zeta = (bb - a) / (2.0 * g);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
        for (int k = 0; k < m; ++k) {
          double up = U[k * 3 + p], uq = U[k * 3 + q];
          U[k * 3 + p] = c * up - s * uq;
          U[k * 3 + q] = s * up + c * uq;
        }
        for (int k = 0; k < 3; ++k) {
          double vp = V[k * 3 + p], vq = V[k * 3 + q];
          V[k * 3 + p] = c * vp - s * vq;
          V[k * 3 + q] = s * vp + c * vq;
        }
      }
    if (!changed) break;
  }
  double w[3], sumw = 0;
  for (int j = 0; j < 3; ++j) {
    double s = 0;
    for (int k = 0; k < m; ++k) s += U[k * 3 + j] * U[k * 3 + j];
    w[j] = std::sqrt(s);
    sumw += w[j];
  }
  const double thr = 2.0 * DBL_EPSILON * sumw;
  x[0] = x[1] = x[2] = 0;
  for (int j = 0; j < 3; ++j) {
    if (w[j] <= thr) continue;
    double utb = 0;
    for (int k = 0; k < m; ++k) utb += U[k * 3 + j] * b[k];
    const double coef = utb / (w[j] * w[j]);  // (u_j/w_j)^T b / w_j, with u_j = U_j / w_j
    for (int k = 0; k < 3; ++k) x[k] += V[k * 3 + j] * coef;
  }
}

}  // namespace

extern "C" {

int orc_hardware_threads() { return (int)std::max(1u, std::thread::hardware_concurrency()); }

void orc_sym_eigen(int n, const double* A, double* evals, double* evecs_rows) { sym_eigen(n, A, evals, evecs_rows); }

void orc_epipole2(const double* F, double* e2 /*2*/) {
  double e[3];
  epipole2(F, e);
  e2[0] = e[0]; e2[1] = e[1];
}

// ---------------------------------------------------------------------------
// K1 oracle.  MultiH::ComputeLocalHomographies (MH.cpp:696-717) ->
// MultiH::GetHomographyHAF (MH.cpp:850-911).  pts: N x (x1 y1 x2 y2), aff:
// N x (a11 a12 a21 a22), F row-major (x2^T F x1 = 0), e2 = epipole in image 2.
// Out: N x 9, divided by h33 (MH.cpp:910).
// ---------------------------------------------------------------------------
void orc_haf_hypotheses(const double* pts, const double* aff, const double* F, const double* e2, int64_t N, double* H,
                        int threads) {
  parallel_for(N, threads, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i) {
      double R[6][4], M[16], ev[4], evec[16];
      haf_rows(pts + 4 * i, aff + 4 * i, F, e2[0], e2[1], R);
      for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
          double s = 0;
          for (int k = 0; k < 6; ++k) s += R[k][r] * R[k][c];
          M[r * 4 + c] = s;  // A^T A (MH.cpp:890)
        }
      sym_eigen(4, M, ev, evec);
      double* Hi = H + 9 * i;
      haf_assemble(evec + 12, F, e2[0], e2[1], Hi);  // EVec.row(3) (MH.cpp:895)
      const double h33 = Hi[8];
      for (int k = 0; k < 9; ++k) Hi[k] /= h33;  // MH.cpp:910
    }
  });
}

// ---------------------------------------------------------------------------
// Feature vectors.  10-D: MultiH::EstablishStablePointSets (MH.cpp:612-646),
// order x1 x2 x3 y1 y2 y3 + locality*(src.x src.y dst.x dst.y).
// 6-D: MultiH::MergingStep (MH.cpp:359-389), order x1 y1 x2 y2 x3 y3.
// ---------------------------------------------------------------------------
void orc_features10(const double* H, const double* pts, double locality, int64_t N, double* out) {
  for (int64_t i = 0; i < N; ++i) {
    const double* h = H + 9 * i;
    const double s1 = h[8], x1 = h[2] / s1, y1 = h[5] / s1;
    const double s2 = h[6] + h[8], x2 = (h[0] + h[2]) / s2, y2 = (h[3] + h[5]) / s2;
    const double s3 = h[7] + h[8], x3 = (h[1] + h[2]) / s3, y3 = (h[4] + h[5]) / s3;
    double* f = out + 10 * i;
    f[0] = x1; f[1] = x2; f[2] = x3; f[3] = y1; f[4] = y2; f[5] = y3;
    f[6] = pts[4 * i] * locality; f[7] = pts[4 * i + 1] * locality;
    f[8] = pts[4 * i + 2] * locality; f[9] = pts[4 * i + 3] * locality;
  }
}

void orc_features6(const double* H, int64_t K, double* out) {
  for (int64_t i = 0; i < K; ++i) {
    const double* h = H + 9 * i;
    const double s1 = h[8], x1 = h[2] / s1, y1 = h[5] / s1;
    const double s2 = h[6] + h[8], x2 = (h[0] + h[2]) / s2, y2 = (h[3] + h[5]) / s2;
    const double s3 = h[7] + h[8], x3 = (h[1] + h[2]) / s3, y3 = (h[4] + h[5]) / s3;
    double* f = out + 6 * i;
    f[0] = x1; f[1] = y1; f[2] = x2; f[3] = y2; f[4] = x3; f[5] = y3;
  }
}

// ---------------------------------------------------------------------------
// K3 oracle.  MeanShiftClustering<double>::Cluster (MS.h:22-157), verbatim
// semantics: window test = SUM_j |mean_j - x_ij| (L1) < bw^2 (MS.h:76-85);
// stop when ||mean - old||_2 < 1e-3 bw (MS.h:48,98); merge into the FIRST centre
// with L2 distance < bw/2 by averaging the two centres and adding votes
// (MS.h:100-120); final assignment = most votes, first wins ties (MS.h:133-146).
// Seeds: tempInd = round(rnd * (remaining-1)), rnd = rand()/RAND_MAX (MS.h:54-56).
// The injected generator restates MSVC's rand() (the reference is an MSVC
// program): holdrand = holdrand*214013 + 2531011, (holdrand>>16)&0x7fff,
// RAND_MAX 32767; *rng_state carries holdrand across calls (initially 1).
// metric: 0 = L1_REF (reference), 1 = L2 (||.||_2^2 < bw^2; the tensor-core variant).
// Returns the number of centres C; centres: up to max_c x D; assign: N (−1 never
// occurs unless no votes); votes_out optional.  Also reports trajectory and
// window-iteration counts.
// ---------------------------------------------------------------------------
constexpr int MS_MAX_WINDOW_ITERS = 200;   // = MH_MS_MAX_WINDOW_ITERS of include/multih_b200.h
int orc_meanshift(const double* data, int N, int D, double bw, int metric, uint32_t* rng_state, double* centres,
                  int max_c, int* assign, int64_t* stats /*[2]: trajectories, iterations*/) {
  const double bandSq = bw * bw, stopThresh = 1e-3 * bw;
  std::vector<int> initPtInds(N);
  std::vector<char> visited(N, 0);
  for (int i = 0; i < N; ++i) initPtInds[i] = i;
  std::vector<std::vector<double>> clustCent;
  std::vector<std::vector<int>> clusterVotes;
  int64_t traj = 0, iters = 0;
  uint32_t hold = rng_state ? *rng_state : 1u;
  std::vector<double> myMean(D), oldMean(D), acc(D);
  while (!initPtInds.empty()) {
    hold = hold * 214013u + 2531011u;
    const int r = (int)((hold >> 16) & 0x7fff);
    const double rnd = r / 32767.0;
    const int tempInd = (int)std::round(rnd * (double)(initPtInds.size() - 1));
    const int stInd = initPtInds[tempInd];
    for (int j = 0; j < D; ++j) myMean[j] = data[(size_t)stInd * D + j];
    std::vector<int> votes(N, 0);
    ++traj;
    int window_iters = 0;
    while (true) {
      ++iters;
      ++window_iters;
      oldMean = myMean;
      std::fill(acc.begin(), acc.end(), 0.0);
      int cnt = 0;
      for (int i = 0; i < N; ++i) {
        const double* x = data + (size_t)i * D;
        double s = 0;
        if (metric == 0)
          for (int j = 0; j < D; ++j) { double d = oldMean[j] - x[j]; s += std::sqrt(d * d); }
        else
          for (int j = 0; j < D; ++j) { double d = oldMean[j] - x[j]; s += d * d; }
        if (s < bandSq) {
          ++votes[i]; ++cnt; visited[i] = 1;
          for (int j = 0; j < D; ++j) acc[j] += x[j];
        }
      }
      for (int j = 0; j < D; ++j) myMean[j] = acc[j] / cnt;  // cnt==0 -> NaN as in the reference (cannot occur: seed is a member)
      double n2 = 0;
      for (int j = 0; j < D; ++j) n2 += (myMean[j] - oldMean[j]) * (myMean[j] - oldMean[j]);
      // MS.h:98 loops `while (1)` until the mean stops moving.  With an L1 window and a flat kernel the mean can enter a
      // cycle (e.g. synthetic scene seed 0xB200 + 33, 5000 correspondences) and the reference never returns; oracle and
      // product stop a trajectory after MS_MAX_WINDOW_ITERS window iterations and keep its current mean — a deviation
      // only on inputs on which the reference hangs.
      if (std::sqrt(n2) < stopThresh || window_iters >= MS_MAX_WINDOW_ITERS) {
        int mergeWith = -1;
        for (size_t cn = 0; cn < clustCent.size(); ++cn) {
          double d2 = 0;
          for (int j = 0; j < D; ++j) d2 += (myMean[j] - clustCent[cn][j]) * (myMean[j] - clustCent[cn][j]);
          if (std::sqrt(d2) < bw / 2) { mergeWith = (int)cn; break; }
        }
        if (mergeWith > -1) {
          for (int j = 0; j < D; ++j) clustCent[mergeWith][j] = 0.5 * (clustCent[mergeWith][j] + myMean[j]);
          for (int i = 0; i < N; ++i) clusterVotes[mergeWith][i] += votes[i];
        } else {
          clustCent.push_back(myMean);
          clusterVotes.push_back(votes);
        }
        break;
      }
    }
    initPtInds.clear();
    for (int i = 0; i < N; ++i)
      if (!visited[i]) initPtInds.push_back(i);
  }
  std::vector<int> best(N, 0);
  for (int i = 0; i < N; ++i) assign[i] = -1;
  for (size_t r = 0; r < clusterVotes.size(); ++r)
    for (int i = 0; i < N; ++i)
      if (best[i] < clusterVotes[r][i]) { best[i] = clusterVotes[r][i]; assign[i] = (int)r; }
  const int C = (int)clustCent.size();
  for (int c = 0; c < C && c < max_c; ++c)
    for (int j = 0; j < D; ++j) centres[(size_t)c * D + j] = clustCent[c][j];
  if (rng_state) *rng_state = hold;
  if (stats) { stats[0] = traj; stats[1] = iters; }
  return C;
}

void orc_normalize_points(const double* pts, int n, double* out, double* T) { normalize_points(pts, n, out, T); }

// ---------------------------------------------------------------------------
// The reference's Levenberg-Marquardt polish of the 3PT fit: RefineHomography3PT + Homography_Refine3PTCallback
// (3PTcb.h:7-58, 81-143) driven by its copy of cv::LMSolverImpl::run (Utilities.hpp:762-869, 1000 iterations, eps = FLT_EPSILON),
// restated step for step.  Parameters: the third row (h31, h32, h33) of H in NORMALISED coordinates; residuals: the reprojection
// errors (x2 - x', y2 - y'); the callback's Jacobian is the reference's own approximation (e_x s x1, e_x s y1, e_x s; e_y ...) — it
// ignores the derivative of the projective division and has the opposite sign of d(err)/dh, so most steps are rejected; whatever
// the iteration accepts is what the reference keeps.  (The HAF polish, RefineHomographyHAF, never writes its result back:
// HAFcb.h:58 rebinds a local header — its effect is nil, so there is nothing to restate.)
// ---------------------------------------------------------------------------
namespace {
struct Lm3ptProblem { const double *p1, *p2; int n; const double* F; double ex, ey; };

// Homography_Refine3PTCallback::compute (3PTcb.h:81-143): err [2n], J [2n][3] (row-major) when wanted
void lm3pt_compute(const Lm3ptProblem& P, const double* h, double* err, double* J) {
  for (int i = 0; i < P.n; ++i) {
    const double x1 = P.p1[2 * i], y1 = P.p1[2 * i + 1], x2 = P.p2[2 * i], y2 = P.p2[2 * i + 1];
    double s = h[0] * x1 + h[1] * y1 + h[2];
    s = std::fabs(s) > DBL_EPSILON ? 1. / s : 0;
    const double h21 = P.ey * h[0] - P.F[0], h22 = P.ey * h[1] - P.F[1], h23 = P.ey * h[2] - P.F[2];
    const double h11 = P.ex * h[0] + P.F[3], h12 = P.ex * h[1] + P.F[4], h13 = P.ex * h[2] + P.F[5];
    const double xi = (h11 * x1 + h12 * y1 + h13) * s, yi = (h21 * x1 + h22 * y1 + h23) * s;
    err[2 * i] = x2 - xi;
    err[2 * i + 1] = y2 - yi;
    if (J) {
      double* j = J + 6 * (size_t)i;
      j[0] = P.ex * s * x1; j[1] = P.ex * s * y1; j[2] = P.ex * s;
      j[3] = P.ey * s * x1; j[4] = P.ey * s * y1; j[5] = P.ey * s;
    }
  }
}
// x = A^+ b for a symmetric 3x3 A through its eigen-decomposition, as cv::solve / cv::invert do with DECOMP_EIG (back substitution
// drops eigenvalues <= 2 DBL_EPSILON sum |w|); Ainv optionally returns the pseudo-inverse
void sym3_eig_solve(const double* A, const double* b, double* x, double* Ainv) {
  double w[3], V[9];
  sym_eigen(3, A, w, V);   // eigenvectors in rows
  double sum = 0;
  for (int k = 0; k < 3; ++k) sum += std::fabs(w[k]);
  const double thr = 2 * DBL_EPSILON * sum;
  double inv[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int k = 0; k < 3; ++k) {
    if (!(std::fabs(w[k]) > thr)) continue;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) inv[i * 3 + j] += V[k * 3 + i] * V[k * 3 + j] / w[k];
  }
  if (x)
    for (int i = 0; i < 3; ++i) x[i] = inv[i * 3] * b[0] + inv[i * 3 + 1] * b[1] + inv[i * 3 + 2] * b[2];
  if (Ainv) std::memcpy(Ainv, inv, sizeof(inv));
}
// cv::LMSolverImpl::run (Utilities.hpp:762-869) on the three parameters; returns the iteration count
int lm3pt_run(const Lm3ptProblem& P, double* x, int maxIters = 1000) {
  const int m = 2 * P.n;
  std::vector<double> r(m), rd(m), J(3 * (size_t)m);
  auto normal = [&](double* A, double* v) {   // A = J^T J, v = J^T r
    for (int k = 0; k < 9; ++k) A[k] = 0;
    v[0] = v[1] = v[2] = 0;
    for (int i = 0; i < m; ++i) {
      const double* j = J.data() + 3 * (size_t)i;
      for (int a = 0; a < 3; ++a) {
        v[a] += j[a] * r[i];
        for (int b = 0; b < 3; ++b) A[a * 3 + b] += j[a] * j[b];
      }
    }
  };
  auto sumsq = [&](const std::vector<double>& e) { double s = 0; for (double t : e) s += t * t; return s; };
  auto maxabs = [](const double* e, int n) { double s = 0; for (int i = 0; i < n; ++i) s = std::max(s, std::fabs(e[i])); return s; };
  lm3pt_compute(P, x, r.data(), J.data());
  double S = sumsq(r), A[9], v[3], D[3], Ap[9], d[3], xd[3];
  normal(A, v);
  for (int i = 0; i < 3; ++i) D[i] = A[i * 3 + i];
  const double Rlo = 0.25, Rhi = 0.75, eps = FLT_EPSILON;
  double lambda = 1, lc = 0.75;
  int iter = 0;
  for (;;) {
    std::memcpy(Ap, A, sizeof(Ap));
    for (int i = 0; i < 3; ++i) Ap[i * 3 + i] += lambda * D[i];
    sym3_eig_solve(Ap, v, d, nullptr);
    for (int i = 0; i < 3; ++i) xd[i] = x[i] - d[i];
    lm3pt_compute(P, xd, rd.data(), nullptr);
    const double Sd = sumsq(rd);
    double dS = 0;   // d . (2 v - A d)
    for (int i = 0; i < 3; ++i) dS += d[i] * (2 * v[i] - (A[i * 3] * d[0] + A[i * 3 + 1] * d[1] + A[i * 3 + 2] * d[2]));
    const double R = (S - Sd) / (std::fabs(dS) > DBL_EPSILON ? dS : 1);
    if (R > Rhi) {
      lambda *= 0.5;
      if (lambda < lc) lambda = 0;
    } else if (R < Rlo) {
      const double t = d[0] * v[0] + d[1] * v[1] + d[2] * v[2];
      double nu = (Sd - S) / (std::fabs(t) > DBL_EPSILON ? t : 1) + 2;
      nu = std::min(std::max(nu, 2.), 10.);
      if (lambda == 0) {
        sym3_eig_solve(A, nullptr, nullptr, Ap);
        double maxval = DBL_EPSILON;
        for (int i = 0; i < 3; ++i) maxval = std::max(maxval, std::fabs(Ap[i * 3 + i]));
        lambda = lc = 1. / maxval;
        nu *= 0.5;
      }
      lambda *= nu;
    }
    if (Sd < S) {
      S = Sd;
      std::memcpy(x, xd, sizeof(xd));
      lm3pt_compute(P, x, r.data(), J.data());
      normal(A, v);
    }
    ++iter;
    if (!(iter < maxIters && maxabs(d, 3) >= eps && maxabs(r.data(), m) >= eps)) break;
  }
  return iter;
}
}  // namespace

// ---------------------------------------------------------------------------
// K4 (3PT) oracle.  MultiH::GetHomography3PT (MH.cpp:995-1055); refine = its do_numerical_refinement (the LM polish
// above).  pts1/pts2: n x 2.  Out H (3x3, NOT divided by h33).
// ---------------------------------------------------------------------------
void orc_homography_3pt_ex(const double* pts1, const double* pts2, int n, const double* F, int refine, double* H);
void orc_homography_3pt(const double* pts1, const double* pts2, int n, const double* F, double* H) {
  orc_homography_3pt_ex(pts1, pts2, n, F, 0, H);
}
void orc_homography_3pt_ex(const double* pts1, const double* pts2, int n, const double* F, int refine, double* H) {
  std::vector<double> n1(2 * n), n2(2 * n);
  double T1[9], T2[9], T1i[9], T2i[9];
  normalize_points(pts1, n, n1.data(), T1);
  normalize_points(pts2, n, n2.data(), T2);
  inv_similarity(T1, T1i);
  inv_similarity(T2, T2i);
  double T2it[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T2it[i * 3 + j] = T2i[j * 3 + i];
  double Fn[9];
  mat3_mul(T2it, F, Fn);
  mat3_mul(Fn, T1i, Fn);  // MH.cpp:1009
  double e[3];
  epipole2(Fn, e);  // MH.cpp:1013-1017
  std::vector<double> A(6 * n), b(2 * n);
  for (int i = 0; i < n; ++i) {
    const double x1 = n1[2 * i], y1 = n1[2 * i + 1], x2 = n2[2 * i], y2 = n2[2 * i + 1];
    A[6 * i + 0] = e[0] * x1 - x2 * x1; A[6 * i + 1] = e[0] * y1 - x2 * y1; A[6 * i + 2] = e[0] - x2;
    A[6 * i + 3] = e[1] * x1 - y2 * x1; A[6 * i + 4] = e[1] * y1 - y2 * y1; A[6 * i + 5] = e[1] - y2;
    b[2 * i] = -(x1 * Fn[3] + y1 * Fn[4] + Fn[5]);
    b[2 * i + 1] = (x1 * Fn[0] + y1 * Fn[1] + Fn[2]);
  }
  double h3[3];
  pinv3_solve(A, b, 2 * n, h3);  // MH.cpp:1038
  if (refine) {  // MH.cpp:1052-1053
    const Lm3ptProblem P{n1.data(), n2.data(), n, Fn, e[0], e[1]};
    lm3pt_run(P, h3);
  }
  double Hn[9];
  const double v[4] = {h3[0], h3[1], h3[2], 1.0};  // lambda == 1 (MH.cpp:1045-1050)
  haf_assemble(v, Fn, e[0], e[1], Hn);
  mat3_mul(T2i, Hn, Hn);
  mat3_mul(Hn, T1, H);  // MH.cpp:1054
}

// Batched form used by EstablishStablePointSets (MH.cpp:664-688): one 3PT fit
// per cluster over the members listed in CSR (offsets[C+1], members[]).
// Clusters with < 3 members get keep[c]=0 (MH.cpp:667).
void orc_cluster_3pt_ex(const double* pts, const int* offsets, const int* members, int C, const double* F, int refine, double* H,
                        int* keep);
void orc_cluster_3pt(const double* pts, const int* offsets, const int* members, int C, const double* F, double* H,
                     int* keep) {
  orc_cluster_3pt_ex(pts, offsets, members, C, F, 0, H, keep);
}
void orc_cluster_3pt_ex(const double* pts, const int* offsets, const int* members, int C, const double* F, int refine, double* H,
                        int* keep) {
  for (int c = 0; c < C; ++c) {
    const int n = offsets[c + 1] - offsets[c];
    keep[c] = n >= 3;
    if (n < 3) continue;
    std::vector<double> p1(2 * n), p2(2 * n);
    for (int j = 0; j < n; ++j) {
      const double* p = pts + 4 * (size_t)members[offsets[c] + j];
      p1[2 * j] = p[0]; p1[2 * j + 1] = p[1]; p2[2 * j] = p[2]; p2[2 * j + 1] = p[3];
    }
    orc_homography_3pt_ex(p1.data(), p2.data(), n, F, refine, H + 9 * c);
  }
}

// MergingStep's mode -> homography (MH.cpp:408-427): 3PT on (0,0),(1,0),(0,1)
// and the mode's 6-D feature (x1 y1 x2 y2 x3 y3).
void orc_mode_to_homography_ex(const double* mode6, const double* F, int refine, double* H) {
  const double p1[6] = {0, 0, 1, 0, 0, 1};
  orc_homography_3pt_ex(p1, mode6, 3, F, refine, H);
}
void orc_mode_to_homography(const double* mode6, const double* F, double* H) { orc_mode_to_homography_ex(mode6, F, 0, H); }

// ---------------------------------------------------------------------------
// Post-processing oracle.  MultiH::HomographyCompatibilityCheck (MH.cpp:100-222): a cross-validation filter on the
// clusters the alternating optimisation returned.  Per cluster with N >= max(min_inliers, 4) members, 501 trials
// (MH.cpp:130: MAX(501, MIN(501, .)) == 501): draw 3 members WITHOUT replacement from the cluster's point vector
// (idx = (size-1) * rand()/RAND_MAX truncated, then erase: MH.cpp:142-154), fit GetHomography3PT without refinement
// (:157), take the "median" squared transfer error of the remaining N-3 members (:162-181), put the three back at the
// END of the vector in reverse draw order (:183-194).  The cluster is removed when the median over the trials exceeds
// thr^2 * 81/16 (:200-203), or when it has fewer than min_inliers members (:207-208).  Removed clusters' points become
// outliers and the higher labels move down (:216-230).
// Reproduced as written, including
//  * the distance buffer of size N of which only the first N-3 entries are rewritten per trial while ALL N are sorted
//    (:140, :177): three stale values — the three largest of the previous trial's sorted buffer, zeros in the first
//    trial — take part in every "median";
//  * the even-length median 0.5 (d[n/2] + d[n/2+1]) (:178, :200; off by one);
//  * rand() = the injected MSVC generator, clusters visited in index order (the reference's parallel_for makes the
//    draw order nondeterministic; this is its serial order, USE_CONCURRENCY 0).
// labels: in/out (N, -1 = outlier); H: in/out (K x 9, compacted); returns the new K; medians[K_in] (NaN = not tested),
// removed[K_in].
// ---------------------------------------------------------------------------
int orc_compatibility_check(const double* pts, int64_t N, int32_t* labels, double* H, int K, const double* F, double thr,
                            int min_inliers, uint32_t* rng_state, double* medians, int32_t* removed) {
  uint32_t hold = rng_state ? *rng_state : 1u;
  auto msvc_rand = [&]() { hold = hold * 214013u + 2531011u; return (int)((hold >> 16) & 0x7fff); };
  const double limit = thr * thr * 81.0 / 16.0;
  std::vector<std::vector<std::array<double, 4>>> per(K);
  for (int64_t i = 0; i < N; ++i)
    if (labels[i] > -1 && labels[i] < K) per[labels[i]].push_back({pts[4 * i], pts[4 * i + 1], pts[4 * i + 2], pts[4 * i + 3]});
  std::vector<char> rem(K, 0);
  const int trials = 501;
  for (int c = 0; c < K; ++c) {
    std::vector<std::array<double, 4>>& v = per[c];
    const int n = (int)v.size();
    if (medians) medians[c] = std::nan("");
    if (n >= std::max(min_inliers, 4)) {
      std::vector<double> dist(n, 0.0), distances(trials);
      for (int t = 0; t < trials; ++t) {
        double p1[6], p2[6];
        std::array<double, 4> mss[3];
        for (int j = 0; j < 3; ++j) {
          const int idx = (int)((double)(v.size() - 1) * ((double)msvc_rand() / 32767.0));
          mss[j] = v[idx];
          p1[2 * j] = v[idx][0]; p1[2 * j + 1] = v[idx][1]; p2[2 * j] = v[idx][2]; p2[2 * j + 1] = v[idx][3];
          v.erase(v.begin() + idx);
        }
        double h[9];
        orc_homography_3pt(p1, p2, 3, F, h);
        for (size_t j = 0; j < v.size(); ++j) {
          const double s = h[6] * v[j][0] + h[7] * v[j][1] + h[8];
          const double x1 = (h[0] * v[j][0] + h[1] * v[j][1] + h[2]) / s, y1 = (h[3] * v[j][0] + h[4] * v[j][1] + h[5]) / s;
          const double dx = v[j][2] - x1, dy = v[j][3] - y1;
          dist[j] = dx * dx + dy * dy;
        }
        std::sort(dist.begin(), dist.end());   // all n entries: the last three are stale
        const size_t m = v.size();
        distances[t] = m % 2 ? dist[m / 2] : 0.5 * (dist[m / 2] + dist[m / 2 + 1]);
        v.resize(n);
        for (int j = 0; j < 3; ++j) v[n - j - 1] = mss[j];
      }
      std::sort(distances.begin(), distances.end());
      const double median = trials % 2 ? distances[trials / 2] : 0.5 * (distances[trials / 2] + distances[trials / 2 + 1]);
      if (medians) medians[c] = median;
      rem[c] = median > limit;
    } else if (n < min_inliers) {
      rem[c] = 1;
    }
    if (removed) removed[c] = rem[c];
  }
  int Kn = K;
  for (int c = K - 1; c >= 0; --c)
    if (rem[c]) {
      for (int64_t j = 0; j < N; ++j) {
        if (labels[j] == c) labels[j] = -1;
        else if (labels[j] > c) --labels[j];
      }
      for (int k = c; k + 1 < Kn; ++k) std::memcpy(H + 9 * (size_t)k, H + 9 * (size_t)(k + 1), sizeof(double) * 9);
      --Kn;
    }
  if (rng_state) *rng_state = hold;
  return Kn;
}

// ---------------------------------------------------------------------------
// K2 oracle.  dataEnergy (MH.cpp:473-504) + EnergyDataStruct (MH.h:23-47),
// evaluated densely: out[p*(K+1)+l], l=0 is the outlier label.
// lambda = spatial weight (0.5), thr = homography threshold (2.2 px).
// ---------------------------------------------------------------------------
int orc_data_cost(const double* p /*x1 y1 x2 y2*/, const double* h /*9 or null for l==0*/, double lambda, double thr) {
  const double lam = 100.0 / lambda;              // one_per_energy_lambda (MH.h:42)
  const double T = thr * thr * 81.0 / 16.0;       // truncated_sqr_threshold (MH.h:44)
  if (!h) return c_round(lam * T);                // MH.cpp:478-479
  const double s1 = h[6] * p[0] + h[7] * p[1] + h[8];
  const double x1 = (h[0] * p[0] + h[1] * p[1] + h[2]) / s1;
  const double y1 = (h[3] * p[0] + h[4] * p[1] + h[5]) / s1;
  const double dx = x1 - p[2], dy = y1 - p[3];
  const double distance = dx * dx + dy * dy;
  if (distance < T) return c_round(lam * (1.0f - (distance / T)));  // MH.cpp:501-502
  return 2 * c_round(lam * T);                                       // MH.cpp:503
}

void orc_residuals(const double* pts, int64_t N, const double* H, int K, double* d2 /*N x K*/, int threads) {
  parallel_for(N, threads, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i) {
      const double* p = pts + 4 * i;
      for (int l = 0; l < K; ++l) {
        const double* h = H + 9 * (size_t)l;
        const double s1 = h[6] * p[0] + h[7] * p[1] + h[8];
        const double x1 = (h[0] * p[0] + h[1] * p[1] + h[2]) / s1;
        const double y1 = (h[3] * p[0] + h[4] * p[1] + h[5]) / s1;
        const double dx = x1 - p[2], dy = y1 - p[3];
        d2[i * K + l] = dx * dx + dy * dy;
      }
    }
  });
}

void orc_data_cost_dense(const double* pts, int64_t N, const double* H, int K, double lambda, double thr, int32_t* out,
                         int threads) {
  parallel_for(N, threads, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i) {
      int32_t* o = out + i * (K + 1);
      o[0] = orc_data_cost(pts + 4 * i, nullptr, lambda, thr);
      for (int l = 0; l < K; ++l) o[l + 1] = orc_data_cost(pts + 4 * i, H + 9 * (size_t)l, lambda, thr);
    }
  });
}

// Throughput form for the CPU baseline: evaluates dataEnergy over N x K without
// materialising the matrix; returns a checksum (sum of costs) so the work cannot
// be elided, plus per-point argmin label (0 = outlier) and per-hypothesis inlier
// counts at d2 < thr^2 (the quantities the fused GPU kernel emits).
int64_t orc_data_cost_sweep(const double* pts, int64_t N, const double* H, int K, double lambda, double thr,
                            int32_t* argmin_label /*N or null*/, int64_t* inlier_count /*K or null*/, int threads) {
  std::atomic<int64_t> total{0};
  std::vector<std::vector<int64_t>> counts(threads > 0 ? threads : 1, std::vector<int64_t>(inlier_count ? K : 0, 0));
  const double thr2 = thr * thr;
  int64_t chunk = (N + std::max(threads, 1) - 1) / std::max(threads, 1);
  parallel_for(N, threads, [&](int64_t b, int64_t e) {
    const int tid = chunk > 0 ? (int)(b / chunk) : 0;
    int64_t local = 0;
    for (int64_t i = b; i < e; ++i) {
      const double* p = pts + 4 * i;
      int best = orc_data_cost(p, nullptr, lambda, thr), bestl = 0;
      local += best;
      for (int l = 0; l < K; ++l) {
        const double* h = H + 9 * (size_t)l;
        const int c = orc_data_cost(p, h, lambda, thr);
        local += c;
        if (c < best) { best = c; bestl = l + 1; }
        if (inlier_count) {
          const double s1 = h[6] * p[0] + h[7] * p[1] + h[8];
          const double x1 = (h[0] * p[0] + h[1] * p[1] + h[2]) / s1;
          const double y1 = (h[3] * p[0] + h[4] * p[1] + h[5]) / s1;
          const double dx = x1 - p[2], dy = y1 - p[3];
          if (dx * dx + dy * dy < thr2) ++counts[tid][l];
        }
      }
      if (argmin_label) argmin_label[i] = bestl;
    }
    total += local;
  });
  if (inlier_count)
    for (int l = 0; l < K; ++l) {
      int64_t s = 0;
      for (auto& c : counts) s += c[l];
      inlier_count[l] = s;
    }
  return total.load();
}

int orc_smooth_cost(int l1, int l2, double lambda) { return l1 != l2 ? c_round(100.0 * lambda) : 0; }  // MH.cpp:506-511

// ---------------------------------------------------------------------------
// MergingStep inlier scan + straightness test (MH.cpp:430-463) for K' candidate
// homographies: count[k] = #{d2 < thr^2}; scatter[k] = 6 uniques of
// S = SUM_inl [x y 1]^T [x y 1] (xx xy x yy y n); lambda_min[k] = smallest
// eigenvalue of S; keep[k] = !(lambda_min < straightness || count < 3).
// ---------------------------------------------------------------------------
void orc_inlier_stats(const double* pts, int64_t N, const double* H, int K, double thr, double straightness,
                      int64_t* count, double* scatter6, double* lambda_min, int* keep) {
  const double thr2 = thr * thr;
  for (int k = 0; k < K; ++k) {
    const double* h = H + 9 * (size_t)k;
    double sxx = 0, sxy = 0, sx = 0, syy = 0, sy = 0;
    int64_t n = 0;
    for (int64_t j = 0; j < N; ++j) {
      const double* p = pts + 4 * j;
      const double s = h[6] * p[0] + h[7] * p[1] + h[8];
      const double x1 = (h[0] * p[0] + h[1] * p[1] + h[2]) / s;
      const double y1 = (h[3] * p[0] + h[4] * p[1] + h[5]) / s;
      const double dx = p[2] - x1, dy = p[3] - y1;
      if (dx * dx + dy * dy < thr2) {
        ++n; sxx += p[0] * p[0]; sxy += p[0] * p[1]; sx += p[0]; syy += p[1] * p[1]; sy += p[1];
      }
    }
    const double S[9] = {sxx, sxy, sx, sxy, syy, sy, sx, sy, (double)n};
    double ev[3], evec[9];
    sym_eigen(3, S, ev, evec);
    if (count) count[k] = n;
    if (scatter6) { double* o = scatter6 + 6 * k; o[0] = sxx; o[1] = sxy; o[2] = sx; o[3] = syy; o[4] = sy; o[5] = (double)n; }
    if (lambda_min) lambda_min[k] = ev[2];
    if (keep) keep[k] = !(ev[2] < straightness || n < 3);
  }
}

// ComputeInliersOfHomography (MH.cpp:743-768): labels[i] = idx where d2 < thr^2.
void orc_inliers_of_homography(const double* pts, int64_t N, const double* h, double thr, int idx, int32_t* labels) {
  const double thr2 = thr * thr;
  for (int64_t i = 0; i < N; ++i) {
    const double* p = pts + 4 * i;
    const double s = h[6] * p[0] + h[7] * p[1] + h[8];
    const double x1 = (h[0] * p[0] + h[1] * p[1] + h[2]) / s;
    const double y1 = (h[3] * p[0] + h[4] * p[1] + h[5]) / s;
    const double dx = x1 - p[2], dy = y1 - p[3];
    if (dx * dx + dy * dy < thr2) labels[i] = idx;
  }
}

// ---------------------------------------------------------------------------
// K4 (HAF) oracle.  LabelingStep's per-label gather (MH.cpp:545-584) +
// MultiH::GetHomographyHAFNonminimal (MH.cpp:913-990) with
// do_numerical_refinement=false.  labels: N ints in −1..K−1 (−1 = outlier).
// Out: H[K][9] (NOT divided by h33), M10[K][10] = upper triangle of
// SUM_i A_i^T A_i (row-major uniques 00 01 02 03 11 12 13 22 23 33), count[K].
// Labels with no member keep H untouched (MH.cpp:592-593).
// ---------------------------------------------------------------------------
void orc_refit_haf(const double* pts, const double* aff, const int32_t* labels, int64_t N, int K, const double* F,
                   const double* e2, double* H, double* M10, int64_t* count) {
  std::vector<double> M((size_t)K * 16, 0.0);
  std::vector<int64_t> cnt(K, 0);
  for (int64_t i = 0; i < N; ++i) {
    const int l = labels[i];
    if (l < 0 || l >= K) continue;
    double R[6][4];
    haf_rows(pts + 4 * i, aff + 4 * i, F, e2[0], e2[1], R);
    double* m = M.data() + 16 * (size_t)l;
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) {
        double s = 0;
        for (int k = 0; k < 6; ++k) s += R[k][r] * R[k][c];
        m[r * 4 + c] += s;
      }
    ++cnt[l];
  }
  for (int l = 0; l < K; ++l) {
    const double* m = M.data() + 16 * (size_t)l;
    if (M10) {
      double* o = M10 + 10 * (size_t)l;
      int q = 0;
      for (int r = 0; r < 4; ++r)
        for (int c = r; c < 4; ++c) o[q++] = m[r * 4 + c];
    }
    if (count) count[l] = cnt[l];
    if (cnt[l] == 0) continue;
    double ev[4], evec[16];
    sym_eigen(4, m, ev, evec);
    haf_assemble(evec + 12, F, e2[0], e2[1], H + 9 * (size_t)l);
  }
}

// 4-D neighbourhood restating MH.cpp:231-253: FlannBasedMatcher::radiusMatch on float (x1,y1,x2,y2) with
// maxDistance = 1/locality (OpenCV squares maxDistance for the L2 FLANN index).  The reference uses FLANN's DEFAULT
// index and search parameters: 4 randomised KD-trees searched with checks = 32, and a radius result set reports
// full() == true, so a query stops after examining 32 leaves (= 32 points, the query itself among them): it returns
// (approximately) the 31 nearest other points that lie within the radius, NOT the full radius ball.  The oracle
// restates that behaviour exactly-defined: the `max_neighbours` nearest points (ties by index) among those with
// d^2 <= radius^2; max_neighbours <= 0 means the full ball.  Returns directed CSR with ascending neighbour index;
// self excluded (MH.cpp:537).  Call with adj == null to get the total count first.
int64_t orc_radius_neighbours(const double* pts, int N, double radius, int max_neighbours, int64_t* offsets /*N+1*/,
                              int32_t* adj) {
  const float r2 = (float)(radius * radius);
  int64_t total = 0;
  std::vector<std::pair<float, int>> cand;
  std::vector<int> row;
  for (int i = 0; i < N; ++i) {
    if (offsets) offsets[i] = total;
    const float a0 = (float)pts[4 * i], a1 = (float)pts[4 * i + 1], a2 = (float)pts[4 * i + 2], a3 = (float)pts[4 * i + 3];
    cand.clear();
    for (int j = 0; j < N; ++j) {
      if (j == i) continue;
      const float d0 = a0 - (float)pts[4 * j], d1 = a1 - (float)pts[4 * j + 1], d2 = a2 - (float)pts[4 * j + 2],
                  d3 = a3 - (float)pts[4 * j + 3];
      const float d = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
      if (d <= r2) cand.emplace_back(d, j);
    }
    if (max_neighbours > 0 && (int)cand.size() > max_neighbours) {
      std::partial_sort(cand.begin(), cand.begin() + max_neighbours, cand.end());
      cand.resize(max_neighbours);
    }
    row.clear();
    for (auto& c : cand) row.push_back(c.second);
    std::sort(row.begin(), row.end());
    if (adj) for (size_t k = 0; k < row.size(); ++k) adj[total + k] = row[k];
    total += (int64_t)row.size();
  }
  if (offsets) offsets[N] = total;
  return total;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// Pre-path: the per-correspondence refinement loop of
// MultiH::GetFundamentalMatrixAndRefineData (MH.cpp:786-838) with F given:
// OptimalTriangulation (MH.cpp:1116-1188, Hartley-Sturm with the reference's
// un-normalised epipole "rotations" MH.cpp:801-802), GetAffineConsistency
// (:1057-1090, only distanceError is used, :826), GetBetaScale (:1092-1114) and
// GetOptimalAffineTransformation (:1190-1223).  cv::solvePoly (OpenCV: Durand-
// Kerner from the start values (1+i)^k, leading coefficients <= DBL_EPSILON
// trimmed) is restated with a relative convergence test instead of OpenCV's
// "until the update is exactly zero or 300*n sweeps".
// keep[i] = 1 if the correspondence survives; out_pts/out_aff are written at
// index i (not compacted) — the caller compacts in order, as the reference's
// push_back does.
// ---------------------------------------------------------------------------
namespace {
struct Cx { double re, im; };
inline Cx cmul(Cx a, Cx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
inline Cx cdiv(Cx a, Cx b) { const double d = b.re * b.re + b.im * b.im; return {(a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d}; }
int solve_poly_dk(const double* c /*ascending, 7*/, Cx* roots) {
  int n = 6;
  for (; n > 1; --n) if (std::fabs(c[n]) > DBL_EPSILON) break;
  Cx p{1, 0}; const Cx r{1, 1};
  for (int i = 0; i < n; ++i) { roots[i] = p; p = cmul(p, r); }
  for (int iter = 0; iter < 300 * n; ++iter) {
    double maxDiff = 0, maxRoot = 0;
    for (int i = 0; i < n; ++i) {
      p = roots[i];
      Cx num{c[n], 0}, den{c[n], 0};
      for (int j = 0; j < n; ++j) {
        num = cmul(num, p); num.re += c[n - j - 1];
        if (j != i) { Cx d{p.re - roots[j].re, p.im - roots[j].im}; if (d.re != 0 || d.im != 0) den = cmul(den, d); }
      }
      num = cdiv(num, den);
      roots[i] = {p.re - num.re, p.im - num.im};
      maxDiff = std::max(maxDiff, std::hypot(num.re, num.im));
      maxRoot = std::max(maxRoot, std::hypot(roots[i].re, roots[i].im));
    }
    if (maxDiff <= 1e-16 * std::max(maxRoot, 1e-300)) break;
  }
  return n;
}
inline void mat3_inv(const double* m, double* o) {
  const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
  const double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g, det = a * A + b * B + c * C;
  o[0] = A / det; o[1] = -(b * i - c * h) / det; o[2] = (b * f - c * e) / det;
  o[3] = B / det; o[4] = (a * i - c * g) / det; o[5] = -(a * f - c * d) / det;
  o[6] = C / det; o[7] = -(a * h - b * g) / det; o[8] = (a * e - b * d) / det;
}
inline double beta_scale(const double* F, const double* p1, const double* p2) {  // MH.cpp:1092-1114
  const double l1x = F[0] * p2[0] + F[3] * p2[1] + F[6], l1y = F[1] * p2[0] + F[4] * p2[1] + F[7], l1z = F[2] * p2[0] + F[5] * p2[1] + F[8];
  const double l2x = F[0] * p1[0] + F[1] * p1[1] + F[2], l2y = F[3] * p1[0] + F[4] * p1[1] + F[5];
  const double xn = p1[0] + 1.0, yn = -(l1x * xn + l1z) / l1y;
  double dx = xn - p1[0], dy = yn - p1[1];
  const double nn = std::sqrt(dx * dx + dy * dy);
  dx /= nn; dy /= nn;
  return std::fabs(std::sqrt(l2x * l2x + l2y * l2y) /
                   ((-F[0] * dy + F[1] * dx) * p2[0] + (-F[3] * dy + F[4] * dx) * p2[1] - F[6] * dy + F[7] * dx));
}
}  // namespace

extern "C" void orc_prefilter(const double* pts, const double* aff, const double* F, int64_t N, double* out_pts,
                              double* out_aff, int32_t* keep) {
  double e2[3], e1[3], Ft[9];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Ft[i * 3 + j] = F[j * 3 + i];
  epipole2(F, e2);   // MH.cpp:789-793
  epipole2(Ft, e1);  // eigen(F^T F) last row / z, MH.cpp:795-799
  const double R1[9] = {e1[0], e1[1], 0, -e1[1], e1[0], 0, 0, 0, 1};   // MH.cpp:801
  const double R2[9] = {-e2[0], -e2[1], 0, e2[1], -e2[0], 0, 0, 0, 1}; // MH.cpp:802
  const double f1 = 1.0, f2 = 1.0;  // epipoles are divided by z (MH.cpp:793, 799)
  for (int64_t i = 0; i < N; ++i) {
    keep[i] = 0;
    const double p1[3] = {pts[4 * i], pts[4 * i + 1], 1.0}, p2[3] = {pts[4 * i + 2], pts[4 * i + 3], 1.0};
    // ---- OptimalTriangulation (MH.cpp:1116-1188)
    const double T1i[9] = {1, 0, p1[0], 0, 1, p1[1], 0, 0, 1};      // T1.inv()
    const double T2ti[9] = {1, 0, 0, 0, 1, 0, p2[0], p2[1], 1};     // T2.t().inv()
    double F2[9], F3[9], R1t[9];
    mat3_mul(T2ti, F, F2); mat3_mul(F2, T1i, F2);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R1t[r * 3 + c] = R1[c * 3 + r];
    mat3_mul(R2, F2, F3); mat3_mul(F3, R1t, F3);
    const double a = F3[4], b = F3[5], c = F3[7], d = F3[8];
    const double f14 = f1 * f1 * f1 * f1, adbc = a * d - b * c;
    double co[7];
    co[6] = -a * c * f14 * adbc;
    co[5] = (a * a + f2 * f2 * c * c) * (a * a + f2 * f2 * c * c) - (a * d + b * c) * f14 * adbc;
    co[4] = 2 * (a * a + f2 * f2 * c * c) * (2 * a * b + 2 * c * d * f2 * f2) - d * b * f14 * adbc - 2 * a * c * f1 * f1 * adbc;
    co[3] = (2 * a * b + 2 * c * d * f2 * f2) * (2 * a * b + 2 * c * d * f2 * f2) + 2 * (a * a + f2 * f2 * c * c) * (b * b + f2 * f2 * d * d) - 2 * f1 * f1 * adbc * (a * d + b * c);
    co[2] = 2 * (2 * a * b + 2 * c * d * f2 * f2) * (b * b + f2 * f2 * d * d) - 2 * (f1 * f1 * a * d - f1 * f1 * b * c) * b * d - a * c * adbc;
    co[1] = (b * b + f2 * f2 * d * d) * (b * b + f2 * f2 * d * d) - (a * d + b * c) * adbc;
    co[0] = -adbc * b * d;
    Cx roots[6];
    const int n = solve_poly_dk(co, roots);
    double bestS = (double)INT32_MAX, bestT = 0;
    for (int k = 0; k < n; ++k)
      if (std::fabs(roots[k].im) <= 1e-10) {
        const double t = roots[k].re;
        const double val = t * t / (1 + f1 * f1 * t * t) + ((c * t + d) * (c * t + d)) / ((a * t + b) * (a * t + b) + f2 * f2 * ((c * t + d) * (c * t + d)));
        if (val < bestS) { bestS = val; bestT = t; }
      }
    const double valInf = 1 / (f1 * f1) + (c * c) / (a * a + f2 * f2 * c * c);
    if (valInf < bestS) continue;  // MH.cpp:1170-1175 -> error -> not pushed (MH.cpp:816-817)
    const double q1[3] = {0, bestT, 1};
    const double l0 = F3[0] * q1[0] + F3[1] * q1[1] + F3[2], l1 = F3[3] * q1[0] + F3[4] * q1[1] + F3[5], l2 = F3[6] * q1[0] + F3[7] * q1[1] + F3[8];
    double q2[3] = {-l0 * l2, -l1 * l2, l0 * l0 + l1 * l1};
    q2[0] /= q2[2]; q2[1] /= q2[2]; q2[2] = 1.0;
    const double T1[9] = {1, 0, -p1[0], 0, 1, -p1[1], 0, 0, 1}, T2[9] = {1, 0, -p2[0], 0, 1, -p2[1], 0, 0, 1};
    double M1[9], M2[9], M1i[9], M2i[9];
    mat3_mul(R1, T1, M1); mat3_mul(R2, T2, M2); mat3_inv(M1, M1i); mat3_inv(M2, M2i);
    const double cpt[3] = {M1i[0] * q1[0] + M1i[1] * q1[1] + M1i[2], M1i[3] * q1[0] + M1i[4] * q1[1] + M1i[5], 1.0};
    const double dpt[3] = {M2i[0] * q2[0] + M2i[1] * q2[1] + M2i[2], M2i[3] * q2[0] + M2i[4] * q2[1] + M2i[5], 1.0};
    // ---- GetAffineConsistency (MH.cpp:1057-1090), distanceError only
    const double* A = aff + 4 * i;
    double l1v[3] = {F[0] * dpt[0] + F[3] * dpt[1] + F[6], F[1] * dpt[0] + F[4] * dpt[1] + F[7], F[2] * dpt[0] + F[5] * dpt[1] + F[8]};
    double l2v[3] = {F[0] * cpt[0] + F[1] * cpt[1] + F[2], F[3] * cpt[0] + F[4] * cpt[1] + F[5], F[6] * cpt[0] + F[7] * cpt[1] + F[8]};
    double n1[2] = {l1v[0] / l1v[2], l1v[1] / l1v[2]}, n2[2] = {l2v[0] / l2v[2], l2v[1] / l2v[2]};
    double nn = std::sqrt(n1[0] * n1[0] + n1[1] * n1[1]); n1[0] /= nn; n1[1] /= nn;
    nn = std::sqrt(n2[0] * n2[0] + n2[1] * n2[1]); n2[0] /= nn; n2[1] /= nn;
    const double beta = beta_scale(F, cpt, dpt);
    const double det = A[0] * A[3] - A[1] * A[2];
    // r1 = A^-T n1 : A^-1 = [a22 -a12; -a21 a11]/det, transpose -> [a22 -a21; -a12 a11]/det
    const double r1x = (A[3] * n1[0] - A[2] * n1[1]) / det, r1y = (-A[1] * n1[0] + A[0] * n1[1]) / det;
    const double ex = r1x - beta * n2[0], ey = r1y - beta * n2[1];
    if (std::sqrt(ex * ex + ey * ey) > 1.0) continue;  // MH.cpp:826
    // ---- GetOptimalAffineTransformation (MH.cpp:1190-1223)
    if (n1[0] * n2[0] + n1[1] * n2[1] < 0) { n2[0] = -n2[0]; n2[1] = -n2[1]; }
    const double bx = beta * n2[0], by = beta * n2[1];
    // C x = rhs, C = [I4 G; G^T 0] with G rows (-bx 0), (0 -bx), (-by 0), (0 -by): solved in closed form
    // x = a - G mu, G^T x = (-n1) => mu = (G^T G)^-1 (G^T a + n1), G^T G = (bx^2 + by^2) I2
    const double gta0 = -bx * A[0] - by * A[2], gta1 = -bx * A[1] - by * A[3];
    const double s2 = bx * bx + by * by;
    const double mu0 = (gta0 + n1[0]) / s2, mu1 = (gta1 + n1[1]) / s2;
    out_aff[4 * i + 0] = A[0] + bx * mu0; out_aff[4 * i + 1] = A[1] + bx * mu1;
    out_aff[4 * i + 2] = A[2] + by * mu0; out_aff[4 * i + 3] = A[3] + by * mu1;
    out_pts[4 * i + 0] = cpt[0]; out_pts[4 * i + 1] = cpt[1]; out_pts[4 * i + 2] = dpt[0]; out_pts[4 * i + 3] = dpt[1];
    keep[i] = 1;
  }
}
