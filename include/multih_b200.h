/* ============================================================================
 * multih_b200.h — C ABI of libmultih_b200.so: the B200-native (sm_100a) hot path
 * of Multi-H behind the reference's pipeline surface.
 *
 * The reference (danini/multi-h) has no plugin / FFI layer; its boundary is the
 * C++ class `MultiH` (MultiH/MultiH/MultiH.h:20-149).  Every entry point below
 * cites the reference member it replaces (paths relative to MultiH/MultiH/).
 * A C++ shim with the reference's class and method names lives in
 * include/multih_b200.hpp and forwards to this ABI.
 *
 * Conventions
 *  - plain C: pointers + sizes, no C++/torch types; every call returns an
 *    mh_status (0 = ok); no exception crosses the boundary; the message of the
 *    last failure is mh_last_error(ctx).
 *  - "host" pointers are caller-owned host memory, only touched during the call.
 *    "d_" pointers are DEVICE pointers on the context's device (from mh_alloc,
 *    cudaMalloc or a torch tensor's data_ptr()).
 *  - DEVICE-SPACE data are FP32 and live in the context's normalised image
 *    coordinates (Hartley-style similarity per image, fixed by mh_set_geometry):
 *      d_pts  float4[N]   = (x1', y1', x2', y2')
 *      d_aff  float4[N]   = (a11', a12', a21', a22')  (A' = s2/s1 * A)
 *      d_hyp  float[K][12]= H' = T2 H T1^-1 row-major (9 used, 3 pad)
 *    Host-space data are FP64 in pixels exactly as the reference holds them
 *    (x1 y1 x2 y2 | a11 a12 a21 a22 | 3x3 row-major H, F with x2^T F x1 = 0).
 *    Thresholds are always given in pixels; the library rescales them.
 *  - one context = one CUDA device + one stream; a context is not thread-safe,
 *    distinct contexts are independent.
 *  - there is NO CPU fallback: without a usable CUDA device mh_create fails with
 *    MH_ECUDA and nothing else can be called.
 * ==========================================================================*/
#ifndef MULTIH_B200_H
#define MULTIH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mh_ctx mh_ctx;

typedef enum mh_status {
  MH_OK = 0,
  MH_EINVAL = 1,      /* bad argument (null pointer, N < 8 where the reference refuses: MultiH.cpp:44-50, ...) */
  MH_ECUDA = 2,       /* CUDA runtime / launch failure, or no device */
  MH_ENCCL = 3,       /* NCCL not loadable, or a communicator / collective call failed (mh_comm_*, mh_step_sharded) */
  MH_EDEGENERATE = 4, /* degenerate geometry (||F|| < 1e-5: MultiH.cpp:779; no cluster left) */
  MH_ENOMEM = 5
} mh_status;

/* Mirrors the MultiH constructor (MultiH.h:49-53) and its compile-time constants (MultiH.h:7-18). */
typedef struct mh_params {
  double thr_fundamental; /* _thr_fund_mat  (unused on the hot path: F is an input)            */
  double thr_homography;  /* _thr_hom       threshold_homography, px                           */
  double locality;        /* _locality      locality_lambda                                    */
  double lambda;          /* _lambda        spatial coherence weight (energy_lambda)           */
  int32_t min_inliers;    /* _minimum_inlier_number                                            */
  double straightness;    /* DEFAULT_LINENESS_THRESHOLD 0.005 (MultiH.h:13)                    */
  int32_t max_iterations; /* MAX_ITERATION_NUMBER 500 (MultiH.h:14)                            */
  double convergence;     /* CONVERGENCE_THRESHOLD 1e-5 (MultiH.h:15)                          */
  int32_t meanshift_metric; /* 0 = L1_REF (MeanShiftClustering.h:76-85, the reference), 1 = L2 window, same sequential
                               algorithm; 2 = L2 window, all seeds batched, Gram on the tensor cores (tcgen05; parity T3) */
  uint32_t rng_seed;      /* state of the injected MSVC-rand() restatement (1 = unseeded rand) */
  int32_t max_gc_cycles;  /* expansion(iter, 1000)  (MultiH.cpp:543)                           */
  int32_t max_neighbours; /* 31: FLANN default checks=32 caps radiusMatch (MultiH.cpp:252-253) at the ~31 nearest
                             other sites inside the radius; <= 0 = the full radius ball                      */
  int32_t precise_pipeline; /* mh_process data path: 1 (default) = FP64 pixel-space kernels tracking the reference's
                               arithmetic; 0 = FP32 normalised throughput kernels (same control flow)        */
  int32_t prefilter;        /* mh_process: 1 = run K0 first (MultiH.cpp:807-838: Hartley-Sturm correction, affine
                               consistency test, optimal affine) as Process() does after estimating F; dropped
                               correspondences get label -2.  0 (default) = inputs are already refined             */
  int32_t compatibility_check; /* mh_process: 1 (default) = finish with HomographyCompatibilityCheck (MultiH.cpp:78-86, 100-222)
                                  as Process() does; 0 = return the clusters of the alternating optimisation              */
  int32_t lm_refine;           /* 1 (default) = every GetHomography3PT fit that the reference polishes (cluster fits MultiH.cpp:684,
                                  mode fits :427) goes through its Levenberg-Marquardt iteration (Homography_Refine3PTCallback.h,
                                  Utilities.hpp:762-869), reproduced step for step; 0 = the linear solutions                */
} mh_params;

/* main.cpp:55-59: 2.6 / 2.2 / 0.005 / 0.5 / 20 (the CLI's values; the class defaults of MultiH.h:7-10 are what the MultiH
 * shims of include/multih_b200.hpp and multih_b200.MultiH pass).  Where mh_process deliberately differs from
 * MultiH::Process(), all in one place:
 *   - F is an input; the RANSAC of MultiH.cpp:775 and its inlier mask are upstream;
 *   - prefilter = 0 by default (inputs already refined); the shims set 1, as Process() refines every correspondence;
 *   - the LM polish of the 3PT fits is reproduced (lm_refine); the HAF polish (RefineHomographyHAF) is not run: it never writes
 *     its result back (Homography_RefineHAFCallback.h:58 rebinds a local header), so skipping it changes nothing — and its
 *     callback reads out of bounds (:148-151);
 *   - the neighbourhood is the exact 31 nearest within the radius (max_neighbours), FLANN's is a randomised 32-check search;
 *   - K <= 1 after the loop: the reference discards everything and calls cv::findHomography (HandleDegenerateCase,
 *     MultiH.cpp:88-94, 719-741); mh_process returns the single surviving homography with its inliers (labels 0 / -1,
 *     MultiH.cpp:280-285, 743-768) or K_out = 0 with all labels -1 — labels_out[i] < K_out always holds. */
void mh_default_params(mh_params* p);

/* ---- context ----------------------------------------------------------- */
mh_status mh_create(const mh_params* params, int device, mh_ctx** out); /* MultiH::MultiH  (MultiH.cpp:10-21) */
void mh_destroy(mh_ctx* ctx);                                           /* MultiH::~MultiH (MultiH.cpp:23-30) */
const char* mh_last_error(const mh_ctx* ctx);
const char* mh_version(void);
mh_status mh_set_stream(mh_ctx* ctx, void* cuda_stream); /* run on the caller's stream (e.g. torch's current stream) */
mh_status mh_sync(mh_ctx* ctx);
mh_status mh_alloc(mh_ctx* ctx, uint64_t bytes, void** d_ptr);
mh_status mh_free(mh_ctx* ctx, void* d_ptr);
mh_status mh_host_alloc(mh_ctx* ctx, uint64_t bytes, void** pinned_host_ptr);
mh_status mh_host_free(mh_ctx* ctx, void* pinned_host_ptr);
mh_status mh_memcpy_d2h(mh_ctx* ctx, void* host, const void* d_src, uint64_t bytes);
mh_status mh_memcpy_h2d(mh_ctx* ctx, void* d_dst, const void* host, uint64_t bytes);
int64_t mh_kernel_launches(const mh_ctx* ctx); /* number of OUR kernels launched by this context so far */

/* ---- pair geometry ------------------------------------------------------
 * Replaces the members fundamental_matrix / epipole_2 that
 * GetFundamentalMatrixAndRefineData leaves behind (MultiH.cpp:775-793): F is an
 * INPUT here.  norm1/norm2 = (scale, tx, ty) of T = [s 0 tx; 0 s ty; 0 0 1] per
 * image; pass NULL to derive them from a strided sample of `pts_host` (N x 4). */
mh_status mh_set_geometry(mh_ctx* ctx, const double F[9], const double norm1[3], const double norm2[3],
                          const double* pts_host, int64_t N);
mh_status mh_get_geometry(const mh_ctx* ctx, double F[9], double e2[2], double norm1[3], double norm2[3]);

/* ---- host <-> device-space conversion ----------------------------------- */
/* correspondences in the reference's layout (MultiH.h:78-80) -> normalised float4 arrays.  Either half (points or
 * affines: both of its pointers NULL) may be omitted so that the two uploads can be overlapped with compute. */
mh_status mh_upload_correspondences(mh_ctx* ctx, const double* pts_host, const double* aff_host, int64_t N,
                                    void* d_pts, void* d_aff /* may be NULL */);
mh_status mh_hypotheses_from_host(mh_ctx* ctx, const double* H_host /*K x 9 px*/, int32_t K, void* d_hyp);
mh_status mh_hypotheses_to_host(mh_ctx* ctx, const void* d_hyp, int32_t K, double* H_host /*K x 9 px*/,
                                int32_t divide_by_h33);

/* ---- K0: per-correspondence pre-filter (the step directly before K1) ------------
 * The refinement loop of GetFundamentalMatrixAndRefineData with F given (MultiH.cpp:786-838): OptimalTriangulation
 * (:1116-1188), GetAffineConsistency / GetBetaScale (:1057-1114, drop when distanceError > 1), GetOptimalAffineTransformation
 * (:1190-1223).  FP64 pixels in, survivors out in input order (the reference's push_back order); keep_host (N bytes,
 * optional) flags the survivors; *M_out = their number.  The RANSAC estimation of F itself (MultiH.cpp:775) stays upstream. */
mh_status mh_prefilter(mh_ctx* ctx, const double* pts_host, const double* aff_host, const double F[9], int64_t N,
                       double* pts_out_host, double* aff_out_host, uint8_t* keep_host, int64_t* M_out);
/* same on device-resident FP64 arrays (N x 4 each); d_keep i32 [N] */
mh_status mh_prefilter_device(mh_ctx* ctx, const void* d_pts64, const void* d_aff64, const double F[9], int64_t N,
                              void* d_pts64_out, void* d_aff64_out, void* d_keep, int64_t* M_out);

/* ---- K1: per-correspondence HAF hypotheses ------------------------------
 * MultiH::ComputeLocalHomographies (MultiH.cpp:696-717) -> GetHomographyHAF (:850-911).
 * Solved in pixel coordinates in FP64 (A^T A + cyclic Jacobi, as the reference) — the least-squares estimate is not
 * invariant to normalisation.  `precision` is reserved (0). */
mh_status mh_haf_hypotheses(mh_ctx* ctx, const void* d_pts, const void* d_aff, int64_t N, void* d_hyp,
                            int32_t precision);

/* ---- K2: N x K reprojection residual / data cost ------------------------
 * dataEnergy (MultiH.cpp:473-504) with EnergyDataStruct (MultiH.h:23-47).
 * dense: int32 (elem_bytes=4) or int16 (2) matrix [N][K+1], site-major, column 0 =
 * outlier label — the layout GCO's setDataCost(int*) consumes (GCoptimization.h:336-343). */
mh_status mh_data_cost_dense(mh_ctx* ctx, const void* d_pts, int64_t N, const void* d_hyp, int32_t K, void* d_cost,
                             int32_t elem_bytes);
/* raw squared residuals in px^2, float [N][K] (MultiH.cpp:491-498) — parity/debug aid */
mh_status mh_residuals(mh_ctx* ctx, const void* d_pts, int64_t N, const void* d_hyp, int32_t K, void* d_d2);
/* fused: nothing N x K touches HBM.  Per site: up to `kmax` (label,cost) entries with d2 < T packed as
 * (label << 8 | cost) (label 1-based as in the dense matrix; cost 0..255 — requires 100/lambda <= 255), the number of
 * such entries (may exceed kmax = overflow), and the data-term argmin packed (cost << 32 | label) (label 0 = outlier;
 * what GCO returns without smoothness, GCoptimization.cpp solveSpecialCases).  Per hypothesis: #sites with d2 < thr_H^2
 * (the inlier test of MultiH.cpp:441 / :762).  Any output pointer may be NULL.  d_list_count / d_inlier_count /
 * d_best are (re)initialised by the call. */
mh_status mh_data_cost_fused(mh_ctx* ctx, const void* d_pts, int64_t N, const void* d_hyp, int32_t K, int32_t kmax,
                             void* d_list /*u32 [N][kmax]*/, void* d_list_count /*i32 [N]*/,
                             void* d_best /*u64 [N]*/, void* d_inlier_count /*i32 [K]*/);
/* MergingStep's inlier scan + straightness statistics (MultiH.cpp:430-463): per hypothesis the inlier count and the
 * 6 uniques (xx xy x yy y n) of S = sum [x y 1]^T[x y 1] over inliers in PIXEL coordinates (FP64 [K][6]),
 * and (host side, optional) lambda_min of S and the keep flag (lambda_min >= straightness && count >= 3). */
mh_status mh_inlier_stats(mh_ctx* ctx, const void* d_pts, int64_t N, const void* d_hyp, int32_t K,
                          double* scatter_host /*K x 6*/, double* lambda_min_host /*K*/, int32_t* keep_host /*K*/);
/* ComputeInliersOfHomography (MultiH.cpp:743-768): labels[i] = idx where d2 < thr_H^2 */
mh_status mh_inliers_of_homography(mh_ctx* ctx, const void* d_pts, int64_t N, const void* d_hyp_one, int32_t idx,
                                   void* d_labels /*i32 [N]*/);

/* ---- K3: mean-shift -------------------------------------------------------
 * features: MultiH.cpp:612-646 (10-D, from per-point hypotheses + points) and :359-389 (6-D, from K hypotheses);
 * output FP64 row-major in PIXEL units exactly as the reference builds them. */
mh_status mh_features10(mh_ctx* ctx, const void* d_hyp, const void* d_pts, int64_t N, void* d_feat /*f64 [N][10]*/);
mh_status mh_features6(mh_ctx* ctx, const void* d_hyp, int32_t K, void* d_feat /*f64 [K][6]*/);
/* MeanShiftClustering<double>::Cluster (MeanShiftClustering.h:22-157): sequential-seed flat-kernel mean-shift, run as ONE
 * persistent cooperative kernel in FP64 (seeds drawn by the restated MSVC rand()).  d_centres f64 [max_c][D],
 * d_assign i32 [N] (cluster with most votes, first wins ties).  *C_out = number of centres. */
/* A mean-shift trajectory ends after this many window iterations at the latest (the reference's `while (1)`,
 * MeanShiftClustering.h:62-124, never returns when the mean cycles — possible with its L1 window; observed on synthetic
 * 5000-correspondence scenes). */
#define MH_MS_MAX_WINDOW_ITERS 200

/* The seed generator's state persists across mh_meanshift calls of a context (as rand() does in the reference process);
 * mh_process re-seeds it from params.rng_seed on entry. */
mh_status mh_set_rng_state(mh_ctx* ctx, uint32_t state);
uint32_t mh_get_rng_state(const mh_ctx* ctx);
mh_status mh_meanshift(mh_ctx* ctx, const void* d_feat, int32_t N, int32_t D, double bandwidth, void* d_centres,
                       int32_t max_c, void* d_assign, int32_t* C_out, int64_t* stats_out /*[2] or NULL*/);

/* ---- K4: per-label refit ---------------------------------------------------
 * LabelingStep's gather (MultiH.cpp:545-584) + GetHomographyHAFNonminimal (:913-990, linear solution =
 * do_numerical_refinement=false).  d_labels i32 [N] in -1..K-1.  Labels without members keep d_hyp[l] untouched
 * (MultiH.cpp:592-593).  d_count i32 [K] optional. */
mh_status mh_refit_haf(mh_ctx* ctx, const void* d_pts, const void* d_aff, const void* d_labels, int64_t N, int32_t K,
                       void* d_hyp, void* d_count);
/* The same refit split at its reduction point, for correspondence-sharded multi-GPU runs: accumulate writes, per label,
 * the 10 uniques of SUM A_i^T A_i (FP64, pixel coordinates), the member count ([10]) and a pad ([11]) into
 * d_acc f64 [K][12]; shards all-reduce(sum) that array; solve does the batched 4x4 eigen-solves. */
mh_status mh_refit_haf_accumulate(mh_ctx* ctx, const void* d_pts, const void* d_aff, const void* d_labels, int64_t N,
                                  int32_t K, void* d_acc);
mh_status mh_refit_haf_solve(mh_ctx* ctx, const void* d_acc, int32_t K, void* d_hyp, void* d_count);
/* Glue between K2 and K4 on the device: d_labels[i] = (d_best[i] & 0xffffffff) - 1 (the per-site argmin of
 * mh_data_cost_fused as the -1..K-1 label array of MultiH.h:61-62); and (un)packing of the per-hypothesis inlier counts
 * into the pad column d_acc[k][11], so that one all-reduce carries both statistics of a correspondence shard. */
mh_status mh_labels_from_best(mh_ctx* ctx, const void* d_best, int64_t N, void* d_labels);
mh_status mh_pack_inlier_counts(mh_ctx* ctx, void* d_inlier_count, int32_t K, void* d_acc, int32_t unpack);
/* GetHomography3PT (MultiH.cpp:995-1055, linear solution) per cluster of EstablishStablePointSets (:664-688):
 * d_assign i32 [N] in -1..C-1; clusters with < 3 members get keep=0 (:667). */
mh_status mh_refit_3pt(mh_ctx* ctx, const void* d_pts, const void* d_assign, int64_t N, int32_t C, void* d_hyp,
                       void* d_keep /*i32 [C]*/);
/* MergingStep's mode -> homography (MultiH.cpp:408-427): 3PT on (0,0),(1,0),(0,1) -> 6-D mode (pixel units). */
mh_status mh_modes_to_hypotheses(mh_ctx* ctx, const void* d_modes /*f64 [C][6]*/, int32_t C, void* d_hyp);

/* ---- host combinatorial steps (consume GPU-built costs) ------------------- */
/* 4-D neighbourhood replacing FlannBasedMatcher::radiusMatch (MultiH.cpp:231-253): per site the `max_neighbours`
 * nearest other sites (ties by index) with d^2 <= radius^2 on float (x1,y1,x2,y2); <= 0 = full ball.  Directed CSR out,
 * ascending neighbour index; call with adj_host == NULL to size (or pass N * max_neighbours entries).  ctx may be NULL: host
 * grid search.  With a context and 1 <= max_neighbours <= 64 the search runs on the GPU for N >= 256 (K5):
 * the same set, bit for bit (mh_diag_set_neighbourhood_backend forces either). */
mh_status mh_neighbourhood(mh_ctx* ctx, const double* pts_host, int32_t N, double radius, int32_t max_neighbours,
                           int64_t* offsets_host, int32_t* adj_host, int64_t* total_out);
/* Alpha-expansion (the role of GCoptimizationGeneralGraph in MultiH.cpp:520-543; own implementation, int64 totals):
 * dense site-major int32 costs [N][L], Potts weight per DIRECTED adjacency entry. */
mh_status mh_alpha_expansion(mh_ctx* ctx, const int32_t* cost_host, int32_t N, int32_t L, int32_t potts,
                             const int64_t* offsets_host, const int32_t* adj_host, const int32_t* init_labels,
                             int32_t max_cycles, int32_t* labels_out, int64_t* energy_out);
/* The same with sparse-with-default data costs (cf. GCO's setDataCost(SparseDataCost), GCoptimization.h:220-224): per site up to
 * kmax entries (label << 16 | cost) for labels 1..L-1 and their number; label 0 (the outlier label of MultiH.cpp:488-489) costs
 * cost_label0 at every site, a label that is not listed costs cost_default (MultiH.cpp:499-500: d2 >= T).  This is what
 * mh_process feeds its labelling steps: N x kmax words leave the device instead of the N x (K + 1) matrix.  counts[i] > kmax is
 * MH_EINVAL (truncated list).  Result identical to mh_alpha_expansion on the expanded matrix. */
mh_status mh_alpha_expansion_sparse(mh_ctx* ctx, const uint32_t* lists_host /*[N][kmax]*/, const int32_t* counts_host /*[N]*/,
                                    int32_t kmax, int32_t N, int32_t L, int32_t cost_label0, int32_t cost_default, int32_t potts,
                                    const int64_t* offsets_host, const int32_t* adj_host, const int32_t* init_labels,
                                    int32_t max_cycles, int32_t* labels_out, int64_t* energy_out);

/* ---- whole path ------------------------------------------------------------
 * MultiH::Process (MultiH.cpp:42-98) from ComputeLocalHomographies on, with F supplied:
 * K1 -> 10-D mean-shift -> cluster 3PT -> { merge (6-D mean-shift, inlier/straightness test) <-> label (K2 dense costs
 * -> host alpha-expansion) + refit (K4) } until convergence (MultiH.cpp:224-312).
 * labels_out: N, -1 = outlier (GetLabels, MultiH.h:62), -2 = dropped by the pre-filter (params.prefilter only); H_out: up to Kmax x 9 px (GetHomography, MultiH.h:69);
 * K_out = GetClusterNumber (MultiH.h:67). */
/* HomographyCompatibilityCheck (MultiH.cpp:100-222): cross-validation of the clusters — 501 random three-point fits per
 * cluster (GPU), median of the per-trial median transfer errors against thr_H^2 * 81/16; clusters that fail, and clusters
 * with fewer than min_inliers members, are removed (their points become outliers, higher labels move down).  Serial draw
 * order of the context's rand() restatement.  labels_host [N] and H_host [K][9] (pixel) are updated in place, *K_inout too;
 * medians_out [K] optional (NaN = cluster not tested).  Needs mh_set_geometry (F). */
mh_status mh_compatibility_check(mh_ctx* ctx, const double* pts_host, int32_t N, int32_t* labels_host, double* H_host,
                                 int32_t* K_inout, double* medians_out);
/* Its two host halves (no context; mh_compatibility_check = plan -> GPU order statistics -> decide):
 * plan   replays the reference's sampling (rand() draws without replacement from an evolving point vector, MultiH.cpp:142-154,
 *        183-194): per tested cluster (>= max(min_inliers, 4) members) its members [moff[k], moff[k+1]) in index order and
 *        MH_COMPAT_TRIALS x 3 sampled correspondence indices; removed[c] = 1 for clusters below min_inliers.
 * decide takes per (tested cluster, trial) 8 doubles — the sorted squared transfer errors of the n = members - 3 other members
 *        at indices max(0, n/2 - 3) .. min(n - 1, n/2 + 1) (5 slots, +inf when absent) and the three largest (-inf when absent) —
 *        replays the reference's per-trial "median" with its three stale buffer entries and the off-by-one even case
 *        (MultiH.cpp:140, 177-178), takes the median over the trials, removes clusters above thr_H^2 * 81/16 and relabels. */
#define MH_COMPAT_TRIALS 501 /* MAX(501, MIN(501, n choose 3)), MultiH.cpp:130 */
mh_status mh_compat_plan(const int32_t* labels, int32_t N, int32_t K, int32_t min_inliers, uint32_t* rng_state, int32_t* tested,
                         int32_t* T_out, int32_t* members, int32_t* moff, int32_t* samples, int32_t* removed);
mh_status mh_compat_decide(const int32_t* tested, int32_t T, const int32_t* moff, const double* stats, double thr_homography,
                           int32_t* removed, int32_t N, int32_t* labels, double* H, int32_t* K_inout, double* medians_out);
mh_status mh_process(mh_ctx* ctx, const double* pts_host, const double* aff_host, const double F[9], int32_t N,
                     int32_t* labels_out, double* H_out, int32_t Kmax, int32_t* K_out);
double mh_get_energy(const mh_ctx* ctx);        /* GetEnergy          (MultiH.h:74) */
int32_t mh_get_iterations(const mh_ctx* ctx);   /* GetIterationNumber (MultiH.h:68) */
/* stage timers mirroring the reference's printf timers (MultiH.cpp:68,74,258,310): ms for
 * [0] point-wise homographies [1] stable clusters [2] adjacency [3] alternating optimisation [4] total */
mh_status mh_get_stage_ms(const mh_ctx* ctx, double ms[5]);

/* ---- multi-GPU: correspondences sharded over ranks, hypotheses replicated ---------------------------
 * The reference is one CPU process; its dataEnergy loop (MultiH.cpp:909-924) is independent per correspondence and its
 * refits (HomographyHAFNonminimal, MultiH.cpp:1057-1115) are sums over members, so the path shards over correspondences
 * with one broadcast (hypotheses) and one all-reduce (refit statistics + inlier counts) per pass.  One context = one
 * device = one rank.  NCCL is bound at run time (the libnccl.so.2 already in the process, else dlopen, else $MH_NCCL_LIB);
 * every failure of that layer is MH_ENCCL with the NCCL message in mh_last_error.
 *   mh_comm_unique_id    rank 0 makes the 128-byte id; the caller carries it to the other ranks (MPI, file, socket, ...)
 *   mh_comm_init         collective over `world` contexts; mh_destroy (or mh_comm_destroy) releases the communicator
 *   mh_comm_broadcast / mh_comm_allreduce_sum_f64   the two collectives on the context's stream (building blocks)
 *   mh_step_sharded      one whole sharded pass: broadcast of d_hyp f32 [K][12] from rank 0 (on the communicator's own
 *                        stream, under K1) -> K1 into d_hyp_pt [n_local][12] -> K2 fused into d_best u64 [n_local] ->
 *                        d_labels i32 [n_local] -> K4 statistics of the shard -> ONE all-reduce (statistics + inlier counts)
 *                        -> K4 solves into d_ref f32 [K][12], whole-scene counts into d_inliers i32 [K].  The all-reduce
 *                        of pass i overlaps pass i+1 (double-buffered): d_ref / d_inliers of pass i are complete after
 *                        the next mh_step_sharded or after mh_step_sharded_finish.  Without a communicator the same call
 *                        runs the single-GPU pass and completes in stream order.  ev_k2_*: optional cudaEvent_t recorded
 *                        around the K2 launch. */
mh_status mh_comm_unique_id(void* id128);
mh_status mh_comm_init(mh_ctx* ctx, const void* id128, int32_t rank, int32_t world);
mh_status mh_comm_destroy(mh_ctx* ctx);
int32_t mh_comm_rank(const mh_ctx* ctx);
int32_t mh_comm_world(const mh_ctx* ctx);
mh_status mh_comm_broadcast(mh_ctx* ctx, void* d_buf, uint64_t bytes, int32_t root);
mh_status mh_comm_allreduce_sum_f64(mh_ctx* ctx, void* d_buf, uint64_t count);
mh_status mh_step_sharded(mh_ctx* ctx, const void* d_pts, const void* d_aff, int64_t n_local, void* d_hyp, int32_t K,
                          void* d_hyp_pt, void* d_best, void* d_labels, void* d_inliers, void* d_ref, void* ev_k2_begin,
                          void* ev_k2_end);
mh_status mh_step_sharded_finish(mh_ctx* ctx);

/* ---- diagnostics (measurement aids, not part of the reference surface) -------- */
/* FP32 FMA-pipe peak of this GPU: variant 0 = scalar FFMA, 1 = packed FFMA2; the roofline denominator of K2. */
mh_status mh_diag_fp32_peak(mh_ctx* ctx, int32_t variant, int32_t iters, double* tflops_out, double* ms_out);
/* legacy tensor path probe: mma.sync.m16n8k8 TF32 dense TFLOP/s and warp-MMAs per clock per SM (at 1965 MHz) */
mh_status mh_diag_mma_tf32_peak(mh_ctx* ctx, int32_t iters, double* tflops_out, double* mma_per_clk_per_sm);
/* K2 fused inner loop: 1 = packed FFMA2 (default), 0 = scalar FFMA (A/B evidence only). */
mh_status mh_diag_set_fused_variant(mh_ctx* ctx, int32_t variant);
/* K2 fast path launch shape (threads/CTA x CTAs/SM): 0 = 256x3, 1 = 256x2, 2 = 256x4, 3 = 128x5, 4 = 128x6,
 * 5 = 128x7 (default), 6 = 128x4 — tuning aid. */
mh_status mh_diag_set_fast_config(mh_ctx* ctx, int32_t config);
int32_t mh_diag_get_fast_config(mh_ctx* ctx);
/* 1 = tiled dense-cost kernel with TMA row stores (default; float->int through a denormal product), 2 / 3 = same kernel with
 * the 2^23-magic / F2I conversion, 0 = first-generation scalar-store kernel (A/B evidence) */
mh_status mh_diag_set_dense_variant(mh_ctx* ctx, int32_t variant);
/* mh_neighbourhood backend: 0 = auto, 1 = host, 2 = device */
mh_status mh_diag_set_neighbourhood_backend(mh_ctx* ctx, int32_t backend);
/* where the alternating optimisation of the last mh_process spent its time, ms: [0] mean-shift of the hypotheses
 * [1] mode fit + inlier scan + straightness [2] data-cost matrix (+ D2H) [3] host alpha-expansion [4] refit */
mh_status mh_diag_get_alternating_ms(const mh_ctx* ctx, double ms[5]);

#ifdef __cplusplus
}
#endif
#endif /* MULTIH_B200_H */
