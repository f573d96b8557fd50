/*
 * hop_c_api.h -- C ABI of libhop.so, the B200 (sm_100a) implementation of the pose-hypothesis hot path of
 * wenbowen123/icra20-hand-object-pose.  Plain pointers and sizes only; no C++/torch types cross this boundary.
 *
 * The reference has no FFI: its boundary is a set of C++ methods.  Each entry point below names the reference
 * interface it replaces (paths relative to the reference root):
 *
 *   hop_icp_refine        <- PoseEstimator<PointT>::refineByICP()   src/perception/src/PoseEstimator.cpp:235-275
 *                            (per hypothesis: Utils::runICP<PointT>  src/perception/src/Utils.cpp:188-229)
 *   hop_lcp_score         <- PoseEstimator<PointT>::selectBest()    src/perception/src/PoseEstimator.cpp:465-502
 *                            (per hypothesis: Utils::computeLCP<PointT> src/perception/src/Utils.cpp:372-444)
 *   hop_select_topk       <- the argmax inside selectBest (:491-496) / the 100-candidate cap of refineByICP (:241)
 *   hop_verify_lcp        <- gr::CongruentSetExplorationBase::TryCongruentSet / Verify   (all trials of a frame at once)
 *                            src/OpenGR_4pcs/src/gr/algorithms/congruentSetExplorationBase.hpp:221-340, 346-435
 *                            (+ MatchBase::ComputeRigidTransformation matchBase.hpp:230-377)
 *   hop_hand_overlap      <- objFuncPSO                             src/perception/src/Hand.cpp:10-178
 *   hop_cloud_upload      <- the pcl::PointCloud<PointT>::Ptr arguments of the calls above (PointSurfel /
 *                            PointXYZRGBNormal: xyz + normal + confidence)
 *   hop_pose_rec          <- class PoseHypo                          src/perception/include/PoseHypo.h:7-27
 *
 * Conventions
 *   - every function returns 0 (HOP_OK) or a negative HOP_E* code; hop_last_error() gives the message.
 *   - 4x4 transforms are 16 floats, COLUMN-major (Eigen::Matrix4f::data() passes through unchanged).
 *   - units: metres, degrees (as in config_autodataset.yaml).
 *   - "host" entry points take host pointers and block until results are in the caller's buffers
 *     (they include the host<->device copies); "_dev" entry points take device pointers, enqueue on the
 *     context's stream and return without synchronising.
 *   - there is NO CPU fallback: without a CUDA device hop_create() fails with HOP_ENODEV.
 */
#ifndef HOP_C_API_H_
#define HOP_C_API_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HOP_OK 0
#define HOP_EINVAL (-1)  /* bad argument */
#define HOP_ECUDA (-2)   /* CUDA runtime error (message in hop_last_error) */
#define HOP_ENODEV (-3)  /* no usable CUDA device (this library never computes on the CPU) */
#define HOP_ENOMEM (-4)

typedef struct hop_ctx hop_ctx;     /* one per host thread / GPU; owns a stream and scratch buffers */
typedef struct hop_cloud hop_cloud; /* device-resident point cloud (+ cached nearest-neighbour grids) */

/* wire record for winners; mirrors class PoseHypo (80 bytes, PoseHypo.h:7-27) */
typedef struct hop_pose_rec {
  float pose[16]; /* model2scene, column-major */
  float score;    /* _lcp_score */
  int32_t id;     /* _id (index of the hypothesis in the caller's batch) */
  int32_t frame;  /* caller-defined tag (frame / rank) */
  int32_t pad;
} hop_pose_rec;

/* Utils::runICP arguments (Utils.cpp:188) + the PCL criteria the reference sets (Utils.cpp:207-208) */
typedef struct hop_icp_params {
  int32_t max_iter;     /* 10 in refineByICP (PoseEstimator.cpp:266), 50 in handbaseICP (Hand.cpp:734) */
  float angle_deg;      /* rejection_angle: icp_angle_thres (45) */
  float max_dist;       /* max_corres_dist: icp_dist_thres (0.01) */
  double abs_mse_eps;   /* setAbsoluteMSE(1e-6) */
  int32_t mode;         /* 0 = point-to-plane (the reference's hot ICP); 1 = point-to-point SVD (Utils.cpp:135-184) */
  int32_t solver;       /* mode 0 only: 0 = the reference's own per-iteration solver, PCL's float Levenberg-Marquardt with its forward-
                           difference Jacobian and MINPACK stopping rules, replayed on the 13x13 moments (csrc/lm_replay.cuh): the
                           parity path and the default;  1 = one Gauss-Newton step per iteration;  2 = exact minimiser of every
                           iteration's objective (goes further along weak directions than the reference does: not parity) */
  int32_t team_warps;   /* reserved (was: warps cooperating on one hypothesis); ignored */
  int32_t pipeline;     /* mode 0: 0 = persistent fused (default): the whole ICP of a hypothesis inside one CTA, one launch per batch, a
                           CTA's warps solve its hypotheses in parallel;  1 = iteration-synchronous: per ICP iteration one launch that
                           turns correspondences into the 13x13 moments of every active hypothesis (records stay in shared memory) and
                           one that solves, one warp per hypothesis.  mode 1 always runs iteration-synchronous. */
} hop_icp_params;

/* Utils::computeLCP arguments (Utils.cpp:372) */
typedef struct hop_lcp_params {
  float dist;           /* lcp.dist (0.001) */
  float angle_deg;      /* lcp.normal_angle (10) */
  int32_t use_normal;   /* selectBest passes true,true,true (PoseEstimator.cpp:488) */
  int32_t use_dot_score;
  int32_t use_reciprocal;
  int32_t team_warps;   /* 0 = auto */
} hop_lcp_params;

/* ---- context ------------------------------------------------------------------------------------------- */
int hop_create(int device, hop_ctx **out);
void hop_destroy(hop_ctx *ctx);
const char *hop_last_error(const hop_ctx *ctx); /* ctx may be NULL: error of the last failed hop_create */
/* use the caller's CUDA stream (cudaStream_t as void*) for everything this context enqueues; NULL = own stream */
int hop_set_stream(hop_ctx *ctx, void *cuda_stream);
int hop_sync(hop_ctx *ctx);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
int64_t hop_launch_count(const hop_ctx *ctx);
/* ---- per-kernel device timing (CUDA events on the context's stream around each launch of a kernel family) ------ */
#define HOP_PROF_ICP_CORRESPOND 0 /* icp_moments_kernel / icp_kabsch_sums_kernel: NN + rejection + sums, one launch per ICP iteration (pipeline 1, mode 1) */
#define HOP_PROF_ICP_SOLVE 1      /* icp_solve_kernel: the reference's LM on the moments + convergence, one launch per ICP iteration */
#define HOP_PROF_LCP_SCORE 2      /* lcp_score_kernel (+ its fixed-order reduction) */
#define HOP_PROF_NN_BUILD 3       /* nearest-neighbour grid builds (all kernels of one build = one span) */
#define HOP_PROF_TOPK 4
#define HOP_PROF_VERIFY 5         /* verify_lcp_kernel (K3) */
#define HOP_PROF_HAND 6           /* hand_overlap_kernel (K1) */
#define HOP_PROF_ICP_FUSED 7      /* icp_fused_kernel: the whole ICP of a batch in one launch (default pipeline) */
#define HOP_PROF_S4_PAIRS 8       /* K2a: extract_pairs_kernel + selection (all trials) */
#define HOP_PROF_S4_JOIN 9        /* K2b: prepare_pairs + congruent_join (count, scan, fill) */
#define HOP_PROF_CLUSTER 10       /* hop_cluster_poses_gpu: all block launches of one call = one span */
#define HOP_PROF_FRAME 11         /* hop_frame_to_scene: every launch of one frame's front end = one span */
#define HOP_PROF_SDF 12           /* sdf_kernel / collision_kernel (physics pruning) */
#define HOP_PROF_RENDER 13        /* hop_reject_by_render: bbox + raster + walk kernels of one call = one span */
#define HOP_PROF_KINDS 14
int hop_profile_enable(hop_ctx *ctx, int on);  /* also resets the accumulated numbers */
/* synchronises the stream, folds the finished spans in, returns accumulated milliseconds and span count of `kind` */
int hop_profile_read(hop_ctx *ctx, int kind, double *total_ms, int64_t *spans);
/* diagnostics: K4's inner solver alone -- the replay of the reference's LM (TransformationEstimationPointToPlane, PCL 1.9, as
 * Utils.cpp:188-229 runs it) on n caller-supplied moment sets.  sums: n x 96 floats, the packed upper triangle of the 13x13 moment
 * matrix (91, row-major) + 5 unused; x_out: n x 6 (translation, quaternion vector part); status_out: Eigen's LevenbergMarquardtSpace
 * status, -1 when the translation is unconstrained (the kernel then ends the hypothesis "not converged"); cycles_out (may be NULL):
 * SM clock cycles of each solve.  Host buffers. */
int hop_debug_lm_solve(hop_ctx *ctx, const float *sums, int n, float *x_out, int32_t *nfev_out, int32_t *status_out, int64_t *cycles_out);
void hop_default_icp_params(hop_icp_params *p);
void hop_default_lcp_params(hop_lcp_params *p);

/* ---- memory helpers for FFI callers that have no CUDA binding of their own ------------------------------- */
int hop_malloc(hop_ctx *ctx, size_t bytes, void **dev_ptr);
int hop_free(hop_ctx *ctx, void *dev_ptr);
int hop_host_alloc(hop_ctx *ctx, size_t bytes, void **host_ptr); /* pinned */
int hop_host_free(hop_ctx *ctx, void *host_ptr);
int hop_memcpy_h2d(hop_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes); /* async on the ctx stream */
int hop_memcpy_d2h(hop_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes); /* async on the ctx stream */

/* ---- clouds ------------------------------------------------------------------------------------------------ */
/* xyz: n x 3, nrm: n x 3 or NULL, prob: n (per-point confidence / LCP weight) or NULL.  Host pointers.
 * The cloud is repacked on the device into two float4 streams (x,y,z,prob) and (nx,ny,nz,1/|n|). */
int hop_cloud_upload(hop_ctx *ctx, const float *xyz, const float *nrm, const float *prob, int n, hop_cloud **out);
/* re-use an existing cloud object for a new frame (same or smaller capacity avoids reallocation) */
int hop_cloud_update(hop_ctx *ctx, hop_cloud *cloud, const float *xyz, const float *nrm, const float *prob, int n);
int hop_cloud_free(hop_ctx *ctx, hop_cloud *cloud);
int hop_cloud_size(const hop_cloud *cloud);
/* Build (or fetch from the cloud's cache) the exact nearest-neighbour grid for queries within `radius`.
 * voxel <= 0 picks the voxel edge automatically.  Called implicitly by the entry points below; exposed so a
 * caller can pay the per-model cost up front.  stats (may be NULL): [0]=voxels [1]=candidate entries
 * [2]=max list length [3]=bytes. */
int hop_cloud_prepare_nn(hop_ctx *ctx, hop_cloud *cloud, float radius, float voxel, int64_t *stats);
/* The same build on a second stream of the context: ordered after everything enqueued so far, concurrent with what is enqueued
 * next; the first entry point that needs the grid waits for it on the device (no host synchronisation).  The scene kd-tree that
 * Utils::computeLCP builds on every call (Utils.cpp:378-379) depends on the frame only, not on the hypotheses:
 * hop_cloud_prepare_nn_async(scene, lcp.dist * 1.01f, 0) right after the upload (PoseEstimator::setCurScene) lets hop_lcp_score's
 * scene grid be built while hop_icp_refine runs.  Results are identical to the implicit build. */
int hop_cloud_prepare_nn_async(hop_ctx *ctx, hop_cloud *cloud, float radius, float voxel);
/* The scene grid hop_lcp_score needs for its reciprocal term, built ahead like hop_cloud_prepare_nn_async(scene, dist * 1.01, ...) with the
 * voxel edge chosen for the number of hypotheses that will be scored against this frame (a batch of thousands repays a finer, slower-to-build
 * grid; a frame's hundred do not).  PoseEstimator::setCurScene (PoseEstimator.cpp:22-60 sets the scene once per frame) is where a shim calls it. */
int hop_lcp_prepare_scene_async(hop_ctx *ctx, hop_cloud *scene, const hop_lcp_params *params, int expected_hypotheses);
/* mark the cloud's cached grids stale (allocations are kept): the next use rebuilds them */
int hop_cloud_drop_nn(hop_ctx *ctx, hop_cloud *cloud);
/* Hint: this cloud's contents stay (the object model, loaded once per PoseEstimator -- PoseEstimator.cpp:22-60).  Grids built for it afterwards
 * may use finer voxels (more memory, a longer build, shorter candidate lists in every later query).  A performance hint only: every query
 * is exact either way. */
int hop_cloud_hint_static(hop_ctx *ctx, hop_cloud *cloud, int is_static);
/* exact 1-NN of host queries (nq x 3) within the prepared radius: idx = -1 when none. (test / debug entry) */
int hop_cloud_nn_query(hop_ctx *ctx, hop_cloud *cloud, float radius, const float *queries, int nq, int32_t *idx, float *d2);

/* ---- K4: ICP refinement of a batch of hypotheses ----------------------------------------------------------- */
/* poses_inout: H x 16 model2scene; replaced by T_icp^-1 * pose (PoseEstimator.cpp:267-269).
 * iters_out / converged_out (H, may be NULL): ICP iterations executed / reg.hasConverged(). */
int hop_icp_refine(hop_ctx *ctx, hop_cloud *scene, hop_cloud *model, float *poses_inout, int H,
                   const hop_icp_params *params, int32_t *iters_out, int32_t *converged_out);
int hop_icp_refine_dev(hop_ctx *ctx, hop_cloud *scene, hop_cloud *model, float *d_poses_inout, int H,
                       const hop_icp_params *params, int32_t *d_iters_out, int32_t *d_converged_out);

/* ---- K5: LCP score of a batch of hypotheses ------------------------------------------------------------------ */
/* scene weights are the scene cloud's prob channel when use_weights != 0, else 1 (selectBest uses 1). */
int hop_lcp_score(hop_ctx *ctx, hop_cloud *scene, hop_cloud *model, const float *poses, int H,
                  const hop_lcp_params *params, int use_weights, float *scores_out);
int hop_lcp_score_dev(hop_ctx *ctx, hop_cloud *scene, hop_cloud *model, const float *d_poses, int H,
                      const hop_lcp_params *params, int use_weights, float *d_scores_out);

/* ---- K3: Super4PCS congruent-set verification ------------------------------------------------------------------- */
/* For every congruent quadrilateral m (4 indices into Q) of trial quad_trial[m] (whose base is bases[4 * trial ..],
 * 4 indices into P): the 3-point rigid transform (MatchBase::ComputeRigidTransformation), the rms < delta gate, the
 * LCP of the sampled model Q against the scene P (Verify) and, for lcp > 0, the hypothesis in the un-centred frame
 * (TryCongruentSet's getGlobalTransform).
 *   P_centered : the scene AFTER MatchBase::init removed its centroid (matchBase.hpp:425-432); centroid_P is that centroid
 *   Q_xyz      : nQ x 3, the sampled model points after centring; centroid_Q likewise
 *   poses/lcp/valid : per quadrilateral (M x 16 column-major, M, M); may be NULL in the host entry
 *   hyp_poses/hyp_lcp/n_hyp : the emitted hypotheses, compacted in (trial, quadrilateral) order = the reference's
 *                single-thread push order (congruentSetExplorationBase.hpp:324-333); capacity M; may be NULL
 * The _dev variant compacts in place (poses/lcp) when d_n_valid is not NULL. */
int hop_verify_lcp(hop_ctx *ctx, hop_cloud *P_centered, const float *Q_xyz, int nQ, const int32_t *bases, int T,
                   const int32_t *quads, const int32_t *quad_trial, int M, const float *centroid_P, const float *centroid_Q,
                   float delta, float *poses, float *lcp, int32_t *valid, float *hyp_poses, float *hyp_lcp, int32_t *n_hyp);
int hop_verify_lcp_dev(hop_ctx *ctx, hop_cloud *P_centered, const float *d_Q, int nQ, const int32_t *d_bases, int T,
                       const int32_t *d_quads, const int32_t *d_quad_trial, int M, const float *centroid_P,
                       const float *centroid_Q, float delta, float *d_poses, float *d_lcp, int32_t *d_valid, int32_t *d_n_valid);

/* ---- Super4PCS global registration: PoseEstimator::runSuper4pcs -> pcl::Super4PCS::align ------------------------------ */
/* the options PoseEstimator::runSuper4pcs sets (PoseEstimator.cpp:66-73; config_autodataset.yaml:133-140) */
typedef struct hop_s4pcs_options {
  int32_t sample_size;            /* super4pcs_sample_size (100): points kept of the model Q */
  float overlap;                  /* super4pcs_overlap (0.2); only configures the (unused) terminate threshold */
  float delta;                    /* super4pcs_delta (0.003) */
  float dispersion;               /* super4pcs_dispersion (0.5): sampling weights are multiplied by it once a point is used */
  int32_t success_quadrilaterals; /* super4pcs_success_quadrilaterals (10) */
  float max_normal_difference;    /* super4pcs_max_normal_difference (-1 = off) */
  float max_color_distance;       /* super4pcs_max_color_distance (-1 = off; colours are not carried on this path) */
  int32_t max_trials;             /* 0 = the reference's effective 30 (congruentSetExplorationBase.hpp:77-100, SURVEY quick facts) */
  uint32_t random_seed;           /* 0 = std::mt19937::default_seed (matchBase.h:103) */
  int32_t keep_intermediates;     /* keep pair sets / quadrilaterals on the plan for inspection (tests) */
} hop_s4pcs_options;
void hop_default_s4pcs_options(hop_s4pcs_options *o);

/* Host-side plan of one registration (no GPU needed): MatchBase::init (matchBase.hpp:382-462: first-hit voxel sampling
 * of Q, std::shuffle, centring, diameter) and the RNG-driven base selection of EVERY trial (SelectQuadrilateral
 * match4pcsBase.hpp:107-189, SelectRandomTriangle matchBase.hpp:111-212, TryQuadrilateral :50-101, computePPF
 * matchBase.hpp:31-68).  Base selection never depends on what earlier trials found, so all trials are planned up front
 * and the device then works on all of them at once.  ppf_keys: n_keys x 4 ints, the keys of the model's PPF table
 * (only membership is ever queried).  P = scene (target), Q = model (source). */
typedef struct hop_s4pcs_plan hop_s4pcs_plan;
int hop_s4pcs_plan_create(const float *P_xyz, const float *P_nrm, const float *P_prob, int nP, const float *Q_xyz,
                          const float *Q_nrm, int nQ, const int32_t *ppf_keys, int n_keys, const hop_s4pcs_options *opt,
                          hop_s4pcs_plan **out);
/* The same plan with the planner's O(N) scans answered by the device: "is the PPF key of this pair of scene points in the table?"
 * is computed for all pairs in one launch (a bit matrix; row by row on demand above 6144 scene points) and the host planner, which
 * alone consumes the random streams, reads bits.  Pairs whose angles sit within 2e-4 degrees of an integer are re-evaluated on the host
 * with the reference's libm, so the plan is bit-identical to hop_s4pcs_plan_create's. */
int hop_s4pcs_plan_create_gpu(hop_ctx *ctx, const float *P_xyz, const float *P_nrm, const float *P_prob, int nP, const float *Q_xyz,
                              const float *Q_nrm, int nQ, const int32_t *ppf_keys, int n_keys, const hop_s4pcs_options *opt,
                              hop_s4pcs_plan **out);
/* The model's PPF table (src/perception/src/app/computePPF.cpp:56-107: gr::computePPF of all point pairs of the 5 mm model): the
 * distinct keys, sorted, as the reference's std::map holds them.  keys_out: capacity x 4 ints (NULL + capacity 0 asks for the
 * count); *n_keys = number of distinct keys.  All-pairs kernel + device sort / unique; boundary pairs re-evaluated on the host. */
int hop_ppf_table_build(hop_ctx *ctx, const float *xyz, const float *nrm, int n, int32_t *keys_out, int capacity, int32_t *n_keys);
void hop_s4pcs_plan_destroy(hop_s4pcs_plan *plan);
/* sizes: [0] nP [1] sampled nQ [2] trials planned [3] pairs kept [4] quadrilaterals kept [5] trials executed */
int hop_s4pcs_plan_sizes(const hop_s4pcs_plan *plan, int32_t *sizes);
/* any pointer may be NULL.  Pc/Qc: centred clouds (n x 3); q_ids: index of each sampled Q point in the input Q;
 * centroids: 6 floats (P then Q); misc: [0] diameter [1] unit-cube ratio; trial_i: T x 5 (base found, base[4] into P);
 * trial_f: T x 4 (invariant1, invariant2, distance1, distance2) */
int hop_s4pcs_plan_get(const hop_s4pcs_plan *plan, float *Pc, float *Qc, int32_t *q_ids, float *centroids, float *misc,
                       int32_t *trial_i, float *trial_f);
/* after hop_super4pcs_run with keep_intermediates: trial_ranges T x 6 (pairs1 begin,end, pairs2 begin,end, quads begin,end),
 * pairs n x 2, quads m x 4 (indices into the sampled Q) */
int hop_s4pcs_plan_intermediates(const hop_s4pcs_plan *plan, int32_t *trial_ranges, int32_t *pairs, int32_t *quads);
/* diagnostics: the planner's replica of std::discrete_distribution<int>(w, w + n) drawing from std::mt19937(seed) (matchBase.hpp:120-140
 * builds one per draw; the planner reproduces the draws without building the table) */
int hop_debug_draw_discrete(const float *w, int n, uint32_t seed, int draws, int32_t *out);
/* gr::computePPF for one pair (exposed for tests): key = 4 ints */
void hop_compute_ppf(const float *p1, const float *n1, const float *p2, const float *n2, int32_t *key);

/* The device part of a planned registration: pair extraction (K2a), congruent-set search (K2b) and verification (K3) of
 * all trials, stopping like Perform_N_steps (:129-194) after success_quadrilaterals successful trials.  hyp_poses
 * (capacity x 16, column-major, model -> scene) / hyp_lcp: every congruent quadrilateral with LCP > 0, in (trial,
 * quadrilateral) order; *n_hyp = how many there are (may exceed capacity: the first `capacity` are written; entries of the
 * buffers past min(*n_hyp, capacity) may be overwritten with zeros). */
int hop_super4pcs_run(hop_ctx *ctx, hop_s4pcs_plan *plan, float *hyp_poses, float *hyp_lcp, int capacity, int32_t *n_hyp);

/* ---- pose clustering (host) ------------------------------------------------------------------------------------- */
/* PoseEstimator::clusterPoses (PoseEstimator.cpp:106-233): sort by (score desc, id asc), keep a hypothesis when no kept one
 * lies within dist_diff AND (all symmetry-folded ZYX Euler differences <= angle_diff OR the geodesic distance <=
 * angle_diff).  poses n x 16 column-major; ids NULL = 0..n-1; symmetry_deg = object_symmetry.<model>.{x,y,z} (0 = free
 * axis, negative = no folding).  keep_out (capacity n): indices of the kept hypotheses in cluster order. */
int hop_cluster_poses(const float *poses, const float *scores, const int32_t *ids, int n, float angle_diff_deg, float dist_diff,
                      const float *symmetry_deg, int32_t *keep_out, int32_t *n_keep);
/* The same decisions with the O(n x clusters) comparisons on the device (sorted blocks of 1024: block vs kept clusters on all
 * SMs, then an exact in-block greedy walk over a shared-memory bit matrix).  Identical keep list; worth it from a few
 * thousand hypotheses up (the host loop needs 0.5 s at 20 k spread-out hypotheses, 13 s at 65 k). */
int hop_cluster_poses_gpu(hop_ctx *ctx, const float *poses, const float *scores, const int32_t *ids, int n, float angle_diff_deg,
                          float dist_diff, const float *symmetry_deg, int32_t *keep_out, int32_t *n_keep);

/* ---- K1: hand-state overlap objective (the function the reference's swarm minimises) ------------------------------- */
#define HOP_MAX_FINGER_BINS 32
/* Everything objFuncPSO reads from optim::ArgPasser / the YAML (Hand.cpp:10-178), flattened.  Matrices column-major. */
typedef struct hop_finger_params {
  float model2handbase[16];    /* args->model2handbase: the finger link in the hand-base frame at joint angle 0 (Hand.cpp:646) */
  float finger_out2parent[16]; /* args->finger_out2parent: the distal link in this link's frame (palm-side fingers, :30, :630) */
  float tip1_local[3];         /* palm side: finger_out_property (min_x,max_y,min_z), else finger_property (min_x,max_y,min_z) (:28,:38) */
  float tip2_local[3];         /* palm side: finger_property (min_x,max_y,min_z), else (min_x,max_y,max_z)                       (:33,:41) */
  float pair_tip1_y, pair_tip2_y; /* args->pair_tip1(1), args->pair_tip2(1): the opposite finger's tips, hand-base frame (:48-54) */
  int32_t palm_side;           /* name == finger_1_1 || finger_2_1 (:26) */
  int32_t right_side;          /* name == finger_2_1 || finger_2_2 (:46) */
  float gripper_min_dist;      /* cfg.gripper_min_dist (0.8 * min bbox extent of the object, main_realdata_auto.cpp:41-45) */
  float dist_thres;            /* hand_match.finger{1,2}_dist_thres */
  float normal_angle_deg;      /* hand_match.finger{1,2}_normal_angle */
  int32_t check_normal;        /* hand_match.check_normal */
  int32_t num_division;        /* FingerProperty bins along z (10), <= HOP_MAX_FINGER_BINS */
  float min_z, stride_z;       /* FingerProperty::_min_z, _stride_z */
  float hist_min_y[HOP_MAX_FINGER_BINS]; /* FingerProperty::_hist_alongz(1, bin): the finger's outer face per z bin */
  int32_t max_outter_pts;      /* hand_match.max_outter_pts (300) */
  float outter_pt_dist;        /* hand_match.outter_pt_dist (0.002) */
  float outter_pt_dist_weight; /* hand_match.outter_pt_dist_weight (1) */
} hop_finger_params;

/* cost_out[s] = objFuncPSO(thetas[s]) for S joint angles (radians) of one finger link -- the dense, data-parallel
 * replacement of the 16-particle x 4-evaluation swarm of Hand::matchOneComponentPSO (Hand.cpp:603-672, pso.hpp:146-351).
 *   finger         : args->model (the link's cloud with normals, its own frame)
 *   scene_hand     : scene_hand_region_removed_noise, hand-base frame (what args->kdtree_scene indexes)
 *   scene_normals  : cloud whose NORMALS are read with the neighbour index (args->scene_hand_region, Hand.cpp:94); NULL =
 *                    scene_hand itself
 *   scene_noswivel : args->scene_remove_swivel, hand-base frame
 *   best_out       : (may be NULL) index of the lowest cost, ties -> lowest index */
int hop_hand_overlap(hop_ctx *ctx, hop_cloud *finger, hop_cloud *scene_hand, hop_cloud *scene_normals, hop_cloud *scene_noswivel,
                     const hop_finger_params *params, const double *thetas, int S, double *cost_out, int32_t *best_out);
/* device pointers: d_thetas S doubles; d_half_cs S x 2 floats = (cosf(theta/2), sinf(theta/2)) of float(theta) -- the two
 * numbers Eigen's AngleAxisf -> Quaternionf conversion produces on the host (pass NULL to have the device compute them);
 * d_cost S doubles; d_best one int32 (may be NULL). */
int hop_hand_overlap_dev(hop_ctx *ctx, hop_cloud *finger, hop_cloud *scene_hand, hop_cloud *scene_normals, hop_cloud *scene_noswivel,
                         const hop_finger_params *params, const double *d_thetas, const float *d_half_cs, int S, double *d_cost,
                         int32_t *d_best);

/* HandT42::adjustHandHeight (Hand.cpp:999-1051; main_realdata_auto.cpp:141): for each trial height h (the reference's 13 offsets
 * -0.03 .. 0.03 m of the hand base along its z), the number of points of hand->_hand_cloud (hand-base frame, shifted by h in z)
 * whose nearest point of the hand-region scene (given in the hand-base frame: _scene_hand_region moved by
 * _handbase_in_cam.inverse()) lies within 5 mm with a normal dot product >= cos 45 deg.  *best_index = the first height with the
 * highest non-zero count, or -1 when nothing matches (the reference then leaves _handbase_in_cam unchanged); the caller applies
 * _handbase_in_cam = _handbase_in_cam * offset(heights[best]).  match_counts (may be NULL): n_heights ints. */
int hop_adjust_hand_height(hop_ctx *ctx, hop_cloud *hand_cloud, hop_cloud *scene_handbase, const float *heights, int n_heights,
                           int32_t *match_counts, int32_t *best_index);

/* ---- per-frame front end: depth image -> object-segment cloud (device) --------------------------------------------- */
/* The pre-processing main_realdata_auto.cpp:54-96,144-181 does with OpenCV / PCL between reading the depth PNG and
 * PoseEstimator::setCurScene, in the same order: back-projection (Utils.cpp:78-115) of the pixels with 0.1 m < z < 2 m,
 * VoxelGrid(leaf_dense) (Utils.cpp:333-340), camera -> hand base, crop box (the three PassThroughs), back to the camera frame,
 * normals over normal_radius flipped to the viewpoint, VoxelGrid(leaf_object) with normals, NaN removal, normals flipped
 * towards the camera origin, confidence 1.  (The hand-point removal of Hand.cpp:781-888 needs the hand meshes, which the
 * reference does not ship: every cropped point keeps confidence 1, as in this repository's main_realdata_auto.) */
typedef struct hop_frame_params {
  float fx, fy, cx, cy;          /* cam_K */
  float leaf_dense;              /* 0.001 */
  float cam_in_handbase[16];     /* column-major: handbase_in_cam.inverse() */
  float handbase_in_cam[16];     /* column-major: cam_in_handbase.inverse() as the host computed it (the reference inverts numerically) */
  float box_min[3], box_max[3];  /* crop in the hand-base frame: x [-0.25,-0.07], y [-0.2,0.2], z [-0.12,0.05] */
  float normal_radius;           /* 0.003 */
  float leaf_object;             /* 0.003 */
  float viewpoint[3];            /* for the first normal flip (the camera origin) */
} hop_frame_params;
void hop_default_frame_params(hop_frame_params *p);
/* depth_mm: height x width uint16 millimetres (the PNG's pixels, host memory).  *scene: NULL to create a cloud, or an existing
 * cloud to refill (its cached NN grids are invalidated).  stage_counts (may be NULL): 5 ints = valid pixels, dense leaves,
 * cropped points, object leaves, final points. */
int hop_frame_to_scene(hop_ctx *ctx, const uint16_t *depth_mm, int width, int height, const hop_frame_params *params, hop_cloud **scene,
                       int32_t *stage_counts);
/* The cloud filters of Hand::setCurScene (Hand.cpp:279-334) on device clouds; each writes a NEW cloud (*out NULL) or refills
 * *out (which must not be `in`); points keep their input order; normals and the weight channel travel with the points.
 *   hop_cloud_voxel_grid                   Utils::downsamplePointCloud = pcl::VoxelGrid (Utils.cpp:333-340): leaf centroids in leaf-index order
 *   hop_cloud_transform                    pcl::transformPointCloudWithNormals, T = 16 floats column-major
 *   hop_cloud_pass_through                 pcl::PassThrough on axis 0 / 1 / 2: keeps lo <= v <= hi
 *   hop_cloud_radius_outlier_removal       pcl::RadiusOutlierRemoval: keeps a point with more than min_neighbors points (itself included)
 *                                          within the radius, d^2 <= float(r^2)
 *   hop_cloud_statistical_outlier_removal  pcl::StatisticalOutlierRemoval: mean distance to the mean_k (<= 64) nearest neighbours,
 *                                          kept when <= mean + stddev_mul * stddev over the cloud (statistics in double, in point order) */
int hop_cloud_voxel_grid(hop_ctx *ctx, const hop_cloud *in, float leaf, hop_cloud **out);
int hop_cloud_transform(hop_ctx *ctx, const hop_cloud *in, const float *T, hop_cloud **out);
int hop_cloud_pass_through(hop_ctx *ctx, const hop_cloud *in, int axis, float lo, float hi, hop_cloud **out);
int hop_cloud_radius_outlier_removal(hop_ctx *ctx, const hop_cloud *in, float radius, int min_neighbors, hop_cloud **out);
/* the per-point filter of Hand::handbaseICP (Hand.cpp:704-728) on a cloud in the hand-base frame: drops the points within 15 mm
 * (in the y-z plane) of either proximal finger axis (y1, z1) = _tf_in_parent["finger_1_1"](1..2,3), (y2, z2) = ...["finger_2_1"],
 * and the points between the two axes in y that lie within 10 mm of z1 */
int hop_cloud_handbase_region(hop_ctx *ctx, const hop_cloud *in, float y1, float z1, float y2, float z2, hop_cloud **out);
int hop_cloud_statistical_outlier_removal(hop_ctx *ctx, const hop_cloud *in, int mean_k, float stddev_mul, hop_cloud **out);

/* HandT42::removeSurroundingPointsAndAssignProbability (Hand.cpp:781-888; main_realdata_auto.cpp:144-148): drops the scene points
 * that belong to the hand and gives the others the confidence 1 - exp(-231.049 * min_dist), min_dist = the smallest exact
 * nearest-neighbour distance to the link clouds visited (start value 1.0), then drops the points on the outer side of either
 * distal link (y < 0 and z >= min_z in the link's frame).  scene: camera frame, with normals.  links: hand->_kdtrees' clouds in
 * the hand-base frame, in std::map order of their names (the reference stops at the first link that claims a point);
 * link_kind[k]: 0 = the dist_thres_sq below, 1 = finger_1_1 / finger_2_1 (5 mm), 2 = base / swivel_1 / swivel_2 (20 mm).
 * *out: NULL to create the cloud (camera frame, confidence in the weight channel), or a cloud to refill; points keep their
 * input order (the reference's order is that of its OpenMP critical section). */
typedef struct hop_hand_removal_params {
  float cam_in_handbase[16];          /* column-major: _handbase_in_cam.inverse() */
  float handbase_in_cam[16];
  float handbase_in_finger_1_2[16];   /* getTFHandBase("finger_1_2").inverse() */
  float handbase_in_finger_2_2[16];
  float min_z;                        /* _finger_properties["finger_1_2"]._min_z */
  float dist_thres_sq;                /* near_hand_dist^2 (main_realdata_auto.cpp:147-148) */
} hop_hand_removal_params;
int hop_remove_hand_points(hop_ctx *ctx, const hop_cloud *scene, const hop_cloud *const *links, const int32_t *link_kind, int n_links,
                           const hop_hand_removal_params *params, hop_cloud **out);
/* The two PCL normal estimators around the hand branch of main_realdata_auto (restated from PCL 1.9.1):
 *   hop_frame_organized  Utils::readDepthImage + convert3dOrganizedRGB + Utils::calNormalIntegralImage(cloud, -1, 0.02, 10, true) +
 *                        PassThrough(z, 0.1, 2.0)  (main_realdata_auto.cpp:54-70; Utils.cpp:294-329): the frame's valid pixels in raster
 *                        order with their integral-image normals (pcl::IntegralImageNormalEstimation, SIMPLE_3D_GRADIENT, depth dependent
 *                        smoothing, viewpoint at the origin; NaN where PCL writes none) -- the reference's scene_organized.  Only the
 *                        intrinsics of `params` are read.
 *   hop_cloud_mls        Utils::calNormalMLS(cloud, radius)  (main_realdata_auto.cpp:160; Utils.cpp:268-292): pcl::MovingLeastSquares of
 *                        order 2 -- the points with at least three neighbours within the radius, PROJECTED onto their fitted surfaces,
 *                        with the surface normals there (not oriented, like PCL); input order, the weight channel travels along. */
int hop_frame_organized(hop_ctx *ctx, const uint16_t *depth_mm, int width, int height, const hop_frame_params *params,
                        float max_depth_change_factor, float normal_smoothing_size, hop_cloud **out);
int hop_cloud_mls(hop_ctx *ctx, const hop_cloud *in, float radius, hop_cloud **out);
/* copies a device cloud back (tests, debugging output such as scene_normals.ply); any pointer may be NULL */
int hop_cloud_download(hop_ctx *ctx, const hop_cloud *cloud, float *xyz, float *nrm, float *prob);

/* ---- physics pruning: signed distance to meshes, rejectByCollisionOrNonTouching ------------------------------------- */
/* A triangle mesh with every normal igl::signed_distance builds for SIGNED_DISTANCE_TYPE_PSEUDONORMAL (face, angle-weighted
 * vertex, uniform edge: src/perception/include/igl/signed_distance.cpp:97-100), resident on the device.  Replaces
 * SDFchecker::registerMesh (src/perception/src/SDFchecker.cpp:36-78): V nv x 3 floats, F nf x 3 vertex indices. */
typedef struct hop_mesh hop_mesh;
int hop_mesh_upload(hop_ctx *ctx, const float *V, int nv, const int32_t *F, int nf, hop_mesh **out);
int hop_mesh_free(hop_ctx *ctx, hop_mesh *mesh);
/* SDFchecker::getSignedDistanceMinMaxWithRegistered (SDFchecker.cpp:115-134) for H placements at once:
 * S[h*n+i] = signed distance (negative inside) of point_transforms[h] * pts[i] to the mesh, I = the closest face (first among
 * exact ties), min_out/max_out[h] = S.minCoeff()/maxCoeff(), n_inside[h] = #(S < 0).  point_transforms: H x 16 column-major
 * (NULL = identity, H placements of the same points) -- the INVERSE of the pose SDFchecker::transformMesh would apply to the
 * mesh.  Any output may be NULL.  n = 0: min = FLT_MAX, max = -FLT_MAX. */
int hop_sdf_query(hop_ctx *ctx, const hop_mesh *mesh, const float *pts, int n, const float *point_transforms, int H, float *S,
                  int32_t *I, float *min_out, float *max_out, int32_t *n_inside);

typedef struct hop_collision_params {
  float cam2handbase[16];  /* column-major: hand->_handbase_in_cam.inverse() (PoseEstimator.cpp:567) */
  float model_center[3];   /* _model_center_init: centroid of model001 (PoseEstimator.cpp:12) */
  float ob_diameter;       /* _ob_diameter (PoseEstimator.cpp:20) */
  float collision_dist;    /* min(-_smallest_dim * collision_thres, -0.007) (:563) */
  float inside_ob_dist;    /* min(-_smallest_dim / 5, -0.01) (:564) */
  float non_touch_dist;    /* cfg non_touch_dist */
  float collision_finger_dist;          /* -cfg collision_finger_dist (:567) */
  float collision_finger_volume_ratio;  /* cfg collision_finger_volume_ratio */
  int32_t finger_status[4];             /* hand->_component_status of finger_1_1, finger_1_2, finger_2_1, finger_2_2 */
} hop_collision_params;
/* PoseEstimator::rejectByCollisionOrNonTouching (PoseEstimator.cpp:524-735) for H hypotheses in one launch, every test in the
 * hand-base frame and in the reference's order.  Arrays of 4 follow std::map order: finger_1_1, finger_1_2, finger_2_1,
 * finger_2_2; finger_clouds[k] = the link's cloud already in the hand-base frame (NULL = link disabled, not in
 * finger_cloud_eigens), finger_meshes[k] = the link's convex mesh registered in the hand-base frame (registerHandMesh, :513-521;
 * NULL = not registered).  scene_without_hand = _cloud_withouthand_raw in the hand-base frame after the 5 mm VoxelGrid (:556-559),
 * hand_cloud = hand->_hand_cloud, model = _model (own frame), poses = H x 16 column-major PoseHypo::_pose (model -> camera).
 * keep[h] = 1 when hypothesis h survives (the caller compacts in order; the reference's survivors come out in OpenMP order);
 * reason[h] (may be NULL): 0 kept, 1 scene point deep inside the object, 2 hand point inside the object, 3 a finger cloud
 * penetrates the object, 4 one whole side does not touch, 5 the object penetrates a finger mesh, 6 the object is inside a finger;
 * diag (may be NULL): H x 10 = signed distance of the scene point, of the hand point, min over each finger cloud (4), min of the
 * model over each finger mesh (4); FLT_MAX where a step was not evaluated. */
int hop_reject_by_collision(hop_ctx *ctx, const hop_mesh *object, const hop_mesh *const *finger_meshes, hop_cloud *const *finger_clouds,
                            hop_cloud *scene_without_hand, hop_cloud *hand_cloud, hop_cloud *model, const float *poses, int H,
                            const hop_collision_params *params, int32_t *keep, int32_t *reason, float *diag);
/* the same on device buffers, stream-ordered on the context's stream (no synchronisation) */
int hop_reject_by_collision_dev(hop_ctx *ctx, const hop_mesh *object, const hop_mesh *const *finger_meshes, hop_cloud *const *finger_clouds,
                                hop_cloud *scene_without_hand, hop_cloud *hand_cloud, hop_cloud *model, const float *d_poses, int H,
                                const hop_collision_params *params, int32_t *d_keep, int32_t *d_reason, float *d_diag);

/* ---- render-based rejection: PoseEstimator::rejectByRender ---------------------------------------------------------- */
/* The OpenGL camera of the reference's depth_sim package (src/depth_sim/src/range_likelihood.cpp:391-475: projection from the
 * intrinsics, CV axes; simulation_io.cpp:411-440: depth buffer -> rounded millimetres, image flipped) as a software rasteriser:
 * image pixel (x, y) samples sx = x + 0.5, sy = y + 0.5 of sx = fx X/Z + cx, sy = fy Y/Z + (height - cy); nearest fragment wins;
 * sim = round(1000 Z) mm / 1000 clamped to [0.1, 2.0]; background = z_far. */
typedef struct hop_render_params {
  float fx, fy, cx, cy;      /* cam_K */
  int32_t width, height;     /* 640 x 480 (PoseEstimator.cpp:349) */
  float z_near, z_far;       /* 0.1, 2.0 (simulation_io.cpp:425-426) */
  float roi_weight;          /* render_roi_weight */
  float keep_ratio;          /* render_keep_hypo */
} hop_render_params;
void hop_default_render_params(hop_render_params *p);
/* Per frame: the real depth image (_depth_meters: height x width floats, metres) and every enabled hand mesh already in the
 * camera frame (hand->_handbase_in_cam * getTFHandBase(name) applied, concatenated; Renderer::addObject, PoseEstimator.cpp:362-383;
 * hand_nf may be 0).  Rasterises the hand once and prepares the per-pixel differences the hypotheses share. */
typedef struct hop_render_scene hop_render_scene;
int hop_render_scene_create(hop_ctx *ctx, const hop_render_params *params, const float *depth_m, const float *hand_V, int hand_nv,
                            const int32_t *hand_F, int hand_nf, hop_render_scene **out);
int hop_render_scene_destroy(hop_ctx *ctx, hop_render_scene *scene);
/* Renderer::doRender for one placement of the object mesh (obj_V nv x 3 in the model frame, obj_F nf x 3, pose = 16 floats
 * column-major model -> camera): depth = height x width simulated depth in metres, mask (may be NULL) = 1 where the object is
 * the nearest surface (the blue pixels of color_sims, PoseEstimator.cpp:425). */
int hop_render_depth(hop_ctx *ctx, const hop_render_scene *scene, const float *obj_V, int obj_nv, const int32_t *obj_F, int obj_nf,
                     const float *pose, float *depth, uint8_t *mask);
/* PoseEstimator::rejectByRender (PoseEstimator.cpp:345-463) for H hypotheses: wrong_ratio[h] = roi_weight * roi_diff / roi_cnt +
 * bg_diff / bg_cnt with the reference's row-major float sums; order (may be NULL, room for H) = the hypotheses to keep in the
 * order the reference's priority queue pops them (ascending wrong ratio; ties and NaN -- an object that owns no pixel -- are
 * unspecified there: lower index first, NaN last), *n_keep = min(max(int(keep_ratio * H), 10), H). */
int hop_reject_by_render(hop_ctx *ctx, const hop_render_scene *scene, const float *obj_V, int obj_nv, const int32_t *obj_F, int obj_nf,
                         const float *poses, int H, float *wrong_ratio, int32_t *order, int32_t *n_keep);

/* ---- winners ------------------------------------------------------------------------------------------------- */
/* top-K by score (ties -> lower id), written as K hop_pose_rec (unused slots: id = -1, score = -inf).
 * d_out may be the send slot of a collective (the all-gather of winners). */
int hop_select_topk_dev(hop_ctx *ctx, const float *d_poses, const float *d_scores, int H, int K, int32_t id_offset,
                        int32_t frame, hop_pose_rec *d_out);
int hop_select_topk(hop_ctx *ctx, const float *poses, const float *scores, int H, int K, int32_t id_offset,
                    int32_t frame, hop_pose_rec *out);

/* ---- the one collective: an all-gather of the per-rank winner records --------------------------------------- */
/* The reference spreads its hypotheses over OpenMP threads (PoseEstimator.cpp:247-258, 474-483); here a frame's batch is
 * partitioned over GPUs -- rank r of G takes hypotheses [r H / G, (r + 1) H / G) -- one hop_ctx per GPU (one process per GPU, or one
 * host thread per context), and the only exchange is ONE ncclAllGather of K x 80 bytes per rank.  Every rank then holds the same
 * world x K records and finishes selectBest / clusterPoses identically.  NCCL is bound at run time (libnccl.so.2).
 *   hop_comm_unique_id : rank 0 creates the 128-byte ncclUniqueId and hands it to the other ranks (any side channel)
 *   hop_comm_init      : ncclCommInitRank on the context's device (collective: every rank calls it)
 *   hop_gather_winners : host buffers, local K records -> all world x K records (rank-major), blocking
 *   hop_gather_winners_dev : device buffers; overlap = 0 stream-ordered on the context's stream; overlap = 1 on the communicator's own
 *                        stream, after what the context's stream has enqueued so far, WITHOUT blocking it (the next frame's
 *                        kernels run while the records travel); hop_gather_wait orders the context's stream (block_host = 0) or
 *                        the host (1) after the gather.  d_send is typically the slot hop_select_topk_dev /
 *                        hop_refine_score_select_dev just wrote. */
#define HOP_COMM_ID_BYTES 128
int hop_comm_unique_id(void *id_out_128);
int hop_comm_init(hop_ctx *ctx, const void *id_128, int rank, int world);
int hop_comm_destroy(hop_ctx *ctx);
int hop_comm_rank(const hop_ctx *ctx, int *rank, int *world);
int hop_gather_winners(hop_ctx *ctx, const hop_pose_rec *local, int K, hop_pose_rec *all);
int hop_gather_winners_dev(hop_ctx *ctx, const hop_pose_rec *d_send, int K, hop_pose_rec *d_recv, int overlap);
int hop_gather_wait(hop_ctx *ctx, int block_host);

/* ---- refineByICP + selectBest in one call ------------------------------------------------------------------------- */
/* main_realdata_auto.cpp:199-204 runs PoseEstimator::refineByICP (PoseEstimator.cpp:235-275) and, after the pruning steps,
 * PoseEstimator::selectBest (:465-502) on the same hypotheses; the LCP score of a hypothesis does not depend on the others, so
 * both can be done in one visit to the device: poses up once, K4 -> K5 -> top-K, results down once, ONE synchronisation (the
 * separate entry points upload the poses twice and synchronise twice).
 *   model_icp : the 5 mm model of refineByICP (_model); model_lcp : the 1 mm model of selectBest (_model001), NULL = model_icp
 *   poses_in  : H x 16 model2scene;  poses_out (may be NULL, may alias poses_in) : the refined poses
 *   scores_out / iters_out / converged_out (H, may be NULL) : as hop_lcp_score / hop_icp_refine
 *   winners   : K records, score descending, ties -> lower index (K = 0: none; K = 1: selectBest's arg-max) */
int hop_refine_score_select(hop_ctx *ctx, hop_cloud *scene, hop_cloud *model_icp, hop_cloud *model_lcp, const float *poses_in, int H,
                            const hop_icp_params *icp, const hop_lcp_params *lcp, int use_weights, int K, float *poses_out,
                            float *scores_out, int32_t *iters_out, int32_t *converged_out, hop_pose_rec *winners);
/* device buffers, stream-ordered, no synchronisation; d_winners may be the send slot of the all-gather of winners */
int hop_refine_score_select_dev(hop_ctx *ctx, hop_cloud *scene, hop_cloud *model_icp, hop_cloud *model_lcp, float *d_poses_inout, int H,
                                const hop_icp_params *icp, const hop_lcp_params *lcp, int use_weights, int K, int32_t id_offset,
                                int32_t frame, int32_t *d_iters, int32_t *d_conv, float *d_scores, hop_pose_rec *d_winners);

#ifdef __cplusplus
}
#endif
#endif /* HOP_C_API_H_ */
