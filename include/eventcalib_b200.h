/* eventcalib_b200 — C ABI of the B200-native EventCalib hot path.
 *
 * Plain C, plain pointers and sizes, no exceptions, no torch / Eigen / OpenCV types.  Every entry point
 * returns ECB_OK (0) or a negative ecb_status; ecb_last_error(ctx) gives the text.  A context owns one
 * CUDA device + stream and all device buffers; distinct contexts may be used from distinct host threads
 * (the reference calls DBSCAN::Run from hardware_concurrency()-2 threads, eventCameraCalib.cpp:181-187).
 * There is NO CPU fallback: without a CUDA device every compute call fails with ECB_ERR_CUDA.
 *
 * Reference interfaces replaced (paths relative to /root/reference/modules/camera_calibration/):
 *   ecb_load_events_*      Event::operator>> (event/include/opengv2/event/Event.hpp:41-47) + the load loop
 *                          (event_camera_calib/test/eventCameraCalib.cpp:154-163)
 *   ecb_frontend_run       EventFrame::EventFrame (event/src/EventFrame.cpp:10-36) +
 *                          CirclesEventFrame::extractFeatures up to findCirclesGrid
 *                          (event_camera_calib/src/CirclesEventFrame.cpp:61-312) incl. fitCircle (:361-415)
 *   ecb_dbscan_run[_batch] DBSCAN<Eigen::Vector2d,double>::Run (dbscan/include/dbscan.h:115-177)
 *   ecb_fit_circles        CirclesEventFrame::fitCircle (CirclesEventFrame.cpp:361-415)
 *   ecb_cost_*             EventCalibSpline::optimize association loop (src/EventCalibSpline.cpp:157-192),
 *                          CalibReprojectionError::operator() (include/.../EventCalibSpline.hpp:168-229) as
 *                          Ceres evaluates it (residual, Jacobian, Huber, quaternion local parameterisation)
 *                          and the normal-equation build that feeds the LM step.
 */
#ifndef EVENTCALIB_B200_H
#define EVENTCALIB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ecb_ctx ecb_ctx;

typedef enum {
    ECB_OK = 0,
    ECB_FAILED = 1,            /* mirrors DBSCAN::ERROR_TYPE::FAILED (dbscan.h:46-48,121-123) */
    ECB_ERR_CUDA = -1,         /* no device / CUDA runtime error */
    ECB_ERR_ARG = -2,          /* bad argument */
    ECB_ERR_UNSUPPORTED = -3,  /* input outside what the device path handles (see ecb_last_error) */
    ECB_ERR_STATE = -4         /* call order (e.g. no events loaded) */
} ecb_status;

/* per-problem status bits reported by the device kernels */
#define ECB_PB_OK 0u
#define ECB_PB_DUPLICATE 2u     /* duplicate pixels inside one front-end problem (cannot happen after the per-pixel dedupe);
                                   ecb_dbscan_run* re-runs such inputs on the general path instead of reporting this */
#define ECB_PB_CLUSTER_CAP 4u   /* more kept clusters than max_clusters: tables truncated, labels still exact */
#define ECB_PB_RANGE 8u         /* pixel outside the sensor / bitmap */

/* ---- context ------------------------------------------------------------------------------------- */
/* stream: a cudaStream_t (as void*) to run on, or NULL for a context-owned stream. */
int ecb_ctx_create(int device, void *stream, ecb_ctx **out);
void ecb_ctx_destroy(ecb_ctx *ctx);
const char *ecb_last_error(const ecb_ctx *ctx);
/* number of kernels of this library launched through ctx so far */
uint64_t ecb_launch_count(const ecb_ctx *ctx);
int ecb_synchronize(ecb_ctx *ctx);
/* Per-kernel device timing (CUDA events on the context stream around every launch; no extra syncs).
 * ecb_stage_ms: durations of the most recent launch of each stage, out[ECB_N_STAGES]; synchronizes. */
#define ECB_STAGE_INGEST 0
#define ECB_STAGE_BOUNDS 1
#define ECB_STAGE_WINDOW 2
#define ECB_STAGE_CLUSTER 3
#define ECB_STAGE_PAIR 4
#define ECB_STAGE_ASSOC 5
#define ECB_STAGE_NORMAL_EQ 6
#define ECB_STAGE_COST 7
#define ECB_STAGE_ORDER 8
#define ECB_STAGE_BFS 9
#define ECB_N_STAGES 10
int ecb_set_profiling(ecb_ctx *ctx, int on);
int ecb_stage_ms(ecb_ctx *ctx, float *out);
const char *ecb_version(void);
/* number of CUDA devices visible to the process (0 without a driver / device): a multi-GPU host creates one context per device */
int ecb_device_count(void);

/* ---- a1: event ingest ---------------------------------------------------------------------------- */
/* Sensor size in pixels (Camera.width / Camera.height of the YAML). Must be set before loading events. */
int ecb_set_sensor(ecb_ctx *ctx, int width, int height);
/* records: n packed 25-byte reference records (f64 t, f64 x, f64 y, u8 polarity), time sorted.
 * _host copies host->device inside the call; _device expects a device pointer (16-byte aligned).
 * The records are unpacked to the SoA layout the kernels use (f64 t[n], u32 x|y<<15|pol<<31).
 * Events with non-integer or out-of-sensor coordinates, or an unsorted stream, make the call fail
 * with ECB_ERR_UNSUPPORTED (the reference would accept them; this path does not, loudly). */
int ecb_load_events_host(ecb_ctx *ctx, const void *records, int64_t n);
int ecb_load_events_device(ecb_ctx *ctx, const void *d_records, int64_t n);
int64_t ecb_num_events(const ecb_ctx *ctx);

/* ---- a2-a5: batched window front end --------------------------------------------------------------- */
typedef struct {
    double dbscan_eps;          /* CirclesEventFrame::Params (CirclesEventFrame.cpp:35-48) */
    uint32_t dbscan_min_pts;    /* dbscan_startMinSample */
    uint32_t cluster_min;       /* clusterMinSample */
    int32_t knn_num;
    int32_t fit_circle;         /* 0: mutual-nearest medians (example.yaml default), 1: fitted circles */
    double radius_threshold;    /* circleRadiusThreshold_ (CirclesEventFrame.cpp:16-33) */
    uint32_t rows_cols;         /* pattern rows*cols (need that many kept clusters per polarity, :127-129) */
    int32_t order_mode;         /* 0: pid = first-arrival order; 1: libstdc++ unordered_set iteration order
                                   (the reference's order, EventFrame.cpp:12-35) */
    uint32_t max_clusters;      /* kept-cluster table capacity per (window,polarity).  0 (recommended): automatic — the
                                   tables grow until every window fits (the reference has no cap,
                                   CirclesEventFrame.cpp:89-117) and the context keeps the capacity for later runs;
                                   > 0: fixed capacity, windows with more kept clusters are truncated and flagged
                                   ECB_PB_CLUSTER_CAP */
    uint32_t median_mode;       /* cluster centre = member with the median norm (CirclesEventFrame.cpp:137-147):
                                   0: slot size/2 of the members sorted by (norm, pid) — order independent;
                                   1: the reference's pick: std::nth_element (libstdc++) over the members in DBSCAN's
                                      BFS pop order — differs from 0 only when several members share the median norm */
} ecb_frontend_params;

typedef struct {
    int64_t ev_lo, ev_hi;       /* event index range of the CLOSED window [t0,t1] */
    int32_t n_points[2];        /* [0]=negative, [1]=positive unique surviving pixels (pid count) */
    int32_t n_clusters[2];      /* raw DBSCAN clusters */
    int32_t n_kept[2];          /* after the clusterMinSample filter */
    int32_t n_candidates;       /* candidate circles (pairs) */
    uint32_t status;            /* ECB_PB_* bits */
    int64_t point_offset[2];    /* offset of this window's points in the flat per-polarity arrays */
} ecb_window_summary;

/* windows: n_win closed intervals [t0,t1] (host doubles, interleaved).  Runs window selection, per-pixel
 * dedupe, +/- cancellation, DBSCAN on both polarities, the cluster-size filter, medians, pairing and
 * circle fit for every window in one batch. Results stay on the device until fetched. */
int ecb_frontend_run(ecb_ctx *ctx, const double *windows, int n_win, const ecb_frontend_params *params);
int ecb_frontend_summary(ecb_ctx *ctx, ecb_window_summary *out, int n_win);
/* total number of point slots per polarity (size of the flat arrays below) */
int64_t ecb_frontend_total_points(ecb_ctx *ctx, int polarity);
/* flat per-polarity arrays: xy (2 doubles per point, pid order per window) and labels (cluster id in
 * discovery order, -1 = Noise) */
int ecb_frontend_points(ecb_ctx *ctx, int polarity, double *xy, int32_t *labels);
/* per window up to max_cand candidates: pi, ni (kept-cluster indices), cx, cy, r  -> out[n_win][max_cand][5] */
int ecb_frontend_candidates(ecb_ctx *ctx, double *out, int max_cand);
/* kept-cluster table of one window/polarity: raw cluster id, size, median pid  (each up to cap entries) */
int ecb_frontend_clusters(ecb_ctx *ctx, int window, int polarity, int32_t *raw_id, int32_t *size,
                          int32_t *median_pid, int cap);
/* a6: CirclesEventFrame::rectifyFeatures (CirclesEventFrame.cpp:417-609) for n_frames windows of the last run, batched over
 * frames x board circles.  image_points[n_frames][n_circles][5][2]: the projected circle centre and its four quadrant
 * points (:431-456; cv::projectPoints stays on the caller's side, it needs the OpenCV initialisation).  Per circle: radius
 * search over the window's points, quadrant-wise +-inlier_threshold band, expansion to whole DBSCAN clusters, circle fit,
 * sanity gates.  out[n_frames][n_circles][3] = rectified cx, cy, r (r < 0: feature deleted, :461,:560,:573);
 * frame_ok[n_frames] (optional) = the frame verdict of :585-609 (edge scores — skipped when fit_circle — and the 20 % rule)
 * for a rows x cols (a)symmetric pattern.  inlier_threshold is 3 px in the reference (:475). */
int ecb_frontend_rectify(ecb_ctx *ctx, const int32_t *window_index, int n_frames, int n_circles,
                         const double *image_points, double inlier_threshold, int rows, int cols, int asymmetric,
                         double *out, int32_t *frame_ok);
/* device pointers of the results of the last run (for callers that keep working on the GPU) */
int ecb_frontend_device_ptrs(ecb_ctx *ctx, void **d_summary, void **d_candidates, int *cand_stride);

/* ---- a3: the DBSCAN::Run boundary ---------------------------------------------------------------- */
/* xy: n x 2 doubles in pid order — ANY finite doubles, duplicates included, any eps (the reference's template takes every
 * T with operator[], dbscan.h:40,70).  Distinct integer pixels with 1 <= eps <= 15 (what the calibration front end
 * produces) run on the sensor-plane bitmap kernel; everything else runs on the general path (grid hash: cell =
 * floor((p - min) / eps), in-tree radix sort by cell key, 3 x 3 cell search, the kd query's strict pruning rule replayed on
 * the emulated insertion tree for the pairs it can affect).  Both give the reference's labels.
 * labels[n]: cluster id in the reference's discovery order, -1 = Noise.
 * Returns ECB_FAILED for n<1 or min_pts<1 exactly like the reference; ECB_ERR_UNSUPPORTED only for NaN / infinite
 * coordinates, NaN eps and problems of 2^21 points or more. */
int ecb_dbscan_run(ecb_ctx *ctx, const double *xy, int n, double eps, uint32_t min_pts, int32_t *labels,
                   int32_t *n_clusters);
/* batch: problem k is xy[offsets[k] .. offsets[k+1]); labels is flat; n_clusters[n_problems]; status
 * (optional) receives ECB_PB_* bits per problem */
int ecb_dbscan_run_batch(ecb_ctx *ctx, const double *xy, const int64_t *offsets, int n_problems, double eps,
                         uint32_t min_pts, int32_t *labels, int32_t *n_clusters, uint32_t *status);

/* Same, plus `Clusters` as the reference builds them (dbscan.h:92,143-162,229-259): cluster c (discovery order) holds
 * cluster_sizes[c] members, listed consecutively in `members` in the reference's order (the seed, then core points in
 * the order expandCluster's FIFO pops them; neighbours enumerated like kd_nearest_range does).  Noise = the pids with
 * label -1, ascending.  cluster_sizes and members need room for n entries. */
int ecb_dbscan_run_ordered(ecb_ctx *ctx, const double *xy, int n, double eps, uint32_t min_pts, int32_t *labels,
                           int32_t *n_clusters, int32_t *cluster_sizes, uint32_t *members);
/* batch form: problem k's sizes start at cluster_sizes[offsets[k]] (n_clusters[k] entries) and its member lists at
 * members[offsets[k]] (one entry per core point) */
int ecb_dbscan_run_batch_ordered(ecb_ctx *ctx, const double *xy, const int64_t *offsets, int n_problems, double eps,
                                 uint32_t min_pts, int32_t *labels, int32_t *n_clusters, uint32_t *status,
                                 int32_t *cluster_sizes, uint32_t *members);

/* DBSCAN<T,Float>::Run(V, dim, eps, min) for dim = 1 .. 4 (dbscan.h:115-177): pts = n x dim doubles.  cluster_sizes and
 * members are optional (both or neither): ordered `Clusters` as above.  dim < 1 returns ECB_FAILED like the reference. */
int ecb_dbscan_run_nd(ecb_ctx *ctx, const double *pts, int n, int dim, double eps, uint32_t min_pts, int32_t *labels,
                      int32_t *n_clusters, int32_t *cluster_sizes, uint32_t *members);

/* ---- a5: batched circle fit ---------------------------------------------------------------------- */
/* set k = points xy[offsets[k]..offsets[k+1]) (union of a + and a - index set); out[k] = cx, cy, r */
int ecb_fit_circles(ecb_ctx *ctx, const double *xy, const int64_t *offsets, int n_sets, double *out);

/* ---- a7-a12: cost evaluation of the dynamic-calibration objective ------------------------------------ */
/* Spline structure (EventCalibSpline ctor, src/EventCalibSpline.cpp:63-91): n_splines segments, n_cp[s] control
 * points each, knots = concatenation of the (n_cp[s] + 4) clamped cubic knot vectors.  Control points are passed to
 * the evaluation calls as flat arrays over all segments: rot_cp 4 doubles each (x,y,z,w), trans_cp 3 doubles each.
 * huber_delta = 0.2 * circle_radius in the reference (:197). */
int ecb_cost_setup(ecb_ctx *ctx, int n_splines, const int32_t *n_cp, const double *knots, double circle_radius,
                   double huber_delta);
/* Rotation model of the residual blocks (the reference's `useSO3` switch, eventCameraCalib.cpp:204-209):
 *   0 (default) CalibReprojectionError      — normalised quaternion B-spline, EigenQuaternionParameterization (hpp:158-250)
 *   1           CalibReprojectionError_SO3 — cumulative SO(3) B-spline R0 * prod exp(beta_j log(R_{j-1}^-1 R_j)) with
 *               LocalParameterizationSO3 (hpp:65-156, BsplineSO3.hpp:190-221); rot_cp are then Sophus::SO3d coefficients (x,y,z,w)
 * Also selects the Plus operation of the host LM step (ecb_lm_options.rotation_model must match). */
int ecb_cost_set_rotation_model(ecb_ctx *ctx, int use_so3);
/* total control points, total span blocks (sum of n_cp-3), residual count, doubles in the packed result */
int ecb_cost_layout(ecb_ctx *ctx, int32_t *total_cp, int32_t *total_spans, int64_t *n_residuals, int64_t *out_doubles);
/* Residual blocks from the loaded events (association loop of optimize(), :157-192, with findCenter,
 * CirclesEventFrame.hpp:50-65): kf_time[K] ascending key-frame stamps, kf_circles[K][n_circles][3] = (cx, cy, r) of
 * each frame's features (r < 0 marks an absent feature), landmarks_xyz[n_circles][3] the board points. */
int ecb_cost_associate(ecb_ctx *ctx, const double *kf_time, const double *kf_circles, int n_keyframes, int n_circles,
                       const double *landmarks_xyz, double motion_time_step, int64_t *n_residuals);
/* The same with the three tables already in device memory (the counterpart of ecb_load_events_device: a caller that keeps the
 * key frames on the GPU across optimisations does not pay their upload per call; the tables are read during this call only). */
int ecb_cost_associate_device(ecb_ctx *ctx, const double *d_kf_time, const double *d_kf_circles, int n_keyframes, int n_circles,
                              const double *d_landmarks_xyz, double motion_time_step, int64_t *n_residuals);
int ecb_cost_get_association(ecb_ctx *ctx, int64_t *event_index, int32_t *circle_id, int64_t cap);
/* ... or explicit residual blocks (host arrays, ordered by (spline, time)) */
int ecb_cost_set_residuals(ecb_ctx *ctx, const double *obs_xy, const double *lm_xyz, const double *t,
                           const int32_t *spline, int64_t n);
/* cost = sum 1/2 rho(r^2)  (Ceres Evaluate without Jacobians; LM step acceptance) */
int ecb_cost_eval(ecb_ctx *ctx, const double *intrinsics, const double *rot_cp, const double *trans_cp, double *cost);
/* Normal equations in the tangent space, packed per knot span s (33 local parameters: intrinsics 9 | rotation
 * tangent of control points s..s+3, 3 each | translation of the same control points, 3 each):
 *   out[s*1122 .. +1089) = J^T J of the span's residual blocks (33x33, full symmetric, row major)
 *   out[s*1122+1089 .. +33) = J^T r ;  out[n_spans*1122] = cost
 * d_out: device buffer to fill (e.g. one a collective will all-reduce), or NULL for the context's own;
 * h_out: optional host copy; cost: optional.  Deterministic (fixed-order reductions). */
int ecb_cost_normal_eq(ecb_ctx *ctx, const double *intrinsics, const double *rot_cp, const double *trans_cp, void *d_out,
                       double *h_out, double *cost);

/* Multi-GPU (one process per GPU): the same evaluation with the inter-GPU sum fused in.  Every rank owns a receive
 * buffer of ecb_exchange_buffer_bytes() bytes, ZERO-FILLED once, in device memory its peers can write (same process:
 * any cudaMalloc pointer with peer access; other processes: ecb_device_alloc + ecb_ipc_export / ecb_ipc_open).
 * recv_buffers[p] = rank p's buffer as seen from this process (own buffer at [rank]).  The span reduction writes this rank's
 * result into its slot of every peer's buffer over NVLink, publishes (epoch, span range, cost), waits for all ranks'
 * slots of this epoch and sums them in rank order into d_out (layout of ecb_cost_normal_eq; bit-identical on all ranks).
 * epoch: 1, 2, 3, ... — the same sequence on every rank; `cost` (optional) synchronises the stream.
 * phases: ECB_EXCHANGE_BOTH normally.  The receive side spins (one warp, bounded) until the peers' slots arrive, so
 * "ranks" that share ONE device must not enqueue it before every rank's send side (CUDA does not promise that two streams
 * of a device run concurrently): issue ECB_EXCHANGE_SEND on all of them first, then ECB_EXCHANGE_RECV. */
#define ECB_MAX_PEERS 16
#define ECB_EXCHANGE_SEND 1
#define ECB_EXCHANGE_RECV 2
#define ECB_EXCHANGE_BOTH 3
size_t ecb_exchange_buffer_bytes(ecb_ctx *ctx, int n_ranks);
int ecb_cost_normal_eq_exchange(ecb_ctx *ctx, const double *intrinsics, const double *rot_cp, const double *trans_cp, int rank,
                                int n_ranks, void *const *recv_buffers, uint64_t epoch, int phases, void *d_out, double *cost);
/* plumbing for the receive buffers: device allocation (zero-filled) that can be exported to other processes of the node */
int ecb_device_alloc(ecb_ctx *ctx, size_t bytes, void **d_ptr);
int ecb_device_free(ecb_ctx *ctx, void *d_ptr);
int ecb_ipc_export(ecb_ctx *ctx, void *d_ptr, void *handle64);          /* 64-byte cudaIpcMemHandle_t */
int ecb_ipc_open(ecb_ctx *ctx, const void *handle64, void **d_ptr);     /* maps a peer's buffer (enables peer access) */
int ecb_ipc_close(ecb_ctx *ctx, void *d_ptr);
/* several GPUs driven from ONE process (one context per device): lets ctx's kernels store into ecb_device_alloc buffers of
 * peer_device (cudaDeviceEnablePeerAccess); receive buffers are then passed to the exchange as plain device pointers */
int ecb_enable_peer_access(ecb_ctx *ctx, int peer_device);

/* ---- a12: the LM step around the GPU normal equations (EventCalibSpline::optimize, src/EventCalibSpline.cpp:197-247) ---- */
typedef struct {
    int32_t max_iterations;        /* 50 (Ceres default; BASELINE config C4) */
    int32_t jacobi_scaling;        /* 1 */
    int32_t fixed_iterations;      /* != 0: convergence tests off, exactly max_iterations LM iterations (benchmark C4) */
    int32_t rotation_model;        /* 0: x+ = dq(delta) * x (EigenQuaternionParameterization); 1: x+ = x * exp(delta) (LocalParameterizationSO3) */
    double function_tolerance;     /* 1e-10 (EventCalibSpline.cpp:239-240) */
    double gradient_tolerance;     /* 1e-10 */
    double parameter_tolerance;    /* 1e-8 */
    double initial_radius, max_radius, min_radius;      /* 1e4, 1e16, 1e-32 */
    double min_relative_decrease;  /* 1e-3 */
    double min_lm_diagonal, max_lm_diagonal;            /* 1e-6, 1e32 */
} ecb_lm_options;

typedef struct {
    int32_t iterations, successful_steps, termination, reserved;
    double initial_cost, final_cost, gradient_max_norm, radius;
} ecb_lm_summary;

#define ECB_LM_RUNNING 0
#define ECB_LM_NO_CONVERGENCE 2       /* max_iterations reached */
#define ECB_LM_FUNCTION_TOLERANCE 3
#define ECB_LM_GRADIENT_TOLERANCE 4
#define ECB_LM_PARAMETER_TOLERANCE 5
#define ECB_LM_MIN_RADIUS 6
#define ECB_LM_FAILURE 7

typedef struct ecb_lm ecb_lm;
void ecb_lm_default_options(ecb_lm_options *o);
/* Host-only LM state machine (no CUDA): the caller evaluates, possibly all-reducing the packed normal equations
 * across GPUs in between.   begin(x, packed) -> { propose(candidate) -> [caller: cost(candidate)] -> feedback(cost)
 *   -> 1: accepted, caller evaluates packed at the candidate -> update(packed) | 0: rejected } until != RUNNING */
ecb_lm *ecb_lm_create(int n_splines, const int32_t *n_cp, const ecb_lm_options *opt);
void ecb_lm_destroy(ecb_lm *lm);
int ecb_lm_dimension(const ecb_lm *lm);
int ecb_lm_begin(ecb_lm *lm, const double *intrinsics, const double *rot_cp, const double *trans_cp, const double *packed);
int ecb_lm_propose(ecb_lm *lm, double *cand_intrinsics, double *cand_rot_cp, double *cand_trans_cp);
int ecb_lm_feedback(ecb_lm *lm, double candidate_cost);
int ecb_lm_update(ecb_lm *lm, const double *packed);
int ecb_lm_state(const ecb_lm *lm, double *intrinsics, double *rot_cp, double *trans_cp, ecb_lm_summary *summary);
/* rows of (cost, gradient_max_norm, radius, accepted) per recorded evaluation; returns the row count */
int ecb_lm_trace(const ecb_lm *lm, double *out, int cap_rows);
/* single-GPU driver of the whole loop; parameters are updated in place */
int ecb_calibrate(ecb_ctx *ctx, int n_splines, const int32_t *n_cp, double *intrinsics, double *rot_cp, double *trans_cp,
                  const ecb_lm_options *opt, ecb_lm_summary *summary, double *trace, int trace_rows);

/* ---- the same loop with the state machine AND the linear solve on the device --------------------------------------------
 * The host only enqueues kernels: per iteration a band-arrow Cholesky of the damped system (one CTA per spline segment, the 9
 * intrinsics and the right-hand side as arrow rows), the candidate's cost, the accept / reject decision and — only if the step
 * was accepted, decided by a device flag — the normal equations at the new point.  Nothing crosses PCIe until the result is
 * read.  Several GPUs (one process per GPU, or several contexts of one process): every rank runs the same replicated state
 * machine on its own share of the residuals; the packed normal equations and the candidate costs are summed over the ranks
 * inside the kernels through the peer-mapped receive buffers of ecb_cost_normal_eq_exchange (sized by
 * ecb_exchange_buffer_bytes, zero-filled once), in rank order, so all ranks take bit-identical decisions.
 * fixed_iterations != 0 runs exactly max_iterations iterations (benchmark config C4): no convergence test, no MIN_RADIUS exit.
 * Call order: create (after ecb_cost_setup + association) -> [set_exchange] -> begin -> iterate(n) ... -> result. */
typedef struct ecb_lm_device ecb_lm_device;
int ecb_lm_device_create(ecb_ctx *ctx, int n_splines, const int32_t *n_cp, const ecb_lm_options *opt, ecb_lm_device **out);
void ecb_lm_device_destroy(ecb_lm_device *lm);
int ecb_lm_device_set_exchange(ecb_lm_device *lm, int rank, int n_ranks, void *const *recv_buffers);
/* uploads x, evaluates the normal equations at x, initialises the trust region (asynchronous) */
int ecb_lm_device_begin(ecb_lm_device *lm, const double *intrinsics, const double *rot_cp, const double *trans_cp);
/* enqueues n_iterations LM iterations (asynchronous; no-ops on the device once the state machine has terminated) */
int ecb_lm_device_iterate(ecb_lm_device *lm, int n_iterations);
/* synchronises: 1 while the state machine is running, 0 when it has terminated, < 0 on error */
int ecb_lm_device_running(ecb_lm_device *lm);
/* synchronises and reads back the parameters, the summary and up to trace_rows rows of (cost, gradient_max_norm, radius,
 * accepted 1 / rejected 0 / invalid -1) */
int ecb_lm_device_result(ecb_lm_device *lm, double *intrinsics, double *rot_cp, double *trans_cp, ecb_lm_summary *summary,
                         double *trace, int trace_rows);
/* begin + iterations until termination + result; parameters are updated in place */
int ecb_calibrate_device(ecb_lm_device *lm, double *intrinsics, double *rot_cp, double *trans_cp, ecb_lm_summary *summary,
                         double *trace, int trace_rows);

#ifdef __cplusplus
}
#endif
#endif /* EVENTCALIB_B200_H */
