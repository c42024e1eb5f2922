/*
 * lvio2d.h — C ABI of the B200-native front-end sliding-window solver.
 *
 * This is the drop-in boundary for ONE hot path of LittleDang/2DLIW-SLAM (lvio_2d):
 * the front-end optimiser `lvio_2d::solver` (reference src/factor/solver.h:28-80) with its
 * factors (src/factor/{laser,imu,wheel,ground,marginalization}_factor.h) and the two
 * preintegrators that produce the factors' constants (src/factor/imu_preintegraption.h,
 * src/factor/wheel_odom_preintegration.h).  The reference has no FFI of its own; the entry
 * points below are what a `lvio_2d::solver` shim binds (see INTEGRATION.md and
 * shim/lvio2d_solver_shim.h).  Every entry point cites the reference code it replaces.
 *
 * Conventions
 *   - plain C, no C++/torch types; all matrices row-major; all reals are IEEE double
 *     (the reference is double everywhere, src/trajectory/trajectory_type.h:23-26).
 *   - per-frame state is 15 doubles in the solver's column order [p(3) q(3) v(3) bs(6)]
 *     (src/factor/solver.cpp:332-342); q is the angle-axis vector of the IMU orientation and must
 *     satisfy |q| <= pi (what lie::normalize_so3 guarantees, src/utilies/common.h:121-135).
 *   - preintegration order is [alpha beta gamma ba bw] (src/factor/factor_common.h:16-20).
 *   - every function returns 0 (LVIO2D_OK) or a negative lvio2d_status; nothing throws across
 *     the ABI.  The caller owns all host buffers; they are copied during the call.
 *   - a context is single-threaded and non-reentrant like the reference's solver, which is only
 *     touched by the dispatch thread (src/trajectory/dispatch.h:104,240).
 *   - the library is CUDA-only: lvio2d_create fails with LVIO2D_ERR_NO_DEVICE when no sm_100
 *     device is present.  There is no CPU fallback.
 */
#ifndef LVIO2D_H_
#define LVIO2D_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LVIO2D_ABI_VERSION 1
#define LVIO2D_STATE_DIM 15
/* imu_preint_result flattened: X[15] | J[15x15] | sqrt_inverse_P[15x15] | Dt
 * (src/factor/imu_preintegraption.h:45-52) */
#define LVIO2D_IMU_BLOB 466
/* wheel_odom_preint_result flattened: delta_Tij[3x4] | diag(sqrt_inverse_P)[3]
 * (src/factor/wheel_odom_preintegration.h:25-33; only the diagonal is read, wheel_factor.h:59-69) */
#define LVIO2D_WHEEL_BLOB 15
/* per-frame laser normal-equation block: 21 upper-triangular entries of the 6x6 J^T J block over
 * (p,q), 6 entries of J^T r, 1 sum of squared residuals */
#define LVIO2D_LASER_BLOCK 28

typedef enum lvio2d_status {
    LVIO2D_OK = 0,
    LVIO2D_ERR_INVALID_ARG = -1,
    LVIO2D_ERR_NO_DEVICE = -2,   /* no CUDA device / not sm_100: the product never falls back to CPU */
    LVIO2D_ERR_CUDA = -3,
    LVIO2D_ERR_NO_WINDOW = -4,   /* solve/linearize called before set_windows */
    LVIO2D_ERR_DOMAIN = -5,      /* |q| > pi, n_frames out of range, too many lines per frame ... */
    LVIO2D_ERR_ALLOC = -6
} lvio2d_status;

/* bits of const_mask[frame]: which parameter blocks ceres::Problem::SetParameterBlockConstant was
 * called on (src/factor/solver.cpp:787-794) */
#define LVIO2D_CONST_P 1u
#define LVIO2D_CONST_Q 2u
#define LVIO2D_CONST_V 4u
#define LVIO2D_CONST_BS 8u

#define LVIO2D_ASSOC_FIXED 0
#define LVIO2D_ASSOC_NEAREST 1

/* termination codes of lvio2d_summary.termination (Ceres TerminationType semantics) */
#define LVIO2D_TERM_NO_CONVERGENCE 0  /* max_iters reached */
#define LVIO2D_TERM_CONVERGENCE_FUNCTION 1
#define LVIO2D_TERM_CONVERGENCE_PARAMETER 2
#define LVIO2D_TERM_CONVERGENCE_GRADIENT 3
#define LVIO2D_TERM_CONVERGENCE_RADIUS 4
#define LVIO2D_TERM_FAILURE 5         /* >= 5 consecutive invalid steps, or NaN */

/* The values of param::manager the hot path reads (src/utilies/params.h:7, params.cpp:92-175).
 * T_imu_to_laser / T_imu_to_wheel are the already re-orthonormalised isometries
 * (lie::normalize_tf, src/utilies/params.cpp:52) as row-major 3x4 [R | t]. */
typedef struct lvio2d_params {
    int32_t abi_version;          /* LVIO2D_ABI_VERSION */
    int32_t device;               /* CUDA device ordinal */
    double T_imu_to_laser[12];
    double T_imu_to_wheel[12];
    double g;                     /* PARAM(g), imu_factor.h:41 */
    double line_to_line_sigma;    /* laser_factor.h:21 */
    double manifold_p_sigma;      /* ground_factor.h:20 */
    double manifold_q_sigma;      /* ground_factor.h:21 */
    double imu_noise_acc_sigma[3];   /* imu_preintegraption.h:26-42 */
    double imu_bias_acc_sigma[3];
    double imu_noise_gyro_sigma[3];
    double imu_bias_gyro_sigma[3];
    double wheel_sigma[3];        /* wheel_odom_preintegration.h:19-22 */
    int32_t max_iters;            /* ceres max_num_iterations: 50 default, 10 in fast_mode (solver.cpp:800-801) */
    int32_t assoc_mode;           /* LVIO2D_ASSOC_FIXED (reference: correspondences frozen before the solve,
                                     trajectory.cpp:210) or LVIO2D_ASSOC_NEAREST (BASELINE config 3) */
    double huber_delta;           /* <= 0 or inf: no robust loss = the reference (solver.cpp:635); otherwise Ceres'
                                     HuberLoss(delta) on every laser point residual */
    /* ceres::Solver::Options defaults the reference never changes; <= 0 selects the default */
    double function_tolerance;    /* 1e-6 */
    double gradient_tolerance;    /* 1e-10 */
    double parameter_tolerance;   /* 1e-8 */
    double initial_trust_region_radius; /* 1e4 */
    /* LVIO2D_ASSOC_NEAREST: at every evaluation a point with point_line >= 0 is matched to the line of its frame's
     * local map with the smallest perpendicular distance among the lines whose extent (plus assoc_gate metres at
     * both ends) contains the point's foot; no such line, or a distance above assoc_max_dist, drops the point. */
    double assoc_gate;            /* <= 0: 0.1 m */
    double assoc_max_dist;        /* <= 0: 0.5 m */
} lvio2d_params;

/* A batch of B independent sliding windows with the same number of frames.  One window is what
 * solver::solve / solver::init_solve receive as `std::deque<frame_info::ptr>` (solver.h:72-76),
 * flattened to SoA.  Laser terms are "points against lines": residual w*dis_from_line(C, A1, A2)
 * (src/utilies/common.h:86-95) with C the point moved by the frame's pose and (A1,A2) the line moved
 * by its reference pose — the reference's laser_factor (laser_factor.h:45-89) is the special case
 * of 2 points (the segment end points) per matched line pair with the pair weight
 * (laser_factor.h:38-42); "beam mode" feeds every scan point.
 *
 * Index f = window*n_frames + frame.  All pointers are host pointers for lvio2d_set_windows and
 * device pointers for lvio2d_bind_windows. */
typedef struct lvio2d_window_batch {
    int32_t n_windows;            /* B >= 1 */
    int32_t n_frames;             /* n >= 1, frames per window */
    const double* states;         /* [B*n][15]  p q v bs, read at set time, solution via lvio2d_get_states */
    const uint8_t* const_mask;    /* [B*n]  LVIO2D_CONST_* bits */
    /* laser */
    const int64_t* point_offset;  /* [B*n+1] prefix offsets into points / point_line / point_weight */
    const double* points;         /* [N][2]  x,y in the frame's own laser frame (z = 0, common.cpp:22-24) */
    const int32_t* point_line;    /* [N]  index into the frame's line list, < 0 = no correspondence */
    const double* point_weight;   /* [N] or NULL: per-point weight multiplying 1/line_to_line_sigma
                                     (laser_factor::sum, laser_factor.h:40-42); NULL = 1 */
    const int64_t* line_offset;   /* [B*n+1] prefix offsets into lines */
    const double* lines;          /* [L][4]  p1.x p1.y p2.x p2.y in the REFERENCE laser frame */
    const int32_t* ref_frame;     /* [B*n]  -1: lines live under the external pose ref_pose[f], constant
                                     (tracking, solver.cpp:685-691);  k>=0: under frame k of this window,
                                     a free block (initialisation, solver.cpp:93-106) */
    const double* ref_pose;       /* [B*n][6]  p1,q1 of laser_match (laser_type.h:83), used when ref_frame<0 */
    /* imu / wheel factors between frame i-1 and i, i = 1..n-1 */
    const double* imu;            /* [B*(n-1)][LVIO2D_IMU_BLOB] */
    const double* wheel;          /* [B*(n-1)][LVIO2D_WHEEL_BLOB] */
    /* ground factors: the reference adds both factors `n` times per frame (solver.cpp:727-743) */
    int32_t ground_multiplicity;  /* reference behaviour: n_frames */
    /* marginalisation prior r = J (X - X0) on frame prior_frame (solver.cpp:744-785,
     * marginalization_factor.h:22-53).  prior_frame < 0: no prior */
    int32_t prior_frame;          /* reference: n_frames - 2 */
    const double* prior_X0;       /* [B][15] */
    const double* prior_J;        /* [B][15x15] */
} lvio2d_window_batch;

typedef struct lvio2d_summary {
    int32_t iterations;           /* LM iterations executed (Ceres iteration counter, excludes iteration 0) */
    int32_t termination;          /* LVIO2D_TERM_* */
    int32_t num_successful_steps;
    int32_t num_unsuccessful_steps;
    double initial_cost;          /* 1/2 sum r^2 over active residual blocks */
    double final_cost;
    double final_radius;
    double reserved;
} lvio2d_summary;

typedef struct lvio2d_ctx lvio2d_ctx;

/* ---- lifetime: replaces solver::solver() (src/factor/solver.cpp:43-48) and the PARAM()/noise
 * singletons the functors read on every call (laser_factor.h:12-24, ground_factor.h:11-22) ---- */
int lvio2d_create(lvio2d_ctx** out, const lvio2d_params* params);
void lvio2d_destroy(lvio2d_ctx* ctx);
const char* lvio2d_strerror(int status);
/* last CUDA error string recorded by the context (empty if none) */
const char* lvio2d_last_error(const lvio2d_ctx* ctx);
/* the CUDA stream (cudaStream_t) all work of this context is enqueued on */
void* lvio2d_stream(lvio2d_ctx* ctx);

/* ---- problem upload: replaces ceres::Problem assembly, solver.cpp:636-794 / :50-159 ---- */
int lvio2d_set_windows(lvio2d_ctx* ctx, const lvio2d_window_batch* host_batch);
/* zero-copy variant: every pointer of the batch is a device pointer that stays valid until the next
 * set/bind; states are copied into context-owned buffers (they are mutated by solve) */
int lvio2d_bind_windows(lvio2d_ctx* ctx, const lvio2d_window_batch* device_batch);
/* asynchronous flavours for pipelining several contexts (host buffers should be pinned): the copies are only
 * enqueued on lvio2d_stream(); the caller's buffers must stay valid and unchanged until lvio2d_sync() returns */
int lvio2d_set_windows_async(lvio2d_ctx* ctx, const lvio2d_window_batch* host_batch);
int lvio2d_get_states_async(lvio2d_ctx* ctx, double* host_states /* [B*n][15] */);
/* ceres::Solver::Options::max_num_iterations of the following solves.  The reference sets it per call: solver::solve
 * lowers it to 10 in fast_mode (src/factor/solver.cpp:800-801), solver::do_init_solve never does (:161-168), so an
 * initialisation runs Ceres' default 50 iterations whatever fast_mode says.  max_iters <= 0 selects 50. */
int lvio2d_set_max_iterations(lvio2d_ctx* ctx, int32_t max_iters);
/* Compact wire encoding of the laser input of a beam-mode batch: what the sensor delivers (one sensor_msgs/LaserScan per
 * frame, float32) plus a 16-bit line index per beam, instead of 16-byte points and 32-bit indices — 6 instead of 20 bytes
 * per beam across PCIe.  The device rebuilds the points like convert::laser_to_point_times does (src/utilies/common.cpp:6-24:
 * float32 angle_min + k * angle_increment, cos / sin in double, times the float32 range) without its 1 cm thinning;
 * beams the reference rejects (NaN, inf, <= 0.1 m) and beams with index 0xFFFF take no part.  Every frame has n_beams
 * point slots.  In `host_batch` points / point_line / point_weight / point_offset are ignored; lines, line_offset,
 * ref_frame, ref_pose and everything else are as for lvio2d_set_windows. */
typedef struct lvio2d_scan_wire {
    int32_t n_beams;
    const float* ranges;          /* [B*n][n_beams] */
    const float* angle;           /* [B*n][2]  angle_min, angle_increment */
    const uint16_t* beam_line;    /* [B*n][n_beams]  index into the frame's line list, 0xFFFF = no correspondence */
    /* optional (NULL: host_batch->imu is used): the 190 doubles of every imu_preint_result that imu_factor::operator()
     * reads (src/factor/imu_factor.h:52-86) instead of the 466 of the full blob:
     *   X[15] | J(k, 9..14), k = 0..8, row-major [9][6] | the upper triangle of sqrt_inverse_P row by row (120) | Dt */
    const double* imu_compact;    /* [B*(n-1)][LVIO2D_IMU_COMPACT] */
    /* 1: the frames of a window were matched against the SAME sub-map (laser_match::lines1 of consecutive frames are
     * shared_ptrs into one laser_submap, laser_manager.cpp:544-545) and its lines travel once per window:
     * host_batch->line_offset has n_windows + 1 entries and every frame of window w uses the lines
     * [line_offset[w], line_offset[w + 1]); beam_line indexes that list.  The device replicates them into the per-frame
     * layout.  0: line_offset has n_windows * n_frames + 1 entries as for lvio2d_set_windows. */
    int32_t shared_lines;
    int32_t reserved;
    /* optional, replaces beam_line when not NULL: the same indices in 8 bits (0xFF = no correspondence), for local maps of
     * at most 255 lines — 5 instead of 6 bytes per beam */
    const uint8_t* beam_line8;    /* [B*n][n_beams] */
} lvio2d_scan_wire;
#define LVIO2D_IMU_COMPACT 190
int lvio2d_set_windows_wire(lvio2d_ctx* ctx, const lvio2d_window_batch* host_batch, const lvio2d_scan_wire* wire, int32_t async);
/* overwrite the states only (same shapes): re-arm a bound batch for another solve */
int lvio2d_reset_states(lvio2d_ctx* ctx, const double* host_states);

/* ---- solver::solve / solver::do_init_solve: ceres::Solve with the reference's options
 * (solver.cpp:795-802, :161-168): trust-region Levenberg–Marquardt, Jacobi scaling, exact step ---- */
int lvio2d_solve(lvio2d_ctx* ctx, lvio2d_summary* summaries /* [B] host, may be NULL */);
/* same, but only enqueues on lvio2d_stream(); summaries and states are read later */
int lvio2d_solve_async(lvio2d_ctx* ctx);
int lvio2d_sync(lvio2d_ctx* ctx);
int lvio2d_get_summaries(lvio2d_ctx* ctx, lvio2d_summary* summaries /* [B] host */);
/* frame_infos[i]->{p,q,v,bs} after the solve (written in place by Ceres, solver.cpp:685-707) */
int lvio2d_get_states(lvio2d_ctx* ctx, double* host_states /* [B*n][15] */);

/* ---- split-phase solve for point-sharded multi-GPU runs (one allreduce per LM iteration).
 * Each rank owns points [rank*N_f/world, (rank+1)*N_f/world) of every frame.  Loop:
 *   lvio2d_solve_begin; repeat { lvio2d_eval_laser; allreduce(sum) of the reduce buffer;
 *   lvio2d_lm_step(&active) } until active == 0. ---- */
int lvio2d_set_point_shard(lvio2d_ctx* ctx, int32_t rank, int32_t world);
int lvio2d_solve_begin(lvio2d_ctx* ctx);
int lvio2d_eval_laser(lvio2d_ctx* ctx);
/* device pointer + element count (doubles) of the per-frame laser blocks to be summed over ranks;
 * the caller may instead hand in its own device buffer (e.g. a torch tensor) with set_reduce_buffer */
int lvio2d_reduce_buffer(lvio2d_ctx* ctx, void** device_ptr, int64_t* count);
int lvio2d_set_reduce_buffer(lvio2d_ctx* ctx, void* device_ptr, int64_t count);
int lvio2d_lm_step(lvio2d_ctx* ctx, int32_t* n_active /* host out, may be NULL (no sync) */);

/* ---- measurement hooks (bench.py): with profiling on, every scan-match and window-step launch is bracketed by
 * CUDA events on lvio2d_stream().  lvio2d_get_profile syncs and returns, accumulated since profiling was last
 * switched on: out[0] scan-match ms, out[1] scan-match launches, out[2] window-step ms, out[3] window-step launches,
 * out[4] all kernel launches of the context, out[5] scan-match algorithmic bytes (points/lines/tables read + blocks
 * written by the frames that were actually processed is not tracked on the device; this is the per-launch figure
 * for all active frames x launches), out[6] factor-kernel ms, out[7] factor-kernel launches. ---- */
int lvio2d_set_profiling(lvio2d_ctx* ctx, int32_t on);
int lvio2d_get_profile(lvio2d_ctx* ctx, double* out /* [8] */);
/* fp64 roofline denominators measured on this device now (no reference counterpart; BASELINE.md §1 asks the build to
 * measure one): out[0] vector-pipe DFMA TFLOP/s (16 independent chains per thread, all SMs), out[1] tensor-pipe DMMA
 * (mma.sync.m8n8k4.f64) TFLOP/s, out[2], out[3] the two kernel durations in ms. */
int lvio2d_measure_fp64_peak(lvio2d_ctx* ctx, double* out /* [4] */);

/* ---- one-shot linearisation (test hook + what solver::marginalization builds, solver.cpp:367-380).
 * mode 0: the solver's reduced program (constant blocks have zero rows/cols, inactive residual
 *         blocks dropped); mode 1: the marginalisation program of solver.cpp:257-380 (no constant
 *         blocks, laser Jacobian w.r.t. the frame's own pose only).
 * H is dense [B][15n][15n], g = J^T r [B][15n], cost = 1/2 sum r^2 [B]; host buffers. ---- */
int lvio2d_linearize(lvio2d_ctx* ctx, int32_t mode, double* H, double* g, double* cost);

/* ---- solver::marginalization (solver.cpp:257-442 with marginalization_matrix :4-40): Schur
 * complement onto the last frame, eigen-decomposition, sqrt-information prior.
 * X0 [B][15], J_lin [B][15x15], r_lin [B][15], host buffers. ---- */
int lvio2d_marginalize(lvio2d_ctx* ctx, double* X0, double* J_lin, double* r_lin);

/* ---- preintegration (the factors' constants) ----
 * imu_preintegraption::{reset_imu_measure, update, get_preintegraption_result}
 * (imu_preintegraption.h:113-124, :170-208, :147-152), one interval per work item.
 * samples: [M][7] = dt, acc(3), gyro(3): the step `update(dt)` integrates with the PREVIOUS
 * sample's acc/gyro (last_info, :179-180), so row k holds (dt_k, acc_{k-1}, gyro_{k-1}).
 * bias0: [n_intervals][6] (ba,bw) at reset.  out: [n_intervals][LVIO2D_IMU_BLOB]. */
int lvio2d_imu_preintegrate(lvio2d_ctx* ctx, int32_t n_intervals, const int64_t* sample_offset,
                            const double* samples, const double* bias0, double* out_blobs);
/* wheel_odom_preintegration::{update_by_v, get_preintegraption_result}
 * (wheel_odom_preintegration.h:141-152, :111-125).  steps: [M][7] = dt, v(3), omega(3).
 * out: [n_intervals][LVIO2D_WHEEL_BLOB]. */
int lvio2d_wheel_preintegrate(lvio2d_ctx* ctx, int32_t n_intervals, const int64_t* step_offset,
                              const double* steps, double* out_blobs);

/* ---- per-factor evaluation on the device = auto_diff::compute_res_and_jacobi
 * (src/utilies/common.h:201-217): residual + row-major Jacobian per parameter block, concatenated
 * column-wise in the functor's parameter order.  Host buffers. ---- */
/* laser_factor(l1_p1,l1_p2,l2_p1,l2_p2)(p_i,q_i,p_j,q_j): res[2], jac[2][12] (laser_factor.h:31-89);
 * each l*_p* is 3 doubles */
int lvio2d_eval_laser_factor(lvio2d_ctx* ctx, const double* l1_p1, const double* l1_p2,
                             const double* l2_p1, const double* l2_p2, const double* pose_i,
                             const double* pose_j, double* res, double* jac);
/* imu_factor(blob)(p_i,q_i,v_i,bs_i,p_j,q_j,v_j,bs_j): res[15], jac[15][30] (imu_factor.h:13-89) */
int lvio2d_eval_imu_factor(lvio2d_ctx* ctx, const double* imu_blob, const double* state_i,
                           const double* state_j, double* res, double* jac);
/* wheel_odom_factor(blob)(p_i,q_i,p_j,q_j): res[3], jac[3][12] (wheel_factor.h:12-73) */
int lvio2d_eval_wheel_factor(lvio2d_ctx* ctx, const double* wheel_blob, const double* pose_i,
                             const double* pose_j, double* res, double* jac);
/* ground_factor_p / ground_factor_q (p,q): res[2] = (res_p, res_q), jac[2][6] (ground_factor.h:27-82) */
int lvio2d_eval_ground_factors(lvio2d_ctx* ctx, const double* pose, double* res, double* jac);

/* ---- laser front-end, step 1: scan points -> line segments.
 * Replaces, for a batch of scans, laser_manager::spawn_scan (src/trajectory/laser_manager.cpp:350-422) with
 * scan::add_line's fit and filters (:137-154; fit_line_by_least_square :19-36, create_line :62-94): continuity split,
 * corner response cos(angle) with step 3, non-maximum suppression, tolerance-angle merge, least-squares fit
 * (smallest right singular vector of [x y 1]), max-distance / min-length filters, grid-validity test of
 * scan::xy_to_index (src/trajectory/laser_type.h:34-41).  The output is scan::lines in the reference's order. ---- */
typedef struct lvio2d_line_params {
    double line_continuous_threshold;     /* config/corridor.yaml:84 */
    double line_max_tolerance_angle_deg;  /* :89 (degrees, converted like convert::angle_to_rad) */
    double line_max_dis;                  /* :86 */
    double line_min_len;                  /* :85 */
    double laser_resolution;              /* :81 */
    double w_laser_each_scan;             /* :79 */
    double h_laser_each_scan;             /* :80 */
} lvio2d_line_params;
/* points: [N][2] (x, y) in the laser frame, scan s owns rows point_offset[s] .. point_offset[s+1]-1 (the de-skewed
 * points of sensor::laser, z = 0).  Outputs per scan, `max_lines` slots each: n_lines[s] (lines found; when it
 * exceeds max_lines only the first max_lines are stored), lines [S][max_lines][4] = (p1.x, p1.y, p2.x, p2.y),
 * abc [S][max_lines][3] (the fitted (a, b, c), sign normalised so that the component of largest magnitude is
 * positive), index_range [S][max_lines][2] = (index1, index2) into the scan's points.
 * on_device = 0: every pointer is a host buffer (copied in and out, synchronous).  on_device = 1: every pointer is a
 * device buffer; the kernel is enqueued on the context's stream and the call returns without synchronising. */
int lvio2d_extract_lines(lvio2d_ctx* ctx, const lvio2d_line_params* lp, int32_t n_scans, const int64_t* point_offset,
                         const int32_t* point_count /* [S] or NULL: scan s has point_count[s] points from point_offset[s] */,
                         const double* points, const double* point_z /* [N] or NULL (z = 0) */, int32_t max_lines,
                         int32_t* n_lines, double* lines, double* abc, int32_t* index_range, int32_t on_device);

/* ---- laser front-end, step 0: LaserScan ranges -> (de-skewed) points.
 * Replaces convert::laser_to_point_times (src/utilies/common.cpp:4-40: polar -> Cartesian in the laser frame, readings
 * that are NaN / inf / <= 0.1 m dropped, a point closer than 0.01 m to the last KEPT point dropped) followed by
 * sensor::laser::correct (src/trajectory/sensor.h:51-94: p <- make_tf(dt*linear, dt*angular) * p with
 * dt = time_of_beam - stamp), which trajectory::add_laser calls before spawn_scan (trajectory.cpp:147, :195). ---- */
typedef struct lvio2d_scan_header {
    float angle_min, angle_increment, time_increment, reserved;  /* sensor_msgs/LaserScan fields, float32 on the wire */
    double stamp;        /* header.stamp.toSec() */
    double linear[3];    /* correct(linear, angular): velocity of the laser frame expressed in the laser frame */
    double angular[3];
} lvio2d_scan_header;
/* ranges: [S][n_beams] float32.  Outputs use a fixed stride of n_beams slots per scan: point_count[s] points at
 * points[s*n_beams ...][2] (x, y), point_z[s*n_beams ...] (0 unless de-skewed) and, when not NULL,
 * point_time[s*n_beams ...] (the beam time stamps, times_ptr of sensor::laser).  deskew = 0 skips correct().
 * The fixed-stride layout feeds lvio2d_extract_lines directly (point_offset[s] = s*n_beams, point_count).
 * on_device as in lvio2d_extract_lines. */
int lvio2d_scan_to_points(lvio2d_ctx* ctx, int32_t n_scans, int32_t n_beams, const float* ranges,
                          const lvio2d_scan_header* headers, int32_t deskew, int32_t* point_count, double* points,
                          double* point_z, double* point_time, int32_t on_device);

/* ---- laser front-end, step 2: segment association between a reference scan and the
 * current scan.  Replaces laser_manager::do_match (src/trajectory/laser_manager.cpp:244-348) for a batch of scan
 * pairs: every line of scan 2 is taken to scan 1's frame (T_1_2 from the two IMU poses and T_imu_to_laser), the lines
 * rasterised into the (2*(1+kk)+1)^2 cell neighbourhood of its mid point in scan 1's line_map are the candidates, the
 * candidate of smallest direction angle wins (first one on ties, in the reference's cell / insertion order), pairs
 * above 10 degrees are dropped, then pairs whose mean end-point distance is >= 1.2 x the average are dropped.
 * scan 1's line_map is never materialised: a line covers a cell when one of its own points (points1 / index_range1,
 * the spawn_scan flavour of scan::add_line, :155-191) or, with points1 == NULL, one of the samples taken every 0.05 m
 * along the segment (the sub-map flavour add_line(p1, p2, false), :192-212) falls into it.
 * Layouts as produced by lvio2d_extract_lines: lines* [P][max_lines*][4], n_lines* [P], index_range1 [P][max_lines1][2];
 * scan 1's points as [point_offset1[p] .. +point_count1[p]) (point_count1 == NULL: offsets are [P+1]).
 * pose1 / pose2: [P][6] (p, q) of the IMU.  Outputs: n_match [P], match [P][max_lines2][2] = (line of scan 1, line of
 * scan 2) in the reference's order.  on_device as above. */
int lvio2d_match_lines(lvio2d_ctx* ctx, const lvio2d_line_params* lp, int32_t n_pairs, int32_t kk,
                       const int64_t* point_offset1, const int32_t* point_count1, const double* points1,
                       int32_t max_lines1, const int32_t* n_lines1, const double* lines1, const int32_t* index_range1,
                       int32_t max_lines2, const int32_t* n_lines2, const double* lines2,
                       const double* pose1, const double* pose2, int32_t* n_match, int32_t* match, int32_t on_device);

/* ---- laser front-end, step 3: the reference sub-map, resident in device memory.
 * Replaces the state of laser_manager::add_scan (src/trajectory/laser_manager.cpp:424-496: ref_submap_ptr,
 * spawnning_ref_submap_ptr, last_add_tf, current_count) and laser_manager::match_with_ref (:531-546) for a batch of
 * `n_managers` independent laser managers (one per robot / stream).  add_scan applies the motion filter
 * (ref_motion_filter_p / ref_motion_filter_q, config/corridor.yaml:120-121), founds the reference sub-map with the
 * first scan, takes every later scan's lines to the frame of the reference sub-map and of the one being spawned
 * (T_il^-1 (make_tf(sub)^-1 make_tf(current)) T_il, Isometry inverses) and appends them through the filters of
 * scan::add_line(p1, p2, false) (:215-223: |z| of the transformed end points against line_max_dis, length against
 * line_min_len, one 0.05 m sample on the grid), founds the spawning sub-map after ref_n_accumulation / 2 accepted
 * scans and hands it over at ref_n_accumulation (:472-489).  Lines are kept in the reference's append order, so
 * lvio2d_submap_match returns the pairs laser_manager::do_match would.
 * line_cap: slots per sub-map (a corridor.yaml run appends up to ~100 lines per scan over ref_n_accumulation = 100
 * scans); lines beyond it are counted in n_lines but not stored. ---- */
typedef struct lvio2d_submap lvio2d_submap;
int lvio2d_submap_create(lvio2d_ctx* ctx, int32_t n_managers, int32_t line_cap, const lvio2d_line_params* lp,
                         double ref_motion_filter_p, double ref_motion_filter_q, int32_t ref_n_accumulation,
                         lvio2d_submap** out);
/* (safe after lvio2d_destroy of its context; every other lvio2d_submap_* call needs the context alive) */
void lvio2d_submap_destroy(lvio2d_submap* sm);
/* back to "no scan seen" for every manager */
int lvio2d_submap_reset(lvio2d_submap* sm);
/* One scan per manager: n_lines [M] (a negative count = no scan for this manager in this call), lines
 * [M][max_lines][4] as produced by lvio2d_extract_lines, pose [M][6] = (current_p, current_q) of the IMU.
 * on_device = 0: host buffers, synchronous.  on_device = 1: device buffers, enqueued on the context's stream. */
int lvio2d_submap_add_scan(lvio2d_submap* sm, int32_t max_lines, const int32_t* n_lines, const double* lines,
                           const double* pose, int32_t on_device);
/* Host copy of the state (tests, visualisation): which = 0 reference sub-map, 1 the one being spawned.
 * meta [M][4] = (has reference, has spawning, current_count, last add_scan passed the motion filter), pose [M][6],
 * n_lines [M], lines [M][line_cap][4]; any output may be NULL. */
int lvio2d_submap_get(lvio2d_submap* sm, int32_t which, int32_t* meta, double* pose, int32_t* n_lines, double* lines);
/* Device pointers of the state, for chaining with lvio2d_match_lines(on_device = 1) / lvio2d_set_windows(bind):
 * meta [M][4], ref_pose [2][M][6], ref_n_lines [2][M], ref_lines [2][M][line_cap][4] (slot 0 = reference sub-map). */
int lvio2d_submap_device(lvio2d_submap* sm, const int32_t** meta, const double** ref_pose, const int32_t** ref_n_lines,
                         const double** ref_lines);
/* laser_manager::match_with_ref for every manager: the current scans (n_lines2 [M], lines2 [M][max_lines2][4], pose2
 * [M][6]) against the resident reference sub-maps.  n_match [M], match [M][max_lines2][2] = (line of the sub-map, line
 * of the scan); a manager without a reference sub-map returns no pairs.  Optional outputs (NULL to skip), which make a
 * copy of the sub-map unnecessary: matched_lines1 [M][max_lines2][4] = end points of the sub-map line of every pair
 * (laser_match::lines1), ref_pose [M][6] = the sub-map's pose (laser_match::p1, q1).  on_device as above. */
int lvio2d_submap_match(lvio2d_submap* sm, int32_t kk, int32_t max_lines2, const int32_t* n_lines2, const double* lines2,
                        const double* pose2, int32_t* n_match, int32_t* match, double* matched_lines1, double* ref_pose,
                        int32_t on_device);

/* ---- back-end pose graph.  Replaces keyframe_manager::solve
 * (src/trajectory/keyframe_manager.cpp:722-838): one ceres::Problem over the key-frame poses (p, q) with an edge_factor
 * (src/factor/edge_factor.h:79-126) per sequential and per loop edge, ground_factor_p / ground_factor_q on every key
 * frame (use_ground_p_factor / use_ground_q_factor), the first edge's index1 held constant (:744-748), ceres::Solve
 * with default options (LM, exact linear solve; the iteration cap and tolerances are the context's lvio2d_params —
 * give the back-end its own context, it runs on its own thread in the reference, keyframe_manager.cpp:91).
 * poses: [n_poses][6] (p, q angle-axis), in/out, host memory.  edge_index: [n_edges][2] = (index1, index2), any order;
 * edges between neighbours (|index1 - index2| = 1) form the block-tridiagonal part, all others (loop_edges) enter as
 * rank-6 updates.  edge_tf: [n_edges][12] tf12 row-major 3x4.  edge_weight: [n_edges] (1 for seq_edges, loop_edge_k
 * for loop_edges, :734, :780).  sqrt_info: edge_noise::J row-major 6x6 (edge_factor.h:14-26; the caller builds it,
 * including the reference's J(1,2) slip if it wants the reference's numbers).  fixed_pose: the key frame held constant, -1
 * for none — the reference calls SetParameterBlockConstant on seq_edges[0]->index1 inside its seq_edges loop only
 * (:744-748), so a graph without sequential edges has none.  Ceres drops the residual blocks whose parameter blocks are
 * all constant from the reduced program: the ground factors of the fixed key frame count as fixed cost, not in the
 * function-tolerance test.  LVIO2D_ERR_INVALID_ARG for an index out of range or a self edge, LVIO2D_ERR_DOMAIN for a
 * non-finite pose, transform, weight or sqrt_info entry. ---- */
int lvio2d_pose_graph_solve(lvio2d_ctx* ctx, int32_t n_poses, double* poses, int32_t n_edges, const int32_t* edge_index,
                            const double* edge_tf, const double* edge_weight, const double* sqrt_info, int32_t ground_p,
                            int32_t ground_q, int32_t fixed_pose, lvio2d_summary* summary);
/* edge_factor through auto_diff::compute_res_and_jacobi (common.h:201-217): res[6], jac[6][12] row-major over
 * (p_i, q_i, p_j, q_j). */
int lvio2d_eval_edge_factor(lvio2d_ctx* ctx, const double* tf12, double weight, const double* sqrt_info,
                            const double* pose_i, const double* pose_j, double* res, double* jac);

#ifdef __cplusplus
}
#endif
#endif /* LVIO2D_H_ */
