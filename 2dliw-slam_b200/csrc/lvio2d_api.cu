// lvio2d_api.cu — host side of the C ABI (include/lvio2d.h): context, device buffers, kernel launches.
// Everything runs on one CUDA stream owned by the context; there is no CPU compute path.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/lvio2d.h"
#include "aux_kernels.cuh"
#include "scan_lines.cuh"
#include "submap.cuh"
#include "pose_graph_segments.cuh"
#include "peaks.cuh"

using namespace lv;

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    bool ensure(size_t bytes) {
        if (bytes <= cap) return true;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        if (cudaMalloc(&p, bytes) != cudaSuccess) return false;
        cap = bytes;
        return true;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// page-locked host staging array: copies out of it are truly asynchronous (a cudaMemcpyAsync from pageable memory
// blocks the calling thread until the stream has drained up to it, which serialises chunked uploads with the host)
template <class T>
struct PinnedVec {
    T* p = nullptr;
    size_t cap = 0, n = 0;
    bool assign(size_t count, T value) {
        if (count > cap) {
            if (p) cudaFreeHost(p);
            p = nullptr; cap = 0;
            if (cudaHostAlloc(reinterpret_cast<void**>(&p), std::max<size_t>(count, 64) * sizeof(T), cudaHostAllocDefault) != cudaSuccess) return false;
            cap = std::max<size_t>(count, 64);
        }
        n = count;
        std::fill(p, p + count, value);
        return true;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = n = 0; }
    T* data() { return p; }
    T& operator[](size_t i) { return p[i]; }
    const T& operator[](size_t i) const { return p[i]; }
};

}  // namespace

struct lvio2d_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    lvio2d_params params;
    Consts C;
    LMOptions opt;
    char err[512] = {0};
    int sm_count = 148;
    double huber = 0.0;

    // problem
    bool have = false, bound = false;
    int B = 0, n = 0, ground_mult = 0, prior_frame = -1;
    bool arrow = false, has_weight = false, has_imu = false, has_wheel = false, assoc_grid = false;
    int imu_stride = LVIO2D_IMU_BLOB;   // LVIO2D_IMU_COMPACT when the batch came with lvio2d_scan_wire::imu_compact
    int64_t N = 0, L = 0;
    int tiles = 1, line_cap = 1, npad = kPadTrack;
    int uniform_pts = 0, uniform_lines = 0;   // > 0: all frames have this many points / lines (scan-match fast prologue)
    int shard_rank = 0, shard_world = 1;
    // inputs (owned copies, or borrowed device pointers when bound)
    DevBuf b_points, b_pline, b_pweight, b_poff, b_loff, b_lines, b_ref, b_refpose, b_imu, b_wheel, b_pX0, b_pJ, b_pH, b_cmask, b_wr, b_wa, b_wl, b_wi, b_agp, b_agm;
    const double2* points = nullptr; const int32_t* point_line = nullptr; const double* point_weight = nullptr;
    const int64_t* point_offset = nullptr; const int64_t* line_offset = nullptr; const double4* lines = nullptr;
    const int32_t* ref_frame = nullptr; const double* ref_pose = nullptr; const double* imu = nullptr; const double* wheel = nullptr;
    const double* prior_X0 = nullptr; const double* prior_J = nullptr; const uint8_t* const_mask = nullptr;
    // work buffers
    DevBuf b_x0, b_x, b_xc, b_scale, b_ftab, b_reftab, b_wlines, b_wlen, b_part, b_lb, b_items, b_vec, b_fac, b_state, b_status, b_active, b_active1, b_reduce;
    DevBuf b_tmp[8];
    DevBuf b_ln[14];   // lvio2d_extract_lines: inputs / workspace / outputs
    DevBuf b_sp[7];    // lvio2d_scan_to_points
    DevBuf b_ml[16];   // lvio2d_match_lines
    int32_t* match_bbox = nullptr;   // set by lvio2d_submap_match around its call: the sub-map's own boxes of scan 1's lines
    DevBuf b_pg[2];    // lvio2d_pose_graph_solve: int arena, double arena
    PinnedVec<double> h_pg;   // its per-iteration read-back (scalars + flags)
    double* ext_reduce = nullptr; int64_t ext_reduce_count = 0;
    int step_calls = 0;
    int window_threads = 0;   // 0 = automatic
    int factor_paired = -1;    // LVIO2D_FACTOR_PAIRED=0 | 1 forces the one-item / two-items-per-warp factor kernel; -1 = by batch size
    bool fused_small = true;   // LVIO2D_FUSED_SMALL=0: batches of <= #SM windows also go through the three-kernel loop
    bool have_solution = false;
    // host staging of the small index arrays (kept alive until the next upload so that copies can stay asynchronous)
    PinnedVec<int64_t> h_poff, h_loff, h_woff;
    cudaEvent_t ev_staged = nullptr;
    // the factor kernel of a trip can run on a second stream beside the scan-match kernel (both only read the candidate; fork /
    // join with events).  Measured: nothing for a batch that fills the machine (26.16 vs 26.20 ms per step: either kernel
    // occupies every SM), 0.84 vs 0.88 ms for a single window — so it is on for small batches; LVIO2D_OVERLAP=0|1 forces it
    cudaStream_t stream2 = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr; int overlap = -1;   // -1: small batches only   // recorded behind the last copy out of the staging vectors of an upload
    DevBuf b_wsl, b_wso;      // lvio2d_scan_wire::shared_lines: the per-window line lists and their offsets
    PinnedVec<int32_t> h_rf;
    PinnedVec<uint8_t> h_cm, h_active, h_active1;
    // measurement
    bool profiling = false;
    std::vector<cudaEvent_t> ev_scan, ev_win, ev_fac;   // begin/end pairs
    size_t ev_scan_used = 0, ev_win_used = 0, ev_fac_used = 0;
    double launches = 0;
    double scan_bytes_per_launch = 0;
};

namespace {

int fail(lvio2d_ctx* c, int code, const char* what, cudaError_t e = cudaSuccess) {
    if (c) snprintf(c->err, sizeof(c->err), "%s%s%s", what, e != cudaSuccess ? ": " : "", e != cudaSuccess ? cudaGetErrorString(e) : "");
    return code;
}
#define CK(call)                                                              \
    do {                                                                      \
        cudaError_t e_ = (call);                                              \
        if (e_ != cudaSuccess) return fail(ctx, LVIO2D_ERR_CUDA, #call, e_); \
    } while (0)

Consts make_consts(const lvio2d_params& p) {
    Consts C;
    std::memcpy(C.T_il, p.T_imu_to_laser, sizeof(C.T_il));
    std::memcpy(C.T_io, p.T_imu_to_wheel, sizeof(C.T_io));
    C.g = p.g;
    C.laser_sqrt_info = 1.0 / p.line_to_line_sigma;
    C.ground_p_sqrt_info = 1.0 / p.manifold_p_sigma;
    C.ground_q_sqrt_info = 1.0 / p.manifold_q_sigma;
    for (int i = 0; i < 3; ++i) {
        C.Q[i] = p.imu_noise_acc_sigma[i] * p.imu_noise_acc_sigma[i];
        C.Q[3 + i] = p.imu_noise_gyro_sigma[i] * p.imu_noise_gyro_sigma[i];
        C.Q[6 + i] = p.imu_bias_acc_sigma[i] * p.imu_bias_acc_sigma[i];
        C.Q[9 + i] = p.imu_bias_gyro_sigma[i] * p.imu_bias_gyro_sigma[i];
        C.wheel_cov[i] = p.wheel_sigma[i] * p.wheel_sigma[i];
    }
    C.huber_delta = p.huber_delta;
    return C;
}
LMOptions make_opts(const lvio2d_params& p) {
    LMOptions o;
    o.max_iters = p.max_iters > 0 ? p.max_iters : 50;
    o.function_tolerance = p.function_tolerance > 0 ? p.function_tolerance : 1e-6;
    o.gradient_tolerance = p.gradient_tolerance > 0 ? p.gradient_tolerance : 1e-10;
    o.parameter_tolerance = p.parameter_tolerance > 0 ? p.parameter_tolerance : 1e-8;
    o.initial_radius = p.initial_trust_region_radius > 0 ? p.initial_trust_region_radius : 1e4;
    o.max_radius = 1e16; o.min_radius = 1e-32; o.min_relative_decrease = 1e-3;
    o.min_lm_diagonal = 1e-6; o.max_lm_diagonal = 1e32; o.max_consecutive_invalid = 5;
    return o;
}

// copy (host->device) or alias (device pointer) one input array
template <class T>
int take(lvio2d_ctx* ctx, bool bind, DevBuf& buf, const T*& dst, const void* src, size_t count) {
    if (!src || count == 0) { dst = nullptr; return LVIO2D_OK; }
    if (bind) { dst = reinterpret_cast<const T*>(src); return LVIO2D_OK; }
    if (!buf.ensure(count * sizeof(T))) return fail(ctx, LVIO2D_ERR_ALLOC, "cudaMalloc(input)");
    CK(cudaMemcpyAsync(buf.p, src, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    dst = buf.as<T>();
    return LVIO2D_OK;
}

cudaEvent_t next_event(std::vector<cudaEvent_t>& pool, size_t& used) {
    if (used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
    return pool[used++];
}

int scan_row(bool arrow, bool assoc) { return arrow ? kRowFree : (assoc ? kRowTrackAssoc : kRowTrack); }

size_t window_smem_bytes(const lvio2d_ctx* c) { return window_smem_doubles(c->n, c->arrow) * sizeof(double); }

ScanMatchArgs scan_args(lvio2d_ctx* ctx, int mode);

int launch_scan_match(lvio2d_ctx* ctx, int mode = 0) {
    ScanMatchArgs a = scan_args(ctx, mode);
    if (ctx->N == 0) return LVIO2D_OK;
    const int wpc = LV_SCAN_WPC;
    const int grid = (a.n_items + wpc - 1) / wpc;
    const bool assoc = ctx->params.assoc_mode == LVIO2D_ASSOC_NEAREST;
    const bool huber = ctx->huber > 0;
    const size_t smem = (size_t)wpc * (ctx->line_cap * scan_row(ctx->arrow, assoc) + ((assoc && !ctx->arrow) ? kAssocGridDoubles : 0)) * sizeof(double);
    if (ctx->profiling) cudaEventRecord(next_event(ctx->ev_scan, ctx->ev_scan_used), ctx->stream);
#define LAUNCH_SM(RF, HW, AS, HU)                                                                                            \
    do {                                                                                                                     \
        CK(cudaFuncSetAttribute(scan_match_kernel<RF, HW, AS, HU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        scan_match_kernel<RF, HW, AS, HU><<<grid, wpc * 32, smem, ctx->stream>>>(a);                                         \
    } while (0)
#define DISPATCH_HU(RF, HW, AS) do { if (huber) LAUNCH_SM(RF, HW, AS, true); else LAUNCH_SM(RF, HW, AS, false); } while (0)
#define DISPATCH_AS(RF, HW) do { if (assoc) DISPATCH_HU(RF, HW, true); else DISPATCH_HU(RF, HW, false); } while (0)
    if (ctx->arrow) { if (ctx->has_weight) DISPATCH_AS(true, true); else DISPATCH_AS(true, false); }
    else { if (ctx->has_weight) DISPATCH_AS(false, true); else DISPATCH_AS(false, false); }
#undef DISPATCH_AS
#undef DISPATCH_HU
#undef LAUNCH_SM
    if (ctx->profiling) cudaEventRecord(next_event(ctx->ev_scan, ctx->ev_scan_used), ctx->stream);
    ctx->launches += 1;
    CK(cudaGetLastError());
    return LVIO2D_OK;
}

ScanMatchArgs scan_args(lvio2d_ctx* ctx, int mode) {
    ScanMatchArgs a;
    a.points = ctx->points; a.point_line = ctx->point_line; a.point_weight = ctx->point_weight;
    a.point_offset = ctx->point_offset; a.line_offset = ctx->line_offset;
    a.wlines = ctx->b_wlines.as<double4>(); a.wlen = ctx->b_wlen.as<double>(); a.lines = ctx->lines; a.ref_frame = ctx->ref_frame;
    a.frame_tab = ctx->b_ftab.as<double>(); a.frame_active = (mode == 1 ? ctx->b_active1 : ctx->b_active).as<uint8_t>();
    a.win_status = ctx->b_status.as<int32_t>(); a.partial = ctx->b_part.as<double>();
    a.n_frames = ctx->n; a.tiles = ctx->tiles; a.n_items = ctx->B * ctx->n * ctx->tiles;
    a.line_cap = ctx->line_cap; a.shard_rank = ctx->shard_rank; a.shard_world = ctx->shard_world;
    a.uniform_pts = ctx->uniform_pts; a.uniform_lines = ctx->uniform_lines;
    a.huber_delta = ctx->huber; a.laser_sqrt_info = ctx->C.laser_sqrt_info;
    a.assoc_gate = ctx->params.assoc_gate > 0 ? ctx->params.assoc_gate : 0.1;
    a.assoc_max_dist = ctx->params.assoc_max_dist > 0 ? ctx->params.assoc_max_dist : 0.5;
    a.assoc_grid_par = ctx->assoc_grid ? ctx->b_agp.as<double>() : nullptr;
    a.assoc_grid_mask = ctx->assoc_grid ? ctx->b_agm.as<unsigned long long>() : nullptr;
    return a;
}

WindowArgs window_args(lvio2d_ctx* ctx, int mode) {
    WindowArgs a;
    std::memset(&a, 0, sizeof(a));
    a.C = ctx->C; a.opt = ctx->opt;
    a.n_windows = ctx->B; a.n_frames = ctx->n; a.tiles = ctx->tiles; a.arrow = ctx->arrow; a.mode = mode;
    a.ground_multiplicity = ctx->ground_mult; a.prior_frame = ctx->prior_frame;
    a.has_imu = ctx->has_imu; a.has_wheel = ctx->has_wheel; a.imu_stride = ctx->imu_stride;
    a.const_mask = ctx->const_mask; a.frame_active = (mode == 1 ? ctx->b_active1 : ctx->b_active).as<uint8_t>(); a.ref_frame = ctx->ref_frame;
    a.imu = ctx->imu; a.wheel = ctx->wheel; a.prior_X0 = ctx->prior_X0; a.prior_J = ctx->prior_J; a.prior_H = ctx->b_pH.as<double>();
    a.partial = ctx->b_part.as<double>();
    a.x = ctx->b_x.as<double>(); a.xc = ctx->b_xc.as<double>(); a.scale = ctx->b_scale.as<double>();
    a.laser_blocks = ctx->b_lb.as<double>(); a.frame_tab = ctx->b_ftab.as<double>();
    a.items = ctx->b_items.as<double>(); a.vec = ctx->b_vec.as<double>(); a.fac = ctx->b_fac.as<double>();
    a.state = ctx->b_state.as<LMState>(); a.win_status = ctx->b_status.as<int32_t>();
    return a;
}

int launch_factors(lvio2d_ctx* ctx, const WindowArgs& a, cudaStream_t on = nullptr) {
    const cudaStream_t st = on ? on : ctx->stream;
    const int items = ctx->B * ctx->n;
    // two items per warp is the throughput shape; when the items do not even fill the machine one item per warp is the
    // shorter chain (measured: 19 vs 22 us per launch up to ~1100 items, 53 vs 45 us at 4440)
    const bool paired = ctx->factor_paired < 0 ? items > 8 * ctx->sm_count : ctx->factor_paired != 0;
    const int wpc = paired ? LV_PAIR_WPC : 4;
    if (ctx->profiling) cudaEventRecord(next_event(ctx->ev_fac, ctx->ev_fac_used), st);
    ctx->launches += 1;
    if (paired) {
        // two items per warp (half-warp each for the dual-number part, see factor_pair_kernel)
        const int pairs = (items + 1) / 2;
        const size_t smem = (size_t)wpc * kPairSmem * sizeof(double);
        CK(cudaFuncSetAttribute(factor_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        factor_pair_kernel<<<(pairs + wpc - 1) / wpc, wpc * 32, smem, st>>>(a);
    } else {
        const size_t smem = (size_t)wpc * kFactorSmem * sizeof(double);
        factor_kernel<<<(items + wpc - 1) / wpc, wpc * 32, smem, st>>>(a);
    }
    if (ctx->profiling) cudaEventRecord(next_event(ctx->ev_fac, ctx->ev_fac_used), st);
    CK(cudaGetLastError());
    return LVIO2D_OK;
}

int launch_window(lvio2d_ctx* ctx, const WindowArgs& a) {
    int per_window = (int)window_smem_doubles(ctx->n, ctx->arrow);
    // thread group per window: one warp when the batch fills the machine, four warps when it fits 4 CTAs per SM,
    // eight warps when it fits 2 CTAs per SM — the same arithmetic spread four ways, for the latency of small batches
    // (LVIO2D_WINDOW_THREADS = 32 | 128 | 256 forces one of them; used by the tests to cover both)
    int nt = ctx->B <= 2 * ctx->sm_count ? 256 : (ctx->B <= 4 * ctx->sm_count ? 128 : 32);
    if (ctx->window_threads == 32 || ctx->window_threads == 128 || ctx->window_threads == 256) nt = ctx->window_threads;
    // small batches of the tracking topology: sixteen warps per window and cyclic reduction over the frames (sequential
    // depth log2 n instead of n), when the window's blocks fit the CTA's shared memory (n <= 50 or so)
    const size_t cr_base = ((window_smem_doubles(ctx->n, false) + 1) & ~(size_t)1) * sizeof(double);
    const size_t cr_budget = 231000;   // 227 KB minus the kernel's static shared memory
    const bool cr_fits = !ctx->arrow && a.mode == 0 && !a.dense_H && cr_base + window_cr_doubles(ctx->n, 0) * sizeof(double) <= cr_budget;
    // (measured, C2: 512 beats the next best shape up to 3 x 148 windows — 3.89 vs 4.18 ms per solve at 444 — and loses from 4 x 148)
    if (cr_fits && ((ctx->B <= 3 * ctx->sm_count && ctx->window_threads == 0) || ctx->window_threads == 512)) nt = 512;
    if (ctx->profiling) cudaEventRecord(next_event(ctx->ev_win, ctx->ev_win_used), ctx->stream);
    ctx->launches += 1;
#define LAUNCH_WIN(AR, NT, GRID, BLOCK, SMEM)                                                                                 \
    do {                                                                                                                      \
        CK(cudaFuncSetAttribute(window_kernel<AR, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM)));            \
        window_kernel<AR, NT><<<GRID, BLOCK, SMEM, ctx->stream>>>(a, per_window);                                             \
    } while (0)
    if (nt == 512) {
        WindowArgs b = a;
        b.cr_scratch = (int)std::min<size_t>(13, (cr_budget - cr_base - window_cr_doubles(ctx->n, 0) * sizeof(double)) / (kTB * sizeof(double)));
        const size_t smem = cr_base + window_cr_doubles(ctx->n, b.cr_scratch) * sizeof(double);
        CK(cudaFuncSetAttribute(window_kernel<false, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        window_kernel<false, 512><<<ctx->B, 512, smem, ctx->stream>>>(b, per_window);
    } else if (nt == 32) {
        const int wpc = LV_WINDOW_WPC;
        // the fast path (tracking topology, solver program) needs 8.6 KB per window; everything else the generic layout
        const bool fast = !ctx->arrow && a.mode == 0 && !a.dense_H;
        if (fast) per_window = kWindowFastSmem;
        const size_t smem = (size_t)wpc * per_window * sizeof(double);
        const int grid = (ctx->B + wpc - 1) / wpc;
        if (ctx->arrow) LAUNCH_WIN(true, 32, grid, wpc * 32, smem); else LAUNCH_WIN(false, 32, grid, wpc * 32, smem);
    } else if (nt == 128) {
        const size_t smem = (size_t)per_window * sizeof(double);
        if (ctx->arrow) LAUNCH_WIN(true, 128, ctx->B, 128, smem); else LAUNCH_WIN(false, 128, ctx->B, 128, smem);
    } else {
        const size_t smem = (size_t)per_window * sizeof(double);
        if (ctx->arrow) LAUNCH_WIN(true, 256, ctx->B, 256, smem); else LAUNCH_WIN(false, 256, ctx->B, 256, smem);
    }
#undef LAUNCH_WIN
    if (ctx->profiling) cudaEventRecord(next_event(ctx->ev_win, ctx->ev_win_used), ctx->stream);
    CK(cudaGetLastError());
    return LVIO2D_OK;
}

// (re)initialise the minimiser at the initial states (b_x0) or, for linearize/marginalize after a solve, at the solution (b_x)
int begin_solve(lvio2d_ctx* ctx, bool from_solution = false) {
    const size_t ns = (size_t)ctx->B * ctx->n * 15;
    const void* src = from_solution ? ctx->b_x.p : ctx->b_x0.p;
    CK(cudaMemcpyAsync(ctx->b_xc.p, src, ns * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    if (!from_solution) CK(cudaMemcpyAsync(ctx->b_x.p, src, ns * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    init_state_kernel<<<(ctx->B + 127) / 128, 128, 0, ctx->stream>>>(ctx->b_state.as<LMState>(), ctx->b_status.as<int32_t>(), ctx->B,
                                                                     ctx->opt.initial_radius);
    ctx->launches += 1;
    const int F = ctx->B * ctx->n;
    frame_table_kernel<<<(F + 127) / 128, 128, 0, ctx->stream>>>(ctx->C, ctx->b_xc.as<double>(), 15, ctx->b_ftab.as<double>(), F);
    CK(cudaGetLastError());
    ctx->launches += 1;
    ctx->step_calls = 0;
    return LVIO2D_OK;
}

int setup_batch(lvio2d_ctx* ctx, const lvio2d_window_batch* b, bool bind, bool async = false, const lvio2d_scan_wire* wire = nullptr) {
    if (!ctx || !b) return LVIO2D_ERR_INVALID_ARG;
    if (b->n_windows < 1 || b->n_frames < 1 || !b->states) return fail(ctx, LVIO2D_ERR_INVALID_ARG, "n_windows/n_frames/states");
    if (b->n_frames > 64) return fail(ctx, LVIO2D_ERR_DOMAIN, "n_frames > 64");
    CK(cudaSetDevice(ctx->device));
    ctx->have = false;
    const int B = b->n_windows, n = b->n_frames, F = B * n;
    ctx->B = B; ctx->n = n; ctx->bound = bind;
    ctx->ground_mult = b->ground_multiplicity;
    ctx->prior_frame = (b->prior_frame >= 0 && b->prior_X0 && b->prior_J) ? b->prior_frame : -1;
    if (ctx->prior_frame >= n) return fail(ctx, LVIO2D_ERR_INVALID_ARG, "prior_frame >= n_frames");
    ctx->has_imu = (b->imu != nullptr || (wire && wire->imu_compact)) && n > 1;
    ctx->has_wheel = b->wheel != nullptr && n > 1;
    ctx->has_weight = b->point_weight != nullptr;

    // host-side views of the small index arrays (offsets, masks, ref frames) to derive the launch shape
    // a previous asynchronous upload may still be reading the staging vectors: wait for ITS copies, not for the whole
    // stream — the solve enqueued behind them may keep running while the host prepares the next upload (a host that waits
    // for the solve leaves this context idle until it comes round again, which cost the chunked end-to-end arm a third
    // of its throughput once PCIe was no longer the bottleneck)
    if (ctx->ev_staged) CK(cudaEventSynchronize(ctx->ev_staged));
    else CK(cudaEventCreateWithFlags(&ctx->ev_staged, cudaEventDisableTiming));
    PinnedVec<int64_t>&poff = ctx->h_poff, &loff = ctx->h_loff;
    PinnedVec<int32_t>& rf = ctx->h_rf;
    PinnedVec<uint8_t>& cm = ctx->h_cm;
    PinnedVec<uint8_t>&active = ctx->h_active, &active1 = ctx->h_active1;
    if (!poff.assign(F + 1, 0) || !loff.assign(F + 1, 0) || !rf.assign(F, -1) || !cm.assign(F, 0) || !active.assign(F, 0) || !active1.assign(F, 0))
        return fail(ctx, LVIO2D_ERR_ALLOC, "cudaHostAlloc(staging)");
    if (wire && (bind || wire->n_beams < 1 || !wire->ranges || !wire->angle || (!wire->beam_line && !wire->beam_line8)))
        return fail(ctx, LVIO2D_ERR_INVALID_ARG, "scan wire: host arrays ranges / angle / beam_line");
    const bool has_laser = wire ? (b->line_offset && b->lines) : (b->point_offset && b->points && b->point_line && b->line_offset && b->lines);
    const bool shared_lines = wire && wire->shared_lines != 0 && has_laser;
    if (bind) {
        if (has_laser) {
            CK(cudaMemcpy(poff.data(), b->point_offset, sizeof(int64_t) * (F + 1), cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(loff.data(), b->line_offset, sizeof(int64_t) * (F + 1), cudaMemcpyDeviceToHost));
            if (b->ref_frame) CK(cudaMemcpy(rf.data(), b->ref_frame, sizeof(int32_t) * F, cudaMemcpyDeviceToHost));
        }
        if (b->const_mask) CK(cudaMemcpy(cm.data(), b->const_mask, F, cudaMemcpyDeviceToHost));
    } else {
        if (has_laser) {
            if (wire) for (int f = 0; f <= F; ++f) poff[f] = (int64_t)f * wire->n_beams;   // every beam is a point slot
            else std::memcpy(poff.data(), b->point_offset, sizeof(int64_t) * (F + 1));
            if (shared_lines) {
                // one line list per window: the per-frame offsets of the replicated layout
                if (!ctx->h_woff.assign(B + 1, 0)) return fail(ctx, LVIO2D_ERR_ALLOC, "cudaHostAlloc(staging)");
                std::memcpy(ctx->h_woff.data(), b->line_offset, sizeof(int64_t) * (B + 1));
                for (int w = 0; w < B; ++w) {
                    const int64_t nl = ctx->h_woff[w + 1] - ctx->h_woff[w];
                    if (nl < 0) return fail(ctx, LVIO2D_ERR_INVALID_ARG, "offsets must be non-decreasing");
                    for (int k = 0; k < n; ++k) loff[(size_t)w * n + k + 1] = loff[(size_t)w * n + k] + nl;
                }
            } else {
                std::memcpy(loff.data(), b->line_offset, sizeof(int64_t) * (F + 1));
            }
            if (b->ref_frame) std::memcpy(rf.data(), b->ref_frame, sizeof(int32_t) * F);
        }
        if (b->const_mask) std::memcpy(cm.data(), b->const_mask, F);
    }
    ctx->N = has_laser ? poff[F] : 0;
    ctx->L = has_laser ? loff[F] : 0;
    bool arrow = false;
    int line_cap = 1;
    for (int f = 0; f < F; ++f) {
        if (!has_laser) break;
        const int64_t np = poff[f + 1] - poff[f], nl = loff[f + 1] - loff[f];
        if (np < 0 || nl < 0) return fail(ctx, LVIO2D_ERR_INVALID_ARG, "offsets must be non-decreasing");
        if (np == 0) continue;
        active1[f] = 1;
        line_cap = std::max<int>(line_cap, (int)nl);
        const int w0 = (f / n) * n;
        bool act = (cm[f] & 3) != 3;
        if (rf[f] >= 0) {
            // the arrow solver keeps fill-in bounded only when every in-window reference frame is frame 0
            if (rf[f] != 0 || f == w0) return fail(ctx, LVIO2D_ERR_DOMAIN, "ref_frame must be -1 or 0 (and not the frame itself)");
            arrow = true;
            act = act || (cm[w0] & 3) != 3;
        }
        active[f] = act ? 1 : 0;
    }
    if (line_cap > 512) return fail(ctx, LVIO2D_ERR_DOMAIN, "more than 512 lines in one local map");
    // host buffers of moderate size: every correspondence must point into its own frame's line list.  (Bound device
    // buffers and large uploads are not scanned on the host; the kernel skips out-of-range indices instead of reading
    // outside its line table.)
    if (!bind && !wire && has_laser && ctx->N <= ((int64_t)1 << 22)) {
        for (int f = 0; f < F; ++f) {
            const int64_t nl = loff[f + 1] - loff[f];
            for (int64_t p = poff[f]; p < poff[f + 1]; ++p)
                if (b->point_line[p] >= nl) return fail(ctx, LVIO2D_ERR_INVALID_ARG, "point_line index beyond the frame's line list");
        }
    }
    {
        // fixed-size scans: offsets are arithmetic, the scan-match prologue needs no dependent offset loads
        bool up = has_laser && F > 0 && poff[0] == 0, ul = has_laser && F > 0 && loff[0] == 0;
        const int64_t P0 = up ? poff[1] - poff[0] : 0, L0 = ul ? loff[1] - loff[0] : 0;
        for (int f = 0; f < F && (up || ul); ++f) {
            up = up && poff[f + 1] - poff[f] == P0;
            ul = ul && loff[f + 1] - loff[f] == L0;
        }
        ctx->uniform_pts = (up && P0 > 0 && P0 < (1 << 30)) ? (int)P0 : 0;
        ctx->uniform_lines = (ul && L0 > 0 && L0 < (1 << 30)) ? (int)L0 : 0;
    }
    ctx->arrow = arrow;
    ctx->line_cap = line_cap;
    ctx->npad = arrow ? kPadFree : kPadTrack;
    // tiles: enough warps to fill the machine a few times over, at least ~64 points per tile
    {
        const int64_t target = (int64_t)ctx->sm_count * 32;
        int tiles = (int)std::min<int64_t>(16, std::max<int64_t>(1, (target + F - 1) / F));
        const int64_t avg = F > 0 ? ctx->N / F : 0;
        tiles = (int)std::max<int64_t>(1, std::min<int64_t>(tiles, avg / 64));
        ctx->tiles = tiles;
    }
    const size_t smem_scan = (size_t)LV_SCAN_WPC * (line_cap * scan_row(arrow, ctx->params.assoc_mode == LVIO2D_ASSOC_NEAREST) + kAssocGridDoubles) * sizeof(double);
    if (smem_scan > 200 * 1024) return fail(ctx, LVIO2D_ERR_DOMAIN, "local map too large for shared memory");
    if (window_smem_bytes(ctx) > 200 * 1024) return fail(ctx, LVIO2D_ERR_DOMAIN, "n_frames too large for shared memory");

    int rc;
    const float* wire_r = nullptr; const float* wire_a = nullptr; const uint16_t* wire_l = nullptr; const uint8_t* wire_l8 = nullptr;
    const double4* shared_src = nullptr; const int64_t* shared_off = nullptr;
    if (wire && has_laser) {
        // compact wire encoding (lvio2d_set_windows_wire): float32 ranges + uint16 line indices travel, the double points
        // and int32 indices the scan-match kernel streams are rebuilt on the device
        const size_t N = (size_t)ctx->N;
        const float* d_r; const float* d_a; const uint16_t* d_l = nullptr; const uint8_t* d_l8 = nullptr;
        if ((rc = take<float>(ctx, false, ctx->b_wr, d_r, wire->ranges, N))) return rc;
        if ((rc = take<float>(ctx, false, ctx->b_wa, d_a, wire->angle, (size_t)F * 2))) return rc;
        if (wire->beam_line8) {
            if (line_cap > 255) return fail(ctx, LVIO2D_ERR_INVALID_ARG, "beam_line8 needs local maps of at most 255 lines");
            if ((rc = take<uint8_t>(ctx, false, ctx->b_wl, d_l8, wire->beam_line8, N))) return rc;
        } else if ((rc = take<uint16_t>(ctx, false, ctx->b_wl, d_l, wire->beam_line, N))) return rc;
        if (!ctx->b_points.ensure(N * sizeof(double2)) || !ctx->b_pline.ensure(N * sizeof(int32_t))) return fail(ctx, LVIO2D_ERR_ALLOC, "cudaMalloc(points)");
        // (launched below, behind every host -> device copy of this upload: a kernel in the middle of the copy sequence waits
        // for SMs that other contexts' solves keep busy, and the copies queued behind it leave PCIe idle meanwhile)
        wire_r = d_r; wire_a = d_a; wire_l = d_l; wire_l8 = d_l8;
        ctx->points = ctx->b_points.as<double2>(); ctx->point_line = ctx->b_pline.as<int32_t>(); ctx->point_weight = nullptr;
        ctx->has_weight = false;
    } else {
        if ((rc = take<double2>(ctx, bind, ctx->b_points, ctx->points, has_laser ? b->points : nullptr, (size_t)ctx->N))) return rc;
        if ((rc = take<int32_t>(ctx, bind, ctx->b_pline, ctx->point_line, has_laser ? b->point_line : nullptr, (size_t)ctx->N))) return rc;
        if ((rc = take<double>(ctx, bind, ctx->b_pweight, ctx->point_weight, has_laser ? b->point_weight : nullptr, (size_t)ctx->N))) return rc;
    }
    if ((rc = take<int64_t>(ctx, false, ctx->b_poff, ctx->point_offset, poff.data(), (size_t)F + 1))) return rc;
    if ((rc = take<int64_t>(ctx, false, ctx->b_loff, ctx->line_offset, loff.data(), (size_t)F + 1))) return rc;
    if (shared_lines) {
        const double4* d_sl; const int64_t* d_wo;
        if ((rc = take<double4>(ctx, false, ctx->b_wsl, d_sl, b->lines, (size_t)ctx->h_woff[B]))) return rc;
        if ((rc = take<int64_t>(ctx, false, ctx->b_wso, d_wo, ctx->h_woff.data(), (size_t)B + 1))) return rc;
        if (!ctx->b_lines.ensure(std::max<size_t>(1, (size_t)ctx->L) * sizeof(double4))) return fail(ctx, LVIO2D_ERR_ALLOC, "cudaMalloc(lines)");
        ctx->lines = ctx->b_lines.as<double4>();
        if (ctx->L > 0 && d_sl) { shared_src = d_sl; shared_off = d_wo; }
    } else if ((rc = take<double4>(ctx, bind, ctx->b_lines, ctx->lines, has_laser ? b->lines : nullptr, (size_t)ctx->L))) return rc;
    if ((rc = take<int32_t>(ctx, false, ctx->b_ref, ctx->ref_frame, rf.data(), (size_t)F))) return rc;
    if ((rc = take<uint8_t>(ctx, false, ctx->b_cmask, ctx->const_mask, cm.data(), (size_t)F))) return rc;
    if ((rc = take<double>(ctx, bind, ctx->b_refpose, ctx->ref_pose, has_laser ? b->ref_pose : nullptr, (size_t)F * 6))) return rc;
    ctx->imu_stride = LVIO2D_IMU_BLOB;
    if (wire && wire->imu_compact && ctx->has_imu) {
        // the factor kernels read the compact records in place (no expansion pass, 1.5 instead of 3.7 KB per item from HBM)
        if ((rc = take<double>(ctx, false, ctx->b_wi, ctx->imu, wire->imu_compact, (size_t)B * (n - 1) * LVIO2D_IMU_COMPACT))) return rc;
        ctx->imu_stride = LVIO2D_IMU_COMPACT;
    } else if ((rc = take<double>(ctx, bind, ctx->b_imu, ctx->imu, ctx->has_imu ? b->imu : nullptr, (size_t)B * (n - 1) * LVIO2D_IMU_BLOB))) return rc;
    if ((rc = take<double>(ctx, bind, ctx->b_wheel, ctx->wheel, ctx->has_wheel ? b->wheel : nullptr, (size_t)B * (n - 1) * LVIO2D_WHEEL_BLOB))) return rc;
    if ((rc = take<double>(ctx, bind, ctx->b_pX0, ctx->prior_X0, ctx->prior_frame >= 0 ? b->prior_X0 : nullptr, (size_t)B * 15))) return rc;
    if ((rc = take<double>(ctx, bind, ctx->b_pJ, ctx->prior_J, ctx->prior_frame >= 0 ? b->prior_J : nullptr, (size_t)B * 225))) return rc;

    const size_t ns = (size_t)F * 15 * sizeof(double);
    bool ok = ctx->b_x0.ensure(ns) && ctx->b_x.ensure(ns) && ctx->b_xc.ensure(ns) && ctx->b_scale.ensure(ns) &&
              ctx->b_ftab.ensure((size_t)F * kFrameTab * sizeof(double)) && ctx->b_reftab.ensure((size_t)F * kFrameTab * sizeof(double)) &&
              ctx->b_wlines.ensure(std::max<size_t>(1, (size_t)ctx->L) * sizeof(double4)) && ctx->b_wlen.ensure(std::max<size_t>(1, (size_t)ctx->L) * sizeof(double)) &&
              ctx->b_part.ensure((size_t)F * ctx->tiles * ctx->npad * sizeof(double)) && ctx->b_lb.ensure((size_t)2 * F * ctx->npad * sizeof(double)) &&
              ctx->b_items.ensure((size_t)2 * F * kItem * sizeof(double)) && ctx->b_vec.ensure((size_t)B * 3 * n * 15 * sizeof(double)) &&
              ctx->b_fac.ensure((size_t)F * 3 * kBlk * sizeof(double)) && ctx->b_state.ensure(sizeof(LMState) * B) &&
              ctx->b_status.ensure(sizeof(int32_t) * B) && ctx->b_active.ensure(F) && ctx->b_active1.ensure(F) && ctx->b_reduce.ensure((size_t)F * ctx->npad * sizeof(double));
    if (!ok) return fail(ctx, LVIO2D_ERR_ALLOC, "cudaMalloc(work buffers)");
    CK(cudaMemcpyAsync(ctx->b_x0.p, b->states, ns, bind ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->b_active.p, active.data(), F, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->b_active1.p, active1.data(), F, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaEventRecord(ctx->ev_staged, ctx->stream));   // every copy out of the staging vectors is enqueued
    if (wire_r) {
        const size_t N = (size_t)ctx->N;
        if (wire_l8)
            expand_wire_kernel<uint8_t><<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(wire_r, wire_a, wire_l8, wire->n_beams, (int64_t)N, ctx->b_points.as<double2>(), ctx->b_pline.as<int32_t>());
        else
            expand_wire_kernel<uint16_t><<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(wire_r, wire_a, wire_l, wire->n_beams, (int64_t)N, ctx->b_points.as<double2>(), ctx->b_pline.as<int32_t>());
        CK(cudaGetLastError());
        ctx->launches += 1;
    }
    if (shared_src) {
        expand_shared_lines_kernel<<<F, 128, 0, ctx->stream>>>(shared_src, shared_off, ctx->line_offset, n, F, ctx->b_lines.as<double4>());
        CK(cudaGetLastError());
        ctx->launches += 1;
    }
    CK(cudaMemsetAsync(ctx->b_part.p, 0, (size_t)F * ctx->tiles * ctx->npad * sizeof(double), ctx->stream));
    if (ctx->prior_frame >= 0) {
        // J^T J of the marginalisation prior: constant over the solve, so the factor kernel adds it instead of recomputing it
        if (!ctx->b_pH.ensure((size_t)B * kBlk * sizeof(double))) return fail(ctx, LVIO2D_ERR_ALLOC, "cudaMalloc(prior information)");
        prior_info_kernel<<<B, 64, 0, ctx->stream>>>(ctx->prior_J, ctx->b_pH.as<double>(), B);
        CK(cudaGetLastError());
    }
    if (has_laser && ctx->L > 0) {
        // world lines of every local map that hangs under an external constant pose
        frame_table_kernel<<<(F + 127) / 128, 128, 0, ctx->stream>>>(ctx->C, ctx->ref_pose, 6, ctx->b_reftab.as<double>(), F);
        world_lines_kernel<<<F, 128, 0, ctx->stream>>>(ctx->lines, ctx->line_offset, ctx->ref_frame,
                                                       ctx->b_reftab.as<double>(), ctx->b_wlines.as<double4>(), ctx->b_wlen.as<double>(), F);
        CK(cudaGetLastError());
        // in-kernel re-association: the candidate grid of every frame's local map (LVIO2D_ASSOC_GRID=0: all-lines loop)
        ctx->assoc_grid = false;
        const char* ag = std::getenv("LVIO2D_ASSOC_GRID");
        if (ctx->params.assoc_mode == LVIO2D_ASSOC_NEAREST && !arrow && !(ag && std::atoi(ag) == 0)) {
            if (!ctx->b_agp.ensure((size_t)F * 4 * sizeof(double)) || !ctx->b_agm.ensure((size_t)F * kAssocGridDoubles * sizeof(unsigned long long)))
                return fail(ctx, LVIO2D_ERR_ALLOC, "cudaMalloc(association grid)");
            const double gate = ctx->params.assoc_gate > 0 ? ctx->params.assoc_gate : 0.1, md = ctx->params.assoc_max_dist > 0 ? ctx->params.assoc_max_dist : 0.5;
            assoc_grid_kernel<<<F, 128, 0, ctx->stream>>>(ctx->b_wlines.as<double4>(), ctx->b_wlen.as<double>(), ctx->line_offset, ctx->ref_frame, gate, md,
                                                          ctx->b_agp.as<double>(), ctx->b_agm.as<unsigned long long>(), F);
            CK(cudaGetLastError());
            ctx->assoc_grid = true;
        }
    }
    if (!async) CK(cudaStreamSynchronize(ctx->stream));  // synchronous flavour: the caller's buffers are free again on return
    {
        // algorithmic bytes one scan-match launch moves when every active frame is processed
        double pts = 0, lns = 0, frames = 0;
        for (int f = 0; f < F; ++f)
            if (active[f]) { pts += (double)(poff[f + 1] - poff[f]); lns += (double)(loff[f + 1] - loff[f]); frames += 1; }
        ctx->scan_bytes_per_launch = pts * (16 + 4 + (ctx->has_weight ? 8 : 0)) + lns * 32 +
                                     frames * (kFrameTab * 8.0 + (double)ctx->tiles * ctx->npad * 8.0);
    }
    ctx->ext_reduce = nullptr;
    ctx->have = true;
    ctx->have_solution = false;
    return LVIO2D_OK;
}

int run_linearize(lvio2d_ctx* ctx, int mode, WindowArgs& a) {
    int rc = begin_solve(ctx, ctx->have_solution);
    if (rc) return rc;
    if ((rc = launch_scan_match(ctx, mode))) return rc;
    a.x = ctx->b_x.as<double>();
    if ((rc = launch_factors(ctx, a))) return rc;
    return launch_window(ctx, a);
}

}  // namespace

extern "C" {

const char* lvio2d_strerror(int status) {
    switch (status) {
        case LVIO2D_OK: return "ok";
        case LVIO2D_ERR_INVALID_ARG: return "invalid argument";
        case LVIO2D_ERR_NO_DEVICE: return "no CUDA sm_100 device (the library has no CPU path)";
        case LVIO2D_ERR_CUDA: return "CUDA error";
        case LVIO2D_ERR_NO_WINDOW: return "no window batch set";
        case LVIO2D_ERR_DOMAIN: return "input outside the supported domain";
        case LVIO2D_ERR_ALLOC: return "device allocation failed";
        default: return "unknown status";
    }
}

const char* lvio2d_last_error(const lvio2d_ctx* ctx) { return ctx ? ctx->err : ""; }

int lvio2d_create(lvio2d_ctx** out, const lvio2d_params* params) {
    if (!out || !params) return LVIO2D_ERR_INVALID_ARG;
    *out = nullptr;
    if (params->abi_version != LVIO2D_ABI_VERSION) return LVIO2D_ERR_INVALID_ARG;
    if (!(params->line_to_line_sigma > 0) || !(params->manifold_p_sigma > 0) || !(params->manifold_q_sigma > 0)) return LVIO2D_ERR_INVALID_ARG;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= params->device || params->device < 0) return LVIO2D_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, params->device) != cudaSuccess) return LVIO2D_ERR_NO_DEVICE;
    if (prop.major != 10) return LVIO2D_ERR_NO_DEVICE;  // kernels are built for sm_100a only
    lvio2d_ctx* ctx = new (std::nothrow) lvio2d_ctx();
    if (!ctx) return LVIO2D_ERR_ALLOC;
    ctx->device = params->device;
    ctx->params = *params;
    ctx->C = make_consts(*params);
    ctx->opt = make_opts(*params);
    ctx->huber = (params->huber_delta > 0 && std::isfinite(params->huber_delta)) ? params->huber_delta : 0.0;
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaSetDevice(ctx->device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return LVIO2D_ERR_CUDA;
    }
    if (const char* wt = std::getenv("LVIO2D_WINDOW_THREADS")) ctx->window_threads = std::atoi(wt);
    if (const char* fp = std::getenv("LVIO2D_FACTOR_PAIRED")) ctx->factor_paired = std::atoi(fp) != 0 ? 1 : 0;
    if (const char* fs = std::getenv("LVIO2D_FUSED_SMALL")) ctx->fused_small = std::atoi(fs) != 0;
    if (cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess) {
        lvio2d_destroy(ctx);
        return LVIO2D_ERR_CUDA;
    }
    if (const char* ov = std::getenv("LVIO2D_OVERLAP")) ctx->overlap = std::atoi(ov) != 0 ? 1 : 0;
    *out = ctx;
    return LVIO2D_OK;
}

void lvio2d_destroy(lvio2d_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    DevBuf* all[] = {&ctx->b_points, &ctx->b_pline, &ctx->b_pweight, &ctx->b_poff, &ctx->b_loff, &ctx->b_lines, &ctx->b_ref, &ctx->b_refpose,
                     &ctx->b_imu, &ctx->b_wheel, &ctx->b_pX0, &ctx->b_pJ, &ctx->b_pH, &ctx->b_cmask, &ctx->b_wr, &ctx->b_wa, &ctx->b_wl, &ctx->b_wi, &ctx->b_wsl, &ctx->b_wso, &ctx->b_agp, &ctx->b_agm, &ctx->b_x0, &ctx->b_x, &ctx->b_xc, &ctx->b_scale,
                     &ctx->b_ftab, &ctx->b_reftab, &ctx->b_wlines, &ctx->b_wlen, &ctx->b_part, &ctx->b_lb, &ctx->b_items, &ctx->b_vec, &ctx->b_fac, &ctx->b_state,
                     &ctx->b_status, &ctx->b_active, &ctx->b_active1, &ctx->b_reduce};
    for (DevBuf* b : all) b->release();
    for (auto& b : ctx->b_tmp) b.release();
    for (auto& b : ctx->b_ln) b.release();
    for (auto& b : ctx->b_sp) b.release();
    for (auto& b : ctx->b_ml) b.release();
    for (auto& b : ctx->b_pg) b.release();
    ctx->h_pg.release();
    ctx->h_poff.release(); ctx->h_loff.release(); ctx->h_woff.release(); ctx->h_rf.release(); ctx->h_cm.release(); ctx->h_active.release(); ctx->h_active1.release();
    for (auto e : ctx->ev_scan) cudaEventDestroy(e);
    for (auto e : ctx->ev_win) cudaEventDestroy(e);
    for (auto e : ctx->ev_fac) cudaEventDestroy(e);
    if (ctx->ev_staged) cudaEventDestroy(ctx->ev_staged);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

void* lvio2d_stream(lvio2d_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int lvio2d_set_windows(lvio2d_ctx* ctx, const lvio2d_window_batch* host_batch) { return setup_batch(ctx, host_batch, false); }
int lvio2d_bind_windows(lvio2d_ctx* ctx, const lvio2d_window_batch* device_batch) { return setup_batch(ctx, device_batch, true); }
int lvio2d_set_windows_async(lvio2d_ctx* ctx, const lvio2d_window_batch* host_batch) { return setup_batch(ctx, host_batch, false, true); }
int lvio2d_set_windows_wire(lvio2d_ctx* ctx, const lvio2d_window_batch* host_batch, const lvio2d_scan_wire* wire, int32_t async) {
    if (!wire) return LVIO2D_ERR_INVALID_ARG;
    return setup_batch(ctx, host_batch, false, async != 0, wire);
}

int lvio2d_set_max_iterations(lvio2d_ctx* ctx, int32_t max_iters) {
    if (!ctx) return LVIO2D_ERR_INVALID_ARG;
    ctx->opt.max_iters = max_iters > 0 ? max_iters : 50;
    ctx->params.max_iters = ctx->opt.max_iters;
    return LVIO2D_OK;
}

int lvio2d_reset_states(lvio2d_ctx* ctx, const double* host_states) {
    if (!ctx || !host_states) return LVIO2D_ERR_INVALID_ARG;
    if (!ctx->have) return fail(ctx, LVIO2D_ERR_NO_WINDOW, "reset_states");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(ctx->b_x0.p, host_states, (size_t)ctx->B * ctx->n * 15 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ctx->have_solution = false;
    return LVIO2D_OK;
}

int lvio2d_set_profiling(lvio2d_ctx* ctx, int32_t on) {
    if (!ctx) return LVIO2D_ERR_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->profiling = on != 0;
    ctx->ev_scan_used = ctx->ev_win_used = ctx->ev_fac_used = 0;
    ctx->launches = 0;
    return LVIO2D_OK;
}

int lvio2d_get_profile(lvio2d_ctx* ctx, double* out) {
    if (!ctx || !out) return LVIO2D_ERR_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 8; ++i) out[i] = 0.0;
    for (size_t i = 0; i + 1 < ctx->ev_scan_used; i += 2) { float ms = 0; cudaEventElapsedTime(&ms, ctx->ev_scan[i], ctx->ev_scan[i + 1]); out[0] += ms; out[1] += 1; }
    for (size_t i = 0; i + 1 < ctx->ev_win_used; i += 2) { float ms = 0; cudaEventElapsedTime(&ms, ctx->ev_win[i], ctx->ev_win[i + 1]); out[2] += ms; out[3] += 1; }
    for (size_t i = 0; i + 1 < ctx->ev_fac_used; i += 2) { float ms = 0; cudaEventElapsedTime(&ms, ctx->ev_fac[i], ctx->ev_fac[i + 1]); out[6] += ms; out[7] += 1; }
    out[4] = ctx->launches;
    out[5] = ctx->scan_bytes_per_launch;
    return LVIO2D_OK;
}

int lvio2d_measure_fp64_peak(lvio2d_ctx* ctx, double* out) {
    if (!ctx || !out) return LVIO2D_ERR_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    const int grid = ctx->sm_count * 8, block = 256, iters = 4096;
    if (!ctx->b_tmp[0].ensure((size_t)grid * block * sizeof(double))) return fail(ctx, LVIO2D_ERR_ALLOC, "cudaMalloc(peak)");
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int which = 0; which < 2; ++which) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {   // first repetition warms up; best of the other three
            CK(cudaEventRecord(e0, ctx->stream));
            if (which == 0) dfma_peak_kernel<<<grid, block, 0, ctx->stream>>>(ctx->b_tmp[0].as<double>(), iters, 1.0000001, 1e-9);
            else dmma_peak_kernel<<<grid, block, 0, ctx->stream>>>(ctx->b_tmp[0].as<double>(), iters, 1.0000001, 1e-9);
            CK(cudaEventRecord(e1, ctx->stream));
            CK(cudaEventSynchronize(e1));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0) best = std::min(best, ms);
        }
        // DFMA: 64 FMAs per thread per iteration; DMMA: 32 mma per warp per iteration, 8*8*4 FMAs each
        const double fma = which == 0 ? (double)grid * block * iters * 64.0 : (double)grid * (block / 32) * iters * 32.0 * 256.0;
        out[which] = 2.0 * fma / (best * 1e-3) * 1e-12;   // TFLOP/s
        out[2 + which] = best;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    CK(cudaGetLastError());
    return LVIO2D_OK;
}

int lvio2d_set_point_shard(lvio2d_ctx* ctx, int32_t rank, int32_t world) {
    if (!ctx || world < 1 || rank < 0 || rank >= world) return LVIO2D_ERR_INVALID_ARG;
    ctx->shard_rank = rank; ctx->shard_world = world;
    return LVIO2D_OK;
}

int lvio2d_solve_begin(lvio2d_ctx* ctx) {
    if (!ctx) return LVIO2D_ERR_INVALID_ARG;
    if (!ctx->have) return fail(ctx, LVIO2D_ERR_NO_WINDOW, "solve_begin");
    CK(cudaSetDevice(ctx->device));
    return begin_solve(ctx);
}

int lvio2d_eval_laser(lvio2d_ctx* ctx) {
    if (!ctx) return LVIO2D_ERR_INVALID_ARG;
    if (!ctx->have) return fail(ctx, LVIO2D_ERR_NO_WINDOW, "eval_laser");
    CK(cudaSetDevice(ctx->device));
    int rc = launch_scan_match(ctx);
    if (rc) return rc;
    const int F = ctx->B * ctx->n;
    double* out = ctx->ext_reduce ? ctx->ext_reduce : ctx->b_reduce.as<double>();
    const int total = F * ctx->npad;
    reduce_tiles_kernel<<<(total + 255) / 256, 256, 0, ctx->stream>>>(ctx->b_part.as<double>(), ctx->b_active.as<uint8_t>(), out, F, ctx->tiles, ctx->npad);
    CK(cudaGetLastError());
    return LVIO2D_OK;
}

int lvio2d_reduce_buffer(lvio2d_ctx* ctx, void** device_ptr, int64_t* count) {
    if (!ctx || !device_ptr || !count) return LVIO2D_ERR_INVALID_ARG;
    if (!ctx->have) return fail(ctx, LVIO2D_ERR_NO_WINDOW, "reduce_buffer");
    *device_ptr = ctx->ext_reduce ? (void*)ctx->ext_reduce : ctx->b_reduce.p;
    *count = (int64_t)ctx->B * ctx->n * ctx->npad;
    return LVIO2D_OK;
}

int lvio2d_set_reduce_buffer(lvio2d_ctx* ctx, void* device_ptr, int64_t count) {
    if (!ctx) return LVIO2D_ERR_INVALID_ARG;
    if (!ctx->have) return fail(ctx, LVIO2D_ERR_NO_WINDOW, "set_reduce_buffer");
    if (device_ptr && count < (int64_t)ctx->B * ctx->n * ctx->npad) return fail(ctx, LVIO2D_ERR_INVALID_ARG, "reduce buffer too small");
    ctx->ext_reduce = reinterpret_cast<double*>(device_ptr);
    ctx->ext_reduce_count = count;
    return LVIO2D_OK;
}

int lvio2d_lm_step(lvio2d_ctx* ctx, int32_t* n_active) {
    if (!ctx) return LVIO2D_ERR_INVALID_ARG;
    if (!ctx->have) return fail(ctx, LVIO2D_ERR_NO_WINDOW, "lm_step");
    CK(cudaSetDevice(ctx->device));
    WindowArgs a = window_args(ctx, 0);
    // the (all-reduced) per-frame blocks stand in for the tile partials
    a.partial = ctx->ext_reduce ? ctx->ext_reduce : ctx->b_reduce.as<double>();
    a.tiles = 1;
    int rc = launch_factors(ctx, a);
    if (rc) return rc;
    rc = launch_window(ctx, a);
    if (rc) return rc;
    ++ctx->step_calls;
    ctx->have_solution = true;
    if (n_active) {
        std::vector<int32_t> st(ctx->B);
        CK(cudaMemcpyAsync(st.data(), ctx->b_status.p, sizeof(int32_t) * ctx->B, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        int32_t act = 0;
        for (int32_t s : st) act += (s == 0);
        *n_active = act;
    }
    return LVIO2D_OK;
}

int lvio2d_solve_async(lvio2d_ctx* ctx) {
    if (!ctx) return LVIO2D_ERR_INVALID_ARG;
    if (!ctx->have) return fail(ctx, LVIO2D_ERR_NO_WINDOW, "solve");
    CK(cudaSetDevice(ctx->device));
    int rc = begin_solve(ctx);
    if (rc) return rc;
    WindowArgs a = window_args(ctx, 0);
    // small batches: the whole loop in one launch, one CTA per window (solve_small_kernel)
    {
        const bool assoc = ctx->params.assoc_mode == LVIO2D_ASSOC_NEAREST;
        const size_t smem = std::max({(size_t)8 * ctx->line_cap * scan_row(ctx->arrow, false) * sizeof(double), (size_t)8 * kPairSmem * sizeof(double),
                                      window_smem_bytes(ctx)});
        // (measured: 2 frames x 16 points 2.84 -> 2.47 ms, 10 frames x 176 points 6.41 -> 6.13 ms per 50-iteration solve; one
        // 30-frame x 1081-beam window is faster spread over the machine by the three-kernel loop: 2.3 vs 3.8 ms)
        const bool tiny = ctx->n <= 12 && ctx->N <= (int64_t)4096 * ctx->B;
        if (ctx->fused_small && tiny && ctx->B <= ctx->sm_count && !assoc && !(ctx->huber > 0) && ctx->shard_world == 1 && !ctx->profiling &&
            smem <= 200 * 1024) {
            const ScanMatchArgs sa = scan_args(ctx, 0);
            const int trips = ctx->opt.max_iters + 1;
#define LAUNCH_SMALL(AR, HW)                                                                                            \
    do {                                                                                                                \
        CK(cudaFuncSetAttribute(solve_small_kernel<AR, HW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
        solve_small_kernel<AR, HW><<<ctx->B, 256, smem, ctx->stream>>>(sa, a, trips);                                   \
    } while (0)
            if (ctx->arrow) { if (ctx->has_weight) LAUNCH_SMALL(true, true); else LAUNCH_SMALL(true, false); }
            else { if (ctx->has_weight) LAUNCH_SMALL(false, true); else LAUNCH_SMALL(false, false); }
#undef LAUNCH_SMALL
            ctx->launches += 1;
            CK(cudaGetLastError());
            ctx->have_solution = true;
            return LVIO2D_OK;
        }
    }
    // trip 0 linearises the initial point; trips 1..max_iters each judge one candidate
    const bool overlap = ctx->N > 0 && !ctx->profiling && (ctx->overlap < 0 ? ctx->B <= 3 * ctx->sm_count : ctx->overlap != 0);
    for (int it = 0; it <= ctx->opt.max_iters; ++it) {
        if (overlap) {
            CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
            CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
            if ((rc = launch_factors(ctx, a, ctx->stream2))) return rc;
            CK(cudaEventRecord(ctx->ev_join, ctx->stream2));
            if ((rc = launch_scan_match(ctx))) return rc;
            CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
        } else {
            if ((rc = launch_scan_match(ctx))) return rc;
            if ((rc = launch_factors(ctx, a))) return rc;
        }
        if ((rc = launch_window(ctx, a))) return rc;
    }
    ctx->have_solution = true;
    return LVIO2D_OK;
}

int lvio2d_sync(lvio2d_ctx* ctx) {
    if (!ctx) return LVIO2D_ERR_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return LVIO2D_OK;
}

int lvio2d_get_summaries(lvio2d_ctx* ctx, lvio2d_summary* out) {
    if (!ctx || !out) return LVIO2D_ERR_INVALID_ARG;
    if (!ctx->have) return fail(ctx, LVIO2D_ERR_NO_WINDOW, "get_summaries");
    CK(cudaSetDevice(ctx->device));
    std::vector<LMState> st(ctx->B);
    CK(cudaMemcpyAsync(st.data(), ctx->b_state.p, sizeof(LMState) * ctx->B, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int w = 0; w < ctx->B; ++w) {
        out[w].iterations = st[w].iteration;
        out[w].termination = st[w].termination;
        out[w].num_successful_steps = st[w].n_success;
        out[w].num_unsuccessful_steps = st[w].n_unsuccess;
        out[w].initial_cost = st[w].initial_cost;
        out[w].final_cost = st[w].cost;
        out[w].final_radius = st[w].radius;
        out[w].reserved = 0.0;
    }
    return LVIO2D_OK;
}

int lvio2d_solve(lvio2d_ctx* ctx, lvio2d_summary* summaries) {
    int rc = lvio2d_solve_async(ctx);
    if (rc) return rc;
    if (summaries) return lvio2d_get_summaries(ctx, summaries);
    return lvio2d_sync(ctx);
}

int lvio2d_get_states(lvio2d_ctx* ctx, double* host_states) {
    if (!ctx || !host_states) return LVIO2D_ERR_INVALID_ARG;
    if (!ctx->have) return fail(ctx, LVIO2D_ERR_NO_WINDOW, "get_states");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(host_states, ctx->b_x.p, (size_t)ctx->B * ctx->n * 15 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return LVIO2D_OK;
}

int lvio2d_get_states_async(lvio2d_ctx* ctx, double* host_states) {
    if (!ctx || !host_states) return LVIO2D_ERR_INVALID_ARG;
    if (!ctx->have) return fail(ctx, LVIO2D_ERR_NO_WINDOW, "get_states_async");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(host_states, ctx->b_x.p, (size_t)ctx->B * ctx->n * 15 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return LVIO2D_OK;
}

int lvio2d_linearize(lvio2d_ctx* ctx, int32_t mode, double* H, double* g, double* cost) {
    if (!ctx || !H || !g || !cost || (mode != 0 && mode != 1)) return LVIO2D_ERR_INVALID_ARG;
    if (!ctx->have) return fail(ctx, LVIO2D_ERR_NO_WINDOW, "linearize");
    CK(cudaSetDevice(ctx->device));
    const size_t dim = 15 * (size_t)ctx->n;
    if (!ctx->b_tmp[0].ensure(ctx->B * dim * dim * sizeof(double)) || !ctx->b_tmp[1].ensure(ctx->B * dim * sizeof(double)) ||
        !ctx->b_tmp[2].ensure(ctx->B * sizeof(double)))
        return fail(ctx, LVIO2D_ERR_ALLOC, "cudaMalloc(linearize)");
    WindowArgs a = window_args(ctx, mode);
    a.dense_H = ctx->b_tmp[0].as<double>(); a.dense_g = ctx->b_tmp[1].as<double>(); a.dense_cost = ctx->b_tmp[2].as<double>();
    int rc = run_linearize(ctx, mode, a);
    if (rc) return rc;
    CK(cudaMemcpyAsync(H, a.dense_H, ctx->B * dim * dim * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(g, a.dense_g, ctx->B * dim * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(cost, a.dense_cost, ctx->B * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return LVIO2D_OK;
}

int lvio2d_marginalize(lvio2d_ctx* ctx, double* X0, double* J_lin, double* r_lin) {
    if (!ctx || !X0 || !J_lin || !r_lin) return LVIO2D_ERR_INVALID_ARG;
    if (!ctx->have) return fail(ctx, LVIO2D_ERR_NO_WINDOW, "marginalize");
    CK(cudaSetDevice(ctx->device));
    const int B = ctx->B;
    if (!ctx->b_tmp[3].ensure((size_t)B * 225 * sizeof(double)) || !ctx->b_tmp[4].ensure((size_t)B * 15 * sizeof(double)) ||
        !ctx->b_tmp[5].ensure((size_t)B * 15 * sizeof(double)) || !ctx->b_tmp[6].ensure((size_t)B * 225 * sizeof(double)) ||
        !ctx->b_tmp[7].ensure((size_t)B * 15 * sizeof(double)))
        return fail(ctx, LVIO2D_ERR_ALLOC, "cudaMalloc(marginalize)");
    WindowArgs a = window_args(ctx, 1);
    a.marg_H = ctx->b_tmp[3].as<double>(); a.marg_g = ctx->b_tmp[4].as<double>();
    int rc = run_linearize(ctx, 1, a);
    if (rc) return rc;
    marginal_prior_kernel<<<(B + 63) / 64, 64, 0, ctx->stream>>>(B, ctx->n, a.marg_H, a.marg_g, ctx->b_x.as<double>(), ctx->b_tmp[5].as<double>(),
                                                                ctx->b_tmp[6].as<double>(), ctx->b_tmp[7].as<double>());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(X0, ctx->b_tmp[5].p, (size_t)B * 15 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(J_lin, ctx->b_tmp[6].p, (size_t)B * 225 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(r_lin, ctx->b_tmp[7].p, (size_t)B * 15 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return LVIO2D_OK;
}

int lvio2d_imu_preintegrate(lvio2d_ctx* ctx, int32_t n_intervals, const int64_t* sample_offset, const double* samples, const double* bias0,
                            double* out_blobs) {
    if (!ctx || n_intervals < 0 || !sample_offset || !bias0 || !out_blobs) return LVIO2D_ERR_INVALID_ARG;
    if (n_intervals == 0) return LVIO2D_OK;
    CK(cudaSetDevice(ctx->device));
    const int64_t M = sample_offset[n_intervals];
    if (M > 0 && !samples) return LVIO2D_ERR_INVALID_ARG;
    if (!ctx->b_tmp[0].ensure(sizeof(int64_t) * (n_intervals + 1)) || !ctx->b_tmp[1].ensure(std::max<size_t>(8, sizeof(double) * 7 * M)) ||
        !ctx->b_tmp[2].ensure(sizeof(double) * 6 * n_intervals) || !ctx->b_tmp[3].ensure(sizeof(double) * 466 * (size_t)n_intervals))
        return fail(ctx, LVIO2D_ERR_ALLOC, "cudaMalloc(imu_preintegrate)");
    CK(cudaMemcpyAsync(ctx->b_tmp[0].p, sample_offset, sizeof(int64_t) * (n_intervals + 1), cudaMemcpyHostToDevice, ctx->stream));
    if (M > 0) CK(cudaMemcpyAsync(ctx->b_tmp[1].p, samples, sizeof(double) * 7 * M, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->b_tmp[2].p, bias0, sizeof(double) * 6 * n_intervals, cudaMemcpyHostToDevice, ctx->stream));
    const int wpc = 4;
    const size_t smem = (size_t)wpc * kImuPreSmem * sizeof(double);
    imu_preintegrate_kernel<<<(n_intervals + wpc - 1) / wpc, wpc * 32, smem, ctx->stream>>>(
        ctx->C, n_intervals, ctx->b_tmp[0].as<int64_t>(), ctx->b_tmp[1].as<double>(), ctx->b_tmp[2].as<double>(), ctx->b_tmp[3].as<double>());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_blobs, ctx->b_tmp[3].p, sizeof(double) * 466 * (size_t)n_intervals, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return LVIO2D_OK;
}

int lvio2d_wheel_preintegrate(lvio2d_ctx* ctx, int32_t n_intervals, const int64_t* step_offset, const double* steps, double* out_blobs) {
    if (!ctx || n_intervals < 0 || !step_offset || !out_blobs) return LVIO2D_ERR_INVALID_ARG;
    if (n_intervals == 0) return LVIO2D_OK;
    CK(cudaSetDevice(ctx->device));
    const int64_t M = step_offset[n_intervals];
    if (M > 0 && !steps) return LVIO2D_ERR_INVALID_ARG;
    if (!ctx->b_tmp[0].ensure(sizeof(int64_t) * (n_intervals + 1)) || !ctx->b_tmp[1].ensure(std::max<size_t>(8, sizeof(double) * 7 * M)) ||
        !ctx->b_tmp[3].ensure(sizeof(double) * 15 * (size_t)n_intervals))
        return fail(ctx, LVIO2D_ERR_ALLOC, "cudaMalloc(wheel_preintegrate)");
    CK(cudaMemcpyAsync(ctx->b_tmp[0].p, step_offset, sizeof(int64_t) * (n_intervals + 1), cudaMemcpyHostToDevice, ctx->stream));
    if (M > 0) CK(cudaMemcpyAsync(ctx->b_tmp[1].p, steps, sizeof(double) * 7 * M, cudaMemcpyHostToDevice, ctx->stream));
    wheel_preintegrate_kernel<<<(n_intervals + 127) / 128, 128, 0, ctx->stream>>>(ctx->C, n_intervals, ctx->b_tmp[0].as<int64_t>(),
                                                                                  ctx->b_tmp[1].as<double>(), ctx->b_tmp[3].as<double>());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_blobs, ctx->b_tmp[3].p, sizeof(double) * 15 * (size_t)n_intervals, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return LVIO2D_OK;
}

// ---- single-factor hooks: inputs staged into one scratch buffer, outputs read back
static int eval_hook(lvio2d_ctx* ctx, const double* const* in, const int* in_len, int n_in, double* res, int n_res, double* jac, int n_jac, int which) {
    if (!ctx || !res || !jac) return LVIO2D_ERR_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    int total = 0;
    for (int i = 0; i < n_in; ++i) { if (!in[i]) return LVIO2D_ERR_INVALID_ARG; total += in_len[i]; }
    std::vector<double> stage(total);
    int off[8], o = 0;
    for (int i = 0; i < n_in; ++i) { off[i] = o; std::memcpy(stage.data() + o, in[i], sizeof(double) * in_len[i]); o += in_len[i]; }
    if (!ctx->b_tmp[0].ensure(sizeof(double) * (total + n_res + n_jac))) return fail(ctx, LVIO2D_ERR_ALLOC, "cudaMalloc(eval)");
    double* d = ctx->b_tmp[0].as<double>();
    CK(cudaMemcpyAsync(d, stage.data(), sizeof(double) * total, cudaMemcpyHostToDevice, ctx->stream));
    double* dres = d + total;
    double* djac = dres + n_res;
    switch (which) {
        case 0: eval_laser_factor_kernel<<<1, 32, 0, ctx->stream>>>(ctx->C, d + off[0], d + off[1], d + off[2], d + off[3], d + off[4], d + off[5], dres, djac); break;
        case 1: eval_imu_factor_kernel<<<1, 32, 0, ctx->stream>>>(ctx->C, d + off[0], d + off[1], d + off[2], dres, djac); break;
        case 2: eval_wheel_factor_kernel<<<1, 32, 0, ctx->stream>>>(ctx->C, d + off[0], d + off[1], d + off[2], dres, djac); break;
        default: eval_ground_factors_kernel<<<1, 32, 0, ctx->stream>>>(ctx->C, d + off[0], dres, djac); break;
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(res, dres, sizeof(double) * n_res, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(jac, djac, sizeof(double) * n_jac, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return LVIO2D_OK;
}

int lvio2d_eval_laser_factor(lvio2d_ctx* ctx, const double* l1_p1, const double* l1_p2, const double* l2_p1, const double* l2_p2,
                             const double* pose_i, const double* pose_j, double* res, double* jac) {
    const double* in[6] = {l1_p1, l1_p2, l2_p1, l2_p2, pose_i, pose_j};
    const int len[6] = {3, 3, 3, 3, 6, 6};
    return eval_hook(ctx, in, len, 6, res, 2, jac, 24, 0);
}
int lvio2d_eval_imu_factor(lvio2d_ctx* ctx, const double* imu_blob, const double* state_i, const double* state_j, double* res, double* jac) {
    const double* in[3] = {imu_blob, state_i, state_j};
    const int len[3] = {466, 15, 15};
    return eval_hook(ctx, in, len, 3, res, 15, jac, 450, 1);
}
int lvio2d_eval_wheel_factor(lvio2d_ctx* ctx, const double* wheel_blob, const double* pose_i, const double* pose_j, double* res, double* jac) {
    const double* in[3] = {wheel_blob, pose_i, pose_j};
    const int len[3] = {15, 6, 6};
    return eval_hook(ctx, in, len, 3, res, 3, jac, 36, 2);
}
int lvio2d_extract_lines(lvio2d_ctx* ctx, const lvio2d_line_params* lp, int32_t n_scans, const int64_t* point_offset, const int32_t* point_count,
                         const double* points, const double* point_z, int32_t max_lines, int32_t* n_lines, double* lines, double* abc,
                         int32_t* index_range, int32_t on_device) {
    if (!ctx || !lp || n_scans < 0 || !point_offset || max_lines < 1 || !n_lines || !lines || !abc || !index_range) return LVIO2D_ERR_INVALID_ARG;
    if (!(lp->laser_resolution > 0.0)) return fail(ctx, LVIO2D_ERR_INVALID_ARG, "laser_resolution must be positive");
    if (n_scans == 0) return LVIO2D_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t S = (size_t)n_scans;
    // N = extent of the point arrays the scans address (with point_count: the end of the last scan's slots)
    int64_t N = 0;
    const int n_off = point_count ? n_scans : n_scans + 1;
    if (on_device) {
        int64_t last = 0;
        int32_t last_cnt = 0;
        CK(cudaMemcpyAsync(&last, point_offset + n_off - 1, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        if (point_count) CK(cudaMemcpyAsync(&last_cnt, point_count + n_scans - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        N = last + last_cnt;
    } else {
        for (int32_t s = 0; s < n_scans; ++s) {
            const int64_t end = point_count ? point_offset[s] + point_count[s] : point_offset[s + 1];
            if (end < point_offset[s] || point_offset[s] < 0) return fail(ctx, LVIO2D_ERR_INVALID_ARG, "offsets must be non-decreasing");
            N = std::max(N, end);
        }
    }
    if (N < 0) return fail(ctx, LVIO2D_ERR_INVALID_ARG, "negative point count");
    if (N > 0 && !points) return LVIO2D_ERR_INVALID_ARG;
    DevBuf* B = ctx->b_ln;
    const size_t nn = (size_t)std::max<int64_t>(N, 1);
    bool ok = B[0].ensure(nn * sizeof(double)) && B[1].ensure(nn * sizeof(int32_t)) && B[2].ensure(nn * sizeof(int32_t)) &&
              B[3].ensure((2 * nn + 2 * S) * sizeof(int32_t)) && B[4].ensure((2 * nn + 2 * S) * sizeof(int32_t)) && B[5].ensure((nn + S) * sizeof(int32_t));
    if (!on_device)
        ok = ok && B[6].ensure((S + 1) * sizeof(int64_t)) && B[7].ensure(nn * sizeof(double2)) && B[8].ensure(S * sizeof(int32_t)) &&
             B[9].ensure(S * max_lines * 4 * sizeof(double)) && B[10].ensure(S * max_lines * 3 * sizeof(double)) && B[11].ensure(S * max_lines * 2 * sizeof(int32_t)) &&
             B[12].ensure(S * sizeof(int32_t)) && B[13].ensure(nn * sizeof(double));
    if (!ok) return fail(ctx, LVIO2D_ERR_ALLOC, "cudaMalloc(extract_lines)");
    ScanLinesArgs a;
    a.n_scans = n_scans; a.max_lines = max_lines;
    a.continuous_threshold = lp->line_continuous_threshold;
    a.max_tolerance_angle = lp->line_max_tolerance_angle_deg / 180.0 * M_PI;   // convert::angle_to_rad (common.h:48)
    a.max_dis = lp->line_max_dis; a.min_len = lp->line_min_len; a.resolution = lp->laser_resolution;
    a.w = (int)(lp->w_laser_each_scan / lp->laser_resolution + 1);             // laser_manager.cpp:231-236
    a.h = (int)(lp->h_laser_each_scan / lp->laser_resolution + 1);
    a.resp = B[0].as<double>(); a.seg_s = B[1].as<int32_t>(); a.seg_e = B[2].as<int32_t>();
    a.cand = B[3].as<int32_t>(); a.lstart = B[4].as<int32_t>(); a.seg_first = B[5].as<int32_t>();
    if (on_device) {
        a.point_offset = point_offset; a.points = reinterpret_cast<const double2*>(points);
        a.point_count = point_count; a.point_z = point_z;
        a.n_lines = n_lines; a.lines = lines; a.abc = abc; a.index_range = index_range;
    } else {
        CK(cudaMemcpyAsync(B[6].p, point_offset, (size_t)n_off * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
        if (N > 0) CK(cudaMemcpyAsync(B[7].p, points, (size_t)N * sizeof(double2), cudaMemcpyHostToDevice, ctx->stream));
        if (point_count) CK(cudaMemcpyAsync(B[12].p, point_count, S * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        if (point_z && N > 0) CK(cudaMemcpyAsync(B[13].p, point_z, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        a.point_offset = B[6].as<int64_t>(); a.points = B[7].as<double2>();
        a.point_count = point_count ? B[12].as<int32_t>() : nullptr; a.point_z = point_z ? B[13].as<double>() : nullptr;
        a.n_lines = B[8].as<int32_t>(); a.lines = B[9].as<double>(); a.abc = B[10].as<double>(); a.index_range = B[11].as<int32_t>();
        CK(cudaMemsetAsync(B[9].p, 0, S * max_lines * 4 * sizeof(double), ctx->stream));
        CK(cudaMemsetAsync(B[10].p, 0, S * max_lines * 3 * sizeof(double), ctx->stream));
        CK(cudaMemsetAsync(B[11].p, 0, S * max_lines * 2 * sizeof(int32_t), ctx->stream));
    }
    const int wpc = 4;
    scan_lines_kernel<<<(n_scans + wpc - 1) / wpc, wpc * 32, 0, ctx->stream>>>(a);
    ctx->launches += 1;
    CK(cudaGetLastError());
    if (!on_device) {
        CK(cudaMemcpyAsync(n_lines, B[8].p, S * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(lines, B[9].p, S * max_lines * 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(abc, B[10].p, S * max_lines * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(index_range, B[11].p, S * max_lines * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return LVIO2D_OK;
}

int lvio2d_scan_to_points(lvio2d_ctx* ctx, int32_t n_scans, int32_t n_beams, const float* ranges, const lvio2d_scan_header* headers, int32_t deskew,
                          int32_t* point_count, double* points, double* point_z, double* point_time, int32_t on_device) {
    if (!ctx || n_scans < 0 || n_beams < 1 || !ranges || !headers || !point_count || !points || !point_z) return LVIO2D_ERR_INVALID_ARG;
    if (n_scans == 0) return LVIO2D_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t S = (size_t)n_scans, NB = S * (size_t)n_beams;
    DevBuf* B = ctx->b_sp;
    bool ok = B[0].ensure(NB * sizeof(int32_t));
    if (!on_device)
        ok = ok && B[1].ensure(NB * sizeof(float)) && B[2].ensure(S * sizeof(lvio2d_scan_header)) && B[3].ensure(S * sizeof(int32_t)) &&
             B[4].ensure(NB * sizeof(double2)) && B[5].ensure(NB * sizeof(double)) && B[6].ensure(NB * sizeof(double));
    if (!ok) return fail(ctx, LVIO2D_ERR_ALLOC, "cudaMalloc(scan_to_points)");
    ScanPointsArgs a;
    a.n_scans = n_scans; a.n_beams = n_beams; a.deskew = deskew;
    a.beam_index = B[0].as<int32_t>();
    if (on_device) {
        a.ranges = ranges; a.headers = headers; a.point_count = point_count; a.points = reinterpret_cast<double2*>(points);
        a.point_z = point_z; a.point_time = point_time;
    } else {
        CK(cudaMemcpyAsync(B[1].p, ranges, NB * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(B[2].p, headers, S * sizeof(lvio2d_scan_header), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemsetAsync(B[4].p, 0, NB * sizeof(double2), ctx->stream));
        CK(cudaMemsetAsync(B[5].p, 0, NB * sizeof(double), ctx->stream));
        CK(cudaMemsetAsync(B[6].p, 0, NB * sizeof(double), ctx->stream));
        a.ranges = B[1].as<float>(); a.headers = B[2].as<lvio2d_scan_header>(); a.point_count = B[3].as<int32_t>();
        a.points = B[4].as<double2>(); a.point_z = B[5].as<double>(); a.point_time = point_time ? B[6].as<double>() : nullptr;
    }
    const int wpc = 4;
    scan_points_kernel<<<(n_scans + wpc - 1) / wpc, wpc * 32, 0, ctx->stream>>>(a);
    ctx->launches += 1;
    CK(cudaGetLastError());
    if (!on_device) {
        CK(cudaMemcpyAsync(point_count, B[3].p, S * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(points, B[4].p, NB * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(point_z, B[5].p, NB * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (point_time) CK(cudaMemcpyAsync(point_time, B[6].p, NB * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return LVIO2D_OK;
}

int lvio2d_match_lines(lvio2d_ctx* ctx, const lvio2d_line_params* lp, int32_t n_pairs, int32_t kk, const int64_t* point_offset1,
                       const int32_t* point_count1, const double* points1, int32_t max_lines1, const int32_t* n_lines1, const double* lines1,
                       const int32_t* index_range1, int32_t max_lines2, const int32_t* n_lines2, const double* lines2, const double* pose1,
                       const double* pose2, int32_t* n_match, int32_t* match, int32_t on_device) {
    if (!ctx || !lp || n_pairs < 0 || kk < 0 || kk > 2 || max_lines1 < 1 || max_lines2 < 1 || !n_lines1 || !lines1 || !n_lines2 || !lines2 ||
        !pose1 || !pose2 || !n_match || !match)
        return LVIO2D_ERR_INVALID_ARG;
    if (points1 && (!point_offset1 || !index_range1)) return fail(ctx, LVIO2D_ERR_INVALID_ARG, "points1 needs point_offset1 and index_range1");
    if (!(lp->laser_resolution > 0.0)) return fail(ctx, LVIO2D_ERR_INVALID_ARG, "laser_resolution must be positive");
    if (n_pairs == 0) return LVIO2D_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t P = (size_t)n_pairs;
    int64_t N1 = 0;
    const int n_off = point_count1 ? n_pairs : n_pairs + 1;
    if (points1) {
        if (on_device) {
            int64_t last = 0;
            int32_t last_cnt = 0;
            CK(cudaMemcpyAsync(&last, point_offset1 + n_off - 1, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
            if (point_count1) CK(cudaMemcpyAsync(&last_cnt, point_count1 + n_pairs - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            N1 = last + last_cnt;
        } else {
            for (int32_t p = 0; p < n_pairs; ++p) {
                const int64_t end = point_count1 ? point_offset1[p] + point_count1[p] : point_offset1[p + 1];
                if (end < point_offset1[p] || point_offset1[p] < 0) return fail(ctx, LVIO2D_ERR_INVALID_ARG, "offsets must be non-decreasing");
                N1 = std::max(N1, end);
                // every line's point range must lie inside its scan (host buffers only; device buffers are clamped by the kernel)
                const int64_t cnt = end - point_offset1[p];
                const int32_t nl = std::min(n_lines1[p], max_lines1);
                for (int32_t l = 0; l < nl; ++l) {
                    const int32_t* r = index_range1 + ((size_t)p * max_lines1 + l) * 2;
                    if (r[0] < 0 || r[1] < r[0] || r[1] >= cnt) return fail(ctx, LVIO2D_ERR_INVALID_ARG, "index_range1 outside the scan's points");
                }
            }
        }
    }
    DevBuf* B = ctx->b_ml;
    const size_t nn = (size_t)std::max<int64_t>(N1, 1);
    bool ok = B[0].ensure(nn * sizeof(int32_t)) && B[1].ensure(P * max_lines2 * sizeof(double)) && B[2].ensure(P * max_lines2 * 2 * sizeof(int32_t)) &&
              B[15].ensure(P * max_lines1 * 4 * sizeof(int32_t));
    if (!on_device)
        ok = ok && B[3].ensure((P + 1) * sizeof(int64_t)) && B[4].ensure(P * sizeof(int32_t)) && B[5].ensure(nn * sizeof(double2)) &&
             B[6].ensure(P * sizeof(int32_t)) && B[7].ensure(P * max_lines1 * sizeof(double4)) && B[8].ensure(P * max_lines1 * 2 * sizeof(int32_t)) &&
             B[9].ensure(P * sizeof(int32_t)) && B[10].ensure(P * max_lines2 * sizeof(double4)) && B[11].ensure(P * 6 * sizeof(double)) &&
             B[12].ensure(P * 6 * sizeof(double)) && B[13].ensure(P * sizeof(int32_t)) && B[14].ensure(P * max_lines2 * 2 * sizeof(int32_t));
    if (!ok) return fail(ctx, LVIO2D_ERR_ALLOC, "cudaMalloc(match_lines)");
    MatchLinesArgs a;
    std::memset(&a, 0, sizeof(a));
    a.max_lines1 = max_lines1; a.max_lines2 = max_lines2; a.n_pairs = n_pairs; a.kk = kk;
    std::memcpy(a.T_il, ctx->params.T_imu_to_laser, sizeof(a.T_il));
    a.resolution = lp->laser_resolution;
    a.w = (int)(lp->w_laser_each_scan / lp->laser_resolution + 1);
    a.h = (int)(lp->h_laser_each_scan / lp->laser_resolution + 1);
    a.cells = B[0].as<int32_t>(); a.diss = B[1].as<double>(); a.prov = B[2].as<int32_t>(); a.bbox = B[15].as<int32_t>();
    if (ctx->match_bbox) { a.bbox = ctx->match_bbox; a.bbox_ready = 1; }
    if (on_device) {
        a.point_offset1 = point_offset1; a.point_count1 = point_count1; a.points1 = reinterpret_cast<const double2*>(points1);
        a.n_lines1 = n_lines1; a.lines1 = reinterpret_cast<const double4*>(lines1); a.index_range1 = index_range1;
        a.n_lines2 = n_lines2; a.lines2 = reinterpret_cast<const double4*>(lines2); a.pose1 = pose1; a.pose2 = pose2;
        a.n_match = n_match; a.match = match;
    } else {
        auto up = [&](DevBuf& b, const void* src, size_t bytes) -> cudaError_t {
            return (src && bytes) ? cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream) : cudaSuccess;
        };
        CK(up(B[3], points1 ? point_offset1 : nullptr, (size_t)n_off * sizeof(int64_t)));
        CK(up(B[4], points1 ? point_count1 : nullptr, P * sizeof(int32_t)));
        CK(up(B[5], points1, (size_t)N1 * sizeof(double2)));
        CK(up(B[6], n_lines1, P * sizeof(int32_t)));
        CK(up(B[7], lines1, P * max_lines1 * sizeof(double4)));
        CK(up(B[8], index_range1, P * max_lines1 * 2 * sizeof(int32_t)));
        CK(up(B[9], n_lines2, P * sizeof(int32_t)));
        CK(up(B[10], lines2, P * max_lines2 * sizeof(double4)));
        CK(up(B[11], pose1, P * 6 * sizeof(double)));
        CK(up(B[12], pose2, P * 6 * sizeof(double)));
        CK(cudaMemsetAsync(B[14].p, 0, P * max_lines2 * 2 * sizeof(int32_t), ctx->stream));
        a.point_offset1 = points1 ? B[3].as<int64_t>() : nullptr; a.point_count1 = (points1 && point_count1) ? B[4].as<int32_t>() : nullptr;
        a.points1 = points1 ? B[5].as<double2>() : nullptr;
        a.n_lines1 = B[6].as<int32_t>(); a.lines1 = B[7].as<double4>(); a.index_range1 = index_range1 ? B[8].as<int32_t>() : nullptr;
        a.n_lines2 = B[9].as<int32_t>(); a.lines2 = B[10].as<double4>(); a.pose1 = B[11].as<double>(); a.pose2 = B[12].as<double>();
        a.n_match = B[13].as<int32_t>(); a.match = B[14].as<int32_t>();
    }
    const int wpc = 4;
    // few pairs: the four warps of a CTA share one pair (a quarter of the latency); many pairs: one warp each
    if (n_pairs <= 2 * ctx->sm_count) match_lines_kernel<4><<<n_pairs, wpc * 32, 0, ctx->stream>>>(a);
    else match_lines_kernel<1><<<(n_pairs + wpc - 1) / wpc, wpc * 32, 0, ctx->stream>>>(a);
    ctx->launches += 1;
    CK(cudaGetLastError());
    if (!on_device) {
        CK(cudaMemcpyAsync(n_match, B[13].p, P * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(match, B[14].p, P * max_lines2 * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return LVIO2D_OK;
}

// ---- device-resident reference sub-map (submap.cuh)
struct lvio2d_submap {
    lvio2d_ctx* ctx = nullptr;
    int device = 0;            // copies: lvio2d_submap_destroy must not touch a context that is already gone
    int32_t n_managers = 0, line_cap = 0, n_accumulation = 0;
    lvio2d_line_params lp{};
    double filter_p = 0, filter_q = 0;
    DevBuf meta, sub_pose, last_pose, sub_n, sub_lines, sub_bbox;   // state
    DevBuf in_n, in_lines, in_pose;                       // staging of host scans
};

static int submap_zero(lvio2d_submap* sm) {
    lvio2d_ctx* ctx = sm->ctx;
    CK(cudaMemsetAsync(sm->meta.p, 0, sm->meta.cap, ctx->stream));
    CK(cudaMemsetAsync(sm->sub_pose.p, 0, sm->sub_pose.cap, ctx->stream));
    CK(cudaMemsetAsync(sm->last_pose.p, 0, sm->last_pose.cap, ctx->stream));
    CK(cudaMemsetAsync(sm->sub_n.p, 0, sm->sub_n.cap, ctx->stream));
    CK(cudaMemsetAsync(sm->sub_lines.p, 0, sm->sub_lines.cap, ctx->stream));   // (slots beyond n_lines read as zero in lvio2d_submap_get)
    CK(cudaMemsetAsync(sm->sub_bbox.p, 0, sm->sub_bbox.cap, ctx->stream));
    return LVIO2D_OK;
}

int lvio2d_submap_create(lvio2d_ctx* ctx, int32_t n_managers, int32_t line_cap, const lvio2d_line_params* lp, double ref_motion_filter_p,
                         double ref_motion_filter_q, int32_t ref_n_accumulation, lvio2d_submap** out) {
    if (!ctx || !lp || !out || n_managers < 1 || line_cap < 1 || ref_n_accumulation < 1) return LVIO2D_ERR_INVALID_ARG;
    if (!(lp->laser_resolution > 0.0)) return fail(ctx, LVIO2D_ERR_INVALID_ARG, "laser_resolution must be positive");
    if (!std::isfinite(ref_motion_filter_p) || !std::isfinite(ref_motion_filter_q)) return fail(ctx, LVIO2D_ERR_INVALID_ARG, "motion filter thresholds must be finite");
    CK(cudaSetDevice(ctx->device));
    lvio2d_submap* sm = new (std::nothrow) lvio2d_submap;
    if (!sm) return fail(ctx, LVIO2D_ERR_ALLOC, "new lvio2d_submap");
    sm->ctx = ctx; sm->device = ctx->device; sm->n_managers = n_managers; sm->line_cap = line_cap; sm->n_accumulation = ref_n_accumulation;
    sm->lp = *lp; sm->filter_p = ref_motion_filter_p; sm->filter_q = ref_motion_filter_q;
    const size_t M = (size_t)n_managers;
    const bool ok = sm->meta.ensure(M * 4 * sizeof(int32_t)) && sm->sub_pose.ensure(2 * M * 6 * sizeof(double)) && sm->last_pose.ensure(M * 6 * sizeof(double)) &&
                    sm->sub_n.ensure(2 * M * sizeof(int32_t)) && sm->sub_lines.ensure(2 * M * (size_t)line_cap * sizeof(double4)) &&
                    sm->sub_bbox.ensure(2 * M * (size_t)line_cap * sizeof(int4));
    if (!ok) { lvio2d_submap_destroy(sm); return fail(ctx, LVIO2D_ERR_ALLOC, "cudaMalloc(submap)"); }
    const int rc = submap_zero(sm);
    if (rc != LVIO2D_OK) { lvio2d_submap_destroy(sm); return rc; }
    *out = sm;
    return LVIO2D_OK;
}

void lvio2d_submap_destroy(lvio2d_submap* sm) {
    if (!sm) return;
    cudaSetDevice(sm->device);
    cudaDeviceSynchronize();   // (not the context's stream: the context may have been destroyed first)
    DevBuf* all[] = {&sm->meta, &sm->sub_pose, &sm->last_pose, &sm->sub_n, &sm->sub_lines, &sm->sub_bbox, &sm->in_n, &sm->in_lines, &sm->in_pose};
    for (DevBuf* b : all) b->release();
    delete sm;
}

int lvio2d_submap_reset(lvio2d_submap* sm) {
    if (!sm) return LVIO2D_ERR_INVALID_ARG;
    lvio2d_ctx* ctx = sm->ctx;
    CK(cudaSetDevice(ctx->device));
    return submap_zero(sm);
}

int lvio2d_submap_add_scan(lvio2d_submap* sm, int32_t max_lines, const int32_t* n_lines, const double* lines, const double* pose, int32_t on_device) {
    if (!sm) return LVIO2D_ERR_INVALID_ARG;
    lvio2d_ctx* ctx = sm->ctx;
    if (max_lines < 1 || !n_lines || !lines || !pose) return fail(ctx, LVIO2D_ERR_INVALID_ARG, "submap_add_scan: null argument or max_lines < 1");
    CK(cudaSetDevice(ctx->device));
    const size_t M = (size_t)sm->n_managers;
    SubmapArgs a;
    std::memset(&a, 0, sizeof(a));
    a.n_managers = sm->n_managers; a.line_cap = sm->line_cap; a.max_lines = max_lines; a.n_accumulation = sm->n_accumulation;
    a.filter_p = sm->filter_p; a.filter_q = sm->filter_q;
    std::memcpy(a.T_il, ctx->params.T_imu_to_laser, sizeof(a.T_il));
    a.line_max_dis = sm->lp.line_max_dis; a.line_min_len = sm->lp.line_min_len; a.resolution = sm->lp.laser_resolution;
    a.w = (int)(sm->lp.w_laser_each_scan / sm->lp.laser_resolution + 1);
    a.h = (int)(sm->lp.h_laser_each_scan / sm->lp.laser_resolution + 1);
    a.meta = sm->meta.as<int32_t>(); a.sub_pose = sm->sub_pose.as<double>(); a.last_pose = sm->last_pose.as<double>();
    a.sub_n = sm->sub_n.as<int32_t>(); a.sub_lines = sm->sub_lines.as<double4>(); a.sub_bbox = sm->sub_bbox.as<int4>();
    if (on_device) {
        a.n_lines = n_lines; a.lines = reinterpret_cast<const double4*>(lines); a.pose = pose;
    } else {
        for (size_t m = 0; m < M; ++m)
            for (int k = 0; k < 6; ++k)
                if (n_lines[m] >= 0 && !std::isfinite(pose[6 * m + k])) return fail(ctx, LVIO2D_ERR_INVALID_ARG, "submap_add_scan: pose is not finite");
        if (!sm->in_n.ensure(M * sizeof(int32_t)) || !sm->in_lines.ensure(M * (size_t)max_lines * sizeof(double4)) || !sm->in_pose.ensure(M * 6 * sizeof(double)))
            return fail(ctx, LVIO2D_ERR_ALLOC, "cudaMalloc(submap scan)");
        CK(cudaMemcpyAsync(sm->in_n.p, n_lines, M * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(sm->in_lines.p, lines, M * (size_t)max_lines * sizeof(double4), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(sm->in_pose.p, pose, M * 6 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        a.n_lines = sm->in_n.as<int32_t>(); a.lines = sm->in_lines.as<double4>(); a.pose = sm->in_pose.as<double>();
    }
    const int wpc = 4;
    submap_add_scan_kernel<<<(sm->n_managers + wpc - 1) / wpc, wpc * 32, 0, ctx->stream>>>(a);
    ctx->launches += 1;
    CK(cudaGetLastError());
    if (!on_device) CK(cudaStreamSynchronize(ctx->stream));   // the staging buffers are the caller's again
    return LVIO2D_OK;
}

int lvio2d_submap_get(lvio2d_submap* sm, int32_t which, int32_t* meta, double* pose, int32_t* n_lines, double* lines) {
    if (!sm) return LVIO2D_ERR_INVALID_ARG;
    lvio2d_ctx* ctx = sm->ctx;
    if (which < 0 || which > 1) return fail(ctx, LVIO2D_ERR_INVALID_ARG, "submap_get: which must be 0 (reference) or 1 (spawning)");
    CK(cudaSetDevice(ctx->device));
    const size_t M = (size_t)sm->n_managers;
    if (meta) CK(cudaMemcpyAsync(meta, sm->meta.p, M * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (pose) CK(cudaMemcpyAsync(pose, sm->sub_pose.as<double>() + which * M * 6, M * 6 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (n_lines) CK(cudaMemcpyAsync(n_lines, sm->sub_n.as<int32_t>() + which * M, M * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (lines)
        CK(cudaMemcpyAsync(lines, sm->sub_lines.as<double4>() + which * M * sm->line_cap, M * (size_t)sm->line_cap * sizeof(double4), cudaMemcpyDeviceToHost,
                           ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return LVIO2D_OK;
}

int lvio2d_submap_device(lvio2d_submap* sm, const int32_t** meta, const double** ref_pose, const int32_t** ref_n_lines, const double** ref_lines) {
    if (!sm) return LVIO2D_ERR_INVALID_ARG;
    if (meta) *meta = sm->meta.as<int32_t>();
    if (ref_pose) *ref_pose = sm->sub_pose.as<double>();
    if (ref_n_lines) *ref_n_lines = sm->sub_n.as<int32_t>();
    if (ref_lines) *ref_lines = sm->sub_lines.as<double>();
    return LVIO2D_OK;
}

int lvio2d_submap_match(lvio2d_submap* sm, int32_t kk, int32_t max_lines2, const int32_t* n_lines2, const double* lines2, const double* pose2,
                        int32_t* n_match, int32_t* match, double* matched_lines1, double* ref_pose, int32_t on_device) {
    if (!sm) return LVIO2D_ERR_INVALID_ARG;
    lvio2d_ctx* ctx = sm->ctx;
    if (kk < 0 || kk > 2 || max_lines2 < 1 || !n_lines2 || !lines2 || !pose2 || !n_match || !match)
        return fail(ctx, LVIO2D_ERR_INVALID_ARG, "submap_match: null argument, max_lines2 < 1 or kk outside 0..2");
    CK(cudaSetDevice(ctx->device));
    const size_t P = (size_t)sm->n_managers;
    const double4* sub_lines = sm->sub_lines.as<double4>();
    if (on_device) {   // everything is in device memory already: the sampled flavour of lvio2d_match_lines on the resident arrays
        ctx->match_bbox = sm->sub_bbox.as<int32_t>();
        const int rc = lvio2d_match_lines(ctx, &sm->lp, sm->n_managers, kk, nullptr, nullptr, nullptr, sm->line_cap, sm->sub_n.as<int32_t>(),
                                          sm->sub_lines.as<double>(), nullptr, max_lines2, n_lines2, lines2, sm->sub_pose.as<double>(), pose2, n_match,
                                          match, 1);
        ctx->match_bbox = nullptr;
        if (rc != LVIO2D_OK) return rc;
        if (matched_lines1 || ref_pose) {
            submap_gather_kernel<<<sm->n_managers, 128, 0, ctx->stream>>>(sm->n_managers, sm->line_cap, max_lines2, n_match, match, sub_lines,
                                                                          sm->sub_pose.as<double>(), reinterpret_cast<double4*>(matched_lines1), ref_pose);
            ctx->launches += 1;
            CK(cudaGetLastError());
        }
        return LVIO2D_OK;
    }
    // scan 2 and its pose go up, the pairs (and the matched sub-map lines) come back; scan 1 is the resident reference sub-map
    DevBuf* B = ctx->b_ml;
    if (!B[9].ensure(P * sizeof(int32_t)) || !B[10].ensure(P * max_lines2 * sizeof(double4)) || !B[12].ensure(P * 6 * sizeof(double)) ||
        !B[13].ensure(P * sizeof(int32_t)) || !B[14].ensure(P * max_lines2 * 2 * sizeof(int32_t)) || !B[7].ensure(P * max_lines2 * sizeof(double4)) ||
        !B[11].ensure(P * 6 * sizeof(double)))
        return fail(ctx, LVIO2D_ERR_ALLOC, "cudaMalloc(submap_match)");
    CK(cudaMemcpyAsync(B[9].p, n_lines2, P * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(B[10].p, lines2, P * max_lines2 * sizeof(double4), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(B[12].p, pose2, P * 6 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(B[14].p, 0, P * max_lines2 * 2 * sizeof(int32_t), ctx->stream));
    ctx->match_bbox = sm->sub_bbox.as<int32_t>();
    const int rc = lvio2d_match_lines(ctx, &sm->lp, sm->n_managers, kk, nullptr, nullptr, nullptr, sm->line_cap, sm->sub_n.as<int32_t>(),
                                      sm->sub_lines.as<double>(), nullptr, max_lines2, B[9].as<int32_t>(), B[10].as<double>(), sm->sub_pose.as<double>(),
                                      B[12].as<double>(), B[13].as<int32_t>(), B[14].as<int32_t>(), 1);
    ctx->match_bbox = nullptr;
    if (rc != LVIO2D_OK) return rc;
    if (matched_lines1 || ref_pose) {
        submap_gather_kernel<<<sm->n_managers, 128, 0, ctx->stream>>>(sm->n_managers, sm->line_cap, max_lines2, B[13].as<int32_t>(), B[14].as<int32_t>(), sub_lines,
                                                                      sm->sub_pose.as<double>(), matched_lines1 ? B[7].as<double4>() : nullptr,
                                                                      ref_pose ? B[11].as<double>() : nullptr);
        ctx->launches += 1;
        CK(cudaGetLastError());
        if (matched_lines1) CK(cudaMemcpyAsync(matched_lines1, B[7].p, P * max_lines2 * sizeof(double4), cudaMemcpyDeviceToHost, ctx->stream));
        if (ref_pose) CK(cudaMemcpyAsync(ref_pose, B[11].p, P * 6 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaMemcpyAsync(n_match, B[13].p, P * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(match, B[14].p, P * max_lines2 * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return LVIO2D_OK;
}

int lvio2d_eval_ground_factors(lvio2d_ctx* ctx, const double* pose, double* res, double* jac) {
    const double* in[1] = {pose};
    const int len[1] = {6};
    return eval_hook(ctx, in, len, 1, res, 2, jac, 12, 3);
}

}  // extern "C"

// ---- back-end pose graph (pose_graph.cuh)
namespace {
struct PgDeviceLauncher {
    lvio2d_ctx* ctx;
    cudaError_t err = cudaSuccess;
    template <int KID> bool go(const pg::Args& a) {
        const int n = pg::kernel_threads(a, KID);
        if (n <= 0) return true;
        pg::pg_kernel<KID><<<(n + 127) / 128, 128, 0, ctx->stream>>>(a, n);
        ctx->launches += 1;
        return (err = cudaGetLastError()) == cudaSuccess;
    }
    bool run(int kid, const pg::Args& a) {
        switch (kid) {
            case pg::K_COLUMNS: return go<pg::K_COLUMNS>(a);
            case pg::K_ASSEMBLE: return go<pg::K_ASSEMBLE>(a);
            case pg::K_SCALE: return go<pg::K_SCALE>(a);
            case pg::K_TRISOLVE: return go<pg::K_TRISOLVE>(a);
            case pg::K_CAPACITANCE: return go<pg::K_CAPACITANCE>(a);
            case pg::K_COMBINE: return go<pg::K_COMBINE>(a);
            case pg::K_MODEL: return go<pg::K_MODEL>(a);
            case pg::K_COST: return go<pg::K_COST>(a);
        }
        return false;
    }
    bool factor(const pg::Args& a) {
        pg::pg_factor_kernel<<<1, 64, 0, ctx->stream>>>(a);
        ctx->launches += 1;
        return (err = cudaGetLastError()) == cudaSuccess;
    }
    // the opt-in partitioned solve (pose_graph_segments.cuh)
    bool chain_factor(const pg::Args& a, int reduced) {
        pg::pg_chain_factor_kernel<<<reduced ? 1 : a.P, 64, 0, ctx->stream>>>(a, reduced);
        ctx->launches += 1;
        return (err = cudaGetLastError()) == cudaSuccess;
    }
    template <int KID> bool seg_go(const pg::Args& a) {
        const int n = pg::seg_kernel_threads(a, KID);
        if (n <= 0) return true;
        pg::pg_seg_kernel<KID><<<(n + 127) / 128, 128, 0, ctx->stream>>>(a, n);
        ctx->launches += 1;
        return (err = cudaGetLastError()) == cudaSuccess;
    }
    bool seg(int kid, const pg::Args& a) {
        if (kid == pg::KS_TRISOLVE && a.stage) {
            const dim3 grid(a.P, (pg::ncol_x(a) + 127) / 128);
            pg::pg_seg_trisolve_staged_kernel<<<grid, 128, 0, ctx->stream>>>(a);
            ctx->launches += 1;
            return (err = cudaGetLastError()) == cudaSuccess;
        }
        switch (kid) {
            case pg::KS_TRISOLVE: return seg_go<pg::KS_TRISOLVE>(a);
            case pg::KS_REDUCED_BLOCKS: return seg_go<pg::KS_REDUCED_BLOCKS>(a);
            case pg::KS_REDUCED_RHS: return seg_go<pg::KS_REDUCED_RHS>(a);
            case pg::KS_REDUCED_TRISOLVE: return seg_go<pg::KS_REDUCED_TRISOLVE>(a);
            case pg::KS_BACKSUB: return seg_go<pg::KS_BACKSUB>(a);
        }
        return false;
    }
    bool partitioned(const pg::Args& a) { return pg::pg_partitioned_solve(*this, a); }
    bool dense(const pg::Args& a) {
        pg::pg_dense_kernel<<<1, 256, 0, ctx->stream>>>(a);
        ctx->launches += 1;
        return (err = cudaGetLastError()) == cudaSuccess;
    }
    bool reduce(const pg::Args& a) {
        pg::pg_reduce_kernel<<<pg::S_COUNT, 256, 0, ctx->stream>>>(a);
        ctx->launches += 1;
        return (err = cudaGetLastError()) == cudaSuccess;
    }
    // the synchronisation point of an iteration: S_COUNT scalars + the two flags
    bool read(const pg::Args& a, double* scal, int32_t* flags) {
        double* h = ctx->h_pg.data();
        if ((err = cudaMemcpyAsync(h, a.scal, sizeof(double) * pg::S_COUNT, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) return false;
        if ((err = cudaMemcpyAsync(h + pg::S_COUNT, a.flags, sizeof(int32_t) * 2, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) return false;
        if ((err = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return false;
        std::memcpy(scal, h, sizeof(double) * pg::S_COUNT);
        std::memcpy(flags, h + pg::S_COUNT, sizeof(int32_t) * 2);
        return true;
    }
};

// uploads the graph into the two arenas and binds every pointer of `a`
int pg_setup(lvio2d_ctx* ctx, pg::Args& a, int32_t n_poses, const double* poses, int32_t n_edges, const int32_t* edge_index, const double* edge_tf,
             const double* edge_weight, const double* sqrt_info, int32_t ground_p, int32_t ground_q, int fixed) {
    std::vector<int32_t> ints;
    std::memset(&a, 0, sizeof(a));
    a.K = n_poses; a.E = n_edges;
    // the chain is cut into P segments solved concurrently (pose_graph_segments.cuh), P = sqrt(1.5 K) by default, and the
    // segment solves stage their blocks in shared memory: measured on a B200 at 0.55 / 1.21 ms per LM iteration for
    // 1000 / 4000 key frames against 4.7 / 18.7 ms on the single-chain path (profiles/r2_pose_graph.md).
    // LVIO2D_PG_SEGMENTS=<P>|0 and LVIO2D_PG_STAGE=0|1 override (0 = the single-chain path), for A/B runs and the tests.
    const char* seg_env = std::getenv("LVIO2D_PG_SEGMENTS");
    int want_segments = (seg_env && std::strcmp(seg_env, "auto") != 0) ? std::atoi(seg_env) : pg::pg_auto_segments(n_poses);
    a.P = pg::pg_segments(n_poses, want_segments);
    const char* stage_env = std::getenv("LVIO2D_PG_STAGE");
    a.stage = (a.P > 1 && (!stage_env || std::atoi(stage_env) != 0)) ? 1 : 0;
    if (!pg::pg_topology(n_poses, n_edges, edge_index, ints, &a.L, a.P)) return fail(ctx, LVIO2D_ERR_INVALID_ARG, "pose graph: edge index out of range or self edge");
    a.fixed = fixed;
    a.ground_p = ground_p != 0; a.ground_q = ground_q != 0;
    a.C = ctx->C;
    std::memcpy(a.Jn, sqrt_info, sizeof(a.Jn));
    const size_t n_dbl = pg::pg_bind(a, nullptr, nullptr);
    if (n_dbl > (size_t)16 << 27) return fail(ctx, LVIO2D_ERR_DOMAIN, "pose graph: work space above 16 GiB (too many loop edges x key frames)");
    if (!ctx->b_pg[0].ensure(sizeof(int32_t) * ints.size()) || !ctx->b_pg[1].ensure(sizeof(double) * n_dbl)) return fail(ctx, LVIO2D_ERR_ALLOC, "cudaMalloc(pose graph)");
    if (!ctx->h_pg.assign(pg::S_COUNT + 1, 0.0)) return fail(ctx, LVIO2D_ERR_ALLOC, "cudaHostAlloc(pose graph)");
    pg::pg_bind(a, ctx->b_pg[0].as<int32_t>(), ctx->b_pg[1].as<double>());
    CK(cudaMemsetAsync(ctx->b_pg[1].p, 0, sizeof(double) * n_dbl, ctx->stream));
    CK(cudaMemcpyAsync(ctx->b_pg[0].p, ints.data(), sizeof(int32_t) * ints.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(const_cast<double*>(a.edge_tf), edge_tf, sizeof(double) * 12 * (size_t)n_edges, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(const_cast<double*>(a.edge_weight), edge_weight, sizeof(double) * (size_t)n_edges, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(a.x, poses, sizeof(double) * 6 * (size_t)n_poses, cudaMemcpyHostToDevice, ctx->stream));
    // the host arrays above are pageable: the copies have consumed them when the calls return, `ints` included
    CK(cudaStreamSynchronize(ctx->stream));
    return LVIO2D_OK;
}
}  // namespace

extern "C" {
int lvio2d_pose_graph_solve(lvio2d_ctx* ctx, int32_t n_poses, double* poses, int32_t n_edges, const int32_t* edge_index, const double* edge_tf,
                            const double* edge_weight, const double* sqrt_info, int32_t ground_p, int32_t ground_q, int32_t fixed_pose,
                            lvio2d_summary* summary) {
    if (!ctx) return LVIO2D_ERR_INVALID_ARG;
    if (n_poses <= 0 || n_edges < 0 || !poses || !sqrt_info || (n_edges > 0 && (!edge_index || !edge_tf || !edge_weight)))
        return fail(ctx, LVIO2D_ERR_INVALID_ARG, "pose graph: null or empty argument");
    if (fixed_pose < -1 || fixed_pose >= n_poses) return fail(ctx, LVIO2D_ERR_INVALID_ARG, "pose graph: fixed_pose out of range");
    for (int64_t i = 0; i < 6 * (int64_t)n_poses; ++i)
        if (!std::isfinite(poses[i])) return fail(ctx, LVIO2D_ERR_DOMAIN, "pose graph: non-finite pose");
    for (int64_t i = 0; i < 12 * (int64_t)n_edges; ++i)
        if (!std::isfinite(edge_tf[i])) return fail(ctx, LVIO2D_ERR_DOMAIN, "pose graph: non-finite edge transform");
    for (int64_t i = 0; i < n_edges; ++i)
        if (!std::isfinite(edge_weight[i])) return fail(ctx, LVIO2D_ERR_DOMAIN, "pose graph: non-finite edge weight");
    for (int i = 0; i < 36; ++i)
        if (!std::isfinite(sqrt_info[i])) return fail(ctx, LVIO2D_ERR_DOMAIN, "pose graph: non-finite sqrt_info");
    CK(cudaSetDevice(ctx->device));
    pg::Args a;
    const int rc = pg_setup(ctx, a, n_poses, poses, n_edges, edge_index, edge_tf, edge_weight, sqrt_info, ground_p, ground_q, fixed_pose);
    if (rc != LVIO2D_OK) return rc;
    pg::Options opt;
    opt.max_iters = ctx->opt.max_iters; opt.function_tolerance = ctx->opt.function_tolerance; opt.gradient_tolerance = ctx->opt.gradient_tolerance;
    opt.parameter_tolerance = ctx->opt.parameter_tolerance; opt.initial_radius = ctx->opt.initial_radius; opt.max_radius = ctx->opt.max_radius;
    opt.min_radius = ctx->opt.min_radius; opt.min_relative_decrease = ctx->opt.min_relative_decrease; opt.min_lm_diagonal = ctx->opt.min_lm_diagonal;
    opt.max_lm_diagonal = ctx->opt.max_lm_diagonal; opt.max_consecutive_invalid = ctx->opt.max_consecutive_invalid;
    PgDeviceLauncher Lr{ctx};
    lvio2d_summary S;
    if (!pg::pg_minimize(Lr, a, opt, &S)) return fail(ctx, LVIO2D_ERR_CUDA, "pose graph kernel", Lr.err);
    CK(cudaMemcpyAsync(poses, a.x, sizeof(double) * 6 * (size_t)n_poses, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (summary) *summary = S;
    return LVIO2D_OK;
}

int lvio2d_eval_edge_factor(lvio2d_ctx* ctx, const double* tf12, double weight, const double* sqrt_info, const double* pose_i, const double* pose_j,
                            double* res, double* jac) {
    if (!ctx || !tf12 || !sqrt_info || !pose_i || !pose_j || !res || !jac) return LVIO2D_ERR_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    double x[12];
    std::memcpy(x, pose_i, sizeof(double) * 6);
    std::memcpy(x + 6, pose_j, sizeof(double) * 6);
    const int32_t idx[2] = {0, 1};
    pg::Args a;
    const int rc = pg_setup(ctx, a, 2, x, 1, idx, tf12, &weight, sqrt_info, 0, 0, -1);
    if (rc != LVIO2D_OK) return rc;
    PgDeviceLauncher Lr{ctx};
    if (!Lr.run(pg::K_COLUMNS, a)) return fail(ctx, LVIO2D_ERR_CUDA, "pose graph kernel", Lr.err);
    double EJ[78];
    CK(cudaMemcpyAsync(EJ, a.EJ, sizeof(EJ), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int r = 0; r < 6; ++r) {
        res[r] = EJ[72 + r];
        for (int c = 0; c < 12; ++c) jac[r * 12 + c] = EJ[c * 6 + r];
    }
    return LVIO2D_OK;
}
}  // extern "C"
