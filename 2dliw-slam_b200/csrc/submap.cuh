// lvio2d_b200 — the reference sub-map of the laser front-end, resident in device memory.
//
// Replaces the state machine of laser_manager::add_scan (reference src/trajectory/laser_manager.cpp:424-496) for a
// batch of independent laser managers (one per robot / stream), one warp per manager:
//   * the motion filter against last_add_tf (:431-438: |dp| < ref_motion_filter_p and |log(dR)| < ref_motion_filter_q
//     drops the scan),
//   * the first scan founds the reference sub-map at the current pose (:439-451),
//   * every later scan's lines are taken to the frame of the reference sub-map and of the one being spawned,
//     T = T_il^-1 (make_tf(sub)^-1 make_tf(current)) T_il with Eigen's Isometry inverse (R^T, -R^T t), and appended
//     through scan::add_line(p1, p2, false) (:452-470 -> :215-223, :137-154, :197-213),
//   * at ref_n_accumulation / 2 accepted scans the spawning sub-map is founded (:472-478), at ref_n_accumulation it
//     becomes the reference, a new one is founded and the count falls back to ref_n_accumulation / 2 (:479-489).
// add_line(p1, p2, false) refits a line to the three collinear fake points p1, mid, p2 and projects them back: the end
// points return unchanged in x, y with z = 0 (tests/test_ref_frontend.py measures 1e-15 against the reference text), so
// what the kernel keeps of it are its filters — max |z| of the fake points against line_max_dis, the length against
// line_min_len, and at least one 0.05 m sample on the grid — and the append in scan order (warp ballot compaction).
#pragma once
#include "lv_math.cuh"

namespace lv {

struct SubmapArgs {
    int32_t n_managers, line_cap, max_lines, n_accumulation;
    double filter_p, filter_q;
    double T_il[12];
    double line_max_dis, line_min_len, resolution;
    int32_t w, h;
    // state
    int32_t* meta;        // [M][4]: has_ref, has_spawn, current_count, 1 when the last add_scan passed the motion filter
    double* sub_pose;     // [2][M][6]: pose of the reference sub-map, then of the spawning one
    double* last_pose;    // [M][6]: the pose behind last_add_tf
    int32_t* sub_n;       // [2][M]
    double4* sub_lines;   // [2][M][line_cap]
    int4* sub_bbox;       // [2][M][line_cap]: (rmin, rmax, cmin, cmax) of the grid cells the line's 0.05 m samples cover
    // one scan per manager
    const int32_t* n_lines;   // [M]; a negative count skips the manager (no scan for it in this call)
    const double4* lines;     // [M][max_lines]
    const double* pose;       // [M][6]
};

struct IsoD { M3<double> R; V3<double> t; };
__device__ __forceinline__ IsoD iso_from_pose(const double* s) {
    IsoD T;
    T.R = exp_so3(v3<double>(s[3], s[4], s[5]));
    T.t = v3<double>(s[0], s[1], s[2]);
    return T;
}
__device__ __forceinline__ IsoD iso_inv(const IsoD& T) {   // Eigen::Isometry inverse
    IsoD o;
    o.R = transpose(T.R);
    o.t = neg(mul(o.R, T.t));
    return o;
}
__device__ __forceinline__ IsoD iso_mul(const IsoD& A, const IsoD& B) {
    IsoD o;
    o.R = mul(A.R, B.R);
    o.t = mul(A.R, B.t) + A.t;
    return o;
}

// scan::add_line(p1, p2, false) for the `n` lines of the scan, taken through T (or as they are), appended to dst in
// scan order.  Returns the new line count; lines beyond line_cap are counted, not stored.
__device__ __forceinline__ int submap_append(const SubmapArgs& a, double4* dst, int4* dbox, int count, const double4* L, int n, const IsoD* T, int lane) {
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        bool keep = false;
        double4 o = make_double4(0, 0, 0, 0);
        int rmin = 0x7fffffff, rmax = -1, cmin = 0x7fffffff, cmax = -1;
        if (i < n) {
            const double4 l = L[i];
            V3<double> p1 = v3<double>(l.x, l.y, 0.0), p2 = v3<double>(l.z, l.w, 0.0);
            if (T) { p1 = mul(T->R, p1) + T->t; p2 = mul(T->R, p2) + T->t; }
            const double zmax = fmax(fmax(fabs(p1.z), fabs(p2.z)), fabs(0.5 * (p1.z + p2.z)));
            const double dx = p2.x - p1.x, dy = p2.y - p1.y;
            const double len = sqrt(dx * dx + dy * dy);
            if (!(zmax > a.line_max_dis) && !(len < a.line_min_len)) {
                const double ux = dx / len, uy = dy / len;
                // the same walk lvio2d_match_lines takes over the line (sampled flavour): its box is kept with the line, so a
                // match against the sub-map does not walk ~800 lines again for every scan
                for (double tr = 0.0; tr <= len; tr += 0.05) {
                    const double qx = p1.x + ux * tr, qy = p1.y + uy * tr;
                    const int c = (int)(qx / a.resolution + a.w / 2), r = (int)(qy / a.resolution + a.h / 2);
                    if (r >= 0 && r < a.h && c >= 0 && c < a.w) {
                        keep = true;
                        rmin = min(rmin, r); rmax = max(rmax, r); cmin = min(cmin, c); cmax = max(cmax, c);
                    }
                }
                o = make_double4(p1.x, p1.y, p2.x, p2.y);
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int slot = count + __popc(m & ((1u << lane) - 1u));
            if (slot < a.line_cap) { dst[slot] = o; dbox[slot] = make_int4(rmin, rmax, cmin, cmax); }
        }
        count += __popc(m);
    }
    return count;
}

__global__ void __launch_bounds__(128) submap_add_scan_kernel(SubmapArgs a) {
    const int lane = threadIdx.x & 31;
    const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (m >= a.n_managers) return;
    const int M = a.n_managers;
    int32_t* meta = a.meta + 4 * (size_t)m;
    if (a.n_lines[m] < 0) { if (lane == 0) meta[3] = 0; return; }
    const int n = min(a.n_lines[m], a.max_lines);
    const double4* L = a.lines + (size_t)m * a.max_lines;
    const double* cur = a.pose + 6 * (size_t)m;
    double* ref_pose = a.sub_pose + 6 * (size_t)m;
    double* spawn_pose = a.sub_pose + 6 * ((size_t)M + m);
    double* last = a.last_pose + 6 * (size_t)m;
    double4* ref_lines = a.sub_lines + (size_t)m * a.line_cap;
    double4* spawn_lines = a.sub_lines + ((size_t)M + m) * a.line_cap;
    int4* ref_box = a.sub_bbox + (size_t)m * a.line_cap;
    int4* spawn_box = a.sub_bbox + ((size_t)M + m) * a.line_cap;
    int has_ref = meta[0], has_spawn = meta[1], count = meta[2];
    int n_ref = a.sub_n[m], n_spawn = a.sub_n[M + m];
    __syncwarp();   // every lane has read the state before lane 0 rewrites it

    const IsoD Tc = iso_from_pose(cur);
    auto set_pose = [&](double* dst) { if (lane < 6) dst[lane] = cur[lane]; };
    auto commit = [&](int added) {
        __syncwarp();
        if (lane == 0) {
            meta[0] = has_ref; meta[1] = has_spawn; meta[2] = count; meta[3] = added;
            a.sub_n[m] = n_ref; a.sub_n[M + m] = n_spawn;
        }
    };
    if (!has_ref) {
        set_pose(ref_pose);
        set_pose(last);
        n_ref = submap_append(a, ref_lines, ref_box, 0, L, n, nullptr, lane);
        has_ref = 1; count = 1;
        commit(1);
        return;
    }
    {
        const IsoD d = iso_mul(iso_inv(iso_from_pose(last)), Tc);
        if (norm(d.t) < a.filter_p && norm(log_so3(d.R)) < a.filter_q) { commit(0); return; }
    }
    IsoD Til;
    { const Iso t = load_iso(a.T_il); Til.R = t.R; Til.t = t.t; }
    const IsoD Tli = iso_inv(Til);
    {
        const IsoD T = iso_mul(iso_mul(Tli, iso_mul(iso_inv(iso_from_pose(ref_pose)), Tc)), Til);
        n_ref = submap_append(a, ref_lines, ref_box, n_ref, L, n, &T, lane);
    }
    if (has_spawn) {
        const IsoD T = iso_mul(iso_mul(Tli, iso_mul(iso_inv(iso_from_pose(spawn_pose)), Tc)), Til);
        n_spawn = submap_append(a, spawn_lines, spawn_box, n_spawn, L, n, &T, lane);
    }
    ++count;
    if (!has_spawn && count == a.n_accumulation / 2) {
        __syncwarp();
        set_pose(spawn_pose);
        n_spawn = submap_append(a, spawn_lines, spawn_box, 0, L, n, nullptr, lane);
        has_spawn = 1;
    }
    if (count == a.n_accumulation) {
        // the spawning sub-map becomes the reference (a null one, only possible with ref_n_accumulation < 4, leaves the
        // manager without a reference: the next scan founds it again, as the reference's null pointer does)
        __syncwarp();
        const int stored = min(n_spawn, a.line_cap);
        for (int i = lane; i < stored; i += 32) { ref_lines[i] = spawn_lines[i]; ref_box[i] = spawn_box[i]; }
        if (lane < 6) ref_pose[lane] = spawn_pose[lane];
        has_ref = has_spawn;
        n_ref = has_spawn ? n_spawn : 0;
        __syncwarp();
        set_pose(spawn_pose);
        n_spawn = submap_append(a, spawn_lines, spawn_box, 0, L, n, nullptr, lane);
        has_spawn = 1;
        count = a.n_accumulation / 2;
    }
    __syncwarp();
    set_pose(last);
    commit(1);
}

// the matched sub-map lines' end points and the sub-map pose, so that match_with_ref needs no copy of the sub-map itself
__global__ void submap_gather_kernel(int n_managers, int line_cap, int max_lines2, const int32_t* n_match, const int32_t* match, const double4* sub_lines,
                                     const double* sub_pose, double4* lines1_out, double* pose_out) {
    const int m = blockIdx.x;
    if (m >= n_managers) return;
    if (pose_out && threadIdx.x < 6) pose_out[6 * m + threadIdx.x] = sub_pose[6 * m + threadIdx.x];
    if (!lines1_out) return;
    const int n = min(n_match[m], max_lines2);
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const int j = match[((size_t)m * max_lines2 + k) * 2];
        lines1_out[(size_t)m * max_lines2 + k] = (j >= 0 && j < line_cap) ? sub_lines[(size_t)m * line_cap + j] : make_double4(0, 0, 0, 0);
    }
}

}  // namespace lv
