// scan_lines.cuh — scan points -> line segments for a batch of scans, one warp per scan.
//
// Replaces laser_manager::spawn_scan (reference src/trajectory/laser_manager.cpp:350-422) together with the fit and
// the filters of scan::add_line (:137-154; fit_line_by_least_square :19-36, create_line :62-94, project_to_line
// :8-18, is_continuous / clac_cos / clac_angle :96-120) and the grid-validity test of scan::xy_to_index /
// is_index_valid (src/trajectory/laser_type.h:34-41).  Output = scan::lines in the reference's order.
//
// The reference walks each scan sequentially; here every stage that has no loop-carried dependency is spread over
// the 32 lanes (points are read through L1, per-point scratch lives in a global workspace that stays L1/L2-resident):
//   (A) continuity split      break flags -> segment start / end of every point by a max-scan / min-scan over the scan
//   (B) corner response       cos of the angle at point i between points i-3 and i+3 (clipped to the segment)
//   (C) non-maximum suppression; the reference's "i += step" after a maximum is a no-op (two strict maxima of a
//       +-3 window cannot lie within 3 of each other), so the flags are exact without the sequential walk
//   (D) candidate list        [segment start, maxima..., segment end] per segment, by ordered warp compaction
//   (E) tolerance-angle merge sequential per segment (loop-carried `last_end_index`): one lane per segment
//   (F) fit + filters         one lane per candidate line: moments of [x y 1] -> smallest eigenvector of the 3x3 Gram
//       matrix by cyclic Jacobi (= V.col(2) of the reference's JacobiSVD), max distance, projections, length, grid test
// Bound: fp64 ALU / latency (sqrt, div, acos per point); 16 B in and < 1 B out per point, far below the HBM roofline.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lvio2d.h"
#include "lv_math.cuh"

namespace lv {

struct ScanLinesArgs {
    const int64_t* point_offset;   // [S+1] (or [S] with point_count)
    const int32_t* point_count;    // [S] or nullptr
    const double2* points;         // [N]
    const double* point_z;         // [N] or nullptr (z = 0)
    int32_t n_scans, max_lines;
    double continuous_threshold, max_tolerance_angle, max_dis, min_len, resolution;
    int32_t w, h;
    // workspace, indexed like the points (x2 for the candidate arrays)
    double* resp;                  // [N]
    int32_t* seg_s;                // [N]
    int32_t* seg_e;                // [N]
    int32_t* cand;                 // [2N + 2S]
    int32_t* lstart;               // [2N + 2S]
    int32_t* seg_first;            // [N + S]
    // outputs
    int32_t* n_lines;              // [S]
    double* lines;                 // [S][max_lines][4]
    double* abc;                   // [S][max_lines][3]
    int32_t* index_range;          // [S][max_lines][2]
};

constexpr double kLineEpsilo = 0.0008;   // laser_manager.cpp:3

struct P3d { double x, y, z; };
__device__ __forceinline__ double line_clac_cos(P3d pj, P3d pi, P3d pk) {
    const double ax = pi.x - pj.x, ay = pi.y - pj.y, az = pi.z - pj.z, bx = pk.x - pj.x, by = pk.y - pj.y, bz = pk.z - pj.z;
    const double na = sqrt(ax * ax + ay * ay + az * az), nb = sqrt(bx * bx + by * by + bz * bz);
    if (na < kLineEpsilo) return -1.0;
    if (nb < kLineEpsilo) return -1.0;
    return (ax / na) * (bx / nb) + (ay / na) * (by / nb) + (az / na) * (bz / nb);
}
__device__ __forceinline__ double lines_warp_sum(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
__device__ __forceinline__ double lines_warp_max(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, d));
    return v;
}

// eigenvector of the smallest eigenvalue of the symmetric 3x3 matrix (m00 m01 m02; . m11 m12; . . m22), cyclic Jacobi
__device__ __forceinline__ void smallest_eigvec3(double m00, double m01, double m02, double m11, double m12, double m22, double* v) {
    double A[3][3] = {{m00, m01, m02}, {m01, m11, m12}, {m02, m12, m22}};
    double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
#pragma unroll 1
    for (int sweep = 0; sweep < 24; ++sweep) {
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        if (off == 0.0) break;
        bool any = false;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
            const double apq = A[p][q];
            if (fabs(apq) <= 1e-19 * sqrt(fabs(A[p][p] * A[q][q])) || apq == 0.0) continue;
            any = true;
            const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
            const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
            const int r = 3 - p - q;
            const double app = A[p][p], aqq = A[q][q], arp = A[r][p], arq = A[r][q];
            A[p][p] = app - t * apq;
            A[q][q] = aqq + t * apq;
            A[p][q] = A[q][p] = 0.0;
            A[r][p] = A[p][r] = c * arp - s * arq;
            A[r][q] = A[q][r] = s * arp + c * arq;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double vp = V[i][p], vq = V[i][q];
                V[i][p] = c * vp - s * vq;
                V[i][q] = s * vp + c * vq;
            }
        }
        if (!any) break;
    }
    int k = 0;
    double dk = A[0][0];
    if (A[1][1] < dk) { k = 1; dk = A[1][1]; }
    if (A[2][2] < dk) k = 2;
    v[0] = k == 0 ? V[0][0] : (k == 1 ? V[0][1] : V[0][2]);
    v[1] = k == 0 ? V[1][0] : (k == 1 ? V[1][1] : V[1][2]);
    v[2] = k == 0 ? V[2][0] : (k == 1 ? V[2][1] : V[2][2]);
}

__global__ void __launch_bounds__(128) scan_lines_kernel(ScanLinesArgs a) {
    const int lane = threadIdx.x & 31;
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= a.n_scans) return;
    const int64_t p0 = a.point_offset[s];
    const int n = a.point_count ? a.point_count[s] : (int)(a.point_offset[s + 1] - p0);
    const double2* P = a.points + p0;
    const double* PZ = a.point_z ? a.point_z + p0 : nullptr;
    auto pt = [&](int i) -> P3d { const double2 q = P[i]; P3d r; r.x = q.x; r.y = q.y; r.z = PZ ? PZ[i] : 0.0; return r; };
    double* resp = a.resp + p0;
    int32_t* seg_s = a.seg_s + p0;
    int32_t* seg_e = a.seg_e + p0;
    int32_t* cand = a.cand + 2 * p0 + 2 * s;
    int32_t* lstart = a.lstart + 2 * p0 + 2 * s;
    int32_t* seg_first = a.seg_first + p0 + s;
    if (n == 0) { if (lane == 0) a.n_lines[s] = 0; return; }

    // ---- (A) segment start of every point (forward max-scan of the break positions) ...
    int carry = 0;
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        bool brk = false;
        if (i < n) {
            if (i == 0) brk = true;
            else {
                const P3d u = pt(i - 1), v = pt(i);
                const double dx = u.x - v.x, dy = u.y - v.y, dz = u.z - v.z;
                brk = !(sqrt(dx * dx + dy * dy + dz * dz) <= a.continuous_threshold);
            }
        }
        int v = brk ? i : -1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v = max(v, t); }
        v = max(v, carry);
        if (i < n) seg_s[i] = v;
        carry = __shfl_sync(0xffffffffu, v, 31);
    }
    __syncwarp();
    // ... and its segment end (backward min-scan of the end positions)
    carry = n - 1;
    for (int base = ((n - 1) / 32) * 32; base >= 0; base -= 32) {
        const int i = base + lane;
        const bool is_end = i < n && (i == n - 1 || seg_s[i + 1] == i + 1);
        int v = is_end ? i : 0x7fffffff;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_down_sync(0xffffffffu, v, d); if (lane + d < 32) v = min(v, t); }
        v = min(v, carry);
        if (i < n) seg_e[i] = v;
        carry = __shfl_sync(0xffffffffu, v, 0);
    }
    __syncwarp();

    // ---- (B) corner responses (laser_manager.cpp:378-382); end points of a segment keep -1
    for (int i = lane; i < n; i += 32) {
        const int ss = seg_s[i], ee = seg_e[i];
        double r = -1.0;
        if (i > ss && i < ee) r = line_clac_cos(pt(i), pt(max(i - 3, ss)), pt(min(i + 3, ee)));
        resp[i] = r;
    }
    __syncwarp();

    // ---- (C) + (D) strict local maxima and the ordered candidate list [s, maxima..., e] of every segment
    int total = 0, n_seg = 0;
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        bool is_s = false, is_e = false, is_max = false;
        if (i < n) {
            const int ss = seg_s[i], ee = seg_e[i];
            is_s = i == ss; is_e = i == ee;
            if (i > ss && i < ee) {
                is_max = true;
                const double ri = resp[i];
                const int bj = max(i - 3, ss + 1), ej = min(i + 3, ee - 1);
                for (int j = bj; j <= ej; ++j)
                    if (j != i && resp[j] >= ri) { is_max = false; break; }
            }
        }
        const int cnt = (is_s ? 1 : 0) + (is_max ? 1 : 0) + (is_e ? 1 : 0);
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
        int pos = total + incl - cnt;
        const unsigned smask = __ballot_sync(0xffffffffu, is_s);
        if (is_s) { seg_first[n_seg + __popc(smask & ((1u << lane) - 1u))] = pos; cand[pos++] = i; }
        if (is_max) cand[pos++] = i;
        if (is_e) cand[pos++] = i;
        total += __shfl_sync(0xffffffffu, incl, 31);
        n_seg += __popc(smask);
    }
    __syncwarp();

    // ---- (E) tolerance-angle merge (laser_manager.cpp:404-413): lstart[pos] = index1 of the line that ends at
    // cand[pos], or -1
    for (int k = lane; k < n_seg; k += 32) {
        const int fa = seg_first[k], fb = (k + 1 < n_seg) ? seg_first[k + 1] : total;
        int last = 0;
        lstart[fa] = -1;
        for (int i = 1; i + 1 < fb - fa; ++i) {
            const double angle = acos(line_clac_cos(pt(cand[fa + i]), pt(cand[fa + last]), pt(cand[fa + i + 1])));
            if (fabs(angle) < a.max_tolerance_angle) { lstart[fa + i] = cand[fa + last]; last = i; }
            else lstart[fa + i] = -1;
        }
        lstart[fb - 1] = cand[fa + last];
    }
    __syncwarp();

    // ---- (F) fit + filters: one lane per candidate line (32 lines at a time; a line is 3 .. a few hundred points),
    // accepted lines appended in the reference's order by a ballot compaction
    int count = 0;
    double* out_lines = a.lines + (size_t)s * a.max_lines * 4;
    double* out_abc = a.abc + (size_t)s * a.max_lines * 3;
    int32_t* out_rng = a.index_range + (size_t)s * a.max_lines * 2;
    for (int base = 0; base < total; base += 32) {
        const int pos = base + lane;
        const int i1 = pos < total ? lstart[pos] : -1;
        const int i2 = pos < total ? cand[pos] : -1;
        bool keep = false;
        double v[3] = {0.0, 0.0, 0.0}, e1x = 0.0, e1y = 0.0, e2x = 0.0, e2y = 0.0;
        if (i1 >= 0 && i2 - i1 >= 2) {                                   // scan::add_line: at least 3 points
            double sxx = 0, sxy = 0, sx = 0, syy = 0, sy = 0;
            for (int i = i1; i <= i2; ++i) {
                const double2 p = P[i];
                sxx += p.x * p.x; sxy += p.x * p.y; sx += p.x; syy += p.y * p.y; sy += p.y;
            }
            smallest_eigvec3(sxx, sxy, sx, syy, sy, (double)(i2 - i1 + 1), v);
            {   // sign convention of the ABI: the component of largest magnitude is positive
                int k = 0;
                if (fabs(v[1]) > fabs(v[k])) k = 1;
                if (fabs(v[2]) > fabs(v[k])) k = 2;
                if (v[k] < 0.0) { v[0] = -v[0]; v[1] = -v[1]; v[2] = -v[2]; }
            }
            // create_line: two points of the fitted line
            double q1x, q1y, q2x, q2y;
            if (fabs(v[1]) < 0.5) { q1y = 0.0; q1x = -v[2] / v[0]; q2y = 1.0; q2x = (-v[2] - v[1]) / v[0]; }
            else { q1x = 0.0; q2x = 1.0; q1y = -v[2] / v[1]; q2y = (-v[2] - v[0]) / v[1]; }
            // e_laser::dis_from_line (common.h:86-95): rejection of p - q2 from the unit direction
            double ux = q2x - q1x, uy = q2y - q1y;
            const double ulen = sqrt(ux * ux + uy * uy);
            if (ulen * ulen > 0.0) { ux /= ulen; uy /= ulen; }
            double err = 0.0;
            bool on_grid = false;
            for (int i = i1; i <= i2; ++i) {
                const double2 p = P[i];
                const double rx = p.x - q2x, ry = p.y - q2y;
                const double t = ux * rx + uy * ry;
                const double ex = rx - t * ux, ey = ry - t * uy, ez = PZ ? PZ[i] : 0.0;   // the fitted line lies in z = 0
                err = fmax(err, sqrt(ex * ex + ey * ey + ez * ez));
                const int c = (int)(p.x / a.resolution + a.w / 2), r = (int)(p.y / a.resolution + a.h / 2);
                on_grid = on_grid || (r >= 0 && r < a.h && c >= 0 && c < a.w);
            }
            // project_to_line of the first and last point
            const double2 pa = P[i1], pb = P[i2];
            if (ulen < kLineEpsilo) { e1x = pa.x; e1y = pa.y; e2x = pb.x; e2y = pb.y; }
            else {
                const double ta = (pa.x - q1x) * ux + (pa.y - q1y) * uy, tb = (pb.x - q1x) * ux + (pb.y - q1y) * uy;
                e1x = q1x + ta * ux; e1y = q1y + ta * uy; e2x = q1x + tb * ux; e2y = q1y + tb * uy;
            }
            const double len = sqrt((e1x - e2x) * (e1x - e2x) + (e1y - e2y) * (e1y - e2y));
            keep = !(err > a.max_dis) && !(len < a.min_len) && on_grid;
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int slot = count + __popc(m & ((1u << lane) - 1u));
            if (slot < a.max_lines) {
                out_lines[4 * slot] = e1x; out_lines[4 * slot + 1] = e1y; out_lines[4 * slot + 2] = e2x; out_lines[4 * slot + 3] = e2y;
                out_abc[3 * slot] = v[0]; out_abc[3 * slot + 1] = v[1]; out_abc[3 * slot + 2] = v[2];
                out_rng[2 * slot] = i1; out_rng[2 * slot + 1] = i2;
            }
        }
        count += __popc(m);
    }
    if (lane == 0) a.n_lines[s] = count;
}

// =====================================================================================================
// LaserScan ranges -> (de-skewed) points: convert::laser_to_point_times (reference src/utilies/common.cpp:4-40) +
// sensor::laser::correct (src/trajectory/sensor.h:51-94).  One warp per scan, fixed output stride of n_beams slots.
//   (1) every lane converts beams (float angle arithmetic as declared in the reference, double cos/sin) into candidate
//       points, written in place at their beam slot;
//   (2) the "closer than 1 cm to the last KEPT point" filter is the only loop-carried step: a chunk of 32 beams whose
//       readings are all valid and all >= 1 cm from their immediate predecessor is kept wholesale (then the
//       predecessor IS the last kept point and the decision is exact); any other chunk is walked by one lane;
//   (3) de-skew of the kept points, in parallel.
struct ScanPointsArgs {
    const float* ranges;                 // [S][n_beams]
    const lvio2d_scan_header* headers;   // [S]
    int32_t n_scans, n_beams, deskew;
    int32_t* point_count;                // [S]
    double2* points;                     // [S][n_beams]
    double* point_z;                     // [S][n_beams]
    double* point_time;                  // [S][n_beams] or nullptr
    int32_t* beam_index;                 // [S][n_beams] workspace: beam of every kept point
};

__global__ void __launch_bounds__(128) scan_points_kernel(ScanPointsArgs a) {
    const int lane = threadIdx.x & 31;
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= a.n_scans) return;
    const lvio2d_scan_header h = a.headers[s];
    const int nb = a.n_beams;
    const float* rg = a.ranges + (size_t)s * nb;
    double2* out = a.points + (size_t)s * nb;
    double* outz = a.point_z + (size_t)s * nb;
    int32_t* bidx = a.beam_index + (size_t)s * nb;
    int kept = 0;
    double lastx = 0.0, lasty = 0.0;   // last kept point (uniform across the warp)
    for (int base = 0; base < nb; base += 32) {
        const int i = base + lane;
        bool valid = false;
        double x = 0.0, y = 0.0;
        if (i < nb) {
            const float r = rg[i];
            valid = !isnan(r) && !isinf(r) && (double)r > 0.1;
            const float ang = __fadd_rn(h.angle_min, __fmul_rn((float)i, h.angle_increment));   // two roundings, no FMA
            double sn, cs;
            sincos((double)ang, &sn, &cs);
            x = cs * (double)r; y = sn * (double)r;
        }
        // distance to the immediate predecessor (previous lane, or the last kept point for lane 0)
        double px = __shfl_up_sync(0xffffffffu, x, 1), py = __shfl_up_sync(0xffffffffu, y, 1);
        if (lane == 0) { px = lastx; py = lasty; }
        const bool has_prev = lane > 0 || kept > 0;
        const bool far = !has_prev || !(sqrt((x - px) * (x - px) + (y - py) * (y - py)) < 0.01);
        const unsigned in_mask = __ballot_sync(0xffffffffu, i < nb);
        const unsigned ok_mask = __ballot_sync(0xffffffffu, valid && far);
        if (ok_mask == in_mask) {
            // the whole chunk is kept
            if (i < nb) { out[kept + lane] = make_double2(x, y); bidx[kept + lane] = i; }
            const int cnt = __popc(in_mask);
            lastx = __shfl_sync(0xffffffffu, x, cnt - 1); lasty = __shfl_sync(0xffffffffu, y, cnt - 1);
            kept += cnt;
        } else {
            // sequential walk of this chunk (every lane runs it redundantly on shuffled values: no divergence, no smem)
            const unsigned vmask = __ballot_sync(0xffffffffu, valid);
            for (int j = 0; j < 32; ++j) {
                const double xj = __shfl_sync(0xffffffffu, x, j), yj = __shfl_sync(0xffffffffu, y, j);
                if (!((vmask >> j) & 1u)) continue;
                if (kept > 0 && sqrt((xj - lastx) * (xj - lastx) + (yj - lasty) * (yj - lasty)) < 0.01) continue;
                if (lane == 0) { out[kept] = make_double2(xj, yj); bidx[kept] = base + j; }
                lastx = xj; lasty = yj;
                ++kept;
            }
        }
        __syncwarp();
    }
    if (lane == 0) a.point_count[s] = kept;
    __syncwarp();
    // times and de-skew of the kept points
    const V3<double> lin = v3<double>(h.linear[0], h.linear[1], h.linear[2]), ang = v3<double>(h.angular[0], h.angular[1], h.angular[2]);
    for (int k = lane; k < kept; k += 32) {
        const int i = bidx[k];
        const double t = h.stamp + (double)((float)(size_t)i * h.time_increment);
        double2 p = out[k];
        double z = 0.0;
        if (a.deskew) {
            const double dt = t - h.stamp;
            const M3<double> R = exp_so3(scale(ang, dt));
            const V3<double> q = mul(R, v3<double>(p.x, p.y, 0.0));
            p.x = q.x + dt * lin.x; p.y = q.y + dt * lin.y; z = q.z + dt * lin.z;
            out[k] = p;
        }
        outz[k] = z;
        if (a.point_time) a.point_time[(size_t)s * nb + k] = t;
    }
}

// =====================================================================================================
// Segment association: laser_manager::do_match (reference src/trajectory/laser_manager.cpp:244-348), one warp per scan
// pair.  scan 1's line_map (a 1001 x 1001 grid of line lists in the reference) is never built: for a line of scan 2
// only the (2a+1)^2 cells around its transformed mid point matter, so every lane takes lines of scan 1, walks their
// rasterisation (packed cell indices of the line's own points, precomputed once per pair; or the 0.05 m samples of the
// sub-map flavour) and records which of those cells the line covers as a bit mask.  Candidate order of the reference =
// (cell in raster order, line index), so "first candidate of smallest angle" = lexicographic minimum of
// (angle, lowest covered cell, line index), reduced across the warp.
struct MatchLinesArgs {
    const int64_t* point_offset1; const int32_t* point_count1; const double2* points1;   // points1 == nullptr: sampled flavour
    int32_t max_lines1; const int32_t* n_lines1; const double4* lines1; const int32_t* index_range1;
    int32_t max_lines2; const int32_t* n_lines2; const double4* lines2;
    const double* pose1; const double* pose2;    // [P][6]
    double T_il[12];
    int32_t n_pairs, kk, w, h;
    double resolution;
    int32_t* cells;     // workspace [N1]: r << 16 | c of every scan-1 point, -1 when off the grid
    int32_t* bbox;      // workspace [P][max_lines1][4]: (rmin, rmax, cmin, cmax) of the cells every scan-1 line covers
    int32_t bbox_ready; // 1: `bbox` already holds them (the resident sub-map keeps the box of every line it appended)
    double* diss;       // workspace [P][max_lines2]
    int32_t* prov;      // workspace [P][max_lines2][2]: pairs before the distance filter
    int32_t* n_match;   // [P]
    int32_t* match;     // [P][max_lines2][2]
};

__device__ __forceinline__ int match_cell(const MatchLinesArgs& a, double x, double y) {
    const int c = (int)(x / a.resolution + a.w / 2), r = (int)(y / a.resolution + a.h / 2);
    return (r >= 0 && r < a.h && c >= 0 && c < a.w) ? ((r << 16) | c) : -1;
}
// e_laser::dis_from_line (common.h:86-95) for a 3-D point against a line in the plane z = 0
__device__ __forceinline__ double match_dis_from_line(const V3<double>& p, double4 l) {
    double ux = l.z - l.x, uy = l.w - l.y;
    const double n = sqrt(ux * ux + uy * uy);
    if (n * n > 0.0) { ux /= n; uy /= n; }
    const double rx = p.x - l.z, ry = p.y - l.w, rz = p.z;
    const double t = ux * rx + uy * ry;
    const double ex = rx - t * ux, ey = ry - t * uy;
    return sqrt(ex * ex + ey * ey + rz * rz);
}

// WPP = warps per pair: 1 (one warp per pair, four pairs per CTA: the batched shape) or 4 (the CTA shares one pair: its warps
// take the lines of scan 2 round robin and warp 0 gathers the results in order — the same pairs, a quarter of the latency,
// for the few-robots case where one warp walking a 1600-line sub-map 47 times is the whole critical path)
template <int WPP>
__global__ void __launch_bounds__(128) match_lines_kernel(MatchLinesArgs a) {
    __shared__ int queue_all[4][64];
    int* queue = queue_all[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
    const int p = WPP == 1 ? blockIdx.x * (blockDim.x >> 5) + wid : blockIdx.x;
    if (p >= a.n_pairs) return;
    const int tl = WPP == 1 ? lane : threadIdx.x;          // thread index within the group that owns the pair
    constexpr int TG = 32 * WPP;
    auto group_sync = [&]() { if (WPP == 1) __syncwarp(); else __syncthreads(); };
    const int n1 = min(a.n_lines1[p], a.max_lines1), n2 = min(a.n_lines2[p], a.max_lines2);
    const double4* L1 = a.lines1 + (size_t)p * a.max_lines1;
    const double4* L2 = a.lines2 + (size_t)p * a.max_lines2;
    const int32_t* R1 = a.index_range1 ? a.index_range1 + (size_t)p * a.max_lines1 * 2 : nullptr;
    double* diss = a.diss + (size_t)p * a.max_lines2;
    int32_t* prov = a.prov + (size_t)p * a.max_lines2 * 2;
    int32_t* out = a.match + (size_t)p * a.max_lines2 * 2;
    // T_1_2 = (make_tf(p1, q1) T_il)^-1 (make_tf(p2, q2) T_il)
    const Iso Til = load_iso(a.T_il);
    const double* s1 = a.pose1 + 6 * (size_t)p;
    const double* s2 = a.pose2 + 6 * (size_t)p;
    const M3<double> Ra = mul(exp_so3(v3<double>(s1[3], s1[4], s1[5])), Til.R), Rb = mul(exp_so3(v3<double>(s2[3], s2[4], s2[5])), Til.R);
    const V3<double> ta = mul(exp_so3(v3<double>(s1[3], s1[4], s1[5])), Til.t) + v3<double>(s1[0], s1[1], s1[2]);
    const V3<double> tb = mul(exp_so3(v3<double>(s2[3], s2[4], s2[5])), Til.t) + v3<double>(s2[0], s2[1], s2[2]);
    const M3<double> R12 = mul_tn(Ra, Rb);
    const V3<double> t12 = mul_t(Ra, tb - ta);
    auto tf = [&](double x, double y) -> V3<double> { return mul(R12, v3<double>(x, y, 0.0)) + t12; };

    // packed cells of scan 1's points
    int64_t q0 = 0;
    int np1 = 0;    // points of scan 1: index ranges are clamped to it (device buffers are not validated on the host)
    if (a.points1) {
        q0 = a.point_offset1[p];
        const int np = a.point_count1 ? a.point_count1[p] : (int)(a.point_offset1[p + 1] - q0);
        np1 = np;
        for (int i = tl; i < np; i += TG) { const double2 q = a.points1[q0 + i]; a.cells[q0 + i] = match_cell(a, q.x, q.y); }
        group_sync();
    }
    // bounding box of every scan-1 line's cells: a line whose box misses the neighbourhood cannot cover any of its cells,
    // which prunes almost every (line of scan 2, line of scan 1) pair with four comparisons
    int32_t* bb = a.bbox + (size_t)p * a.max_lines1 * 4;
    for (int j = tl; j < (a.bbox_ready ? 0 : n1); j += TG) {
        int rmin = 0x7fffffff, rmax = -1, cmin = 0x7fffffff, cmax = -1;
        auto grow = [&](int cell) {
            if (cell < 0) return;
            const int r = cell >> 16, c = cell & 0xffff;
            rmin = min(rmin, r); rmax = max(rmax, r); cmin = min(cmin, c); cmax = max(cmax, c);
        };
        if (a.points1) {
            for (int q = max(R1[2 * j], 0); q <= min(R1[2 * j + 1], np1 - 1); ++q) grow(a.cells[q0 + q]);
        } else {
            const double4 l1 = L1[j];
            const double dx = l1.z - l1.x, dy = l1.w - l1.y, len = sqrt(dx * dx + dy * dy);
            double ux = dx, uy = dy;
            if (len * len > 0.0) { ux /= len; uy /= len; }
            for (double tr = 0; tr <= len; tr += 0.05) grow(match_cell(a, l1.x + ux * tr, l1.y + uy * tr));
        }
        bb[4 * j] = rmin; bb[4 * j + 1] = rmax; bb[4 * j + 2] = cmin; bb[4 * j + 3] = cmax;
    }
    group_sync();
    const int aa = 1 + a.kk, side = 2 * aa + 1;
    int count = 0;
    for (int i = (WPP == 1 ? 0 : wid); i < n2; i += WPP) {
        const double4 l2 = L2[i];
        const V3<double> tm = tf((l2.x + l2.z) / 2.0, (l2.y + l2.w) / 2.0);
        const int c = (int)(tm.x / a.resolution + a.w / 2), r = (int)(tm.y / a.resolution + a.h / 2);
        const V3<double> e1 = tf(l2.x, l2.y), e2 = tf(l2.z, l2.w);
        V3<double> v2 = e2 - e1;
        { const double n = norm(v2); if (n * n > 0.0) v2 = scale(v2, 1.0 / n); }   // (Eigen normalized(): v / norm)
        double best_angle = 1e300;
        int best_cell = 64, best_j = 0x7fffffff;
        // Lines of scan 1 whose box reaches the window are rare and unevenly spread over j (ncu on a 1600-line sub-map: 4.3
        // of 32 lanes active): the warp first collects their indices (ballot compaction into a 64-entry queue in shared
        // memory) and evaluates them 32 at a time, one candidate per lane.  The winner is a lexicographic minimum, so the
        // order of evaluation does not matter.
        int queued = 0;
        auto evaluate = [&](const int j) {
            unsigned long long mask = 0ull;
            auto cover = [&](int cell) {
                if (cell < 0) return;
                const int dr = (cell >> 16) - r + aa, dc = (cell & 0xffff) - c + aa;
                if (dr >= 0 && dr < side && dc >= 0 && dc < side) mask |= 1ull << (dr * side + dc);
            };
            const double4 l1 = L1[j];
            if (a.points1) {
                for (int q = max(R1[2 * j], 0); q <= min(R1[2 * j + 1], np1 - 1); ++q) cover(a.cells[q0 + q]);
            } else {
                const double dx = l1.z - l1.x, dy = l1.w - l1.y, len = sqrt(dx * dx + dy * dy);
                double ux = dx, uy = dy;
                if (len * len > 0.0) { ux /= len; uy /= len; }
                // Only samples inside the (2 aa + 1)^2 cell window can set a bit: clip the segment against that window (slab
                // test, window grown by 1 mm) and evaluate only the samples within one sample spacing of the clipped
                // parameter range.  The sample positions themselves stay the accumulated 0.05 m steps of the reference
                // loop, so the covered cells are exactly those of the full walk; a sub-map wall of 10 m costs ~200 additions
                // instead of 200 cell evaluations.
                const double bx0 = (double)(c - aa - a.w / 2) * a.resolution - 1e-3, bx1 = (double)(c + aa + 1 - a.w / 2) * a.resolution + 1e-3;
                const double by0 = (double)(r - aa - a.h / 2) * a.resolution - 1e-3, by1 = (double)(r + aa + 1 - a.h / 2) * a.resolution + 1e-3;
                double t0 = 0.0, t1 = len;
                bool hit = true;
                if (fabs(ux) > 1e-12) {
                    const double ta = (bx0 - l1.x) / ux, tb = (bx1 - l1.x) / ux;
                    t0 = fmax(t0, fmin(ta, tb)); t1 = fmin(t1, fmax(ta, tb));
                } else if (l1.x < bx0 || l1.x > bx1) hit = false;
                if (fabs(uy) > 1e-12) {
                    const double ta = (by0 - l1.y) / uy, tb = (by1 - l1.y) / uy;
                    t0 = fmax(t0, fmin(ta, tb)); t1 = fmin(t1, fmax(ta, tb));
                } else if (l1.y < by0 || l1.y > by1) hit = false;
                if (hit && t0 <= t1) {
                    const double lo = t0 - 0.06, hi = t1 + 0.06;
                    for (double tr = 0; tr <= len; tr += 0.05) {
                        if (tr < lo) continue;
                        if (tr > hi) break;
                        cover(match_cell(a, l1.x + ux * tr, l1.y + uy * tr));
                    }
                }
            }
            if (!mask) return;
            double v1x = l1.z - l1.x, v1y = l1.w - l1.y;
            { const double n = sqrt(v1x * v1x + v1y * v1y); if (n * n > 0.0) { v1x /= n; v1y /= n; } }
            const double angle = acos(fabs(v1x * v2.x + v1y * v2.y));
            const int cell = __ffsll((long long)mask) - 1;
            if (angle < best_angle || (angle == best_angle && (cell < best_cell || (cell == best_cell && j < best_j)))) {
                best_angle = angle; best_cell = cell; best_j = j;
            }
        };
        for (int jb = 0; jb < n1; jb += 32) {
            const int j = jb + lane;
            bool cand = false;
            if (j < n1) {
                const int4 b4 = *reinterpret_cast<const int4*>(bb + 4 * j);
                cand = !(b4.y < r - aa || b4.x > r + aa || b4.w < c - aa || b4.z > c + aa);
            }
            const unsigned m = __ballot_sync(0xffffffffu, cand);
            if (cand) queue[queued + __popc(m & ((1u << lane) - 1u))] = j;
            queued += __popc(m);
            __syncwarp();
            if (queued >= 32) {
                const int jq = queue[lane];
                const int carry = (lane + 32 < queued) ? queue[lane + 32] : 0;
                __syncwarp();
                evaluate(jq);
                queued -= 32;
                if (lane < queued) queue[lane] = carry;
                __syncwarp();
            }
        }
        if (lane < queued) evaluate(queue[lane]);
        __syncwarp();
        // lexicographic minimum of (angle, cell, j) over the warp; lanes without a candidate carry (1e300, 64, INT_MAX)
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const double oa = __shfl_xor_sync(0xffffffffu, best_angle, d);
            const int oc = __shfl_xor_sync(0xffffffffu, best_cell, d), oj = __shfl_xor_sync(0xffffffffu, best_j, d);
            if (oa < best_angle || (oa == best_angle && (oc < best_cell || (oc == best_cell && oj < best_j)))) { best_angle = oa; best_cell = oc; best_j = oj; }
        }
        const bool matched = best_j != 0x7fffffff /* a candidate, and not only NaN angles */ && !(best_angle / kPi * 180.0 > 10.0);
        if (WPP == 1) {
            if (!matched) continue;
            diss[count] = 0.5 * (match_dis_from_line(e1, L1[best_j]) + match_dis_from_line(e2, L1[best_j]));   // same value on every lane
            if (lane == 0) { prov[2 * count] = best_j; prov[2 * count + 1] = i; }
            ++count;
        } else if (lane == 0) {
            // slot i of the workspace; warp 0 compacts the slots in order below
            prov[2 * i] = matched ? best_j : -1;
            prov[2 * i + 1] = i;
            if (matched) diss[i] = 0.5 * (match_dis_from_line(e1, L1[best_j]) + match_dis_from_line(e2, L1[best_j]));
        }
    }
    if (WPP > 1) {
        __syncthreads();
        if (wid != 0) return;
        for (int base = 0; base < n2; base += 32) {       // (slot o <= slot k: a chunk is read completely before it is written)
            const int k = base + lane;
            const bool has = k < n2 && prov[2 * k] >= 0;
            const int bj = k < n2 ? prov[2 * k] : 0;
            const double dk = has ? diss[k] : 0.0;
            const unsigned m = __ballot_sync(0xffffffffu, has);
            __syncwarp();
            if (has) {
                const int o = count + __popc(m & ((1u << lane) - 1u));
                prov[2 * o] = bj; prov[2 * o + 1] = k; diss[o] = dk;
            }
            count += __popc(m);
            __syncwarp();
        }
    }
    __syncwarp();
    // mean end-point distance (summed in the reference's order), then the 1.2 x filter
    double aver = 0.0;
    for (int k = 0; k < count; ++k) aver += diss[k];
    aver /= (double)count;
    int kept = 0;
    for (int base = 0; base < count; base += 32) {
        const int k = base + lane;
        const bool keep = k < count && diss[k] < aver * 1.2;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int o = kept + __popc(m & ((1u << lane) - 1u));
            out[2 * o] = prov[2 * k]; out[2 * o + 1] = prov[2 * k + 1];
        }
        kept += __popc(m);
    }
    if (lane == 0) a.n_match[p] = kept;
}

}  // namespace lv
