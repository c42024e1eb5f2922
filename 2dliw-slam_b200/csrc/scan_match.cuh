// scan_match.cuh — the hot kernel: for every scan point of every frame of every window, point-to-line
// residual + 1x6 (or 1x12) Jacobian + accumulation of the per-frame J^T J / J^T r / r^2 blocks.
//
// Replaces, for all points at once, what Ceres does per residual block with Jets through
// laser_factor::operator() (reference src/factor/laser_factor.h:45-89) and
// e_laser::dis_from_line (src/utilies/common.h:86-95), plus the J^T J accumulation of the normal
// equations that ceres::Solve builds internally (src/factor/solver.cpp:802, :168).
//
// Algebra (verified in tests/test_device_math_host.py): with the frame table
// (M, t, B_k, b_k) of lv_math.cuh, the flattened world point is C = M c + t and dC/dtheta_k = B_k c + b_k.
// For a world line with unit normal n and offset c0 = n.A2 the signed distance is
//     d = n.C - c0 = f0 cx + f1 cy + f2,           f = (n^T M, n.t - c0)
//     dd/dtheta_j,k = e_k0 cx + e_k1 cy + e_k2,     e_k = (n^T B_k, n.b_k)
//     dd/dp_j = (nx, ny, 0)
// The residual is w |d| and its Jacobian w sign(d) (...): sign^2 = 1, so J^T J, J^T r and r^2 never
// need the sign.  Per line the 12 coefficients (f, e_0, e_1, e_2) + n are staged in shared memory once
// per (frame, tile); per point the kernel does 8 FMAs for (d, j_theta) and 21 for the accumulation.
//
// Work decomposition: one warp per (frame, tile), one warp per CTA (finest scheduling granularity, see LV_SCAN_WPC).
// Points of a frame are contiguous, so a warp reads 512 B (32 x double2) + 128 B (32 x int32) fully coalesced per
// step; loads bypass L1 (read once).  Prologue: one DRAM round trip for fixed-size scans, two for ragged batches.
// Reduction: per-lane register accumulators -> recursive-halving warp shuffle (23 exchanges for 21 values
// instead of 105) -> one 8-byte store per value.  No atomics; results are bit-reproducible.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "lv_math.cuh"

#ifndef LV_SCAN_U
#define LV_SCAN_U 4        // batches of 32 points per pipeline stage
#endif
#ifndef LV_SCAN_WPC
#define LV_SCAN_WPC 1      // warps (= frames in flight) per CTA.  One: the finest scheduling granularity — with 8 the
                           // warps of a CTA retire together and their successors sit in the prologue together
                           // (measured 0.747 / 0.783 / 0.794 / 0.804 of HBM peak for 8 / 4 / 2 / 1)
#endif
#ifndef LV_SCAN_WARPS_PER_SM
#define LV_SCAN_WARPS_PER_SM 16   // register budget: 16 -> 128 registers per thread
#endif

namespace lv {

constexpr int kAccTrack = 21;   // [0..2] aa | [3..8] a x bj | [9..14] bj bj^T (upper) | [15,16] d a | [17..19] d bj | [20] d^2
constexpr int kAccFree = 45;    // [0..14] as above | [15..20] a x bi | [21..26] bi bi^T (upper) | [27..35] bj x bi | [36,37] d a |
                                // [38..40] d bj | [41..43] d bi | [44] d^2
constexpr int kPadTrack = 24;
constexpr int kPadFree = 48;
constexpr int kRowTrack = 14;   // f(3) e(9) n(2)
constexpr int kRowTrackAssoc = 20;  // + h(3) (foot coordinate along the line, relative to A1) + length + (d, s1): the line in the map frame
constexpr int kRowFree = 24;    // + h(3) (relative to A2) alpha(3) beta(3) + length

struct ScanMatchArgs {
    const double2* points;          // [N] scan points in the frame's laser frame
    const int32_t* point_line;      // [N] line index inside the frame's local map, < 0: skip
    const double* point_weight;     // [N] or nullptr
    const int64_t* point_offset;    // [F+1]
    const int64_t* line_offset;     // [F+1]
    const double4* wlines;          // [L] (nx, ny, c0, u.A1) world lines, external constant reference pose
    const double* wlen;             // [L] world length of those lines
    const double4* lines;           // [L] raw (a1x a1y a2x a2y) in the reference laser frame
    const int32_t* ref_frame;       // [F] or nullptr
    const double* frame_tab;        // [F][24] at the evaluation point
    const uint8_t* frame_active;    // [F] 1 when the frame's laser residual block is part of the program
    const int32_t* win_status;      // [B] 0 = window still iterating
    double* partial;                // [F][tiles][pad]
    int32_t n_frames;               // frames per window
    int32_t tiles;
    int32_t n_items;                // F * tiles
    int32_t line_cap;               // shared-memory rows per warp
    int32_t shard_rank, shard_world;
    int32_t uniform_pts, uniform_lines;   // > 0: every frame has exactly this many points / lines (offsets are f * count)
    double huber_delta, laser_sqrt_info, assoc_gate, assoc_max_dist;
    // in-kernel re-association (BASELINE config 3): per frame a kAssocG x kAssocG grid over the world bounding box of the
    // frame's local map, every cell holding the bit mask of the lines a point inside it can possibly match (built once
    // per upload by assoc_grid_kernel); null: every point tests every line
    const double* assoc_grid_par;        // [F][4]  x0, y0, 1 / cell_x, 1 / cell_y   (1 / cell_x == 0: no grid for this frame)
    const unsigned long long* assoc_grid_mask;   // [F][kAssocG * kAssocG][2]
};
#ifndef LV_ASSOC_G
#define LV_ASSOC_G 16
#endif
constexpr int kAssocG = LV_ASSOC_G;
constexpr int kAssocGridDoubles = kAssocG * kAssocG * 2;   // shared-memory copy per warp (as 64-bit words)

// streamed (read-once) loads: no L1 allocation.  (An L2 evict_first cache-hint policy on these loads and
// prefetch.global.L2 ahead of the pipeline were both measured and made no difference / were slower; see DESIGN.md.)
__device__ __forceinline__ double2 ld_stream_f64x2(const double2* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ int ld_stream_s32(const int32_t* p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_stream_f64(const double* p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// Recursive-halving reduce-scatter over the warp: N values per lane in, ceil(N/32) per lane out.
// After the call lane holds v[0..len) = full sums of logical indices [base, base+len).
template <int N, int DIST>
struct ReduceScatter {
    static __device__ __forceinline__ void run(double* v, int lane, int& base, int& len) {
        constexpr int H = (N + 1) / 2;
        const bool upper = (lane & DIST) != 0;
#pragma unroll
        for (int i = 0; i < H; ++i) {
            const double hi = (i + H < N) ? v[i + H] : 0.0;
            const double send = upper ? v[i] : hi;
            const double keep = upper ? hi : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, DIST);
        }
        if (upper) { base += H; len = (len > H) ? len - H : 0; } else { len = (len < H) ? len : H; }
        ReduceScatter<H, DIST / 2>::run(v, lane, base, len);
    }
};
template <int N>
struct ReduceScatter<N, 0> {
    static __device__ __forceinline__ void run(double*, int, int&, int&) {}
};
// number of values a lane holds after 5 halvings of N
__host__ __device__ constexpr int halved5(int n) { for (int i = 0; i < 5; ++i) n = (n + 1) / 2; return n; }

// one (frame, tile) item on one warp; `tab` = this warp's line table in shared memory (line_cap rows)
template <bool REF_FREE, bool HAS_WEIGHT, bool ASSOC, bool HUBER>
__device__ __forceinline__ void scan_match_item(const ScanMatchArgs& a, const int item, const int lane, double* tab) {
    constexpr int NACC = REF_FREE ? kAccFree : kAccTrack;
    constexpr int NPAD = REF_FREE ? kPadFree : kPadTrack;
    constexpr int ROW = REF_FREE ? kRowFree : (ASSOC ? kRowTrackAssoc : kRowTrack);
    const int f = item / a.tiles;
    const int tile = item - f * a.tiles;
    // Prologue.  Ragged batches: every scalar this warp depends on is requested before the first one is tested, so the
    // prologue costs two dependent DRAM round trips (these scalars -> points / lines) instead of four.  Uniform batches
    // (every frame the same number of points and lines — fixed-size scans; the host checks the offsets at upload): the
    // offsets are arithmetic, so status / active flag, frame table, the first 128 lines and the first point batches are
    // all requested in ONE round trip and the early-exit test comes after they are in flight.
    // (volatile asm keeps ptxas from sinking loads below the exit tests)
    const bool uni = !REF_FREE && a.uniform_pts > 0 && a.uniform_lines > 0;
    int wstat = 0, fact;
    int64_t p0, p1, l0, l1;
    asm volatile("ld.global.s32 %0, [%1];" : "=r"(wstat) : "l"(a.win_status + f / a.n_frames));
    asm volatile("ld.global.u8 %0, [%1];" : "=r"(fact) : "l"(a.frame_active + f));
    if (uni) {
        p0 = (int64_t)f * a.uniform_pts; p1 = p0 + a.uniform_pts;
        l0 = (int64_t)f * a.uniform_lines; l1 = l0 + a.uniform_lines;
    } else {
        asm volatile("ld.global.s64 %0, [%1];" : "=l"(p0) : "l"(a.point_offset + f));
        asm volatile("ld.global.s64 %0, [%1];" : "=l"(p1) : "l"(a.point_offset + f + 1));
        asm volatile("ld.global.s64 %0, [%1];" : "=l"(l0) : "l"(a.line_offset + f));
        asm volatile("ld.global.s64 %0, [%1];" : "=l"(l1) : "l"(a.line_offset + f + 1));
    }
    const double* ft = a.frame_tab + (size_t)f * kFrameTab;
    double T[kFrameTab];
#pragma unroll
    for (int i = 0; i < kFrameTab; i += 2)
        asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(T[i]), "=d"(T[i + 1]) : "l"(ft + i));   // coherent: solve_small_kernel rewrites frame_tab between trips
    if (!uni) {
        if ((wstat != 0) | (fact == 0) | (p1 <= p0) | (l1 < l0)) return;   // (one test on all six: ptxas keeps the loads together)
    }

    // ---- this warp's slice of the frame: rank shard, then tile
    const int64_t cnt = p1 - p0;
    const int64_t s0 = p0 + (cnt * a.shard_rank) / a.shard_world;
    const int64_t s1 = p0 + (cnt * (a.shard_rank + 1)) / a.shard_world;
    const int64_t per = (s1 - s0 + a.tiles - 1) / a.tiles;
    const int64_t pb = s0 + per * tile;
    const int64_t pe = (pb + per < s1) ? pb + per : s1;

    // software pipeline: the loads of batch k+1 are in flight while batch k is accumulated; batch 0 is issued
    // before the line table is built
    constexpr int U = LV_SCAN_U;
    double2 nc[U];
    int nli[U];
    double nw[U];
    auto issue = [&](int64_t base) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t p = base + 32 * u;
            const bool ok = p < pe;
            nli[u] = ok ? ld_stream_s32(a.point_line + p) : -1;
            nc[u] = ok ? ld_stream_f64x2(a.points + p) : make_double2(0.0, 0.0);
            if constexpr (HAS_WEIGHT) nw[u] = ok ? ld_stream_f64(a.point_weight + p) : 0.0;
        }
    };
    issue(pb + lane);

    // ---- this frame's table (24 doubles, broadcast loads) and the shared-memory line table
    const int nl = (int)(l1 - l0);
    // a lane's lines are requested four at a time before any is consumed: the table build costs one DRAM round trip
    // per 128 lines instead of one per 32; the first group is requested here, ahead of the uniform path's exit test
    double4 wl4[4];
    if constexpr (!REF_FREE) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            wl4[q] = make_double4(0.0, 0.0, 0.0, 0.0);
            if (lane + 32 * q < nl) {
                const double4* src = a.wlines + l0 + lane + 32 * q;
                asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(wl4[q].x), "=d"(wl4[q].y) : "l"(src));
                asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(wl4[q].z), "=d"(wl4[q].w) : "l"(reinterpret_cast<const double2*>(src) + 1));
            }
        }
    }
    if (uni) {
        if ((wstat != 0) | (fact == 0)) return;
    }
    {
        if constexpr (!REF_FREE) {
            for (int lb = lane; lb < nl; lb += 128) {
              if (lb != lane) {
#pragma unroll
                for (int q = 0; q < 4; ++q) wl4[q] = (lb + 32 * q < nl) ? a.wlines[l0 + lb + 32 * q] : make_double4(0.0, 0.0, 0.0, 0.0);
              }
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int l = lb + 32 * q;
                if (l >= nl) break;
                const double4 wl = wl4[q];
                double* r = tab + l * ROW;
                r[0] = wl.x * T[0] + wl.y * T[2];
                r[1] = wl.x * T[1] + wl.y * T[3];
                r[2] = wl.x * T[4] + wl.y * T[5] - wl.z;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double* Bk = T + 6 + 6 * k;
                    r[3 + 3 * k] = wl.x * Bk[0] + wl.y * Bk[2];
                    r[4 + 3 * k] = wl.x * Bk[1] + wl.y * Bk[3];
                    r[5 + 3 * k] = wl.x * Bk[4] + wl.y * Bk[5];
                }
                r[12] = wl.x;
                r[13] = wl.y;
                if constexpr (ASSOC) {
                    const double ux = wl.y, uy = -wl.x;   // n = (-uy, ux)
                    r[14] = ux * T[0] + uy * T[2];
                    r[15] = ux * T[1] + uy * T[3];
                    r[16] = ux * T[4] + uy * T[5] - wl.w;
                    r[17] = a.wlen[l0 + l];
                    r[18] = wl.z;      // the same line in the frame the points are mapped to by (T[0..5]): n.P = d, u.P - s1 in [0, len]
                    r[19] = wl.w;
                }
              }
            }
        } else {
            const int rf = a.ref_frame[f];
            const double* fi = a.frame_tab + ((size_t)(f - f % a.n_frames) + rf) * kFrameTab;
            double Ti[kFrameTab];
#pragma unroll
            for (int i = 0; i < kFrameTab; ++i) Ti[i] = fi[i];   // (plain load: frame_tab is rewritten between the trips of solve_small_kernel)
            for (int lb = lane; lb < nl; lb += 128) {
              double4 ln4[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) ln4[q] = (lb + 32 * q < nl) ? a.lines[l0 + lb + 32 * q] : make_double4(0.0, 0.0, 1.0, 0.0);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int l = lb + 32 * q;
                if (l >= nl) break;
                const double4 ln = ln4[q];
                const double A1x = Ti[0] * ln.x + Ti[1] * ln.y + Ti[4], A1y = Ti[2] * ln.x + Ti[3] * ln.y + Ti[5];
                const double A2x = Ti[0] * ln.z + Ti[1] * ln.w + Ti[4], A2y = Ti[2] * ln.z + Ti[3] * ln.w + Ti[5];
                const double dx = A2x - A1x, dy = A2y - A1y;
                const double len = sqrt(dx * dx + dy * dy), inv = 1.0 / len;
                const double ux = dx * inv, uy = dy * inv, nx = -uy, ny = ux;
                double* r = tab + l * ROW;
                r[0] = nx * T[0] + ny * T[2];
                r[1] = nx * T[1] + ny * T[3];
                r[2] = nx * T[4] + ny * T[5] - (nx * A2x + ny * A2y);
                r[14] = ux * T[0] + uy * T[2];
                r[15] = ux * T[1] + uy * T[3];
                r[16] = ux * T[4] + uy * T[5] - (ux * A2x + uy * A2y);
                const double ex = ln.z - ln.x, ey = ln.w - ln.y;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double* Bk = T + 6 + 6 * k;
                    r[3 + 3 * k] = nx * Bk[0] + ny * Bk[2];
                    r[4 + 3 * k] = nx * Bk[1] + ny * Bk[3];
                    r[5 + 3 * k] = nx * Bk[4] + ny * Bk[5];
                    const double* Ck = Ti + 6 + 6 * k;
                    // alpha_k = n.(B_ik (a2 - a1)) / L,  beta_k = n.(B_ik a2 + b_ik)
                    r[17 + k] = (nx * (Ck[0] * ex + Ck[1] * ey) + ny * (Ck[2] * ex + Ck[3] * ey)) * inv;
                    r[20 + k] = nx * (Ck[0] * ln.z + Ck[1] * ln.w + Ck[4]) + ny * (Ck[2] * ln.z + Ck[3] * ln.w + Ck[5]);
                }
                r[12] = nx;
                r[13] = ny;
                r[23] = len;
              }
            }
        }
    }
    // in-kernel association: this frame's candidate grid next to the line table
    bool use_grid = false;
    double gx0 = 0.0, gy0 = 0.0, gix = 0.0, giy = 0.0;
    unsigned long long* gmask = reinterpret_cast<unsigned long long*>(tab + (size_t)a.line_cap * ROW);
    if constexpr (ASSOC && !REF_FREE) {
        if (a.assoc_grid_par) {
            const double* gp = a.assoc_grid_par + (size_t)f * 4;
            gx0 = gp[0]; gy0 = gp[1]; gix = gp[2]; giy = gp[3];
            use_grid = gix > 0.0;
            if (use_grid) {
                const unsigned long long* src = a.assoc_grid_mask + (size_t)f * kAssocGridDoubles;
#pragma unroll
                for (int q = 0; q < kAssocGridDoubles / 32; ++q) gmask[lane + 32 * q] = __ldg(src + lane + 32 * q);
            }
        }
    }
    __syncwarp();

    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.0;

    // (ASSOC: the trip count is the same on every lane — the candidate union below is a warp collective; lanes past the end
    // carry li = -1)
    for (int64_t base = pb + lane; ASSOC ? (base - lane < pe) : (base < pe); base += 32 * U) {
        double2 c[U];
        int li[U];
        double w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { c[u] = nc[u]; li[u] = nli[u]; if constexpr (HAS_WEIGHT) w[u] = nw[u]; }
        issue(base + 32 * U);
        bool assoc_done = false;
        if constexpr (ASSOC && !REF_FREE) {
            if (use_grid) {
                // Nearest line by perpendicular distance among the lines whose extent (+gate) contains the foot, for the
                // warp's U x 32 points at once.  Only lines registered in a point's grid cell can pass the test for it, so
                // testing MORE lines than a point's own cell lists changes nothing: the warp walks the UNION of the cell
                // masks of all its points (128 consecutive beams: a few cells) in one uniform loop — no divergence, one
                // broadcast load of a candidate for U points — in increasing line order, so that ties resolve exactly like
                // the all-lines loop.
                double X[U], Y[U], bd[U];
                int bl[U];
                unsigned long long m0 = 0ull, m1 = 0ull;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    X[u] = fma(T[0], c[u].x, fma(T[1], c[u].y, T[4]));
                    Y[u] = fma(T[2], c[u].x, fma(T[3], c[u].y, T[5]));
                    const double gx = (X[u] - gx0) * gix, gy = (Y[u] - gy0) * giy;
                    const bool inside = li[u] >= 0 && gx >= 0.0 && gx < (double)kAssocG && gy >= 0.0 && gy < (double)kAssocG;
                    const int cell = inside ? (int)gy * kAssocG + (int)gx : 0;
                    const ulonglong2 mm = *reinterpret_cast<const ulonglong2*>(gmask + 2 * cell);
                    if (inside) { m0 |= mm.x; m1 |= mm.y; }
                    bd[u] = a.assoc_max_dist;
                    bl[u] = inside ? -1 : -2;      // -2: no correspondence, or outside every line's reach
                }
                unsigned un[4];
                un[0] = __reduce_or_sync(0xffffffffu, (unsigned)m0);
                un[1] = __reduce_or_sync(0xffffffffu, (unsigned)(m0 >> 32));
                un[2] = __reduce_or_sync(0xffffffffu, (unsigned)m1);
                un[3] = __reduce_or_sync(0xffffffffu, (unsigned)(m1 >> 32));
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    unsigned m = un[q];
                    while (m) {
                        const int l = 32 * q + __ffs((int)m) - 1;
                        m &= m - 1;
                        const double* qr = tab + l * ROW;
                        const double2 nn = *reinterpret_cast<const double2*>(qr + 12);
                        const double2 ds = *reinterpret_cast<const double2*>(qr + 18);
                        const double hi = qr[17] + a.assoc_gate;
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const double dd = fma(nn.x, X[u], fma(nn.y, Y[u], -ds.x));
                            const double tt = fma(nn.y, X[u], fma(-nn.x, Y[u], -ds.y));
                            const double ad = fabs(dd);
                            if (tt >= -a.assoc_gate && tt <= hi && ad < bd[u]) { bd[u] = ad; bl[u] = (bl[u] == -2) ? -2 : l; }
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) li[u] = bl[u];
                assoc_done = true;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if constexpr (!(ASSOC && !REF_FREE)) { if (li[u] < 0) continue; }
            if constexpr (!ASSOC) { if (li[u] >= nl) continue; }   // an index beyond the frame's line list takes no part (never an out-of-table read)
            if constexpr (ASSOC && !REF_FREE) { if (assoc_done && li[u] < 0) continue; }
            if constexpr (ASSOC) { if (!assoc_done) {
                // nearest line by perpendicular distance among the lines whose extent (+gate) contains the foot (the
                // all-lines loop: initialisation topology, or no candidate grid)
                int best = -1;
                double best_d = a.assoc_max_dist;
                auto test_line = [&](const int l) {
                    const double* q = tab + l * ROW;
                    const double dd = fma(q[0], c[u].x, fma(q[1], c[u].y, q[2]));
                    double tt = fma(q[14], c[u].x, fma(q[15], c[u].y, q[16]));
                    const double len = REF_FREE ? q[23] : q[17];
                    if constexpr (REF_FREE) tt += len;   // h is relative to A2 there
                    const double ad = fabs(dd);
                    if (tt >= -a.assoc_gate && tt <= len + a.assoc_gate && ad < best_d) { best_d = ad; best = l; }
                };
                if constexpr (!REF_FREE) {
                    // the point in the frame the lines live in; a candidate costs 3 wide loads and 4 FMAs there
                    const double X = fma(T[0], c[u].x, fma(T[1], c[u].y, T[4])), Y = fma(T[2], c[u].x, fma(T[3], c[u].y, T[5]));
                    auto test_world = [&](const int l) {
                        const double* q = tab + l * ROW;
                        const double2 nn = *reinterpret_cast<const double2*>(q + 12);
                        const double2 ds = *reinterpret_cast<const double2*>(q + 18);
                        const double len = q[17];
                        const double dd = fma(nn.x, X, fma(nn.y, Y, -ds.x));
                        const double tt = fma(nn.y, X, fma(-nn.x, Y, -ds.y));
                        const double ad = fabs(dd);
                        if (tt >= -a.assoc_gate && tt <= len + a.assoc_gate && ad < best_d) { best_d = ad; best = l; }
                    };
                    if (li[u] < 0) continue;
                    for (int l = 0; l < nl; ++l) test_world(l);
                } else {
                    for (int l = 0; l < nl; ++l) test_line(l);
                }
                if (best < 0) continue;
                li[u] = best;
            } }
            const double* r = tab + li[u] * ROW;
            const double2 r01 = *reinterpret_cast<const double2*>(r);
            const double2 r23 = *reinterpret_cast<const double2*>(r + 2);
            const double2 r45 = *reinterpret_cast<const double2*>(r + 4);
            const double2 r67 = *reinterpret_cast<const double2*>(r + 6);
            const double2 r89 = *reinterpret_cast<const double2*>(r + 8);
            const double2 rab = *reinterpret_cast<const double2*>(r + 10);
            const double2 rn = *reinterpret_cast<const double2*>(r + 12);
            const double cx = c[u].x, cy = c[u].y;
            double d = fma(r01.x, cx, fma(r01.y, cy, r23.x));
            double j2 = fma(r23.y, cx, fma(r45.x, cy, r45.y));
            double j3 = fma(r67.x, cx, fma(r67.y, cy, r89.x));
            double j4 = fma(r89.y, cx, fma(rab.x, cy, rab.y));
            double j0 = rn.x, j1 = rn.y;
            double i2 = 0.0, i3 = 0.0, i4 = 0.0;
            if constexpr (REF_FREE) {
                const double2 h01 = *reinterpret_cast<const double2*>(r + 14);
                const double2 h2a = *reinterpret_cast<const double2*>(r + 16);
                const double2 a12 = *reinterpret_cast<const double2*>(r + 18);
                const double2 b01 = *reinterpret_cast<const double2*>(r + 20);
                const double b2 = r[22];
                const double tt = fma(h01.x, cx, fma(h01.y, cy, h2a.x));
                i2 = -fma(tt, h2a.y, b01.x);
                i3 = -fma(tt, a12.x, b01.y);
                i4 = -fma(tt, a12.y, b2);
            }
            if constexpr (HAS_WEIGHT) {
                const double ww = w[u];
                d *= ww; j0 *= ww; j1 *= ww; j2 *= ww; j3 *= ww; j4 *= ww;
                if constexpr (REF_FREE) { i2 *= ww; i3 *= ww; i4 *= ww; }
            }
            double cterm;
            if constexpr (HUBER) {
                // ceres::HuberLoss on the whitened residual k*|d|: rho' scales J^T J and J^T r, rho replaces r^2
                const double k = a.laser_sqrt_info;
                const double q2 = (k * d) * (k * d);
                if (q2 > a.huber_delta * a.huber_delta) {
                    const double sq = sqrt(q2);
                    const double sc = sqrt(a.huber_delta / sq);
                    cterm = (2.0 * a.huber_delta * sq - a.huber_delta * a.huber_delta) / (k * k);
                    d *= sc; j0 *= sc; j1 *= sc; j2 *= sc; j3 *= sc; j4 *= sc;
                    if constexpr (REF_FREE) { i2 *= sc; i3 *= sc; i4 *= sc; }
                } else {
                    cterm = d * d;
                }
            } else {
                cterm = d * d;
            }
            acc[0] = fma(j0, j0, acc[0]); acc[1] = fma(j0, j1, acc[1]); acc[2] = fma(j1, j1, acc[2]);
            acc[3] = fma(j0, j2, acc[3]); acc[4] = fma(j0, j3, acc[4]); acc[5] = fma(j0, j4, acc[5]);
            acc[6] = fma(j1, j2, acc[6]); acc[7] = fma(j1, j3, acc[7]); acc[8] = fma(j1, j4, acc[8]);
            acc[9] = fma(j2, j2, acc[9]); acc[10] = fma(j2, j3, acc[10]); acc[11] = fma(j2, j4, acc[11]);
            acc[12] = fma(j3, j3, acc[12]); acc[13] = fma(j3, j4, acc[13]); acc[14] = fma(j4, j4, acc[14]);
            if constexpr (!REF_FREE) {
                acc[15] = fma(d, j0, acc[15]); acc[16] = fma(d, j1, acc[16]);
                acc[17] = fma(d, j2, acc[17]); acc[18] = fma(d, j3, acc[18]); acc[19] = fma(d, j4, acc[19]);
                acc[20] += cterm;
            } else {
                acc[15] = fma(j0, i2, acc[15]); acc[16] = fma(j0, i3, acc[16]); acc[17] = fma(j0, i4, acc[17]);
                acc[18] = fma(j1, i2, acc[18]); acc[19] = fma(j1, i3, acc[19]); acc[20] = fma(j1, i4, acc[20]);
                acc[21] = fma(i2, i2, acc[21]); acc[22] = fma(i2, i3, acc[22]); acc[23] = fma(i2, i4, acc[23]);
                acc[24] = fma(i3, i3, acc[24]); acc[25] = fma(i3, i4, acc[25]); acc[26] = fma(i4, i4, acc[26]);
                acc[27] = fma(j2, i2, acc[27]); acc[28] = fma(j2, i3, acc[28]); acc[29] = fma(j2, i4, acc[29]);
                acc[30] = fma(j3, i2, acc[30]); acc[31] = fma(j3, i3, acc[31]); acc[32] = fma(j3, i4, acc[32]);
                acc[33] = fma(j4, i2, acc[33]); acc[34] = fma(j4, i3, acc[34]); acc[35] = fma(j4, i4, acc[35]);
                acc[36] = fma(d, j0, acc[36]); acc[37] = fma(d, j1, acc[37]);
                acc[38] = fma(d, j2, acc[38]); acc[39] = fma(d, j3, acc[39]); acc[40] = fma(d, j4, acc[40]);
                acc[41] = fma(d, i2, acc[41]); acc[42] = fma(d, i3, acc[42]); acc[43] = fma(d, i4, acc[43]);
                acc[44] += cterm;
            }
        }
    }

    int base = 0, len = NACC;
    ReduceScatter<NACC, 16>::run(acc, lane, base, len);
    double* out = a.partial + (size_t)item * NPAD;
    constexpr int OUTN = halved5(NACC);
#pragma unroll
    for (int i = 0; i < OUTN; ++i)
        if (i < len) out[base + i] = acc[i];
}

template <bool REF_FREE, bool HAS_WEIGHT, bool ASSOC, bool HUBER>
__global__ void __launch_bounds__(32 * LV_SCAN_WPC, (REF_FREE ? 8 : LV_SCAN_WARPS_PER_SM) / LV_SCAN_WPC) scan_match_kernel(ScanMatchArgs a) {
    constexpr int ROW = REF_FREE ? kRowFree : (ASSOC ? kRowTrackAssoc : kRowTrack);
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int item = blockIdx.x * (blockDim.x >> 5) + warp;
    if (item >= a.n_items) return;
    scan_match_item<REF_FREE, HAS_WEIGHT, ASSOC, HUBER>(a, item, lane, smem + (size_t)warp * (a.line_cap * ROW + ((ASSOC && !REF_FREE) ? kAssocGridDoubles : 0)));
}

// world lines for frames whose local map hangs under an external constant reference pose: computed once per
// set_windows (the reference pose never changes during a solve, solver.cpp:690-691)
__global__ void world_lines_kernel(const double4* lines, const int64_t* line_offset, const int32_t* ref_frame,
                                   const double* ref_tab /*[F][24]*/, double4* wlines, double* wlen, int n_frames_total) {
    const int f = blockIdx.x;
    if (f >= n_frames_total) return;
    if (ref_frame && ref_frame[f] >= 0) return;
    const double* T = ref_tab + (size_t)f * kFrameTab;
    const int64_t l0 = line_offset[f], l1 = line_offset[f + 1];
    for (int64_t l = l0 + threadIdx.x; l < l1; l += blockDim.x) {
        const double4 ln = lines[l];
        const double A1x = T[0] * ln.x + T[1] * ln.y + T[4], A1y = T[2] * ln.x + T[3] * ln.y + T[5];
        const double A2x = T[0] * ln.z + T[1] * ln.w + T[4], A2y = T[2] * ln.z + T[3] * ln.w + T[5];
        const double dx = A2x - A1x, dy = A2y - A1y;
        const double len = sqrt(dx * dx + dy * dy), inv = 1.0 / len;
        const double ux = dx * inv, uy = dy * inv, nx = -uy, ny = ux;
        wlines[l] = make_double4(nx, ny, nx * A2x + ny * A2y, ux * A1x + uy * A1y);
        wlen[l] = len;
    }
}

// Candidate grid of the in-kernel association, once per upload: one CTA per frame.  A point P can match line l only when
// |n.P - d| < max_dist and -gate <= u.P - s1 <= len + gate — a rectangle around the segment.  Every cell that rectangle
// can touch (separating-axis test of the cell against the two line axes, plus a 1e-6 m safety margin for the rounding of
// the cell index) gets bit l.  Frames with more than 128 lines keep the all-lines loop (par[2] = 0).
__global__ void assoc_grid_kernel(const double4* wlines, const double* wlen, const int64_t* line_offset, const int32_t* ref_frame,
                                  double gate, double max_dist, double* par, unsigned long long* mask, int n_frames_total) {
    const int f = blockIdx.x;
    if (f >= n_frames_total) return;
    __shared__ unsigned long long sm[kAssocGridDoubles];
    __shared__ double red[4][32];
    double* gp = par + (size_t)f * 4;
    const int64_t l0 = line_offset[f];
    const int nl = (int)(line_offset[f + 1] - l0);
    const int tid = threadIdx.x;
    for (int k = tid; k < kAssocGridDoubles; k += blockDim.x) sm[k] = 0ull;
    if (nl <= 0 || nl > 128 || (ref_frame && ref_frame[f] >= 0)) {
        if (tid < 4) gp[tid] = 0.0;
        for (int k = tid; k < kAssocGridDoubles; k += blockDim.x) mask[(size_t)f * kAssocGridDoubles + k] = 0ull;
        return;
    }
    const double eps = 1e-6, md = max_dist + eps, gt = gate + eps;
    // bounding box of every line's rectangle
    double x0 = 1e300, y0 = 1e300, x1 = -1e300, y1 = -1e300;
    double4 wl = make_double4(0, 0, 0, 0);
    double len = 0.0;
    if (tid < nl) {
        wl = wlines[l0 + tid];
        len = wlen[l0 + tid];
        const double nx = wl.x, ny = wl.y, ux = wl.y, uy = -wl.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double s = (k & 1) ? wl.w + len + gt : wl.w - gt, d = (k & 2) ? wl.z + md : wl.z - md;
            const double px = s * ux + d * nx, py = s * uy + d * ny;
            x0 = fmin(x0, px); x1 = fmax(x1, px); y0 = fmin(y0, py); y1 = fmax(y1, py);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        x0 = fmin(x0, __shfl_xor_sync(0xffffffffu, x0, o)); y0 = fmin(y0, __shfl_xor_sync(0xffffffffu, y0, o));
        x1 = fmax(x1, __shfl_xor_sync(0xffffffffu, x1, o)); y1 = fmax(y1, __shfl_xor_sync(0xffffffffu, y1, o));
    }
    if ((tid & 31) == 0) { red[0][tid >> 5] = x0; red[1][tid >> 5] = y0; red[2][tid >> 5] = x1; red[3][tid >> 5] = y1; }
    __syncthreads();
    const int nw = blockDim.x >> 5;
    x0 = red[0][0]; y0 = red[1][0]; x1 = red[2][0]; y1 = red[3][0];
    for (int k = 1; k < nw; ++k) { x0 = fmin(x0, red[0][k]); y0 = fmin(y0, red[1][k]); x1 = fmax(x1, red[2][k]); y1 = fmax(y1, red[3][k]); }
    x0 -= eps; y0 -= eps; x1 += eps; y1 += eps;
    const double cx = (x1 - x0) / kAssocG, cy = (y1 - y0) / kAssocG;
    if (tid < nl) {
        const double nx = wl.x, ny = wl.y, ux = wl.y, uy = -wl.x;
        const double rn = 0.5 * (fabs(nx) * cx + fabs(ny) * cy) + eps, ru = 0.5 * (fabs(ux) * cx + fabs(uy) * cy) + eps;   // half extents of a cell on the two axes
        for (int gy = 0; gy < kAssocG; ++gy)
            for (int gx = 0; gx < kAssocG; ++gx) {
                const double px = x0 + (gx + 0.5) * cx, py = y0 + (gy + 0.5) * cy;
                const double dn = fabs(nx * px + ny * py - wl.z), su = ux * px + uy * py - wl.w;
                if (dn <= md + rn && su >= -gt - ru && su <= len + gt + ru)
                    atomicOr(&sm[2 * (gy * kAssocG + gx) + (tid >> 6)], 1ull << (tid & 63));
            }
    }
    __syncthreads();
    for (int k = tid; k < kAssocGridDoubles; k += blockDim.x) mask[(size_t)f * kAssocGridDoubles + k] = sm[k];
    if (tid == 0) { gp[0] = x0; gp[1] = y0; gp[2] = 1.0 / cx; gp[3] = 1.0 / cy; }
}

}  // namespace lv
