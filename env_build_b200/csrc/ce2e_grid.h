// ce2e_grid.h -- host-side construction of the nearest-waypoint candidate grid.
//
// ReferencePath.find_closest_point (DM:702-715) is a brute-force first-argmin of the fp32
// squared distance over every 10th waypoint (366 / 390 / 312 candidates).  The kernels must
// return EXACTLY that index, but they do not have to look at every candidate: for each cell of
// a uniform grid over the map this builder stores an index range [lo, hi] that provably
// contains the brute-force answer of every query point inside the cell, so the device scans
// only that range (same arithmetic, same ascending order, same strict `<`, hence the same
// first minimum).  Points outside the grid (or NaN) fall back to the full range.
//
// Proof obligation (see DESIGN.md "candidate grid"): waypoint k may be dropped for a cell R only if
// some waypoint j satisfies  d_k(p) - d_j(p) > slack  for ALL p in R (R inflated by `eps` to
// cover the fp32 rounding of the cell index computation).  d_k - d_j is affine in p, so its
// minimum over the rectangle is attained at a corner; slack = 1e-5 * max_R d_k is ~40x the worst
// fp32 evaluation error 4u*(d_k + d_j), u = 2^-24.  Everything here runs in double on the
// exact fp32 waypoint values.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace ce2e {

struct GridSpec {
    float x0, y0;     // origin (integer valued, exact in fp32)
    float inv_h;      // cells per metre (power of two)
    int nx, ny;
};

constexpr double GRID_H = 0.5;          // cell size [m]
constexpr double GRID_MARGIN = 8.0;     // the grid covers the path's bounding box plus this margin
constexpr int GRID_COARSE = 8;          // coarse pre-filter cells are GRID_COARSE x GRID_COARSE fine cells
constexpr double GRID_SLACK = 1e-5;
constexpr int GRID_REFINE_MAX = 96;     // skip the pairwise refinement above this many candidates

namespace grid_detail {

struct Rect {
    double xlo, xhi, ylo, yhi;
};

inline double dmin2(const Rect &r, double wx, double wy) {
    double dx = std::max(std::max(r.xlo - wx, wx - r.xhi), 0.0);
    double dy = std::max(std::max(r.ylo - wy, wy - r.yhi), 0.0);
    return dx * dx + dy * dy;
}
inline double dmax2(const Rect &r, double wx, double wy) {
    double dx = std::max(std::fabs(wx - r.xlo), std::fabs(wx - r.xhi));
    double dy = std::max(std::fabs(wy - r.ylo), std::fabs(wy - r.yhi));
    return dx * dx + dy * dy;
}

// candidates of `r` among `pool` by the min-max rule: k stays iff dmin_k <= (1+slack) * min_j dmax_j
inline void minmax_filter(const Rect &r, const float *wx, const float *wy, const std::vector<int> &pool,
                          std::vector<int> &out) {
    double U = INFINITY;
    for (int k : pool) U = std::min(U, dmax2(r, wx[k], wy[k]));
    const double lim = U * (1.0 + GRID_SLACK) + 1e-20;
    out.clear();
    for (int k : pool)
        if (dmin2(r, wx[k], wy[k]) <= lim) out.push_back(k);
}

// pairwise (bisector) refinement: drop k if some j beats it by `slack` at all four corners
inline void refine(const Rect &r, const float *wx, const float *wy, std::vector<int> &cand) {
    const double cx[4] = {r.xlo, r.xhi, r.xlo, r.xhi}, cy[4] = {r.ylo, r.ylo, r.yhi, r.yhi};
    const size_t n = cand.size();
    std::vector<double> d(n * 4);
    for (size_t a = 0; a < n; ++a)
        for (int c = 0; c < 4; ++c) {
            double dx = cx[c] - wx[cand[a]], dy = cy[c] - wy[cand[a]];
            d[a * 4 + c] = dx * dx + dy * dy;
        }
    std::vector<int> keep;
    for (size_t a = 0; a < n; ++a) {
        const double slack = GRID_SLACK * std::max(std::max(d[a * 4], d[a * 4 + 1]), std::max(d[a * 4 + 2], d[a * 4 + 3])) + 1e-20;
        bool dominated = false;
        for (size_t b = 0; b < n && !dominated; ++b) {
            if (b == a) continue;
            bool all = true;
            for (int c = 0; c < 4; ++c) all = all && (d[a * 4 + c] - d[b * 4 + c] > slack);
            dominated = all;
        }
        if (!dominated) keep.push_back(cand[a]);
    }
    cand.swap(keep);
}

}  // namespace grid_detail

// wx, wy: the N decimated waypoints (exact fp32 values).  cells[iy*nx + ix] = lo | hi << 16.
inline bool build_candidate_grid(const float *wx, const float *wy, int N, GridSpec &g,
                                 std::vector<uint32_t> &cells) {
    using namespace grid_detail;
    if (N < 1 || N > 65535) return false;
    double xmin = wx[0], xmax = wx[0], ymin = wy[0], ymax = wy[0];
    for (int k = 1; k < N; ++k) {
        xmin = std::min(xmin, (double)wx[k]); xmax = std::max(xmax, (double)wx[k]);
        ymin = std::min(ymin, (double)wy[k]); ymax = std::max(ymax, (double)wy[k]);
    }
    if (!std::isfinite(xmin + xmax + ymin + ymax) || std::max(std::fabs(xmin), std::fabs(xmax)) > 1e4 ||
        std::max(std::fabs(ymin), std::fabs(ymax)) > 1e4)
        return false;
    const double h = GRID_H;
    const double x0 = std::floor(xmin - GRID_MARGIN), y0 = std::floor(ymin - GRID_MARGIN);
    const int nx = (int)std::ceil((xmax + GRID_MARGIN - x0) / h), ny = (int)std::ceil((ymax + GRID_MARGIN - y0) / h);
    if ((int64_t)nx * ny > (1 << 22)) return false;
    g.x0 = (float)x0; g.y0 = (float)y0; g.inv_h = (float)(1.0 / h); g.nx = nx; g.ny = ny;
    cells.assign((size_t)nx * ny, 0u);
    const double eps = 0.01 * h;       // >> fp32 rounding of (x - x0) * inv_h (about 1e-5 cells)
    std::vector<int> all(N), coarse, cand;
    for (int k = 0; k < N; ++k) all[k] = k;
    for (int cy = 0; cy < ny; cy += GRID_COARSE)
        for (int cx = 0; cx < nx; cx += GRID_COARSE) {
            const int ex = std::min(cx + GRID_COARSE, nx), ey = std::min(cy + GRID_COARSE, ny);
            Rect R = {x0 + cx * h - eps, x0 + ex * h + eps, y0 + cy * h - eps, y0 + ey * h + eps};
            minmax_filter(R, wx, wy, all, coarse);
            for (int iy = cy; iy < ey; ++iy)
                for (int ix = cx; ix < ex; ++ix) {
                    Rect r = {x0 + ix * h - eps, x0 + (ix + 1) * h + eps, y0 + iy * h - eps, y0 + (iy + 1) * h + eps};
                    minmax_filter(r, wx, wy, coarse, cand);
                    if ((int)cand.size() <= GRID_REFINE_MAX) refine(r, wx, wy, cand);
                    int lo = N - 1, hi = 0;
                    for (int k : cand) { lo = std::min(lo, k); hi = std::max(hi, k); }
                    cells[(size_t)iy * nx + ix] = (uint32_t)lo | ((uint32_t)hi << 16);
                }
        }
    return true;
}

}  // namespace ce2e
