// fri_order.cpp — the order in which the reference hands coefficients to its entropy coder
// (SURVEY.md §8(f) next-1), computed arithmetically on the host for a depth-9 plan.
//
// Restates crates/libfri/src/stages/wavelet_transform.rs:434-448 (get_global_position_map),
// :490-503 (is_pos_in_row_boundary), :505-654 (scan_level), :657-705 (sort_lattice) and the three
// scans of stages/entropy_coding.rs:283-329 that consume the result.  The reference answers
// "is there a level-L node at this position?" with one HashMap per level (511 inserts per tile);
// here the question is answered in O(1) from the geometry: position p holds leaf
// k = LUT[((p.x - ax) + 181 (p.y - ay)) mod 512] of the tile centred at p - offset(k), and it is a
// level-L node position iff the low 9 - L bits of k are zero and that tile is retained.
#include <algorithm>
#include <cstdint>
#include <string>
#include <thread>
#include <vector>

#include "fri_plan.h"

namespace fri {

namespace {

struct P {
    int x, y;
};
inline P operator+(P a, P b) { return P{a.x + b.x, a.y + b.y}; }

}  // namespace

// wavelet_transform.rs:71-90
void nearby_vectors(int depth, Vec2 out[6])
{
    Vec2 zl, zmd;
    if (depth == 1) { zl = {-1, 1}; zmd = {0, 2}; }
    else if (depth == 2) { zl = {-2, 0}; zmd = {0, -2}; }
    else if (depth == 3) { zl = {-3, -1}; zmd = {-1, -3}; }
    else {
        zl = {kLiterals[depth].x, kLiterals[depth].y};
        zmd = {kLiterals[depth + 1].x + zl.x, kLiterals[depth + 1].y + zl.y};
    }
    out[0] = zl;
    out[1] = {zl.x - zmd.x, zl.y - zmd.y};
    out[2] = {-zmd.x, -zmd.y};
    out[3] = {-zl.x, -zl.y};
    out[4] = {zmd.x - zl.x, zmd.y - zl.y};
    out[5] = zmd;
}

int LatticeIndex::tile_of(int cx, int cy) const
{
    constexpr Vec2 l9 = kLiterals[kBaseDepth], l10 = kLiterals[kBaseDepth + 1];
    const int64_t dx = cx - ax, dy = cy - ay;
    const int64_t na_ = dx * l10.y - (int64_t)l10.x * dy, nb_ = (int64_t)l9.x * dy - dx * l9.y;
    if (na_ % 512 != 0 || nb_ % 512 != 0) return -1;
    const int64_t a = na_ / 512 - amin, b = nb_ / 512 - bmin;
    if (a < 0 || b < 0 || a >= na || b >= nb) return -1;
    return tile_at[(size_t)b * na + a];
}

bool LatticeIndex::node_at(int level, int x, int y, int &tile, int &heap) const
{
    const int k = lut[mod512((x - ax) + 181 * (y - ay))];
    const int low = kBaseDepth - level;
    if (k & ((1 << low) - 1)) return false;
    tile = tile_of(x - off[k].x, y - off[k].y);
    if (tile < 0) return false;
    heap = (1 << level) + (k >> low);
    return true;
}

void build_lattice_index(const Plan &plan, LatticeIndex &L)
{
    const Geometry &g = plan.geo;
    const int n_tiles = g.n_fractals;
    L.ax = g.width / 2;
    L.ay = g.height / 2;
    for (unsigned k = 0; k < (unsigned)kTileLeaves; ++k) {
        L.off[k] = digit_sum(k, 0, kBaseDepth);
        L.lut[LatticeIndex::mod512(L.off[k].x + 181 * L.off[k].y)] = (uint16_t)k;
    }
    constexpr Vec2 l9 = kLiterals[kBaseDepth], l10 = kLiterals[kBaseDepth + 1];
    std::vector<std::pair<int, int>> ab(n_tiles);
    int amin = INT32_MAX, amax = INT32_MIN, bmin = INT32_MAX, bmax = INT32_MIN;
    for (int t = 0; t < n_tiles; ++t) {
        const int64_t dx = plan.centers[2 * t] - L.ax, dy = plan.centers[2 * t + 1] - L.ay;
        const int a = (int)((dx * l10.y - (int64_t)l10.x * dy) / 512), b = (int)(((int64_t)l9.x * dy - dx * l9.y) / 512);
        ab[t] = {a, b};
        amin = std::min(amin, a); amax = std::max(amax, a);
        bmin = std::min(bmin, b); bmax = std::max(bmax, b);
    }
    if (n_tiles == 0) { amin = amax = bmin = bmax = 0; }
    L.amin = amin; L.bmin = bmin;
    L.na = amax - amin + 1; L.nb = bmax - bmin + 1;
    L.tile_at.assign((size_t)L.na * L.nb, -1);
    for (int t = 0; t < n_tiles; ++t) L.tile_at[(size_t)(ab[t].second - bmin) * L.na + (ab[t].first - amin)] = t;
}

std::string build_emission_order(const Plan &plan, std::vector<uint32_t> &order)
{
    const Geometry &g = plan.geo;
    if (g.depth != kBaseDepth) return "the emission order is defined for depth 9 only (sort_lattice uses BASE_FRAC_DEPTH maps)";
    const int n_tiles = g.n_fractals;
    order.clear();
    if (n_tiles == 0) return {};

    LatticeIndex L;
    build_lattice_index(plan, L);

    // :663-682 — bounds over the level-8 node positions (the even leaves of every retained tile)
    int min_real = INT32_MAX, max_real = INT32_MIN, min_imag = INT32_MAX, max_imag = INT32_MIN;
    for (int t = 0; t < n_tiles; ++t)
        for (int k = 0; k < kTileLeaves; k += 2) {
            const int x = plan.centers[2 * t] + L.off[k].x, y = plan.centers[2 * t + 1] + L.off[k].y;
            min_real = std::min(min_real, x); max_real = std::max(max_real, x);
            min_imag = std::min(min_imag, y); max_imag = std::max(max_imag, y);
        }
    auto in_box = [&](P p) { return p.y <= max_imag && p.y >= min_imag && p.x <= max_real && p.x >= min_real; };
    const int64_t guard = 64ll * ((int64_t)(max_real - min_real + 64) * (max_imag - min_imag + 64)) + 1024;  // runaway stop

    order.resize((size_t)n_tiles * kTileLeaves);  // DCs, roots, then levels 1..8: every level owns its segment
    const P center{L.ax, L.ay};
    // The levels are scanned independently of one another (one host thread each: level 8 is half of the work).
    auto scan_one_level = [&](int level) -> std::string {
        // ---- scan_level (:505-654), statement for statement
        Vec2 nv[6];
        nearby_vectors(kBaseDepth - level, nv);
        P vec[6];
        for (int i = 0; i < 6; ++i) vec[i] = P{nv[i].x, nv[i].y};
        const P row_dir = vec[3], rev_row_dir = vec[0], col_dir = vec[1], rev_col_dir = vec[4];
        const bool seven = kBaseDepth - level == 2;  // `depth - level != 2` is false: alternating steps
        auto has = [&](P p) { int t, h; return L.node_at(level, p.x, p.y, t, h); };
        int64_t steps = 0;

        P first = center;
        int mod = 0;
        if (!has(center + rev_row_dir) && has(center + P{-1, -1})) mod = 1;
        P last_seen = first;
        auto step_back = [&]() {
            if (!seven) first = first + rev_row_dir;
            else {
                first = first + ((mod % 2 == 0) ? rev_row_dir : P{-1, -1});
                ++mod;
            }
        };
        while (has(first)) {
            last_seen = first;
            step_back();
            if (++steps > guard) return "emission order: scan did not terminate";
        }
        for (;;) {  // find first row
            P fwd = first, bwd = first;
            bool empty = true;
            while ((fwd.y <= max_imag && fwd.y >= min_imag) || (bwd.y <= max_imag && bwd.y >= min_imag) ||
                   (fwd.x <= max_real && fwd.x >= min_real) || (bwd.x <= max_real && bwd.x >= min_real)) {
                fwd = fwd + col_dir;
                bwd = bwd + rev_col_dir;
                if (has(fwd)) { last_seen = fwd; empty = false; break; }
                if (has(bwd)) { last_seen = bwd; empty = false; break; }
                if (++steps > guard) return "emission order: scan did not terminate";
            }
            if (empty) { first = last_seen; break; }
            step_back();
        }
        while (in_box(first)) {  // scanning backwards find first column
            first = first + rev_col_dir;
            if (has(first)) last_seen = first;
        }
        first = last_seen;
        mod = 1;

        const size_t level_begin = level == 0 ? 0 : (size_t)n_tiles << level;  // level 0 fills both leading blocks
        size_t count = 0;
        const size_t expect = (size_t)n_tiles << level;
        bool done = false;
        while (!done) {  // fill plane in sorted order
            P scan = first;
            for (;;) {
                int t, h;
                if (L.node_at(level, scan.x, scan.y, t, h)) {
                    if (count < expect) {
                        if (level == 0) {
                            order[count] = (uint32_t)t * kTileLeaves;                 // first scan: coefficient 0
                            order[(size_t)n_tiles + count] = (uint32_t)t * kTileLeaves + 1;  // second scan: coefficient 1
                        } else {
                            order[level_begin + count] = (uint32_t)t * kTileLeaves + (uint32_t)h;
                        }
                    }
                    ++count;
                }
                if ((scan.y > max_imag || scan.y < min_imag) || (col_dir.y == 0 && (scan.x > max_real || scan.x < min_real))) break;
                scan = scan + col_dir;
                if (++steps > guard) return "emission order: scan did not terminate";
            }
            if (!seven) first = first + row_dir;
            else {
                first = first + ((mod % 2 == 0) ? P{1, 1} : row_dir);
                ++mod;
            }
            while (!has(first)) {
                first = first + col_dir;
                const bool in_row = std::abs(row_dir.x) > std::abs(row_dir.y)
                                        ? (first.y >= min_imag && first.y <= max_imag)
                                        : (first.x >= min_real && first.x <= max_real);  // :490-503
                if (!in_row) { done = true; break; }
                if (++steps > guard) return "emission order: scan did not terminate";
            }
            if (done) break;
            last_seen = first;
            while (in_box(first)) {
                first = first + rev_col_dir;
                if (has(first)) last_seen = first;
            }
            first = last_seen;
        }
        if (count != expect)  // the reference's assert_eq! at :701 — it would panic for this image size
            return "the reference's sort_lattice assertion (wavelet_transform.rs:701) fails for this image size: level " +
                   std::to_string(level) + " scan visits " + std::to_string(count) + " of " + std::to_string(expect) + " nodes";
        return {};
    };
    std::string level_err[kBaseDepth];
    {
        std::vector<std::thread> workers;
        for (int level = 0; level < kBaseDepth; ++level)
            workers.emplace_back([&, level] {
                try {
                    level_err[level] = scan_one_level(level);
                } catch (const std::exception &e) {
                    level_err[level] = std::string("emission order: ") + e.what();
                }
            });
        for (auto &t : workers) t.join();
    }
    for (int level = 0; level < kBaseDepth; ++level)
        if (!level_err[level].empty()) return level_err[level];
    // every (tile, coefficient) must appear exactly once
    std::vector<uint8_t> seen(order.size(), 0);
    for (uint32_t v : order) {
        if (v >= seen.size() || seen[v]) return "emission order: a node was visited twice";
        seen[v] = 1;
    }
    return {};
}

}  // namespace fri
