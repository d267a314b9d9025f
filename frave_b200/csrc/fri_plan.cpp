// fri_plan.cpp — builds the fractal lattice and the CTA work list on the host.
#include "fri_plan.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <unordered_set>

namespace fri {

namespace {

inline uint64_t key_of(int32_t a, int32_t b) { return (uint64_t)(uint32_t)a << 32 | (uint32_t)b; }

struct LatticeFractal {
    int32_t a, b;    // centre = anchor + a*L[depth] + b*L[depth+1]
    int32_t cx, cy;
    int64_t inside;  // number of leaves inside the image
};

struct BaseTile {
    int32_t a, b;  // base-lattice coordinates: centre = anchor + a*L[9] + b*L[10]
    int32_t cx, cy;
    uint32_t unit;  // fractal_index << sub_bits | sub_tile
};

inline int floor_div(int v, int d) { return v >= 0 ? v / d : -((-v + d - 1) / d); }

// Leaves of the base tile centred at (cx, cy) that fall inside the image.
int base_tile_inside(int32_t cx, int32_t cy, int32_t W, int32_t H)
{
    if (cx + kBaseBox.xmin >= 0 && cx + kBaseBox.xmax < W && cy + kBaseBox.ymin >= 0 && cy + kBaseBox.ymax < H)
        return kTileLeaves;
    if (cx + kBaseBox.xmax < 0 || cx + kBaseBox.xmin >= W || cy + kBaseBox.ymax < 0 || cy + kBaseBox.ymin >= H)
        return 0;
    int cnt = 0;
    for (unsigned k = 0; k < (unsigned)kTileLeaves; ++k) {
        Vec2 o = digit_sum(k, 0, kBaseDepth);
        int x = cx + o.x, y = cy + o.y;
        cnt += x >= 0 && y >= 0 && x < W && y < H;
    }
    return cnt;
}

}  // namespace

void fractal_mask(int depth, int32_t cx, int32_t cy, int32_t width, int32_t height, uint32_t *out)
{
    // Some(pos) <=> at least one leaf of pos's subtree is inside the image: try_apply yields
    // None only when both inputs are None (wavelet_transform.rs:14-26), and a leaf is Some
    // exactly when get_pixel's bounds check passes (images.rs:90).  coefficient[0] is the
    // root's low-pass value (:221), Some under the same condition as the root's residue.
    const size_t n = (size_t)1 << depth;
    std::vector<uint8_t> some(2 * n);
    {
        // leaf k sits at centre + sum_j bit_j(k) * LITERALS[j]: build the offsets by doubling
        std::vector<Vec2> off(n);
        off[0] = Vec2{0, 0};
        for (int j = 0; j < depth; ++j)
            for (size_t k = 0; k < ((size_t)1 << j); ++k)
                off[k + ((size_t)1 << j)] = Vec2{off[k].x + kLiterals[j].x, off[k].y + kLiterals[j].y};
        for (size_t k = 0; k < n; ++k) {
            const int x = cx + off[k].x, y = cy + off[k].y;
            some[n + k] = x >= 0 && y >= 0 && x < width && y < height;
        }
    }
    for (size_t pos = n - 1; pos >= 1; --pos) some[pos] = some[2 * pos] | some[2 * pos + 1];
    std::memset(out, 0, sizeof(uint32_t) * (n / 32));
    if (some[1]) out[0] |= 1u;
    for (size_t pos = 1; pos < n; ++pos)
        if (some[pos]) out[pos >> 5] |= 1u << (pos & 31);
}

double gather_conflict_degree(int pitch, int pixel_bytes, int sample_bytes, int *worst_out)
{
    // For half h and leaf slot i every lane reads the sample at
    //   (anchor(lane) + leaf(i) + h*kHalfB) -> y * pitch + x * pixel_bytes (+ channel offset).
    // Shared memory serves one 4-byte word per bank per wavefront; count distinct words per bank.
    double total = 0;
    int cases = 0, worst_all = 0;
    for (int ch = 0; ch < pixel_bytes / sample_bytes; ++ch)
        for (int h = 0; h < 2; ++h)
            for (int i = 0; i < kSubLeaves; ++i) {
                int words[32];
                for (int lane = 0; lane < 32; ++lane) {
                    Vec2 a = lane_anchor(lane), l = sub_leaf(i);
                    int x = a.x + l.x + h * kHalfB.x + 64, y = a.y + l.y + h * kHalfB.y + 32;
                    words[lane] = (y * pitch + x * pixel_bytes + ch * sample_bytes) >> 2;
                }
                int worst = 0;
                for (int bank = 0; bank < 32; ++bank) {
                    int distinct = 0, seen[32];
                    for (int lane = 0; lane < 32; ++lane) {
                        if ((words[lane] & 31) != bank) continue;
                        bool dup = false;
                        for (int s = 0; s < distinct; ++s) dup |= seen[s] == words[lane];
                        if (!dup) seen[distinct++] = words[lane];
                    }
                    worst = std::max(worst, distinct);
                }
                total += worst;
                worst_all = std::max(worst_all, worst);
                ++cases;
            }
    if (worst_out) *worst_out = worst_all;
    return total / cases;
}

std::string build_plan(Plan &plan, uint32_t width, uint32_t height, uint32_t channels, uint32_t depth,
                       uint32_t sample_bytes, int group_a, int group_b)
{
    if (width == 0 || height == 0) return "width and height must be positive";
    if (width > (1u << 20) || height > (1u << 20)) return "image dimension above 2^20 is not supported";
    if (channels != 1 && channels != 3) return "channels must be 1 (Luma) or 3 (RGB / YCbCr)";
    if (sample_bytes != 1 && sample_bytes != 2) return "sample_bytes must be 1 or 2";
    if (depth < (uint32_t)kBaseDepth || depth > (uint32_t)kMaxDepth) return "depth must be in [9, 24]";
    const int csz = (int)(channels * sample_bytes);
    int tiles_per_warp = 2;
    if (const char *env = std::getenv("FRI_TILES_PER_WARP")) tiles_per_warp = std::max(1, std::atoi(env));
    if (group_a == 0 || group_b == 0) {
        if (const char *env = std::getenv("FRI_GROUP")) std::sscanf(env, "%dx%d", &group_a, &group_b);  // tuning knob
    }
    if (group_a == 0 || group_b == 0) {
        // Default CTA group: enough (tile, channel) tasks per CTA, small enough that four CTAs
        // fit in one SM's shared memory.
        group_a = csz == 1 ? 8 : 4;
        group_b = csz == 6 ? 2 : 4;
    }
    if (group_a < 1 || group_b < 1 || group_a * group_b > kMaxGroupTiles) return "bad group shape";

    const int32_t W = (int32_t)width, H = (int32_t)height;
    const int sub_bits = (int)depth - kBaseDepth;
    const Vec2 la = kLiterals[depth], lb = kLiterals[depth + 1];
    // get_nearby_vectors (wavelet_transform.rs:80-88) in lattice coordinates:
    // zl = L[d] -> (1,0); zmd = L[d+1] + L[d] -> (1,1);
    // [zl, zl-zmd, -zmd, -zl, zmd-zl, zmd] = (1,0) (0,-1) (-1,-1) (-1,0) (0,1) (1,1)
    static const int nb[6][2] = {{1, 0}, {0, -1}, {-1, -1}, {-1, 0}, {0, 1}, {1, 1}};
    const int32_t ax = W / 2, ay = H / 2;  // :452

    // ---- fractal_divide (:450-484).  An in-bounds centre (0 <= re <= w, 0 <= im <= h,
    // inclusive, :459-463) is expanded; an out-of-bounds one is kept as a fringe fractal.  The
    // final key set does not depend on queue order, so a plain visited-set BFS reproduces it.
    std::vector<LatticeFractal> built;
    {
        std::unordered_set<uint64_t> seen;
        seen.reserve((size_t)(((uint64_t)W * H >> depth) * 5 / 4 + 64));
        std::deque<std::pair<int32_t, int32_t>> queue;
        queue.emplace_back(0, 0);
        seen.insert(key_of(0, 0));
        while (!queue.empty()) {
            auto [a, b] = queue.front();
            queue.pop_front();
            const int32_t cx = ax + a * la.x + b * lb.x, cy = ay + a * la.y + b * lb.y;
            built.push_back({a, b, cx, cy, 0});
            if (cx < 0 || cy < 0 || cx > W || cy > H) continue;  // boundary fractal: not expanded
            for (auto &d : nb) {
                const int32_t na = a + d[0], nbb = b + d[1];
                if (seen.insert(key_of(na, nbb)).second) queue.emplace_back(na, nbb);
            }
        }
    }
    plan.n_built = (uint32_t)built.size();

    // ---- retain (:415-416): keep fractals whose DC is Some, i.e. with >= 1 leaf inside the image.
    std::vector<LatticeFractal> kept;
    kept.reserve(built.size());
    uint64_t covered = 0;
    uint32_t n_full = 0;
    const int64_t leaves = (int64_t)1 << depth;
    for (auto &t : built) {
        int64_t cnt = 0;
        for (uint32_t s = 0; s < (1u << sub_bits); ++s) {
            Vec2 o = digit_sum(s, kBaseDepth, sub_bits);
            cnt += base_tile_inside(t.cx + o.x, t.cy + o.y, W, H);
        }
        t.inside = cnt;
        if (cnt > 0) {
            kept.push_back(t);
            covered += (uint64_t)cnt;
            n_full += cnt == leaves;
        }
    }
    plan.pixels_covered = covered;
    plan.n_full = n_full;

    // ---- base tiles.  At depth 9 a fractal is one base tile; deeper, fractal f contributes
    // 2^sub_bits base tiles at centre + sum_{j>=9} bit_j * L[j] (wavelet_transform.rs:47-53),
    // all of which lie on the base lattice anchor + a*L[9] + b*L[10].
    const Vec2 l9 = kLiterals[kBaseDepth], l10 = kLiterals[kBaseDepth + 1];
    const int det = l9.x * l10.y - l10.x * l9.y;  // 512
    if (sub_bits > 0)
        std::sort(kept.begin(), kept.end(), [](const LatticeFractal &l, const LatticeFractal &r) {
            return l.cy != r.cy ? l.cy < r.cy : l.cx < r.cx;
        });
    std::vector<BaseTile> tiles;
    tiles.reserve(kept.size() << sub_bits);
    plan.absent_unit.clear();
    for (size_t f = 0; f < kept.size(); ++f)
        for (uint32_t s = 0; s < (1u << sub_bits); ++s) {
            Vec2 o = digit_sum(s, kBaseDepth, sub_bits);
            BaseTile bt;
            bt.cx = kept[f].cx + o.x;
            bt.cy = kept[f].cy + o.y;
            const int dx = bt.cx - ax, dy = bt.cy - ay;
            const int64_t na = (int64_t)dx * l10.y - (int64_t)l10.x * dy;
            const int64_t nbn = (int64_t)l9.x * dy - (int64_t)dx * l9.y;
            if (na % det != 0 || nbn % det != 0) return "internal error: base tile off the base lattice";
            bt.a = (int32_t)(na / det);
            bt.b = (int32_t)(nbn / det);
            bt.unit = (uint32_t)((f << sub_bits) | s);
            if (sub_bits > 0 && base_tile_inside(bt.cx, bt.cy, W, H) == 0) {
                plan.absent_unit.push_back(bt.unit);  // nothing of this base tile is inside the image
                continue;
            }
            tiles.push_back(bt);
        }

    // ---- group base tiles into A x B blocks of base-lattice coordinates, group-major order.
    int32_t amin = 0, bmin = 0;
    for (auto &t : tiles) { amin = std::min(amin, t.a); bmin = std::min(bmin, t.b); }
    auto group_key = [&](const BaseTile &t) {
        return std::pair<int, int>(floor_div(t.b - bmin, group_b), floor_div(t.a - amin, group_a));
    };
    // Groups are ordered top to bottom (by the y, then x, of their first lattice cell), tiles inside a
    // group by (b, a).  A contiguous range of groups then needs a compact range of pixel rows, which
    // is what lets the host entry points stream a frame through the device in bands.
    auto group_origin = [&](const BaseTile &t) {
        const auto gk = group_key(t);
        const int a0 = amin + gk.second * group_a, b0 = bmin + gk.first * group_b;
        return std::pair<int, int>(a0 * l9.y + b0 * l10.y, a0 * l9.x + b0 * l10.x);
    };
    std::sort(tiles.begin(), tiles.end(), [&](const BaseTile &l, const BaseTile &r) {
        auto gl = group_origin(l), gr = group_origin(r);
        if (gl != gr) return gl < gr;
        if (l.b != r.b) return l.b < r.b;
        return l.a < r.a;
    });
    if (sub_bits == 0) {
        // depth 9: fractal order == base tile order, so a CTA's coefficient blocks are adjacent
        std::vector<LatticeFractal> reordered(kept.size());
        for (size_t i = 0; i < tiles.size(); ++i) {
            reordered[i] = kept[tiles[i].unit];
            tiles[i].unit = (uint32_t)i;
        }
        kept.swap(reordered);
    }

    // Region geometry: union bounding box of the A x B tile footprints relative to tile (0,0).
    int gxmin = 0, gxmax = 0, gymin = 0, gymax = 0;
    for (int j = 0; j < group_b; ++j)
        for (int i = 0; i < group_a; ++i) {
            int ox = i * l9.x + j * l10.x, oy = i * l9.y + j * l10.y;
            gxmin = std::min(gxmin, ox); gxmax = std::max(gxmax, ox);
            gymin = std::min(gymin, oy); gymax = std::max(gymax, oy);
        }
    Geometry &g = plan.geo;
    std::memset(&g, 0, sizeof(g));
    g.width = W; g.height = H;
    g.channels = (int32_t)channels; g.sample_bytes = (int32_t)sample_bytes;
    g.depth = (int32_t)depth;
    g.sub_bits = sub_bits;
    g.group_a = group_a; g.group_b = group_b;
    g.tiles_per_warp = tiles_per_warp;
    g.region_w = gxmax - gxmin + kTileCols;
    g.region_h = gymax - gymin + kTileRows;
    g.row_bytes = g.region_w * csz;
    g.chunks_per_row = g.row_bytes / 16 + 2;
    g.own_words = (g.region_w + 31) / 32;
    for (int j = 0; j < group_b; ++j)
        for (int i = 0; i < group_a; ++i) {
            g.tile_rel_x[j * group_a + i] = (int16_t)(i * l9.x + j * l10.x - gxmin - kBaseBox.xmin);
            g.tile_rel_y[j * group_a + i] = (int16_t)(i * l9.y + j * l10.y - gymin - kBaseBox.ymin);
        }
    g.row_stride = (int64_t)W * csz;
    g.frame_bytes = g.row_stride * H;
    // Shared-memory pitch: rows are staged as 16-byte chunks aligned in GLOBAL memory, so a
    // row's shared-memory image keeps the global address modulo 16: pitch == W*C*sz (mod 16).
    // 32 spare bytes keep the aligned covers of adjacent rows disjoint.  Among the 8 residues
    // mod 128 that satisfy this, take the one with the fewest bank conflicts in the gather.
    {
        int minp = g.row_bytes + 32;
        int p0 = minp + (int)(((g.row_stride - minp) % 16 + 16) % 16);
        int best = p0;
        double best_deg = 1e9;
        for (int c = 0; c < 8; ++c) {
            int p = p0 + 16 * c;
            double deg = gather_conflict_degree(p, csz, (int)sample_bytes, nullptr);
            if (deg < best_deg - 1e-9) { best_deg = deg; best = p; }
        }
        g.pitch = best;
    }

    // Ownership bitmap of a full group: which region pixels belong to one of its A x B tiles.
    plan.ownership.assign((size_t)g.region_h * g.own_words, 0u);
    for (int s = 0; s < group_a * group_b; ++s)
        for (unsigned k = 0; k < (unsigned)kTileLeaves; ++k) {
            Vec2 o = digit_sum(k, 0, kBaseDepth);
            int x = g.tile_rel_x[s] + o.x, y = g.tile_rel_y[s] + o.y;
            plan.ownership[(size_t)y * g.own_words + (x >> 5)] |= 1u << (x & 31);
        }

    for (int s = 0; s < group_a * group_b; ++s) g.tile_off[s] = g.tile_rel_y[s] * g.pitch + g.tile_rel_x[s] * csz;

    // Pixels of the tile slots every warp processes first (slot = warp index in a full group).
    const int first_slots = std::min(group_a * group_b, std::max(1, (group_a * group_b + tiles_per_warp - 1) / tiles_per_warp));
    std::vector<uint32_t> own_first((size_t)g.region_h * g.own_words, 0u);
    for (int s = 0; s < first_slots; ++s)
        for (unsigned k = 0; k < (unsigned)kTileLeaves; ++k) {
            Vec2 o = digit_sum(k, 0, kBaseDepth);
            int x = g.tile_rel_x[s] + o.x, y = g.tile_rel_y[s] + o.y;
            own_first[(size_t)y * g.own_words + (x >> 5)] |= 1u << (x & 31);
        }

    // Row spans for the encoder's bulk-copy staging of interior groups.
    g.n_rows_first = 0;
    if (g.region_h <= kMaxRegionRows && g.row_bytes < 65536) {
        std::vector<int> first_rows, rest_rows;
        for (int r = 0; r < g.region_h; ++r) {
            int lo = -1, hi = -1;
            bool in_first = false;
            for (int x = 0; x < g.region_w; ++x) {
                if ((plan.ownership[(size_t)r * g.own_words + (x >> 5)] >> (x & 31)) & 1u) {
                    if (lo < 0) lo = x;
                    hi = x;
                }
                in_first |= (own_first[(size_t)r * g.own_words + (x >> 5)] >> (x & 31)) & 1u;
            }
            g.row_lo[r] = (uint16_t)(lo < 0 ? 0 : lo * csz);
            g.row_hi[r] = (uint16_t)(lo < 0 ? 0 : (hi + 1) * csz);
            (in_first ? first_rows : rest_rows).push_back(r);
        }
        g.n_rows_first = (int32_t)first_rows.size();
        int k = 0;
        for (int r : first_rows) g.row_order[k++] = (uint8_t)r;
        for (int r : rest_rows) g.row_order[k++] = (uint8_t)r;
    }

    // Chunk lists.  Shared-memory byte r*pitch + phase + b holds region byte (r, b); chunks are
    // the 16-byte aligned pieces of shared memory (== aligned pieces of global memory).
    std::vector<std::vector<std::pair<uint32_t, uint16_t>>> lists(16);
    std::vector<std::vector<uint32_t>> stage(16);
    g.list_cap = 0;
    for (int phase = 0; phase < 16; ++phase) {
        std::vector<std::pair<uint32_t, uint16_t>> full, part;
        std::vector<uint32_t> st_first, st_rest;
        for (int r = 0; r < g.region_h; ++r) {
            const int srow = r * g.pitch + phase, sbase = srow & ~15;
            for (int c = 0; c < g.chunks_per_row; ++c) {
                const int b0 = sbase + 16 * c - srow;
                uint32_t m = 0;
                bool first = false;
                for (int j = 0; j < 16; ++j) {
                    const int b = b0 + j;
                    if (b < 0 || b >= g.row_bytes) continue;
                    const int x = b / csz;
                    if ((plan.ownership[(size_t)r * g.own_words + (x >> 5)] >> (x & 31)) & 1u) m |= 1u << j;
                    first |= (own_first[(size_t)r * g.own_words + (x >> 5)] >> (x & 31)) & 1u;
                }
                if (m == 0) continue;
                const uint32_t entry = (uint32_t)r << 16 | (uint32_t)((sbase + 16 * c) >> 4);
                (m == 0xffffu ? full : part).emplace_back(entry, (uint16_t)m);
                (first ? st_first : st_rest).push_back(entry);
            }
        }
        g.list_full[phase] = (int32_t)full.size();
        g.list_all[phase] = (int32_t)(full.size() + part.size());
        g.list_cap = std::max(g.list_cap, g.list_all[phase]);
        lists[phase] = std::move(full);
        lists[phase].insert(lists[phase].end(), part.begin(), part.end());
        g.stage_first[phase] = (int32_t)st_first.size();
        stage[phase] = std::move(st_first);
        stage[phase].insert(stage[phase].end(), st_rest.begin(), st_rest.end());
    }
    if ((size_t)g.region_h * g.pitch > ((size_t)1 << 20) || g.region_h >= 4096) return "group region too large";
    // edge lists: words and samples of the partially owned chunks
    {
        std::vector<std::vector<uint32_t>> edges(16);
        g.edge_cap = 0;
        for (int phase = 0; phase < 16; ++phase) {
            std::vector<uint32_t> words, samples;
            for (size_t k = (size_t)g.list_full[phase]; k < lists[phase].size(); ++k) {
                const uint32_t entry = lists[phase][k].first, m = lists[phase][k].second;
                const uint32_t r = entry >> 16, sb = (entry & 0xffffu) << 4;
                for (uint32_t w = 0; w < 4; ++w) {
                    const uint32_t nib = (m >> (4 * w)) & 15u;
                    if (nib == 15u) {
                        words.push_back(r << 20 | (sb + 4 * w));
                    } else {
                        for (uint32_t j = 0; j < 4; j += sample_bytes)
                            if ((nib >> j) & 1u) samples.push_back(r << 20 | (sb + 4 * w + j));
                    }
                }
            }
            g.edge_words[phase] = (int32_t)words.size();
            g.edge_samples[phase] = (int32_t)samples.size();
            edges[phase] = std::move(words);
            edges[phase].insert(edges[phase].end(), samples.begin(), samples.end());
            g.edge_cap = std::max(g.edge_cap, (int32_t)edges[phase].size());
        }
        plan.edge_list.assign((size_t)16 * std::max(g.edge_cap, 1), 0u);
        for (int phase = 0; phase < 16; ++phase)
            std::copy(edges[phase].begin(), edges[phase].end(), plan.edge_list.begin() + (size_t)phase * g.edge_cap);
    }
    plan.chunk_list.assign((size_t)16 * g.list_cap, 0u);
    plan.chunk_mask.assign((size_t)16 * g.list_cap, 0);
    plan.stage_list.assign((size_t)16 * g.list_cap, 0u);
    for (int phase = 0; phase < 16; ++phase)
        for (size_t k = 0; k < lists[phase].size(); ++k) {
            plan.chunk_list[(size_t)phase * g.list_cap + k] = lists[phase][k].first;
            plan.chunk_mask[(size_t)phase * g.list_cap + k] = lists[phase][k].second;
            plan.stage_list[(size_t)phase * g.list_cap + k] = stage[phase][k];
        }

    plan.centers.clear(); plan.full.clear(); plan.groups.clear(); plan.tile_unit.clear();
    plan.centers.reserve(kept.size() * 2);
    plan.full.reserve(kept.size());
    for (auto &f : kept) {
        plan.centers.push_back(f.cx);
        plan.centers.push_back(f.cy);
        plan.full.push_back(f.inside == leaves);
    }
    plan.tile_unit.reserve(tiles.size());
    for (size_t i = 0; i < tiles.size();) {
        auto gk = group_key(tiles[i]);
        const int a0 = amin + gk.second * group_a, b0 = bmin + gk.first * group_b;
        GroupDesc gd{};
        gd.x0 = ax + a0 * l9.x + b0 * l10.x + gxmin + kBaseBox.xmin;
        gd.y0 = ay + a0 * l9.y + b0 * l10.y + gymin + kBaseBox.ymin;
        gd.tile_base = (uint32_t)i;
        gd.tile_mask = 0;
        size_t e = i;
        while (e < tiles.size() && group_key(tiles[e]) == gk) {
            const BaseTile &t = tiles[e];
            gd.tile_mask |= 1u << ((t.b - b0) * group_a + (t.a - a0));
            plan.tile_unit.push_back(t.unit);
            ++e;
        }
        plan.groups.push_back(gd);
        i = e;
    }
    g.n_groups = (int32_t)plan.groups.size();
    g.n_base_tiles = (int32_t)tiles.size();
    g.n_fractals = (int32_t)kept.size();
    g.coefs_per_frame = (int64_t)kept.size() * channels * leaves;
    return {};
}

}  // namespace fri
