// fri_geometry.h — the tame-twindragon digit vectors and everything derived from them at
// compile time.  Shared by the host plan builder and the device kernels.
//
// Reference: crates/libfri/src/fractal.rs:51-86 (LITERALS, used verbatim: entries 1 and 2 are
// irregular, so the table is not regenerated) and stages/wavelet_transform.rs:47-53 (leaf k of
// a depth-D fractal sits at centre + sum_j bit_j(k) * LITERALS[j]).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define FRI_HD __host__ __device__
#else
#define FRI_HD
#endif

namespace fri {

struct Vec2 {
    int x, y;  // x = re (column), y = im (row)
};

constexpr int kNumLiterals = 30;
constexpr Vec2 kLiterals[kNumLiterals] = {
    {0, 1},        {-1, 1},      {2, 0},        {-3, -1},      {5, -1},        {1, 3},
    {-11, -1},     {9, -5},      {13, 7},       {-31, 3},      {5, -17},       {57, 11},
    {-67, 23},     {-47, -45},   {181, -1},     {-87, 91},     {-275, -89},    {449, -93},
    {101, 271},    {-999, -85},  {797, -457},   {1201, 627},   {-2795, 287},   {393, -1541},
    {5197, 967},   {-5983, 2115}, {-4411, -4049}, {16377, -181}, {-7555, 8279}, {-25199, -7917},
};

constexpr int kBaseDepth = 9;                 // BASE_FRAC_DEPTH, wavelet_transform.rs:39
constexpr int kTileLeaves = 1 << kBaseDepth;  // 512 pixels / coefficients per base tile
constexpr int kMaxDepth = 24;                 // deep-tree extension limit (CENTERS range, fractal.rs:33-49)

// Offset of leaf k (0 <= k < 2^nbits) using digit vectors [first, first + nbits).
FRI_HD constexpr Vec2 digit_sum(unsigned k, int first, int nbits)
{
    Vec2 o{0, 0};
    for (int j = 0; j < nbits; ++j)
        if ((k >> j) & 1u) {
            o.x += kLiterals[first + j].x;
            o.y += kLiterals[first + j].y;
        }
    return o;
}

// Bounding box of a base tile relative to its centre.
struct TileBox {
    int xmin, xmax, ymin, ymax;
};
constexpr TileBox tile_box(int depth)
{
    TileBox b{0, 0, 0, 0};
    for (unsigned k = 0; k < (1u << depth); ++k) {
        Vec2 o = digit_sum(k, 0, depth);
        if (o.x < b.xmin) b.xmin = o.x;
        if (o.x > b.xmax) b.xmax = o.x;
        if (o.y < b.ymin) b.ymin = o.y;
        if (o.y > b.ymax) b.ymax = o.y;
    }
    return b;
}
constexpr TileBox kBaseBox = tile_box(kBaseDepth);  // x in [-15, 30], y in [-8, 12]
static_assert(kBaseBox.xmin == -15 && kBaseBox.xmax == 30 && kBaseBox.ymin == -8 && kBaseBox.ymax == 12,
              "unexpected base tile footprint");
constexpr int kTileRows = kBaseBox.ymax - kBaseBox.ymin + 1;  // 21
constexpr int kTileCols = kBaseBox.xmax - kBaseBox.xmin + 1;  // 46

// Work split inside the warp that owns a (base tile, channel):
//   lane t holds two complete depth-3 subtrees, "A" = leaves 8t .. 8t+7 and "B" = leaves
//   256 + 8t .. 256 + 8t + 7.  With that split every heap-ordered coefficient block a lane
//   produces (or consumes) is part of a warp-contiguous run:
//     level 8: pos 256 + 4t + m (A) / 384 + 4t + m (B), m < 4   -> one 128-bit access each
//     level 7: pos 128 + 2t + m / 192 + 2t + m, m < 2            -> one 64-bit access each
//     level 6: pos 64 + t / 96 + t                               -> one 32-bit access each
//     levels 5..0 + DC (pos 0..63): exchanged across lanes, then one 64-bit access (2*lane).
constexpr int kSubLeaves = 8;  // leaves per lane per half
FRI_HD constexpr Vec2 sub_leaf(int i) { return digit_sum((unsigned)i, 0, 3); }          // digits 0..2
FRI_HD constexpr Vec2 lane_anchor(int lane) { return digit_sum((unsigned)lane, 3, 5); }  // digits 3..7
constexpr Vec2 kHalfB = kLiterals[8];  // offset of half B from half A: (13, 7)

}  // namespace fri
