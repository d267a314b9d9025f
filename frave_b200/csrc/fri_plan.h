// fri_plan.h — host-side lattice plan: which tiles exist, in which order, and how the kernels
// are launched over them.  Pure C++ (no CUDA), so it is unit-testable without a GPU.
//
// Restates crates/libfri/src/stages/wavelet_transform.rs:450-484 (fractal_divide BFS),
// :71-90 (get_nearby_vectors), :415-416 (retain) in lattice coordinates.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "fri_geometry.h"

namespace fri {

constexpr int kMaxRegionRows = 128;  // staged rows covered by the row-span tables
constexpr int kMaxGroupTiles = 32;  // tiles per CTA group (bits of GroupDesc::tile_mask)

// One CTA's worth of work: up to A x B lattice-adjacent base tiles whose pixels are staged
// together through shared memory.
struct GroupDesc {
    int32_t x0, y0;      // image coordinates of the staged region's top-left pixel (may be < 0)
    uint32_t tile_mask;  // bit (j*A + i) set <=> base tile (a0 + i, b0 + j) is present
    uint32_t tile_base;  // plan index of the group's first present base tile
};
static_assert(sizeof(GroupDesc) == 16, "GroupDesc is uploaded verbatim");

// Launch geometry shared by the encode and decode kernels (passed by value to the kernels).
struct Geometry {
    int32_t width, height;
    int32_t channels, sample_bytes;
    int32_t depth;               // fractal depth (9 = reference)
    int32_t sub_bits;            // depth - 9: a fractal is 2^sub_bits base tiles
    int32_t group_a, group_b;    // group shape in base-lattice coordinates
    int32_t region_w, region_h;  // staged region in pixels
    int32_t row_bytes;           // region_w * channels * sample_bytes
    int32_t pitch;               // shared-memory row pitch in bytes (== width*C*sz mod 16)
    int32_t chunks_per_row;      // upper bound on 16-byte chunks covering one staged row
    int32_t own_words;           // 32-bit words per row of the ownership bitmap
    int32_t n_groups;
    int32_t n_base_tiles;        // base tiles processed per frame: those of the retained fractals with a pixel inside the image
    int32_t n_fractals;          // retained fractals per frame (== n_base_tiles at depth 9)
    int32_t list_cap;            // entries per phase in the chunk list
    int32_t tiles_per_warp;      // base tiles each warp of a CTA processes (full group)
    int32_t independent_calls;   // per launch: the kernel does not wait for the previous kernel of its stream (fri_plan_set_independent_calls)
    int64_t row_stride;          // width * channels * sample_bytes
    int64_t frame_bytes;         // height * row_stride
    int64_t coefs_per_frame;     // n_fractals * channels * 2^depth
    int16_t tile_rel_x[kMaxGroupTiles];  // base tile centre relative to the region origin
    int16_t tile_rel_y[kMaxGroupTiles];
    int32_t tile_off[kMaxGroupTiles];    // tile_rel_y * pitch + tile_rel_x * pixel_bytes
    int32_t list_full[16];               // per phase: fully owned 16-byte chunks (listed first)
    int32_t list_all[16];                // per phase: all chunks that hold at least one owned byte
    int32_t stage_first[16];             // per phase: leading stage_list entries needed by the first tile of every warp
    int32_t edge_cap;                    // entries per phase in the edge list
    int32_t edge_words[16];              // per phase: fully owned 32-bit words of the partially owned chunks (listed first)
    int32_t edge_samples[16];            // per phase: owned samples of their remaining, partially owned words
    // Row spans of a full group's footprint, for the encoder's bulk-copy staging of interior groups (one
    // cp.async.bulk per staged row): owned bytes of region row r lie in [row_lo[r], row_hi[r]) (holes
    // inside a span belong to neighbouring groups and are fetched along); row_order lists the rows that
    // hold pixels of every warp's first tile first (n_rows_first of them).  n_rows_first == 0: not available.
    int32_t n_rows_first;
    uint16_t row_lo[kMaxRegionRows], row_hi[kMaxRegionRows];
    uint8_t row_order[kMaxRegionRows];
};

struct Plan {
    Geometry geo{};
    uint32_t n_built = 0;         // fractals the reference's BFS constructs (incl. dropped fringe)
    uint32_t n_full = 0;          // retained fractals with all leaves inside the image
    uint64_t pixels_covered = 0;  // in-image pixels owned by retained fractals
    std::vector<int32_t> centers;     // [n_fractals][2] (re, im), plan order
    std::vector<uint8_t> full;        // [n_fractals]
    std::vector<GroupDesc> groups;    // [n_groups]
    std::vector<uint32_t> tile_unit;  // [n_base_tiles] fractal_index << sub_bits | sub_tile (empty at depth 9)
    // depth > 9: base tiles of retained fractals that lie entirely outside the image.  All their coefficients are
    // `None`; the transform kernels skip them, the encoder zero-fills their slots of the dense array.
    std::vector<uint32_t> absent_unit;
    std::vector<uint32_t> ownership;  // [region_h][own_words]: bit x set <=> region pixel belongs to the group
    // Byte-ownership of the staged region cut into the 16-byte chunks the kernels move, for each
    // of the 16 possible phases (global address of the region's first byte mod 16):
    // chunk_list[phase][k] = row << 16 | (shared-memory offset >> 4) of the k-th chunk holding owned
    // bytes, fully owned chunks first; chunk_mask[phase][k]: bit j <=> byte j of that chunk is owned.
    std::vector<uint32_t> chunk_list;  // [16][list_cap]
    std::vector<uint16_t> chunk_mask;  // [16][list_cap]
    // The partially owned chunks (the group's fractal outline) resolved to what the decoder's write-out of an
    // interior group actually stores: edge_list[phase][k] = row << 20 | shared-memory byte offset of, first, every
    // fully owned 32-bit word, then every owned sample of the words the outline passes through.
    std::vector<uint32_t> edge_list;   // [16][edge_cap]
    // the same chunks in the order the encoder stages them: first those that hold a pixel of tile
    // slots 0 .. warps-1 (every warp's first tile), then the rest
    std::vector<uint32_t> stage_list;  // [16][list_cap]
};

// Returns an empty string on success, else an error message.  group_a/group_b == 0 picks the
// default group shape for the pixel size.
std::string build_plan(Plan &plan, uint32_t width, uint32_t height, uint32_t channels, uint32_t depth,
                       uint32_t sample_bytes, int group_a, int group_b);

// Order in which the reference's entropy coder consumes the coefficients of one channel
// (entropy_coding.rs:283-329 over sort_lattice, wavelet_transform.rs:505-705), depth 9 only:
// order[i] = tile_index * 512 + coefficient_index, n_tiles * 512 entries (None slots included):
// all DCs, all roots, then levels 1..8, each in scan order.  Returns an error message if the
// reference's own scan would fail its assertion (:701) for this image size.
std::string build_emission_order(const Plan &plan, std::vector<uint32_t> &order);

// O(1) answers to the questions the reference asks its per-level HashMaps (global_position_map,
// wavelet_transform.rs:434-448, and Fractal::position_map, :49) for a depth-9 plan: position p holds leaf
// k = lut[((p.x - ax) + 181 (p.y - ay)) mod 512] of the tile centred at p - off[k]; it is a level-L node
// position iff the low 9 - L bits of k are zero and that tile is retained.  Used by the emission-order builder
// (fri_order.cpp) and uploaded for the prediction kernel (fri_predict.cu).
struct LatticeIndex {
    int ax = 0, ay = 0;                     // anchor (w/2, h/2)
    int amin = 0, bmin = 0, na = 0, nb = 0; // extent of the retained tiles in lattice coordinates
    std::vector<int32_t> tile_at;           // [nb][na] plan index of the tile at (a, b), -1 if none
    uint16_t lut[kTileLeaves];              // residue -> leaf index
    Vec2 off[kTileLeaves];                  // leaf index -> offset from the tile centre

    static int mod512(int v) { return ((v % 512) + 512) % 512; }
    int tile_of(int cx, int cy) const;      // plan index of the tile centred at (cx, cy), or -1
    // global_position_map[level].get(p): the owning tile and the node's heap index, or false
    bool node_at(int level, int x, int y, int &tile, int &heap) const;
};
void build_lattice_index(const Plan &plan, LatticeIndex &out);

// get_nearby_vectors (wavelet_transform.rs:71-90), hard-wired small depths included.
void nearby_vectors(int depth, Vec2 out[6]);

// 2^depth-bit Some/None mask of the fractal centred at (cx, cy): bit i <=> coefficient i is
// Some.  out has 2^depth / 32 words.
void fractal_mask(int depth, int32_t cx, int32_t cy, int32_t width, int32_t height, uint32_t *out);

// Average / worst shared-memory conflict degree of the leaf gather for a pitch (lower is better).
double gather_conflict_degree(int pitch, int pixel_bytes, int sample_bytes, int *worst);

}  // namespace fri
