// fri_kernels.cuh — launch interface between the C ABI (fri_api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>

#include "fri_plan.h"

namespace fri {

// Truncating division by a run-time constant (Rust `i32 / i32` with a positive divisor,
// quantization.rs:19 and :37) as a signed multiply-high, an optional add, a shift and a sign
// fix-up (Granlund & Montgomery; Hacker's Delight ch. 10).  Exact for every i32 numerator and
// every divisor >= 2; divisor 1 is never routed here (layers with q == 1 are skipped).
struct Div {
    int32_t magic;    // signed magic multiplier
    int32_t addmask;  // -1 if the numerator is added after the multiply-high (magic < 0), else 0
    int32_t shift;
};

#if defined(__CUDACC__)
#define FRI_HDI __host__ __device__ __forceinline__
#else
#define FRI_HDI inline
#endif

FRI_HDI int32_t mulhi_s32(int32_t a, int32_t b)
{
#if defined(__CUDA_ARCH__)
    return __mulhi(a, b);
#else
    return (int32_t)(((int64_t)a * b) >> 32);
#endif
}

FRI_HDI int32_t trunc_div(int32_t n, Div dv)
{
    int32_t t = (int32_t)((uint32_t)mulhi_s32(n, dv.magic) + (uint32_t)(n & dv.addmask));
    t >>= dv.shift;
    return t + (int32_t)((uint32_t)n >> 31);
}

// Power-of-two divisor 2^k (k >= 1): add 2^k - 1 to negative numerators, then shift.
FRI_HDI int32_t trunc_div_pow2(int32_t n, int k)
{
    return (int32_t)((uint32_t)n + ((uint32_t)(n >> 31) >> (32 - k))) >> k;
}

// Narrow-range variant for the encoder, whose numerators are residues / low-pass values of 8- or
// 16-bit samples (|n| <= 65535): a positive magic below 2^31 always exists for q >= 3, so the
// "add" step disappears: t = mulhi(n, magic) >> shift; t += sign bit of n.
struct SmallDiv {
    int32_t magic;
    int32_t shift;
};
constexpr int32_t kSmallDivRange = 65535;

FRI_HDI int32_t trunc_div_small(int32_t n, SmallDiv dv)
{
    const int32_t t = mulhi_s32(n, dv.magic) >> dv.shift;
    return t + (int32_t)((uint32_t)n >> 31);
}

// Quantization matrix prepared on the host (quantization.rs:3-25).
struct QuantParams {
    int32_t magic[32];
    int32_t addmask[32];
    int32_t shift[32];
    int32_t q[32];
    int32_t small_magic[32];  // encoder: narrow-range magic (valid where bit l of `small` is set)
    int32_t small_shift[32];
    int32_t pow2_shift[32];   // log2(q[l]) where bit l of `pow2` is set
    uint32_t active;   // bit l set <=> q[l] != 1
    uint32_t small;    // bit l set <=> the narrow-range division is available for layer l
    uint32_t pow2;     // bit l set <=> q[l] = 2^pow2_shift[l], pow2_shift[l] >= 1
    uint32_t fix;      // bit l set <=> q[l + 1] != q[l]: the last node of tree level l (which the reference
                       // bins into layer l + 1, quantization.rs:13) needs its own divisor
    int32_t multiply;  // decode only: 1 = multiply (true dequantizer), 0 = divide (reference)
    FRI_HDI Div div(int l) const { return Div{magic[l], addmask[l], shift[l]}; }
    FRI_HDI SmallDiv sdiv(int l) const { return SmallDiv{small_magic[l], small_shift[l]}; }
};

Div make_div(int32_t q);                         // q >= 2
bool make_small_div(int32_t q, SmallDiv &out);   // false if q < 3
void make_quant_params(QuantParams &qp, const int32_t *q, int multiply);

#ifndef FRI_MAX_THREADS
#define FRI_MAX_THREADS 256
#endif
constexpr int kThreads = FRI_MAX_THREADS;  // upper bound on threads per CTA (launch bounds)
constexpr int kMaxWarps = kThreads / 32;
int cta_threads(const Geometry &g);  // threads per CTA for a plan: one warp per two base tiles of a full group
constexpr int kScratchInts = 64;    // per warp and scratch unit: the 64 level-6 low-pass values of a (base tile, channel)
// Scratch units per warp: one per channel, or — 1-channel images — one per tile of the batch of kTileBatch
// tiles whose top levels are folded together (encode_tiles / decode_tiles in fri_kernels.cu).
constexpr int kTileBatch = 4;
FRI_HDI constexpr int scratch_units(int channels) { return channels == 1 ? kTileBatch : channels; }

size_t kernel_smem_bytes(const Geometry &g);

// Device-side tables of a plan.
struct DeviceTables {
    const GroupDesc *groups = nullptr;         // plan order (top to bottom): banded launches, emission kernels
    const GroupDesc *groups_launch = nullptr;  // whole-frame launch order (nullptr = plan order)
    const uint32_t *tile_unit = nullptr;   // nullptr at depth 9
    const uint32_t *absent_unit = nullptr; // depth > 9: base tiles entirely outside the image (encoder zero fill)
    uint32_t n_absent = 0;
    const uint32_t *chunk_list = nullptr;  // [16][list_cap]
    const uint16_t *chunk_mask = nullptr;  // [16][list_cap]
    const uint32_t *edge_list = nullptr;   // [16][edge_cap]
    const uint32_t *stage_list = nullptr;  // [16][list_cap]
};

// One-time per-process kernel attribute setup (max dynamic shared memory).
cudaError_t configure_kernels();

// Enqueue the fused forward transform + quantization for n_frames frames.  [group_begin, group_end)
// restricts the launch to a band of consecutive groups of every frame (default: all groups).
//   d_coefs / half: int32 coefficient array, or (half; depth 9 and 8-bit samples only) int16.
//   d_dc: scratch for the base tiles' low-pass roots, [n_frames][n_fractals][C][2^sub_bits]
//         (depth > 9 only; finished by the coarse kernel).
cudaError_t launch_encode(const Geometry &g, const DeviceTables &t, const QuantParams &qp, const void *d_pixels,
                          uint32_t n_frames, void *d_coefs, bool half, int32_t *d_dc, cudaStream_t stream,
                          uint32_t *launches, int group_begin = 0, int group_end = -1);
// Enqueue the fused dequantization + inverse transform + scatter.
cudaError_t launch_decode(const Geometry &g, const DeviceTables &t, const QuantParams &qp, const void *d_coefs, bool half,
                          uint32_t n_frames, void *d_pixels, int32_t *d_dc, cudaStream_t stream,
                          uint32_t *launches, int group_begin = 0, int group_end = -1);

// Emission tables of a plan (device side): the `Some` coefficient slots partitioned by group, each
// group's slots in increasing emission index.
struct EmitTables {
    const uint32_t *goff = nullptr;  // [n_groups + 1] offsets into dst / loc
    const uint32_t *dst = nullptr;   // [count] emission index of the slot (position in one channel's stream)
    const uint16_t *loc = nullptr;   // [count] (tile - group's first tile) * 512 + coefficient index
};

// Gathers coefficients into emission order: out[frame][ch][dst[k]] = coefficient loc[k] of the group, for every
// group; out is int32 or (half) int16, [n_frames][C][count].  `count` is the stride between two streams in
// elements: the number of Some slots for dense streams, or a padded stride (the packed transport).
cudaError_t launch_emit(const Geometry &g, const DeviceTables &t, const EmitTables &et, uint64_t count, const int32_t *d_coefs,
                        uint32_t n_frames, void *d_out, bool half, cudaStream_t stream, uint32_t *launches);

// The inverse: coefs[frame][tile][ch][i] <- emitted streams (int32 or, half, int16), `None` slots 0.
cudaError_t launch_unemit(const Geometry &g, const DeviceTables &t, const EmitTables &et, uint64_t count, const void *d_in,
                          bool half, uint32_t n_frames, int32_t *d_coefs, cudaStream_t stream, uint32_t *launches);

// Prediction + context bucketing (fri_predict.cu).  Device image of codec::Predictor's
// neighbour tables (fri_codec.h): where a coefficient's neighbours sit, per heap index, and every tile's adjacent tiles.
struct PredictTables {
    const int32_t *adjacent = nullptr;  // [n_tiles][9] plan index of the tile one lattice step away ((db+1)*3 + da+1), -1 if none
    const uint32_t *steps = nullptr;    // [512][14] per heap index: regular[6], alt[4], probe[4] as heap | cell << 16 (0xffff: none)
    int lf_cell[3] = {4, 4, 4};         // adjacency cells of the tiles at centre + v9[4], v9[5], v9[0]
};
struct PredictParams {  // value / width predictor parameters per channel and layer set (prediction.rs:164-178)
    float value[3][3][6];
    float width[3][3][6];
};
cudaError_t configure_predict_kernel();
// For every frame and channel, in emission order (stride `count`): context bucket, prediction, zig-zag symbol of
// (value - prediction) saturated to 16 bits, and the per-context histograms [n_frames][C][10][1024] (accumulated:
// the caller zeroes them); *d_overflow counts symbols outside the 1024-symbol alphabet.
cudaError_t launch_predict(const Geometry &g, const DeviceTables &t, const EmitTables &et, const PredictTables &pt,
                           const PredictParams &prm, uint64_t count, const int32_t *d_coefs, uint32_t n_frames,
                           uint8_t *d_bucket, int32_t *d_pred, uint16_t *d_sym, uint32_t *d_hist, uint32_t *d_overflow,
                           cudaStream_t stream, uint32_t *launches);
// One pass of the predictor parameter fit over one frame: adds the integer normal-equation sums of every
// (channel, layer set) to d_sums[C][3][27] (21 upper-triangle terms + 6 right-hand sides; the caller zeroes
// them).  width_pass = false: value fit; true: width fit with prm.value already solved.
cudaError_t launch_fit(const Geometry &g, const DeviceTables &t, const EmitTables &et, const PredictTables &pt,
                       const PredictParams &prm, bool width_pass, const int32_t *d_coefs, unsigned long long *d_sums,
                       cudaStream_t stream, uint32_t *launches);

// Packed transport of emission-ordered streams (bits = 10 or 9): int16 streams (stride a multiple of 64
// elements, 16-byte aligned) <-> blocks of 64 zig-zag symbols in 8 * bits bytes; n_blocks = total elements / 64.
constexpr int kPackBlock = 64;  // symbols per packed block
constexpr int pack_block_bytes(int bits) { return 8 * bits; }
cudaError_t launch_pack_bits(int bits, const int16_t *d_src, uint8_t *d_dst, size_t n_blocks, cudaStream_t stream, uint32_t *launches);
cudaError_t launch_unpack_bits(int bits, const uint8_t *d_src, int16_t *d_dst, size_t n_blocks, cudaStream_t stream, uint32_t *launches);

// 16-bit transport of the host-buffer entry points: saturating i32 -> i16 repack and its inverse
// (count is a multiple of 8; both pointers 16-byte aligned).
cudaError_t launch_pack16(const int32_t *d_src, int16_t *d_dst, size_t count, cudaStream_t stream, uint32_t *launches);
cudaError_t launch_unpack16(const int16_t *d_src, int32_t *d_dst, size_t count, cudaStream_t stream, uint32_t *launches);

}  // namespace fri
