// fri_predict.cu — prediction + context bucketing of the quantized coefficients on the device, encode side
// (SURVEY.md §8(f) next-2): what the reference computes, coefficient by coefficient on the host, between the
// quantizer and the rANS coder.
//
// Reference (crates/libfri/src/...):
//   stages/prediction.rs:86-149    get_lf_context_bucket: DC and root residue of every tile from the same
//                                  coefficient of three lattice-neighbour tiles (MED-style predictor, width
//                                  bucket from |left - up_right|);
//   stages/prediction.rs:151-207   get_hf_context_bucket: levels 1..8 from six neighbours with a 6-tap f32
//                                  predictor and a 6-term f32 width model (parameter set by level: < 7 / 7 / 8);
//   context_modeling.rs:25-77      get_neighbour_values: left, up-left, up-right on the node's own level, and the
//                                  PARENTS (heap / 2) of right, down-left, down-right;
//   stages/wavelet_transform.rs:97-177  the six neighbour getters with their depth-2 special cases (which probe
//                                  global_position_map[2], the LEVEL-2 map, for level-7 positions — kept);
//   stages/prediction.rs:55-68     assign_bucket;  utils.rs:34-40 pack_signed;
//   stages/prediction.rs:224-323   the per-context symbol histograms.
// The 18 + 18 f32 parameters per channel are inputs (the reference fits them with an un-vendored f32 SVD).
//
// The reference answers "which tile holds a level-L node at position p, and at which heap index" with one
// HashMap per level; here LatticeIndex answers it arithmetically (fri_plan.h).  One CTA per tile group (the
// transform kernels' groups), one thread per emitted coefficient and channel; neighbour coefficients are
// gathered from the dense [tile][channel][512] blocks, which for a group's neighbours sit in nearby HBM / L2.
// f32 arithmetic uses explicit round-to-nearest multiplies and adds in the reference's order (no FMA
// contraction), float -> int conversions saturate and map NaN to 0 like Rust's `as`.
#include "fri_kernels.cuh"

#include <algorithm>

namespace fri {

namespace {

constexpr int kContexts = 10;     // CONTEXT_AMOUNT, prediction.rs:15
constexpr int kAlphabet = 1024;   // ALPHABET_SIZE, entropy_coding.rs:25

__device__ __forceinline__ int assign_bucket_u32(unsigned w)  // prediction.rs:55-68
{
    return w < 3 ? 0 : w < 5 ? 1 : w < 6 ? 2 : w < 8 ? 3 : w < 12 ? 4 : w < 16 ? 5 : w < 20 ? 6 : w < 25 ? 7 : w < 30 ? 8 : 9;
}

// Where the six neighbours of a coefficient sit does not depend on the tile: per heap index the plan holds
// (heap index, step to an adjacent tile) pairs — built on the host by running the reference's lattice queries once
// for a virtual tile (codec::Predictor, fri_codec.cpp) — and per tile the plan indices of its eight neighbours.
// A step is packed as heap | cell << 16, heap == 0xffff: no node at that position in any tile.
constexpr int kStepsPerHeap = 14;  // regular[6], alt[4], probe[4]

__device__ __forceinline__ int step_tile(const int32_t *__restrict__ adj_row, uint32_t st)
{
    return (st & 0xffffu) == 0xffffu ? -1 : __ldg(adj_row + (st >> 16));
}

// Element offsets (channel 0) of the neighbour values of coefficient `heap` of `tile`, -1 where the neighbour does not
// exist: get_lf_context_bucket's three (heap < 2) or get_neighbour_values' six (context_modeling.rs:25-77, with the
// level-7 alternatives of wavelet_transform.rs:115-177).
template <int C>
__device__ __forceinline__ void neighbour_offsets(const PredictTables &pt, int tile, int heap, int off[6])
{
    const int32_t *adj = pt.adjacent + (size_t)tile * 9;
    if (heap < 2) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int t = __ldg(adj + pt.lf_cell[j]);
            off[j] = t >= 0 ? ((t * C) << kBaseDepth) + heap : -1;
        }
        off[3] = off[4] = off[5] = -1;
        return;
    }
    const uint32_t *hs = pt.steps + heap * kStepsPerHeap;
    uint32_t st[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) st[j] = __ldg(hs + j);
    if (heap >= 128 && heap < 256) {  // level 7: probes of the level-2 map, kept as written in the reference
        const bool alt_down = step_tile(adj, __ldg(hs + 10)) < 0 && step_tile(adj, __ldg(hs + 11)) >= 0;
        const bool alt_up = step_tile(adj, __ldg(hs + 12)) < 0 && step_tile(adj, __ldg(hs + 13)) >= 0;
        if (alt_up) { st[1] = __ldg(hs + 6); st[2] = __ldg(hs + 7); }
        if (alt_down) { st[4] = __ldg(hs + 8); st[5] = __ldg(hs + 9); }
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        const int t = step_tile(adj, st[j]);
        const int h = (int)(st[j] & 0xffffu);
        off[j] = t >= 0 ? ((t * C) << kBaseDepth) + (j < 3 ? h : h >> 1) : -1;
    }
}

// One CTA per (group, frame), 1024 threads, one thread per emitted slot of the group and iteration: the neighbour
// offsets are looked up once per slot, then every channel gathers its six values, predicts, and counts its symbol in
// the channel's own 10 x 1024 histogram in shared memory (C x 40 KB); the histograms reach global memory as one atomic
// per non-zero bin.
constexpr int kPredictThreads = 1024;

template <int C>
__global__ void __launch_bounds__(kPredictThreads, 1)
fri_predict_kernel(const __grid_constant__ PredictTables pt, const __grid_constant__ PredictParams prm,
                   const GroupDesc *__restrict__ groups,
                   const uint32_t *__restrict__ goff, const uint32_t *__restrict__ dst, const uint16_t *__restrict__ loc,
                   unsigned long long count, int n_tiles, const int32_t *__restrict__ coefs,
                   uint8_t *__restrict__ bucket_out, int32_t *__restrict__ pred_out, uint16_t *__restrict__ sym_out,
                   uint32_t *__restrict__ hist_out, uint32_t *__restrict__ overflow)
{
    extern __shared__ __align__(16) uint32_t s_hist[];  // [C][10][1024]
    for (int i = threadIdx.x; i < C * kContexts * kAlphabet; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const GroupDesc gd = groups[blockIdx.x];
    const int frame = blockIdx.y;
    const uint32_t k0 = goff[blockIdx.x], k1 = goff[blockIdx.x + 1];
    const int32_t *fc = coefs + (size_t)frame * n_tiles * C * kTileLeaves;

    for (uint32_t k = k0 + threadIdx.x; k < k1; k += blockDim.x) {
        const uint32_t l = __ldg(loc + k);
        const int tile = (int)gd.tile_base + (int)(l >> kBaseDepth), heap = (int)(l & (kTileLeaves - 1));
        int off[6];
        neighbour_offsets<C>(pt, tile, heap, off);
        const int self = ((tile * C) << kBaseDepth) + heap;
        const size_t e0 = (size_t)frame * C * count + __ldg(dst + k);
        const int level = 31 - __clz(max(heap, 1));
        const int layer = level < kBaseDepth - 2 ? 2 : (level == kBaseDepth - 2 ? 1 : 0);
#pragma unroll
        for (int ch = 0; ch < C; ++ch) {
            const int32_t *cc = fc + (ch << kBaseDepth);
            int v[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) v[j] = off[j] >= 0 ? __ldg(cc + off[j]) : 0;
            int bucket, prediction;
            if (heap < 2) {
                // ---- get_lf_context_bucket: the same coefficient of the tiles at centre + v9[4], v9[5], v9[0]
                const unsigned width = (unsigned)abs(v[0] - v[2]);
                bucket = assign_bucket_u32(__float2uint_rz((float)width));
                const int hi = max(v[0], v[2]), lo = min(v[0], v[2]);
                prediction = v[1] >= hi ? hi : (v[1] <= lo ? lo : v[0] + v[2] - v[1]);
            } else {
                // ---- get_hf_context_bucket
                const float *vp = prm.value[ch][layer], *wp = prm.width[ch][layer];
                float width = wp[0];
                width = __fadd_rn(width, __fmul_rn(wp[1], (float)abs(v[0] - v[3])));
                width = __fadd_rn(width, __fmul_rn(wp[2], (float)abs(v[1] - v[2])));
                width = __fadd_rn(width, __fmul_rn(wp[3], (float)abs(v[4] - v[5])));
                width = __fadd_rn(width, __fmul_rn(wp[4], (float)abs(v[1] - v[5])));
                width = __fadd_rn(width, __fmul_rn(wp[5], (float)abs(v[2] - v[4])));
                bucket = assign_bucket_u32(__float2uint_rz(width));  // `as u32`: saturating, NaN -> 0
                float p = __fmul_rn((float)v[0], vp[0]);
#pragma unroll
                for (int j = 1; j < 6; ++j) p = __fadd_rn(p, __fmul_rn((float)v[j], vp[j]));
                prediction = __float2int_rz(p);                      // `as i32`: saturating, NaN -> 0
            }
            const int value = __ldg(cc + self);
            const int residual = (int)((unsigned)value - (unsigned)prediction);
            const unsigned sym = residual >= 0 ? 2u * (unsigned)residual : (unsigned)(-2 * residual - 1);  // pack_signed
            const size_t e = e0 + (size_t)ch * count;
            bucket_out[e] = (uint8_t)bucket;
            if (pred_out) pred_out[e] = prediction;
            sym_out[e] = (uint16_t)min(sym, 0xffffu);
            if (sym < (unsigned)kAlphabet) atomicAdd(&s_hist[(ch * kContexts + bucket) * kAlphabet + sym], 1u);
            else if (overflow) atomicAdd(overflow, 1u);  // the reference indexes freqs[sym] and panics (entropy_coding.rs:99)
        }
    }
    __syncthreads();
    uint32_t *h = hist_out + (size_t)frame * C * kContexts * kAlphabet;
    for (int i = threadIdx.x; i < C * kContexts * kAlphabet; i += blockDim.x) {
        const uint32_t c = s_hist[i];
        if (c) atomicAdd(h + i, c);
    }
}

// ------------------------------------------------------------------------------------------
// Predictor parameter fit (context_modeling.rs:144-214), accumulation side: the 6 x 6 normal equations of
// every (channel, layer set) as exact 64-bit integer sums (fri_codec.h, FitSums) — pass VALUE: regressors the
// six neighbour values, target the coefficient; pass WIDTH: regressors [1, |v0-v3|, |v1-v2|, |v4-v5|, |v1-v5|,
// |v2-v4|], target |coefficient - prediction| in 1/256 fixed point with the value parameters of pass one.  One
// CTA per group; a thread keeps the 27 sums of one layer set in registers while it walks the group's slots of
// that set, then shuffles + shared memory + one atomic per CTA and sum.  Integer sums are order-independent,
// so the parameters the host solves from them are bit-identical to the host fit's.
// ------------------------------------------------------------------------------------------
constexpr int kFitTerms = 27;  // 21 (upper triangle) + 6

template <bool WIDTH>
__global__ void __launch_bounds__(256)
fri_fit_kernel(const __grid_constant__ PredictTables pt, const __grid_constant__ PredictParams prm,
               const GroupDesc *__restrict__ groups, const uint32_t *__restrict__ goff, const uint16_t *__restrict__ loc,
               int channels, const int32_t *__restrict__ coefs, unsigned long long *__restrict__ sums)
{
    __shared__ unsigned long long s_part[8][kFitTerms];
    const GroupDesc gd = groups[blockIdx.x];
    const uint32_t k0 = goff[blockIdx.x], k1 = goff[blockIdx.x + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int ch = 0; ch < channels; ++ch) {
        const int32_t *cc = coefs + (ch << kBaseDepth);
#pragma unroll 1
        for (int set = 0; set < 3; ++set) {
            const float *vp = prm.value[ch][set];
            unsigned long long acc[kFitTerms];
#pragma unroll
            for (int i = 0; i < kFitTerms; ++i) acc[i] = 0;
            for (uint32_t k = k0 + threadIdx.x; k < k1; k += blockDim.x) {
                const uint32_t l = __ldg(loc + k);
                const int heap = (int)(l & (kTileLeaves - 1));
                if (heap < 2) continue;
                const int level = 31 - __clz(heap);
                if ((level < kBaseDepth - 2 ? 2 : (level == kBaseDepth - 2 ? 1 : 0)) != set) continue;
                const int tile = (int)gd.tile_base + (int)(l >> kBaseDepth);
                int off[6], v[6];
                if (channels == 3) neighbour_offsets<3>(pt, tile, heap, off);
                else neighbour_offsets<1>(pt, tile, heap, off);
#pragma unroll
                for (int j = 0; j < 6; ++j) v[j] = off[j] >= 0 ? __ldg(cc + off[j]) : 0;
                const int value = __ldg(cc + ((tile * channels) << kBaseDepth) + heap);
                long long w[6], y;
                if (!WIDTH) {
#pragma unroll
                    for (int j = 0; j < 6; ++j) w[j] = v[j];
                    y = value;
                } else {
                    float p = __fmul_rn((float)v[0], vp[0]);
#pragma unroll
                    for (int j = 1; j < 6; ++j) p = __fadd_rn(p, __fmul_rn((float)v[j], vp[j]));
                    const float r = __fmul_rn(fabsf(__fsub_rn((float)value, p)), 256.0f);
                    y = r < 1.0e12f ? __float2ll_rz(r) : 1ll << 40;
                    auto ad = [](int a, int b) { return llabs((long long)a - (long long)b); };
                    w[0] = 1; w[1] = ad(v[0], v[3]); w[2] = ad(v[1], v[2]); w[3] = ad(v[4], v[5]);
                    w[4] = ad(v[1], v[5]); w[5] = ad(v[2], v[4]);
                }
                int t = 0;
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    acc[21 + i] += (unsigned long long)w[i] * (unsigned long long)y;
#pragma unroll
                    for (int j = i; j < 6; ++j) acc[t++] += (unsigned long long)w[i] * (unsigned long long)w[j];
                }
            }
#pragma unroll
            for (int i = 0; i < kFitTerms; ++i) {
                unsigned long long x = acc[i];
#pragma unroll
                for (int sh = 16; sh > 0; sh >>= 1) x += __shfl_xor_sync(0xffffffffu, x, sh);
                if (lane == 0) s_part[warp][i] = x;
            }
            __syncthreads();
            if (threadIdx.x < kFitTerms) {
                unsigned long long x = 0;
                for (int wp = 0; wp < (int)(blockDim.x >> 5); ++wp) x += s_part[wp][threadIdx.x];
                if (x) atomicAdd(sums + ((size_t)ch * 3 + set) * kFitTerms + threadIdx.x, x);
            }
            __syncthreads();
        }
    }
}

}  // namespace

cudaError_t configure_predict_kernel()
{
    cudaError_t e = cudaFuncSetAttribute(fri_predict_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kContexts * kAlphabet * (int)sizeof(uint32_t));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(fri_predict_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * kContexts * kAlphabet * (int)sizeof(uint32_t));
}

cudaError_t launch_predict(const Geometry &g, const DeviceTables &t, const EmitTables &et, const PredictTables &pt,
                           const PredictParams &prm, uint64_t count,
                           const int32_t *d_coefs, uint32_t n_frames, uint8_t *d_bucket, int32_t *d_pred, uint16_t *d_sym,
                           uint32_t *d_hist, uint32_t *d_overflow, cudaStream_t stream, uint32_t *launches)
{
    if (count == 0 || n_frames == 0 || g.n_groups == 0) return cudaSuccess;
    if (g.channels != 1 && g.channels != 3) return cudaErrorInvalidValue;
    const size_t smem = (size_t)g.channels * kContexts * kAlphabet * sizeof(uint32_t);
    for (uint32_t f0 = 0; f0 < n_frames; f0 += 65535u) {  // gridDim.y limit
        const uint32_t nf = n_frames - f0 < 65535u ? n_frames - f0 : 65535u;
        const dim3 grid((unsigned)g.n_groups, nf);
        const size_t so = (size_t)f0 * g.channels * count;
        const int32_t *fc = d_coefs + (int64_t)f0 * g.coefs_per_frame;
        uint32_t *fh = d_hist + (size_t)f0 * g.channels * kContexts * kAlphabet;
        if (g.channels == 1)
            fri_predict_kernel<1><<<grid, kPredictThreads, smem, stream>>>(pt, prm, t.groups, et.goff, et.dst, et.loc, count, g.n_fractals, fc,
                                                                         d_bucket + so, d_pred ? d_pred + so : nullptr, d_sym + so, fh, d_overflow);
        else
            fri_predict_kernel<3><<<grid, kPredictThreads, smem, stream>>>(pt, prm, t.groups, et.goff, et.dst, et.loc, count, g.n_fractals, fc,
                                                                         d_bucket + so, d_pred ? d_pred + so : nullptr, d_sym + so, fh, d_overflow);
        if (launches) ++*launches;
    }
    return cudaGetLastError();
}

cudaError_t launch_fit(const Geometry &g, const DeviceTables &t, const EmitTables &et, const PredictTables &pt,
                       const PredictParams &prm, bool width_pass, const int32_t *d_coefs, unsigned long long *d_sums,
                       cudaStream_t stream, uint32_t *launches)
{
    if (g.n_groups == 0) return cudaSuccess;
    if (width_pass)
        fri_fit_kernel<true><<<(unsigned)g.n_groups, 256, 0, stream>>>(pt, prm, t.groups, et.goff, et.loc, g.channels, d_coefs, d_sums);
    else
        fri_fit_kernel<false><<<(unsigned)g.n_groups, 256, 0, stream>>>(pt, prm, t.groups, et.goff, et.loc, g.channels, d_coefs, d_sums);
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace fri

