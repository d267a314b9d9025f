// fri_kernels.cu — sm_100a kernels for frave's fractal transform + quantization hot path.
//
// What the kernels compute (reference: crates/libfri/src/...):
//   encode  = stages/wavelet_transform.rs:179-225 (extract_coefficients: gather through the
//             index map of :47-53, then the lifting d = l - r, s = r + d/2 per tree node)
//             fused with stages/quantization.rs:7-25 (coef /= q[floor(log2(i+1))]).
//   decode  = quantization.rs:27-45 fused with wavelet_transform.rs:358-381 (extract_values:
//             r = s - d/2, l = d + r top-down) and images.rs:103-111 (bounds-checked,
//             clamped scatter).
//
// Mapping to the machine:
//   * one CTA per *group* of up to 32 lattice-adjacent base tiles (512 pixels each).  The
//     group's pixel footprint is staged through shared memory with 16-byte cp.async chunks
//     that are aligned in global memory, so HBM only ever sees full-sector, coalesced
//     traffic although a tile's leaf order is a twindragon curve;
//   * one warp per (base tile, channel).  A lane owns two complete depth-3 subtrees (see
//     fri_geometry.h), so levels 8..6 are register-only, levels 5..0 are five warp shuffles,
//     and every heap-ordered coefficient run a warp touches is contiguous: 2 x 128-bit,
//     2 x 64-bit, 2 x 32-bit and one 64-bit access per lane, all full sectors;
//   * out-of-image leaves are staged as zeros.  With l/r := 0 for a missing side the
//     arithmetic of try_apply (wavelet_transform.rs:14-26) is reproduced exactly and a
//     coefficient the reference holds as None comes out as 0; which coefficients are Some is
//     purely geometric and is reported by the plan (fri_plan_masks).
//   * depth > 9 (extension): a fractal is 2^(depth-9) base tiles; the base kernel writes each
//     base tile's levels into the fractal's heap and its low-pass root into a scratch array
//     that the small coarse kernel folds through the remaining depth-9 levels.
#include "fri_kernels.cuh"

#include <cstdint>

namespace fri {

// ------------------------------------------------------------------------------------------
// host helpers
// ------------------------------------------------------------------------------------------
Div make_div(int32_t q)
{
    const uint32_t d = q < 1 ? 1u : (uint32_t)q;
    uint32_t fl = 0;
    while ((d >> (fl + 1)) != 0) ++fl;  // floor(log2 d)
    if ((d & (d - 1)) == 0) return Div{0u, fl};
    const uint64_t num = (uint64_t)1 << (32 + fl);
    uint32_t m = (uint32_t)(num / d);
    const uint32_t rem = (uint32_t)(num - (uint64_t)m * d);
    const uint32_t e = d - rem;
    uint32_t more;
    if (e < (1u << fl)) {
        more = fl;  // a 32-bit magic is exact
    } else {         // 33-bit magic: keep the low 32 bits and fix up with the add step
        m += m;
        const uint32_t twice = rem + rem;
        if (twice >= d || twice < rem) m += 1;
        more = fl | kDivAdd;
    }
    return Div{m + 1u, more};
}

void make_quant_params(QuantParams &qp, const int32_t *q, int multiply)
{
    qp.active = 0;
    qp.multiply = multiply;
    for (int l = 0; l < 32; ++l) {
        const int32_t v = q ? q[l] : 1;
        const Div dv = make_div(v);
        qp.q[l] = v;
        qp.magic[l] = dv.magic;
        qp.more[l] = (uint8_t)dv.more;
        if (v != 1) qp.active |= 1u << l;
    }
}

size_t kernel_smem_bytes(const Geometry &g)
{
    return (((size_t)g.region_h * g.pitch + 15) & ~(size_t)15) + (size_t)kWarps * kScratchInts * sizeof(int32_t);
}

namespace {

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gmem_src)
{
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// wrapping i32 arithmetic (release-mode Rust semantics)
__device__ __forceinline__ int wsub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }
__device__ __forceinline__ int wadd(int a, int b) { return (int)((unsigned)a + (unsigned)b); }

// forward lifting of one node: d = l - r; s = r + d/2 (truncating)   wavelet_transform.rs:211-218
__device__ __forceinline__ void lift(int l, int r, int &d, int &s)
{
    d = wsub(l, r);
    s = wadd(r, d / 2);
}
// inverse lifting of one node: r = s - d/2; l = d + r                 wavelet_transform.rs:366-367
__device__ __forceinline__ void unlift(int s, int d, int &l, int &r)
{
    r = wsub(s, d / 2);
    l = wadd(d, r);
}

// quantization::encode of one coefficient                            quantization.rs:19
__device__ __forceinline__ int quant1(int d, Div dv) { return trunc_div(d, dv); }
// quantization::decode of one coefficient                            quantization.rs:37
__device__ __forceinline__ int dequant1(int d, Div dv, int q, int multiply)
{
    return multiply ? (int)((unsigned)d * (unsigned)q) : trunc_div(d, dv);
}

// floor(log2(pos + 1)): the reference's layer index of heap position pos (quantization.rs:13)
__device__ __forceinline__ int layer_of(uint32_t pos) { return 31 - __clz((int)(pos + 1u)); }

__device__ __forceinline__ int lane_anchor_bytes(int lane, int pitch, int pixel_bytes)
{
    // lane_anchor(lane) of fri_geometry.h: digit vectors 3..7 selected by the lane's bits
    constexpr Vec2 d3 = kLiterals[3], d4 = kLiterals[4], d5 = kLiterals[5], d6 = kLiterals[6], d7 = kLiterals[7];
    int x = 0, y = 0;
    if (lane & 1) { x += d3.x; y += d3.y; }
    if (lane & 2) { x += d4.x; y += d4.y; }
    if (lane & 4) { x += d5.x; y += d5.y; }
    if (lane & 8) { x += d6.x; y += d6.y; }
    if (lane & 16) { x += d7.x; y += d7.y; }
    return y * pitch + x * pixel_bytes;
}

// Per-CTA view of the staged region.
struct RegionView {
    int64_t a0;   // global address of region pixel (0, 0) of this frame (may lie outside the frame)
    int phi0;     // a0 & 15: the region keeps the global 16-byte phase in shared memory
};

template <int PB>
__device__ __forceinline__ RegionView region_view(const Geometry &g, const GroupDesc &gd, const void *frame_base)
{
    RegionView v;
    v.a0 = (int64_t)(uintptr_t)frame_base + (int64_t)gd.y0 * g.row_stride + (int64_t)gd.x0 * PB;
    v.phi0 = (int)(v.a0 & 15);
    return v;
}

// Output/input addressing of one (base tile, channel) task.
struct TaskAddr {
    int64_t block;   // element offset of the fractal-channel coefficient block
    uint32_t node;   // heap index of the base tile's root inside the fractal (1 at depth 9)
    int64_t dc;      // element offset into the low-pass scratch (depth > 9)
    bool last;       // the base tile is the last one of its fractal
};

template <int C>
__device__ __forceinline__ TaskAddr task_addr(const Geometry &g, const uint32_t *tile_unit, int frame, int tile, int ch)
{
    TaskAddr a;
    if (g.sub_bits == 0) {
        a.block = (((int64_t)frame * g.n_fractals + tile) * C + ch) << kBaseDepth;
        a.node = 1;
        a.dc = 0;
        a.last = true;
    } else {
        const uint32_t u = __ldg(tile_unit + tile);
        const uint32_t f = u >> g.sub_bits, sub = u & ((1u << g.sub_bits) - 1u);
        const int64_t fc = ((int64_t)frame * g.n_fractals + f) * C + ch;
        a.block = fc << g.depth;
        a.node = (1u << g.sub_bits) + sub;
        a.dc = (fc << g.sub_bits) + sub;
        a.last = sub == ((1u << g.sub_bits) - 1u);
    }
    return a;
}

// ------------------------------------------------------------------------------------------
// encode: pixels -> quantized coefficients
// ------------------------------------------------------------------------------------------
template <int C, typename S>
__global__ void __launch_bounds__(kThreads, 4)
fri_encode_kernel(const __grid_constant__ Geometry g, const __grid_constant__ QuantParams qp,
                  const GroupDesc *__restrict__ groups, const uint32_t *__restrict__ tile_unit,
                  const uint8_t *__restrict__ pixels, int32_t *__restrict__ coefs, int32_t *__restrict__ dc_out)
{
    constexpr int SB = (int)sizeof(S);
    constexpr int PB = C * SB;
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t *region = smem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t *scratch = reinterpret_cast<int32_t *>(smem + (((size_t)g.region_h * g.pitch + 15) & ~(size_t)15)) + warp * kScratchInts;

    const GroupDesc gd = groups[blockIdx.x];
    const int frame = blockIdx.y;
    const uint8_t *fbase = pixels + (int64_t)frame * g.frame_bytes;
    const RegionView rv = region_view<PB>(g, gd, fbase);

    // ---- stage the group's pixel footprint: one warp per row, one lane per 16-byte chunk.
    for (int r = warp; r < g.region_h; r += kWarps) {
        const int y = gd.y0 + r;
        const int srow = r * g.pitch + rv.phi0;  // shared-memory offset of region pixel (r, 0)
        const int sbase = srow & ~15;
        const int send = srow + g.row_bytes;
        const bool yin = (unsigned)y < (unsigned)g.height;
        const int64_t row_lo = (int64_t)(uintptr_t)fbase + (int64_t)y * g.row_stride;  // in-image bytes of row y
        const int64_t row_hi = row_lo + g.row_stride;
        const int64_t gbase = rv.a0 + (int64_t)r * g.row_stride - (srow & 15);  // floor16(global row start)
        for (int c = lane; c < g.chunks_per_row; c += 32) {
            const int s = sbase + 16 * c;
            if (s >= send) break;
            const int64_t ga = gbase + 16 * c;
            if (yin && ga >= row_lo && ga + 16 <= row_hi) {
                cp_async_16(region + s, reinterpret_cast<const void *>(ga));
            } else if (!yin || ga + 16 <= row_lo || ga >= row_hi) {
                *reinterpret_cast<int4 *>(region + s) = make_int4(0, 0, 0, 0);
            } else {  // chunk straddles the left or right image edge
                const uint8_t *gp = reinterpret_cast<const uint8_t *>(ga);
#pragma unroll 1
                for (int j = 0; j < 16; ++j)
                    region[s + j] = (ga + j >= row_lo && ga + j < row_hi) ? __ldg(gp + j) : (uint8_t)0;
            }
        }
    }
    cp_async_wait_all();
    __syncthreads();

    // ---- one (base tile, channel) task per warp iteration
    const int n_tasks = __popc(gd.tile_mask) * C;
    const int lane_off = lane_anchor_bytes(lane, g.pitch, PB);
    const bool any_q = qp.active != 0;
    for (int task = warp; task < n_tasks; task += kWarps) {
        const int e = task / C, ch = task - e * C;
        const int slot = (gd.tile_mask & (gd.tile_mask + 1u)) ? (int)__fns(gd.tile_mask, 0, e + 1) : e;
        const uint8_t *p0 = region + rv.phi0 + g.tile_rel_y[slot] * g.pitch + g.tile_rel_x[slot] * PB + lane_off + ch * SB;
        const uint8_t *p1 = p0 + g.pitch, *p2 = p1 + g.pitch;
        const int half = kHalfB.y * g.pitch + kHalfB.x * PB;

        // gather: leaf i of a depth-3 subtree sits at sub_leaf(i) from the subtree's first leaf
        int v[8], w[8];
#define FRI_LD(ptr, dx) ((int)*reinterpret_cast<const S *>((ptr) + (dx) * PB))
        v[0] = FRI_LD(p0, 0);  v[1] = FRI_LD(p1, 0);   // (0,0) (0,1)
        v[2] = FRI_LD(p1, -1); v[3] = FRI_LD(p2, -1);  // (-1,1) (-1,2)
        v[4] = FRI_LD(p0, 2);  v[5] = FRI_LD(p1, 2);   // (2,0) (2,1)
        v[6] = FRI_LD(p1, 1);  v[7] = FRI_LD(p2, 1);   // (1,1) (1,2)
        w[0] = FRI_LD(p0 + half, 0);  w[1] = FRI_LD(p1 + half, 0);
        w[2] = FRI_LD(p1 + half, -1); w[3] = FRI_LD(p2 + half, -1);
        w[4] = FRI_LD(p0 + half, 2);  w[5] = FRI_LD(p1 + half, 2);
        w[6] = FRI_LD(p1 + half, 1);  w[7] = FRI_LD(p2 + half, 1);
#undef FRI_LD

        const TaskAddr ta = task_addr<C>(g, tile_unit, frame, gd.tile_base + e, ch);
        int32_t *out = coefs + ta.block;
        const int top = g.sub_bits;              // global level of the base tile's root
        const bool lastB = ta.last && lane == 31;  // this lane holds the last node of every level

        // levels 8, 7, 6 in registers
        int a8[4], b8[4], sa8[4], sb8[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            lift(v[2 * m], v[2 * m + 1], a8[m], sa8[m]);
            lift(w[2 * m], w[2 * m + 1], b8[m], sb8[m]);
        }
        int a7[2], b7[2], sa7[2], sb7[2];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            lift(sa8[2 * m], sa8[2 * m + 1], a7[m], sa7[m]);
            lift(sb8[2 * m], sb8[2 * m + 1], b7[m], sb7[m]);
        }
        int a6, b6, sA, sB;
        lift(sa7[0], sa7[1], a6, sA);
        lift(sb7[0], sb7[1], b6, sB);

        if (any_q) {  // quantization.rs:13 — layer = level, except the last node of a level: level + 1
            const Div m8 = qp.div(top + 8), m7 = qp.div(top + 7), m6 = qp.div(top + 6);
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                a8[m] = quant1(a8[m], m8);
                b8[m] = quant1(b8[m], (m == 3 && lastB) ? qp.div(top + 9) : m8);
            }
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                a7[m] = quant1(a7[m], m7);
                b7[m] = quant1(b7[m], (m == 1 && lastB) ? m8 : m7);
            }
            a6 = quant1(a6, m6);
            b6 = quant1(b6, lastB ? m7 : m6);
        }
        {
            int32_t *o8 = out + ((size_t)ta.node << 8), *o7 = out + ((size_t)ta.node << 7), *o6 = out + ((size_t)ta.node << 6);
            __stcs(reinterpret_cast<int4 *>(o8) + lane, make_int4(a8[0], a8[1], a8[2], a8[3]));
            __stcs(reinterpret_cast<int4 *>(o8 + 128) + lane, make_int4(b8[0], b8[1], b8[2], b8[3]));
            __stcs(reinterpret_cast<int2 *>(o7) + lane, make_int2(a7[0], a7[1]));
            __stcs(reinterpret_cast<int2 *>(o7 + 64) + lane, make_int2(b7[0], b7[1]));
            __stcs(o6 + lane, a6);
            __stcs(o6 + 32 + lane, b6);
        }

        // levels 5..1: lane pairs, one shuffle per chain per level
#pragma unroll
        for (int step = 0; step < 5; ++step) {
            const int m = 5 - step, bit = 1 << step;
            const int oA = __shfl_xor_sync(0xffffffffu, sA, bit), oB = __shfl_xor_sync(0xffffffffu, sB, bit);
            int dA, dB;
            lift(sA, oA, dA, sA);
            lift(sB, oB, dB, sB);
            if ((lane & (2 * bit - 1)) == 0) {
                const int j = lane >> (step + 1);
                scratch[(1 << m) + j] = dA;
                scratch[(1 << m) + (1 << (m - 1)) + j] = dB;
            }
        }
        if (lane == 0) {  // level 0 and the low-pass root (wavelet_transform.rs:221)
            int d0, s0;
            lift(sA, sB, d0, s0);
            scratch[1] = d0;
            scratch[0] = s0;
        }
        __syncwarp();
        int2 t2 = *reinterpret_cast<const int2 *>(scratch + 2 * lane);
        __syncwarp();
        if (g.sub_bits == 0) {
            if (any_q) {
                t2.x = quant1(t2.x, qp.div(layer_of(2 * lane)));
                t2.y = quant1(t2.y, qp.div(layer_of(2 * lane + 1)));
            }
            __stcs(reinterpret_cast<int2 *>(out) + lane, t2);
        } else {
            // local heap position p of the base tile -> fractal heap position (node << m) + p - 2^m
            const int vals[2] = {t2.x, t2.y};
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const uint32_t p = 2 * lane + k;
                if (p == 0) {
                    dc_out[ta.dc] = vals[k];
                } else {
                    const int m = 31 - __clz((int)p);
                    const uint32_t pos = (ta.node << m) + p - (1u << m);
                    out[pos] = any_q ? quant1(vals[k], qp.div(layer_of(pos))) : vals[k];
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// decode: quantized coefficients -> pixels
// ------------------------------------------------------------------------------------------
template <int C, typename S>
__global__ void __launch_bounds__(kThreads, 4)
fri_decode_kernel(const __grid_constant__ Geometry g, const __grid_constant__ QuantParams qp,
                  const GroupDesc *__restrict__ groups, const uint32_t *__restrict__ tile_unit,
                  const uint32_t *__restrict__ ownership, const int32_t *__restrict__ coefs,
                  const int32_t *__restrict__ dc_in, uint8_t *__restrict__ pixels)
{
    constexpr int SB = (int)sizeof(S);
    constexpr int PB = C * SB;
    constexpr int kMaxVal = SB == 1 ? 255 : 65535;
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t *region = smem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t *scratch = reinterpret_cast<int32_t *>(smem + (((size_t)g.region_h * g.pitch + 15) & ~(size_t)15)) + warp * kScratchInts;

    const GroupDesc gd = groups[blockIdx.x];
    const int frame = blockIdx.y;
    uint8_t *fbase = pixels + (int64_t)frame * g.frame_bytes;
    const RegionView rv = region_view<PB>(g, gd, fbase);

    const int n_tasks = __popc(gd.tile_mask) * C;
    const int lane_off = lane_anchor_bytes(lane, g.pitch, PB);
    const bool any_q = qp.active != 0;
    // A lattice tile the reference's BFS never built (possible only next to the image border,
    // e.g. 480x270) still owns its pixels in the ownership bitmap: stage zeros for it, which is
    // what from_wavelet's zero-initialised raster holds there (wavelet_transform.rs:309-317).
    if (__popc(gd.tile_mask) != g.group_a * g.group_b) {
        const int n16 = (g.region_h * g.pitch + 15) >> 4;
        for (int i = threadIdx.x; i < n16; i += kThreads) reinterpret_cast<int4 *>(region)[i] = make_int4(0, 0, 0, 0);
        __syncthreads();
    }
    for (int task = warp; task < n_tasks; task += kWarps) {
        const int e = task / C, ch = task - e * C;
        const int slot = (gd.tile_mask & (gd.tile_mask + 1u)) ? (int)__fns(gd.tile_mask, 0, e + 1) : e;
        const TaskAddr ta = task_addr<C>(g, tile_unit, frame, gd.tile_base + e, ch);
        const int32_t *in = coefs + ta.block;
        const int top = g.sub_bits;
        const bool lastB = ta.last && lane == 31;

        // coefficient loads: all issued before the first use
        const int32_t *i8 = in + ((size_t)ta.node << 8), *i7 = in + ((size_t)ta.node << 7), *i6 = in + ((size_t)ta.node << 6);
        int4 a8 = __ldcs(reinterpret_cast<const int4 *>(i8) + lane);
        int4 b8 = __ldcs(reinterpret_cast<const int4 *>(i8 + 128) + lane);
        int2 a7 = __ldcs(reinterpret_cast<const int2 *>(i7) + lane);
        int2 b7 = __ldcs(reinterpret_cast<const int2 *>(i7 + 64) + lane);
        int a6 = __ldcs(i6 + lane);
        int b6 = __ldcs(i6 + 32 + lane);
        int2 t2;
        if (g.sub_bits == 0) {
            t2 = __ldcs(reinterpret_cast<const int2 *>(in) + lane);
            if (any_q) {
                t2.x = dequant1(t2.x, qp.div(layer_of(2 * lane)), qp.q[layer_of(2 * lane)], qp.multiply);
                t2.y = dequant1(t2.y, qp.div(layer_of(2 * lane + 1)), qp.q[layer_of(2 * lane + 1)], qp.multiply);
            }
        } else {
            int vals[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const uint32_t p = 2 * lane + k;
                if (p == 0) {
                    vals[k] = dc_in[ta.dc];  // produced (already dequantized) by the coarse kernel
                } else {
                    const int m = 31 - __clz((int)p);
                    const uint32_t pos = (ta.node << m) + p - (1u << m);
                    vals[k] = __ldcs(in + pos);
                    if (any_q) vals[k] = dequant1(vals[k], qp.div(layer_of(pos)), qp.q[layer_of(pos)], qp.multiply);
                }
            }
            t2 = make_int2(vals[0], vals[1]);
        }
        __syncwarp();
        *reinterpret_cast<int2 *>(scratch + 2 * lane) = t2;
        __syncwarp();

        if (any_q) {
            const int mul = qp.multiply;
            const int l8 = top + 8, l7 = top + 7, l6 = top + 6;
            const Div d8 = qp.div(l8), d7 = qp.div(l7), d6 = qp.div(l6);
            const int q8 = qp.q[l8], q7 = qp.q[l7], q6 = qp.q[l6];
            a8.x = dequant1(a8.x, d8, q8, mul); a8.y = dequant1(a8.y, d8, q8, mul);
            a8.z = dequant1(a8.z, d8, q8, mul); a8.w = dequant1(a8.w, d8, q8, mul);
            b8.x = dequant1(b8.x, d8, q8, mul); b8.y = dequant1(b8.y, d8, q8, mul);
            b8.z = dequant1(b8.z, d8, q8, mul);
            b8.w = lastB ? dequant1(b8.w, qp.div(l8 + 1), qp.q[l8 + 1], mul) : dequant1(b8.w, d8, q8, mul);
            a7.x = dequant1(a7.x, d7, q7, mul); a7.y = dequant1(a7.y, d7, q7, mul);
            b7.x = dequant1(b7.x, d7, q7, mul);
            b7.y = lastB ? dequant1(b7.y, d8, q8, mul) : dequant1(b7.y, d7, q7, mul);
            a6 = dequant1(a6, d6, q6, mul);
            b6 = lastB ? dequant1(b6, d7, q7, mul) : dequant1(b6, d6, q6, mul);
        }

        // levels 0..5: every lane walks its own root-to-subtree path (broadcast reads)
        int sA, sB;
        unlift(scratch[0], scratch[1], sA, sB);  // children 2 (half A) and 3 (half B)
#pragma unroll
        for (int m = 1; m <= 5; ++m) {
            const int j = lane >> (6 - m);
            int l, r;
            unlift(sA, scratch[(1 << m) + j], l, r);
            sA = ((lane >> (5 - m)) & 1) ? r : l;
            unlift(sB, scratch[(1 << m) + (1 << (m - 1)) + j], l, r);
            sB = ((lane >> (5 - m)) & 1) ? r : l;
        }
        // levels 6, 7, 8 in registers
        int sa7[2], sb7[2], sa8[4], sb8[4], v[8], w[8];
        unlift(sA, a6, sa7[0], sa7[1]);
        unlift(sB, b6, sb7[0], sb7[1]);
        unlift(sa7[0], a7.x, sa8[0], sa8[1]);
        unlift(sa7[1], a7.y, sa8[2], sa8[3]);
        unlift(sb7[0], b7.x, sb8[0], sb8[1]);
        unlift(sb7[1], b7.y, sb8[2], sb8[3]);
        unlift(sa8[0], a8.x, v[0], v[1]);
        unlift(sa8[1], a8.y, v[2], v[3]);
        unlift(sa8[2], a8.z, v[4], v[5]);
        unlift(sa8[3], a8.w, v[6], v[7]);
        unlift(sb8[0], b8.x, w[0], w[1]);
        unlift(sb8[1], b8.y, w[2], w[3]);
        unlift(sb8[2], b8.z, w[4], w[5]);
        unlift(sb8[3], b8.w, w[6], w[7]);

        // scatter into the staged region (clamp: images.rs:109)
        uint8_t *p0 = region + rv.phi0 + g.tile_rel_y[slot] * g.pitch + g.tile_rel_x[slot] * PB + lane_off + ch * SB;
        uint8_t *p1 = p0 + g.pitch, *p2 = p1 + g.pitch;
        const int half = kHalfB.y * g.pitch + kHalfB.x * PB;
#define FRI_ST(ptr, dx, val) (*reinterpret_cast<S *>((ptr) + (dx) * PB) = (S)min(max((val), 0), kMaxVal))
        FRI_ST(p0, 0, v[0]);  FRI_ST(p1, 0, v[1]);
        FRI_ST(p1, -1, v[2]); FRI_ST(p2, -1, v[3]);
        FRI_ST(p0, 2, v[4]);  FRI_ST(p1, 2, v[5]);
        FRI_ST(p1, 1, v[6]);  FRI_ST(p2, 1, v[7]);
        FRI_ST(p0 + half, 0, w[0]);  FRI_ST(p1 + half, 0, w[1]);
        FRI_ST(p1 + half, -1, w[2]); FRI_ST(p2 + half, -1, w[3]);
        FRI_ST(p0 + half, 2, w[4]);  FRI_ST(p1 + half, 2, w[5]);
        FRI_ST(p1 + half, 1, w[6]);  FRI_ST(p2 + half, 1, w[7]);
#undef FRI_ST
    }
    __syncthreads();

    // ---- write-out: one warp per staged row, one lane per 16-byte chunk aligned in global
    // memory.  Only bytes of pixels that belong to this group's tiles (ownership bitmap) and lie
    // inside the image (set_pixel's bounds check, images.rs:104) are written.
    const int vx0 = max(0, -gd.x0), vx1 = min(g.region_w, g.width - gd.x0);  // in-image columns of the region
    for (int r = warp; r < g.region_h; r += kWarps) {
        const int y = gd.y0 + r;
        if ((unsigned)y >= (unsigned)g.height) continue;
        const int srow = r * g.pitch + rv.phi0;
        const int sbase = srow & ~15;
        const int64_t gbase = rv.a0 + (int64_t)r * g.row_stride - (srow & 15);
        const uint32_t *own = ownership + (size_t)r * g.own_words;
        for (int c = lane; c < g.chunks_per_row; c += 32) {
            const int s = sbase + 16 * c;
            const int b0 = s - srow;  // region byte offset of the chunk's first byte (can be < 0)
            if (b0 >= g.row_bytes) break;
            const int lo = max(0, -b0), hi = min(16, g.row_bytes - b0);
            const int px0 = (b0 + lo) / PB, px1 = (b0 + hi - 1) / PB;
            const int q0 = max(px0, vx0), q1 = min(px1, vx1 - 1);
            if (q0 > q1) continue;
            const int n = q1 - q0 + 1;  // <= 16
            const uint32_t w0 = __ldg(own + (q0 >> 5));
            const uint32_t w1 = __ldg(own + min((q0 >> 5) + 1, g.own_words - 1));
            const uint32_t bits = __funnelshift_r(w0, w1, q0 & 31) & ((1u << n) - 1u);
            if (bits == 0) continue;
            uint8_t *gp = reinterpret_cast<uint8_t *>(gbase + 16 * c);
            const uint8_t *sp = region + s;
            if (bits == ((1u << n) - 1u) && q0 == px0 && q1 == px1 && lo == 0 && hi == 16) {
                *reinterpret_cast<int4 *>(gp) = *reinterpret_cast<const int4 *>(sp);
                continue;
            }
            uint32_t bm = 0;  // byte mask of the chunk
            if (PB == 1) {
                bm = bits << (q0 - b0);
            } else {
#pragma unroll 1
                for (int p = q0; p <= q1; ++p)
                    if ((bits >> (p - q0)) & 1u) {
                        const int bs = p * PB - b0;
                        const uint32_t pm = (1u << PB) - 1u;
                        bm |= bs >= 0 ? pm << bs : pm >> (-bs);
                    }
                bm &= 0xffffu;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t nib = (bm >> (4 * k)) & 15u;
                if (nib == 15u) {
                    *reinterpret_cast<uint32_t *>(gp + 4 * k) = *reinterpret_cast<const uint32_t *>(sp + 4 * k);
                } else if (nib) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if ((nib >> j) & 1u) gp[4 * k + j] = sp[4 * k + j];
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// coarse levels (depth > 9): the top depth-9 levels over the base tiles' low-pass roots
// ------------------------------------------------------------------------------------------
constexpr int kCoarseThreads = 256;

// One CTA per (frame, fractal, channel).  dc: [..][2^sub_bits] low-pass roots in base-tile order.
__global__ void __launch_bounds__(kCoarseThreads)
fri_coarse_forward_kernel(const __grid_constant__ QuantParams qp, int sub_bits, int depth,
                          const int32_t *__restrict__ dc, int32_t *__restrict__ coefs)
{
    extern __shared__ __align__(16) int32_t cs[];
    const int n = 1 << sub_bits;
    int32_t *src = cs, *dst = cs + n;
    const int32_t *in = dc + ((int64_t)blockIdx.x << sub_bits);
    int32_t *out = coefs + ((int64_t)blockIdx.x << depth);
    for (int i = threadIdx.x; i < n; i += kCoarseThreads) src[i] = in[i];
    __syncthreads();
    for (int level = sub_bits - 1; level >= 0; --level) {
        const int cnt = 1 << level;
        for (int j = threadIdx.x; j < cnt; j += kCoarseThreads) {
            int d, s;
            lift(src[2 * j], src[2 * j + 1], d, s);
            dst[j] = s;
            const uint32_t pos = (uint32_t)(cnt + j);
            out[pos] = quant1(d, qp.div(layer_of(pos)));
        }
        __syncthreads();
        int32_t *t = src; src = dst; dst = t;
    }
    if (threadIdx.x == 0) out[0] = quant1(src[0], qp.div(0));  // wavelet_transform.rs:221, layer 0
}

__global__ void __launch_bounds__(kCoarseThreads)
fri_coarse_inverse_kernel(const __grid_constant__ QuantParams qp, int sub_bits, int depth,
                          const int32_t *__restrict__ coefs, int32_t *__restrict__ dc)
{
    extern __shared__ __align__(16) int32_t cs[];
    const int n = 1 << sub_bits;
    int32_t *src = cs + n, *dst = cs;  // sizes: level L reads 2^L values, writes 2^(L+1)
    const int32_t *in = coefs + ((int64_t)blockIdx.x << depth);
    int32_t *out = dc + ((int64_t)blockIdx.x << sub_bits);
    if (threadIdx.x == 0) src[0] = dequant1(in[0], qp.div(0), qp.q[0], qp.multiply);
    __syncthreads();
    for (int level = 0; level < sub_bits; ++level) {
        const int cnt = 1 << level;
        for (int j = threadIdx.x; j < cnt; j += kCoarseThreads) {
            const uint32_t pos = (uint32_t)(cnt + j);
            const int d = dequant1(in[pos], qp.div(layer_of(pos)), qp.q[layer_of(pos)], qp.multiply);
            int l, r;
            unlift(src[j], d, l, r);
            dst[2 * j] = l;
            dst[2 * j + 1] = r;
        }
        __syncthreads();
        int32_t *t = src; src = dst; dst = t;
    }
    for (int i = threadIdx.x; i < n; i += kCoarseThreads) out[i] = src[i];
}

template <typename K>
cudaError_t set_smem(K kernel, size_t bytes)
{
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

constexpr size_t kMaxSmem = 227 * 1024;

}  // namespace

cudaError_t configure_kernels()
{
    cudaError_t e;
#define FRI_CFG(k) if ((e = set_smem(k, kMaxSmem)) != cudaSuccess) return e
    FRI_CFG((fri_encode_kernel<1, uint8_t>));
    FRI_CFG((fri_encode_kernel<3, uint8_t>));
    FRI_CFG((fri_encode_kernel<1, uint16_t>));
    FRI_CFG((fri_encode_kernel<3, uint16_t>));
    FRI_CFG((fri_decode_kernel<1, uint8_t>));
    FRI_CFG((fri_decode_kernel<3, uint8_t>));
    FRI_CFG((fri_decode_kernel<1, uint16_t>));
    FRI_CFG((fri_decode_kernel<3, uint16_t>));
    FRI_CFG(fri_coarse_forward_kernel);
    FRI_CFG(fri_coarse_inverse_kernel);
#undef FRI_CFG
    return cudaSuccess;
}

cudaError_t launch_encode(const Geometry &g, const DeviceTables &t, const QuantParams &qp, const void *d_pixels,
                          uint32_t n_frames, int32_t *d_coefs, int32_t *d_dc, cudaStream_t stream,
                          uint32_t *launches)
{
    if (n_frames == 0 || g.n_groups == 0) return cudaSuccess;
    const size_t smem = kernel_smem_bytes(g);
    const uint8_t *px = static_cast<const uint8_t *>(d_pixels);
    for (uint32_t f0 = 0; f0 < n_frames; f0 += 65535u) {  // gridDim.y limit
        const uint32_t nf = n_frames - f0 < 65535u ? n_frames - f0 : 65535u;
        const dim3 grid((unsigned)g.n_groups, nf);
        const uint8_t *p = px + (int64_t)f0 * g.frame_bytes;
        int32_t *c = d_coefs + (int64_t)f0 * g.coefs_per_frame;
        int32_t *dc = d_dc ? d_dc + (((int64_t)f0 * g.n_fractals * g.channels) << g.sub_bits) : nullptr;
        if (g.channels == 1 && g.sample_bytes == 1)
            fri_encode_kernel<1, uint8_t><<<grid, kThreads, smem, stream>>>(g, qp, t.groups, t.tile_unit, p, c, dc);
        else if (g.channels == 3 && g.sample_bytes == 1)
            fri_encode_kernel<3, uint8_t><<<grid, kThreads, smem, stream>>>(g, qp, t.groups, t.tile_unit, p, c, dc);
        else if (g.channels == 1 && g.sample_bytes == 2)
            fri_encode_kernel<1, uint16_t><<<grid, kThreads, smem, stream>>>(g, qp, t.groups, t.tile_unit, p, c, dc);
        else
            fri_encode_kernel<3, uint16_t><<<grid, kThreads, smem, stream>>>(g, qp, t.groups, t.tile_unit, p, c, dc);
        if (launches) ++*launches;
    }
    if (g.sub_bits > 0) {
        const unsigned blocks = (unsigned)((int64_t)n_frames * g.n_fractals * g.channels);
        const size_t cs = ((size_t)3 << g.sub_bits) / 2 * sizeof(int32_t) + 16;
        fri_coarse_forward_kernel<<<blocks, kCoarseThreads, cs, stream>>>(qp, g.sub_bits, g.depth, d_dc, d_coefs);
        if (launches) ++*launches;
    }
    return cudaGetLastError();
}

cudaError_t launch_decode(const Geometry &g, const DeviceTables &t, const QuantParams &qp, const int32_t *d_coefs,
                          uint32_t n_frames, void *d_pixels, int32_t *d_dc, cudaStream_t stream,
                          uint32_t *launches)
{
    if (n_frames == 0 || g.n_groups == 0) return cudaSuccess;
    const size_t smem = kernel_smem_bytes(g);
    if (g.sub_bits > 0) {
        const unsigned blocks = (unsigned)((int64_t)n_frames * g.n_fractals * g.channels);
        const size_t cs = ((size_t)3 << g.sub_bits) / 2 * sizeof(int32_t) + 16;
        fri_coarse_inverse_kernel<<<blocks, kCoarseThreads, cs, stream>>>(qp, g.sub_bits, g.depth, d_coefs, d_dc);
        if (launches) ++*launches;
    }
    uint8_t *px = static_cast<uint8_t *>(d_pixels);
    for (uint32_t f0 = 0; f0 < n_frames; f0 += 65535u) {
        const uint32_t nf = n_frames - f0 < 65535u ? n_frames - f0 : 65535u;
        const dim3 grid((unsigned)g.n_groups, nf);
        uint8_t *p = px + (int64_t)f0 * g.frame_bytes;
        const int32_t *c = d_coefs + (int64_t)f0 * g.coefs_per_frame;
        int32_t *dc = d_dc ? d_dc + (((int64_t)f0 * g.n_fractals * g.channels) << g.sub_bits) : nullptr;
        if (g.channels == 1 && g.sample_bytes == 1)
            fri_decode_kernel<1, uint8_t><<<grid, kThreads, smem, stream>>>(g, qp, t.groups, t.tile_unit, t.ownership, c, dc, p);
        else if (g.channels == 3 && g.sample_bytes == 1)
            fri_decode_kernel<3, uint8_t><<<grid, kThreads, smem, stream>>>(g, qp, t.groups, t.tile_unit, t.ownership, c, dc, p);
        else if (g.channels == 1 && g.sample_bytes == 2)
            fri_decode_kernel<1, uint16_t><<<grid, kThreads, smem, stream>>>(g, qp, t.groups, t.tile_unit, t.ownership, c, dc, p);
        else
            fri_decode_kernel<3, uint16_t><<<grid, kThreads, smem, stream>>>(g, qp, t.groups, t.tile_unit, t.ownership, c, dc, p);
        if (launches) ++*launches;
    }
    return cudaGetLastError();
}

}  // namespace fri
