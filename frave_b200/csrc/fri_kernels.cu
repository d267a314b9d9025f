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

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <utility>

namespace fri {

// ------------------------------------------------------------------------------------------
// host helpers
// ------------------------------------------------------------------------------------------
Div make_div(int32_t q)
{
    // Signed magic number for a divisor d >= 2 (Hacker's Delight, fig. 10-1, positive d only):
    // the smallest p >= 32 with 2^p > nc * (d - 2^p mod d), nc = 2^31 - 1 - (2^31 mod d).
    const uint32_t d = q < 2 ? 2u : (uint32_t)q;
    const uint32_t two31 = 0x80000000u;
    const uint32_t nc = two31 - 1u - two31 % d;
    int p = 31;
    uint32_t q1 = two31 / nc, r1 = two31 - q1 * nc;  // 2^p / nc
    uint32_t q2 = two31 / d, r2 = two31 - q2 * d;    // 2^p / d
    uint32_t delta;
    do {
        ++p;
        q1 *= 2; r1 *= 2;
        if (r1 >= nc) { ++q1; r1 -= nc; }
        q2 *= 2; r2 *= 2;
        if (r2 >= d) { ++q2; r2 -= d; }
        delta = d - r2;
    } while (q1 < delta || (q1 == delta && r1 == 0));
    Div dv;
    dv.magic = (int32_t)(q2 + 1u);
    dv.addmask = dv.magic < 0 ? -1 : 0;
    dv.shift = p - 32;
    return dv;
}

bool make_small_div(int32_t q, SmallDiv &out)
{
    // magic = floor(2^(32+s) / q) + 1 with the smallest s >= 0 that keeps kSmallDivRange * q <=
    // 2^(32+s) (exactness for |n| <= kSmallDivRange); it stays below 2^31 for every q >= 3.
    if (q < 3) return false;
    int s = 0;
    while (((uint64_t)kSmallDivRange + 1) * (uint64_t)q > ((uint64_t)1 << (32 + s))) ++s;
    const uint64_t m = (((uint64_t)1 << (32 + s)) / (uint64_t)q) + 1u;
    if (m >= ((uint64_t)1 << 31)) return false;
    out.magic = (int32_t)m;
    out.shift = s;
    return true;
}

void make_quant_params(QuantParams &qp, const int32_t *q, int multiply)
{
    qp.active = 0;
    qp.small = 0;
    qp.pow2 = 0;
    qp.fix = 0;
    qp.multiply = multiply;
    for (int l = 0; l < 32; ++l) {
        const int32_t v = q ? q[l] : 1;
        const Div dv = make_div(v);
        qp.q[l] = v;
        qp.magic[l] = dv.magic;
        qp.addmask[l] = dv.addmask;
        qp.shift[l] = dv.shift;
        qp.pow2_shift[l] = 0;
        if (v >= 2 && (v & (v - 1)) == 0) {
            qp.pow2 |= 1u << l;
            while ((1 << qp.pow2_shift[l]) < v) ++qp.pow2_shift[l];
        }
        SmallDiv sd{0, 0};
        if (make_small_div(v, sd)) qp.small |= 1u << l;
        qp.small_magic[l] = sd.magic;
        qp.small_shift[l] = sd.shift;
        if (v != 1) qp.active |= 1u << l;
        if (l < 31 && (q ? q[l + 1] : 1) != v) qp.fix |= 1u << l;
    }
}

int cta_threads(const Geometry &g)
{
    const int warps = (g.group_a * g.group_b + g.tiles_per_warp - 1) / g.tiles_per_warp;
    return 32 * (warps < 1 ? 1 : (warps > kMaxWarps ? kMaxWarps : warps));
}

size_t kernel_smem_bytes(const Geometry &g)
{
    return (((size_t)g.region_h * g.pitch + 15) & ~(size_t)15) + (size_t)(cta_threads(g) / 32) * scratch_units(g.channels) * kScratchInts * sizeof(int32_t);
}

#if FRI_TRACE
__device__ unsigned long long g_trace[3 * 16384];
__device__ unsigned long long g_trace2[4 * 16384];
cudaError_t debug_trace(unsigned long long *out, size_t n) { return cudaMemcpyFromSymbol(out, g_trace, n * sizeof(unsigned long long)); }
cudaError_t debug_trace2(unsigned long long *out, size_t n) { return cudaMemcpyFromSymbol(out, g_trace2, n * sizeof(unsigned long long)); }
#endif

namespace {

// Compile-time quantization classes of the register levels (6..8) of a depth-9 tile, so that the two
// configurations that matter do not walk the generic per-level mode tests for every channel:
//   kQuantNone      layers 6..9 all have divisor 1 — the reference's matrix (quantization.rs:3-5);
//   kQuantSmallest  only layers 8 and 9 are active, with one divisor — "dividing the smallest layer
//                   of fractals" (README.md:12; BASELINE.json's divisor sweep);
//   kQuantGeneric   anything else (deep trees: everything but kQuantNone).
// Layers 0..5 are handled by the run-time tests in every class (once per tile, not per channel).
constexpr int kQuantGeneric = 0, kQuantNone = 1, kQuantSmallest = 2;

int quant_class(const QuantParams &qp, const Geometry &g)
{
    const uint32_t hi = (qp.active >> (g.sub_bits + 6)) & 0xfu;  // the register levels' layers sit sub_bits deeper in a deep tree
    if (hi == 0) return kQuantNone;
    if (g.sub_bits != 0) return kQuantGeneric;
    if (hi == 0xcu && qp.q[8] == qp.q[9]) return kQuantSmallest;
    return kQuantGeneric;
}

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gmem_src)
{
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(gmem_src) : "memory");
}
#if FRI_TRACE
__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define FRI_TRACE_MARK(slot) do { if (threadIdx.x == 0 && blockIdx.x < 16384) g_trace[3 * blockIdx.x + (slot)] = gtime(); } while (0)
#else
#define FRI_TRACE_MARK(slot) do { } while (0)
#endif
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_but_one() { asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }

// Plan tables (group descriptors, chunk lists) are a few hundred KB that every CTA of every launch
// reads first, and the loads that depend on them cannot start before they arrive.  They are
// fetched with an L2 evict_last policy so that the ~250 MB of streaming traffic per launch does
// not push them out of the 126 MB L2 between launches.
__device__ __forceinline__ uint64_t table_policy()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint32_t ld_table(const uint32_t *p, uint64_t pol)
{
    uint32_t v;
    asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ uint32_t ld_table(const uint16_t *p, uint64_t pol)
{
    uint16_t v;
    asm volatile("ld.global.nc.L2::cache_hint.u16 %0, [%1], %2;" : "=h"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ GroupDesc ld_group(const GroupDesc *p, uint64_t pol)
{
    GroupDesc g;
    asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(g.x0), "=r"(g.y0), "=r"(g.tile_mask), "=r"(g.tile_base)
                 : "l"(p), "l"(pol));
    return g;
}

// Programmatic dependent launch: every CTA lets the next kernel of the stream start launching as soon as
// all CTAs of this one have started (its CTAs take the slots this kernel's last wave leaves empty and run
// their prologue there), and waits for the previous kernel to complete — memory flushed — before it touches
// frame data.  Plan tables are constant and may be read before the wait.  Without the launch attribute both
// instructions are no-ops.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// wrapping i32 arithmetic (release-mode Rust semantics)
__device__ __forceinline__ int wsub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }
__device__ __forceinline__ int wadd(int a, int b) { return (int)((unsigned)a + (unsigned)b); }

// d / 2 with Rust/C truncation toward zero.  Written in PTX so that the compiler keeps it a
// 32-bit add-sign-bit + arithmetic shift (it otherwise narrows the arithmetic on 8-bit samples
// to 16 bits and pays for it in masks and sign extensions).
__device__ __forceinline__ int half_trunc(int d)
{
    int h;
    asm("{\n\t.reg .s32 t;\n\tshr.u32 t, %1, 31;\n\tadd.s32 t, t, %1;\n\tshr.s32 %0, t, 1;\n\t}" : "=r"(h) : "r"(d));
    return h;
}
// forward lifting of one node: d = l - r; s = r + d/2 (truncating)   wavelet_transform.rs:211-218
__device__ __forceinline__ void lift(int l, int r, int &d, int &s)
{
    d = wsub(l, r);
    s = wadd(r, half_trunc(d));
}
// inverse lifting of one node: r = s - d/2; l = d + r                 wavelet_transform.rs:366-367
__device__ __forceinline__ void unlift(int s, int d, int &l, int &r)
{
    r = wsub(s, half_trunc(d));
    l = wadd(d, r);
}

// floor(log2(pos + 1)): the reference's layer index of heap position pos (quantization.rs:13)
__device__ __forceinline__ int layer_of(uint32_t pos) { return 31 - __clz((int)(pos + 1u)); }

// quantization::encode (quantization.rs:19) / ::decode (:37) of one coefficient of layer l.
// The branches are uniform whenever l is.
// encoder only (|d| <= 65535): prefers the narrow-range division
__device__ __forceinline__ int quant_layer_enc(const QuantParams &qp, int d, int l)
{
    if (!((qp.active >> l) & 1u)) return d;
    return ((qp.small >> l) & 1u) ? trunc_div_small(d, qp.sdiv(l)) : trunc_div(d, qp.div(l));
}
__device__ __forceinline__ int dequant_layer(const QuantParams &qp, int d, int l)
{
    if (!((qp.active >> l) & 1u)) return d;
    if (qp.multiply) return (int)((unsigned)d * (unsigned)qp.q[l]);
    return ((qp.pow2 >> l) & 1u) ? trunc_div_pow2(d, qp.pow2_shift[l]) : trunc_div(d, qp.div(l));
}

// clamp (images.rs:109) + store of one sample to shared memory: max(min(v, 255), 0) is one DPX
// instruction (VIMNMX.RELU) and the byte store takes the low bits of the 32-bit register, where
// cvt.sat.u8.s32 costs a conversion plus a re-masking LOP3 per sample.
template <typename S>
__device__ __forceinline__ void store_clamped(uint32_t saddr, int v)
{
    const int c = __vimin_s32_relu(v, sizeof(S) == 1 ? 255 : 65535);
    if (sizeof(S) == 1)
        asm volatile("st.shared.u8 [%0], %1;" ::"r"(saddr), "r"(c));
    else
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(saddr), "r"(c));
}

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

// Per-CTA view of the staged region.  Shared-memory byte s of the region image corresponds to
// global byte gbase0 + r * delta + s for a byte of staged row r (delta = row_stride - pitch is a
// multiple of 16, gbase0 is 16-byte aligned), so 16-byte chunks are aligned on both sides.
struct RegionView {
    int64_t gbase0;  // global address of shared-memory offset 0 of row 0
    int delta;       // row_stride - pitch
    int phi0;        // shared-memory offset of region pixel (0, 0): its global address & 15
    int xb0;         // byte offset of region column 0 inside its image row (x0 * pixel bytes)
    bool interior;   // the whole staged region lies inside the image
    // Global address of shared-memory byte s of staged row r.  The offset fits 32 bits (a region spans
    // at most region_h rows), so the per-chunk address is one 32-bit multiply-add and one wide add.
    __device__ __forceinline__ uint8_t *gaddr(int r, int s) const
    {
        return reinterpret_cast<uint8_t *>(gbase0) + (ptrdiff_t)(r * delta + s);
    }
};

template <int PB>
__device__ __forceinline__ RegionView region_view(const Geometry &g, const GroupDesc &gd, const void *frame_base)
{
    RegionView v;
    const int64_t a0 = (int64_t)(uintptr_t)frame_base + (int64_t)gd.y0 * g.row_stride + (int64_t)gd.x0 * PB;
    v.phi0 = (int)(a0 & 15);
    v.gbase0 = a0 - v.phi0;
    v.delta = (int)(g.row_stride - g.pitch);
    v.xb0 = gd.x0 * PB;
    v.interior = gd.x0 >= 0 && gd.y0 >= 0 && gd.x0 + g.region_w <= g.width && gd.y0 + g.region_h <= g.height;
    return v;
}

// Output/input addressing of one (base tile, channel) task.
struct TaskAddr {
    int64_t block;   // element offset of the fractal-channel coefficient block
    uint32_t node;   // heap index of the base tile's root inside the fractal (1 at depth 9)
    int64_t dc;      // element offset into the low-pass scratch (depth > 9)
    bool last;       // the base tile is the last one of its fractal
};

// DEEP == false instantiates the depth-9 (reference) case with sub_bits == 0 folded in at compile
// time: node == 1, every tile is "last", no tile_unit / low-pass scratch traffic.
template <int C, bool DEEP>
__device__ __forceinline__ TaskAddr task_addr(const Geometry &g, const uint32_t *tile_unit, int frame, int tile, int ch)
{
    TaskAddr a;
    if (!DEEP) {
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

// Stores the bytes of one 16-byte chunk selected by mask m (bit j = byte j) to global memory.
__device__ __forceinline__ void store_chunk_masked(uint8_t *gp, const uint8_t *sp, uint32_t m)
{
    const int4 v = *reinterpret_cast<const int4 *>(sp);
    if (m == 0xffffu) {
        *reinterpret_cast<int4 *>(gp) = v;
        return;
    }
    const uint32_t w[4] = {(uint32_t)v.x, (uint32_t)v.y, (uint32_t)v.z, (uint32_t)v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t nib = (m >> (4 * k)) & 15u;
        if (nib == 15u) {
            *reinterpret_cast<uint32_t *>(gp + 4 * k) = w[k];
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if ((nib >> j) & 1u) gp[4 * k + j] = (uint8_t)(w[k] >> (8 * j));
        }
    }
}

// Coefficient stores / loads of 4, 2 or 1 consecutive values (streaming), for the two coefficient
// types a device array can have: int32 (the reference's i32) and int16 (every coefficient of an
// 8-bit image fits; 3 B per sample instead of 5 — reported as its own variant).
__device__ __forceinline__ uint32_t pack16(int lo, int hi)
{
    uint32_t r;
    asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(r) : "r"(hi), "r"(lo));  // {sat(hi), sat(lo)}
    return r;
}
__device__ __forceinline__ void st_c4(int32_t *p, int4 v) { __stcs(reinterpret_cast<int4 *>(p), v); }
__device__ __forceinline__ void st_c2(int32_t *p, int2 v) { __stcs(reinterpret_cast<int2 *>(p), v); }
__device__ __forceinline__ void st_c1(int32_t *p, int v) { __stcs(p, v); }
__device__ __forceinline__ void st_c4(int16_t *p, int4 v)
{
    __stcs(reinterpret_cast<int2 *>(p), make_int2((int)pack16(v.x, v.y), (int)pack16(v.z, v.w)));
}
__device__ __forceinline__ void st_c2(int16_t *p, int2 v) { __stcs(reinterpret_cast<int *>(p), (int)pack16(v.x, v.y)); }
__device__ __forceinline__ void st_c1(int16_t *p, int v) { __stcs(p, (short)pack16(v, 0)); }
__device__ __forceinline__ int4 ld_c4(const int32_t *p) { return __ldcs(reinterpret_cast<const int4 *>(p)); }
__device__ __forceinline__ int2 ld_c2(const int32_t *p) { return __ldcs(reinterpret_cast<const int2 *>(p)); }
__device__ __forceinline__ int ld_c1(const int32_t *p) { return __ldcs(p); }
__device__ __forceinline__ int4 ld_c4(const int16_t *p)
{
    const int2 r = __ldcs(reinterpret_cast<const int2 *>(p));
    return make_int4((int)(short)r.x, r.x >> 16, (int)(short)r.y, r.y >> 16);
}
__device__ __forceinline__ int2 ld_c2(const int16_t *p)
{
    const int r = __ldcs(reinterpret_cast<const int *>(p));
    return make_int2((int)(short)r, r >> 16);
}
__device__ __forceinline__ int ld_c1(const int16_t *p) { return (int)__ldcs(p); }

// ------------------------------------------------------------------------------------------
// CTA-level building blocks
// ------------------------------------------------------------------------------------------

constexpr int kListUnroll = 4;
constexpr uint32_t kNoChunk = 0xffffffffu;  // not a valid chunk-list entry (row 65535)

// Enqueues the copy of a group's pixel footprint into `region`: one 16-byte chunk (aligned in
// global and in shared memory) per thread and iteration, taken from the plan's list of chunks
// that hold at least one pixel of the group's tiles.  Completion: cp.async.wait_all + barrier.
__device__ __forceinline__ void stage_group(const Geometry &g, const GroupDesc &gd, const RegionView &rv,
                                            const uint32_t *__restrict__ stage_list, uint8_t *region, int k_lo, int k_hi,
                                            uint64_t pol)
{
    const int n_threads = blockDim.x;
    const uint32_t *cl = stage_list + (size_t)rv.phi0 * g.list_cap + k_lo;
    const int n_all = k_hi - k_lo;
    if (rv.interior) {
        // list entries are fetched four at a time so that their latencies overlap
        for (int k0 = threadIdx.x; k0 < n_all; k0 += kListUnroll * n_threads) {
            uint32_t e[kListUnroll];
#pragma unroll
            for (int u = 0; u < kListUnroll; ++u) e[u] = k0 + u * n_threads < n_all ? ld_table(cl + k0 + u * n_threads, pol) : kNoChunk;
#pragma unroll
            for (int u = 0; u < kListUnroll; ++u)
                if (e[u] != kNoChunk) {
                    const int r = (int)(e[u] >> 16), s = (int)(e[u] & 0xffffu) << 4;
                    cp_async_16(region + s, rv.gaddr(r, s));
                }
        }
    } else {
        const int stride32 = (int)g.row_stride;
        for (int k = threadIdx.x; k < n_all; k += n_threads) {
            const uint32_t e = ld_table(cl + k, pol);
            const int r = (int)(e >> 16), s = (int)(e & 0xffffu) << 4;
            const int y = gd.y0 + r;
            const bool yin = (unsigned)y < (unsigned)g.height;
            const int xb = rv.xb0 + s - (r * g.pitch + rv.phi0);  // byte position of the chunk inside image row y
            const uint8_t *gp = rv.gaddr(r, s);
            if (yin && xb >= 0 && xb + 16 <= stride32) {
                cp_async_16(region + s, gp);
            } else if (!yin || xb + 16 <= 0 || xb >= stride32) {
                *reinterpret_cast<int4 *>(region + s) = make_int4(0, 0, 0, 0);
            } else {  // chunk straddles the left or right image edge
#pragma unroll 1
                for (int j = 0; j < 16; ++j)
                    region[s + j] = (xb + j >= 0 && xb + j < stride32) ? __ldg(gp + j) : (uint8_t)0;
            }
        }
    }
}

// ---- bulk-copy (TMA, cp.async.bulk) staging of an interior group: one copy per staged row, issued by
// one thread each and completed on an mbarrier, instead of ~1700 16-byte cp.async driven by the chunk
// list.  Two barriers: rows holding pixels of every warp's first tile / the rest.
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "FRI_MBAR_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra FRI_MBAR_DONE;\n\t"
        "bra FRI_MBAR_WAIT;\n"
        "FRI_MBAR_DONE:\n\t}" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
                 : "memory");
}

// Thread k < region_h issues the copy of staged row row_order[k] (the 16-byte-aligned cover of its owned
// span) and arrives on bars[0] (first-round rows) or bars[1].
__device__ __forceinline__ void stage_rows_bulk(const Geometry &g, const RegionView &rv, uint8_t *region, uint64_t *bars,
                                                bool two_stage)
{
    const int k = threadIdx.x;
    if (k >= g.region_h) return;
    uint64_t *bar = bars + ((two_stage && k >= g.n_rows_first) ? 1 : 0);
    const int r = g.row_order[k];
    if (g.row_hi[r] == 0) {
        mbar_arrive(bar);
        return;
    }
    const int srow = r * g.pitch + rv.phi0;
    const int s0 = (srow + g.row_lo[r]) & ~15, s1 = (srow + g.row_hi[r] + 15) & ~15;
    mbar_arrive_expect(bar, (uint32_t)(s1 - s0));
    bulk_g2s(region + s0, rv.gaddr(r, s0), (uint32_t)(s1 - s0), bar);
}

// ---- forward transform + quantization, building blocks
//
// Register levels of one (base tile, channel): gather 16 leaves per lane (two depth-3 subtrees), levels
// 8, 7, 6 in registers, quantize, store; the two level-6 low-pass values go to `sc` (64 ints: node 64 + l).
template <typename S, int PB, bool DEEP, int QS, typename CT>
__device__ __forceinline__ void enc_register_levels(const Geometry &g, const QuantParams &qp, const uint8_t *p0, int half,
                                                    int lane, bool lastB, int top, int32_t *sc, CT *__restrict__ out,
                                                    size_t node)
{
    const uint8_t *p1 = p0 + g.pitch, *p2 = p1 + g.pitch;
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
    sc[lane] = sA;       // low-pass of node 64 + lane
    sc[32 + lane] = sB;  // low-pass of node 96 + lane

    // quantization.rs:13 — layer = level, except the last node of a level: level + 1
    if (QS == kQuantSmallest) {
        // only layers 8 and 9 are active and share one divisor: every level-8 node (the last one
        // sits in layer 9) and the last node of level 7 (layer 8)
        int r7 = b7[1];
        if (qp.small & 0x100u) {
            const SmallDiv dv = qp.sdiv(8);
#pragma unroll
            for (int m = 0; m < 4; ++m) { a8[m] = trunc_div_small(a8[m], dv); b8[m] = trunc_div_small(b8[m], dv); }
            r7 = trunc_div_small(r7, dv);
        } else {
            const Div dv = qp.div(8);
#pragma unroll
            for (int m = 0; m < 4; ++m) { a8[m] = trunc_div(a8[m], dv); b8[m] = trunc_div(b8[m], dv); }
            r7 = trunc_div(r7, dv);
        }
        if (lane == 31) b7[1] = r7;
    } else if (QS == kQuantGeneric && ((qp.active >> (top + 6)) & 0xfu)) {
        const int r8 = b8[3], r7 = b7[1], r6 = b6;  // unquantized values of the level-last nodes
#define FRI_QLEVEL(L, N, A, B)                                                              \
        if ((qp.active >> (top + (L))) & 1u) {                                              \
            if ((qp.small >> (top + (L))) & 1u) {                                           \
                const SmallDiv dv = qp.sdiv(top + (L));                                     \
                _Pragma("unroll") for (int m = 0; m < (N); ++m) { A[m] = trunc_div_small(A[m], dv); B[m] = trunc_div_small(B[m], dv); } \
            } else {                                                                        \
                const Div dv = qp.div(top + (L));                                           \
                _Pragma("unroll") for (int m = 0; m < (N); ++m) { A[m] = trunc_div(A[m], dv); B[m] = trunc_div(B[m], dv); } \
            }                                                                               \
        }
        int a6v[1] = {a6}, b6v[1] = {b6};
        FRI_QLEVEL(8, 4, a8, b8)
        FRI_QLEVEL(7, 2, a7, b7)
        FRI_QLEVEL(6, 1, a6v, b6v)
#undef FRI_QLEVEL
        a6 = a6v[0];
        b6 = b6v[0];
        if (lastB && ((qp.fix >> (top + 6)) & 7u)) {  // only where the next layer's divisor differs
            if ((qp.fix >> (top + 8)) & 1u) b8[3] = quant_layer_enc(qp, r8, top + 9);
            if ((qp.fix >> (top + 7)) & 1u) b7[1] = quant_layer_enc(qp, r7, top + 8);
            if ((qp.fix >> (top + 6)) & 1u) b6 = quant_layer_enc(qp, r6, top + 7);
        }
    }
    CT *o8 = out + (node << 8), *o7 = out + (node << 7), *o6 = out + (node << 6);
    st_c4(o8 + 4 * lane, make_int4(a8[0], a8[1], a8[2], a8[3]));
    st_c4(o8 + 128 + 4 * lane, make_int4(b8[0], b8[1], b8[2], b8[3]));
    st_c2(o7 + 2 * lane, make_int2(a7[0], a7[1]));
    st_c2(o7 + 64 + 2 * lane, make_int2(b7[0], b7[1]));
    st_c1(o6 + lane, a6);
    st_c1(o6 + 32 + lane, b6);
}

// Levels 5..0 of one (base tile, channel) by the 8 lanes of a lane group: lane j8 folds the level-6 values
// sp[0 .. 7] (= s6[8 j8 .. 8 j8 + 7]) through levels 5..3 in registers and levels 2..0 with three
// shuffles inside the group (every lane of the warp must call this), quantizes and — if `live` — stores.
template <bool DEEP, typename CT>
__device__ __forceinline__ void enc_top_levels(const QuantParams &qp, const int32_t *sp, const TaskAddr &ta, int j8, int top,
                                               int sub_bits, bool live, CT *__restrict__ out, int32_t *__restrict__ dc_slot)
{
    const int4 x0 = *reinterpret_cast<const int4 *>(sp), x1 = *reinterpret_cast<const int4 *>(sp + 4);
    int d5[4], s5[4], d4[2], s4[2], d3, d2, d1, d0, s3, s2, s1, s0;
    lift(x0.x, x0.y, d5[0], s5[0]);
    lift(x0.z, x0.w, d5[1], s5[1]);
    lift(x1.x, x1.y, d5[2], s5[2]);
    lift(x1.z, x1.w, d5[3], s5[3]);
    lift(s5[0], s5[1], d4[0], s4[0]);
    lift(s5[2], s5[3], d4[1], s4[1]);
    lift(s4[0], s4[1], d3, s3);
    lift(s3, __shfl_xor_sync(0xffffffffu, s3, 1), d2, s2);  // meaningful in lanes j8 % 2 == 0
    lift(s2, __shfl_xor_sync(0xffffffffu, s2, 2), d1, s1);  // j8 % 4 == 0
    lift(s1, __shfl_xor_sync(0xffffffffu, s1, 4), d0, s0);  // j8 == 0
    if ((qp.active >> top) & 0x7fu) {
        // level-L nodes sit in layer top + L, the level's last node in layer top + L + 1
        const bool lastG = ta.last && j8 == 7;
        d5[0] = quant_layer_enc(qp, d5[0], top + 5); d5[1] = quant_layer_enc(qp, d5[1], top + 5);
        d5[2] = quant_layer_enc(qp, d5[2], top + 5); d5[3] = quant_layer_enc(qp, d5[3], top + (lastG ? 6 : 5));
        d4[0] = quant_layer_enc(qp, d4[0], top + 4); d4[1] = quant_layer_enc(qp, d4[1], top + (lastG ? 5 : 4));
        d3 = quant_layer_enc(qp, d3, top + (lastG ? 4 : 3));
        d2 = quant_layer_enc(qp, d2, top + ((ta.last && j8 == 6) ? 3 : 2));
        d1 = quant_layer_enc(qp, d1, top + ((ta.last && j8 == 4) ? 2 : 1));
        if (sub_bits == 0) {
            d0 = quant_layer_enc(qp, d0, 1);  // position 1 is the last node of level 0
            s0 = quant_layer_enc(qp, s0, 0);  // position 0: the low-pass root (wavelet_transform.rs:221)
        } else {
            d0 = quant_layer_enc(qp, d0, top + (ta.last ? 1 : 0));
        }
    }
    if (live) {
        const size_t node = ta.node;
        st_c4(out + (node << 5) + 4 * j8, make_int4(d5[0], d5[1], d5[2], d5[3]));
        st_c2(out + (node << 4) + 2 * j8, make_int2(d4[0], d4[1]));
        st_c1(out + (node << 3) + j8, d3);
        if ((j8 & 1) == 0) st_c1(out + (node << 2) + (j8 >> 1), d2);
        if ((j8 & 3) == 0) st_c1(out + (node << 1) + (j8 >> 2), d1);
        if (j8 == 0) {
            st_c1(out + node, d0);
            if (sub_bits == 0) st_c1(out, s0);
            else *dc_slot = s0;
        }
    }
}

// Forward transform + quantization of the tiles of one staged group; warp w takes tiles
// w, w + n_warps, ...
//   C == 3, per warp iteration one base tile, all channels:
//     phase 1 (per channel): enc_register_levels; phase 2 (all channels at once): lane group lane / 8 owns a
//     channel (enc_top_levels), 24 of 32 lanes busy;
//   C == 1, per warp iteration a batch of up to four base tiles: phase 1 tile after tile, then ONE phase 2
//     in which lane group lane / 8 owns a tile — all 32 lanes busy instead of 8 (the top levels are a
//     quarter of a 1-channel tile's instructions).
template <int C, typename S, bool DEEP, int QS, typename CT>
__device__ __forceinline__ void encode_tiles(const Geometry &g, const QuantParams &qp, const GroupDesc &gd, const RegionView &rv,
                                             const uint32_t *__restrict__ tile_unit, int frame, const uint8_t *region,
                                             int32_t *scratch, CT *__restrict__ coefs, int32_t *__restrict__ dc_out,
                                             bool two_stage, uint64_t *bars)
{
    constexpr int SB = (int)sizeof(S);
    constexpr int PB = C * SB;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const int n_present = __popc(gd.tile_mask);
    const uint8_t *lane_base = region + rv.phi0 + lane_anchor_bytes(lane, g.pitch, PB);
    const int half = kHalfB.y * g.pitch + kHalfB.x * PB;
    const int sub_bits = DEEP ? g.sub_bits : 0, depth = DEEP ? g.depth : kBaseDepth;
    const int top = sub_bits;  // fractal level of a base tile's root
    const bool sparse_group = (gd.tile_mask & (gd.tile_mask + 1u)) != 0;
    const int grp = min(lane >> 3, scratch_units(C) - 1), j8 = lane & 7;  // phase 2 roles
    auto second_round_wait = [&]() {  // the rest of the footprint must have landed
        if (bars) {
            mbar_wait(bars + 1, 0);
        } else {
            cp_async_wait_all();
            __syncthreads();
        }
    };
    if (C == 1) {
        for (int e0 = warp; e0 < n_present; e0 += kTileBatch * n_warps) {
#pragma unroll
            for (int k = 0; k < kTileBatch; ++k) {
                const int e = e0 + k * n_warps;
                if (e >= n_present) break;
                if (two_stage && e == warp + n_warps) second_round_wait();
                const int slot = sparse_group ? (int)__fns(gd.tile_mask, 0, e + 1) : e;
                const TaskAddr ta = task_addr<C, DEEP>(g, tile_unit, frame, gd.tile_base + e, 0);
                enc_register_levels<S, PB, DEEP, QS, CT>(g, qp, lane_base + g.tile_off[slot], half, lane, ta.last && lane == 31, top,
                                                         scratch + k * kScratchInts, coefs + ta.block, ta.node);
            }
            __syncwarp();
            const int eg = e0 + grp * n_warps;
            const bool live = eg < n_present;
            const TaskAddr ta = task_addr<C, DEEP>(g, tile_unit, frame, gd.tile_base + (live ? eg : e0), 0);
            enc_top_levels<DEEP, CT>(qp, scratch + grp * kScratchInts + 8 * j8, ta, j8, top, sub_bits, live, coefs + ta.block,
                                     DEEP ? dc_out + ta.dc : nullptr);
            __syncwarp();
        }
        return;
    }
    const bool grp_live = (lane >> 3) < C;
    for (int e = warp; e < n_present; e += n_warps) {
        if (two_stage && e == warp + n_warps) second_round_wait();
        const int slot = sparse_group ? (int)__fns(gd.tile_mask, 0, e + 1) : e;
        const uint8_t *t0 = lane_base + g.tile_off[slot];
        const TaskAddr ta = task_addr<C, DEEP>(g, tile_unit, frame, gd.tile_base + e, 0);
        const bool lastB = ta.last && lane == 31;  // this lane holds the last node of levels 8..6
#pragma unroll
        for (int ch = 0; ch < C; ++ch)
            enc_register_levels<S, PB, DEEP, QS, CT>(g, qp, t0 + ch * SB, half, lane, lastB, top, scratch + ch * kScratchInts,
                                                     coefs + ta.block + ((int64_t)ch << depth), ta.node);
        __syncwarp();
        // ---- phase 2: levels 5..0 of all channels, 8 lanes per channel
        enc_top_levels<DEEP, CT>(qp, scratch + grp * kScratchInts + 8 * j8, ta, j8, top, sub_bits, grp_live,
                                 coefs + ta.block + ((int64_t)grp << depth),
                                 DEEP ? dc_out + ta.dc + ((int64_t)grp << sub_bits) : nullptr);
        __syncwarp();
    }
}

// Pulls a group's coefficients towards L2: at depth 9 the blocks of a group's tiles are adjacent
// (plan order is group-major), n_present * C * 2 KB in one run.
template <int C, bool DEEP, typename CT>
__device__ __forceinline__ void prefetch_group_coefs(const Geometry &g, const GroupDesc &gd, int frame,
                                                     const CT *__restrict__ coefs, const uint32_t *__restrict__ tile_unit)
{
    if (DEEP) {
        // depth > 9: a base tile's coefficients are nine runs inside its fractal's heap (node << level):
        // 8 + 4 + 2 lines for levels 8, 7, 6 and one line for each level below, per tile and channel
        constexpr int kLines = 8 + 4 + 2 + 6;
        const int total = __popc(gd.tile_mask) * C * kLines;
#pragma unroll 1
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int task = i / kLines, s = i - task * kLines;
            const TaskAddr ta = task_addr<C, true>(g, tile_unit, frame, (int)gd.tile_base + task / C, task % C);
            const int level = s < 8 ? 8 : s < 12 ? 7 : s < 14 ? 6 : 19 - s;
            const int line = s < 8 ? s : s < 12 ? s - 8 : s < 14 ? s - 12 : 0;
            const char *p = reinterpret_cast<const char *>(coefs + ta.block + ((int64_t)ta.node << level)) + line * 128;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
        }
        return;
    }
    const char *first = reinterpret_cast<const char *>(coefs + ((((int64_t)frame * g.n_fractals + gd.tile_base) * C) << kBaseDepth));
    const int lines = __popc(gd.tile_mask) * C * (4 * (int)sizeof(CT));  // 128-byte lines: 512 coefficients per (tile, channel)
#pragma unroll 1
    for (int i = threadIdx.x; i < lines; i += blockDim.x)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(first + (size_t)i * 128));
}

// Timing-only what-if switches (wrong results; variant builds for bounding what a change could buy)
#ifndef FRI_WHATIF_SKIP_MIXED
#define FRI_WHATIF_SKIP_MIXED 0   // decode: partially owned chunks are not written
#endif
#ifndef FRI_WHATIF_SKIP_FULL
#define FRI_WHATIF_SKIP_FULL 0    // decode: fully owned chunks are not written
#endif
#ifndef FRI_WHATIF_SKIP_SCATTER
#define FRI_WHATIF_SKIP_SCATTER 0 // decode: leaves are not stored to the staged region (one store per lane keeps them live)
#endif
// The coefficients of levels 6..8 a lane consumes for one (tile, channel): 2 x 128, 2 x 64, 2 x 32 bits.
struct LaneCoefs {
    int4 a8, b8;
    int2 a7, b7;
    int a6, b6;
};
template <typename CT>
__device__ __forceinline__ LaneCoefs load_lane_coefs(const CT *__restrict__ in, size_t node, int lane)
{
    const CT *i8 = in + (node << 8), *i7 = in + (node << 7), *i6 = in + (node << 6);
    LaneCoefs c;
    c.a8 = ld_c4(i8 + 4 * lane);
    c.b8 = ld_c4(i8 + 128 + 4 * lane);
    c.a7 = ld_c2(i7 + 2 * lane);
    c.b7 = ld_c2(i7 + 64 + 2 * lane);
    c.a6 = ld_c1(i6 + lane);
    c.b6 = ld_c1(i6 + 32 + lane);
    return c;
}

// The top 64 coefficients of one channel as lane j8 of the channel's lane group consumes them.
struct TopCoefs {
    int4 d5;
    int2 d4;
    int d3, d2, d1, d0, s0;
};
template <bool DEEP, typename CT>
__device__ __forceinline__ TopCoefs load_top_coefs(const CT *__restrict__ coefs, const int32_t *__restrict__ dc_in,
                                                   const TaskAddr &ta, int grp, int j8, int depth, int sub_bits)
{
    const CT *in = coefs + ta.block + ((int64_t)grp << depth);
    const size_t node = ta.node;
    TopCoefs t;
    t.d5 = ld_c4(in + (node << 5) + 4 * j8);
    t.d4 = ld_c2(in + (node << 4) + 2 * j8);
    t.d3 = ld_c1(in + (node << 3) + j8);
    t.d2 = ld_c1(in + (node << 2) + (j8 >> 1));
    t.d1 = ld_c1(in + (node << 1) + (j8 >> 2));
    t.d0 = ld_c1(in + node);
    t.s0 = (!DEEP || sub_bits == 0) ? ld_c1(in) : dc_in[ta.dc + ((int64_t)grp << sub_bits)];
    return t;
}

// ---- dequantization + inverse transform, building blocks
//
// Levels 0..5 of one (base tile, channel) by the 8 lanes of a lane group: each lane walks its own
// root-to-subtree path and ends with the eight level-6 low-pass values 8 j8 .. 8 j8 + 7, stored to sp[0..7].
template <bool DEEP>
__device__ __forceinline__ void dec_top_levels(const QuantParams &qp, const TopCoefs &tc, const TaskAddr &ta, int j8, int top,
                                               int sub_bits, bool live, int32_t *sp)
{
    int4 d5 = tc.d5;
    int2 d4 = tc.d4;
    int d3 = tc.d3, d2 = tc.d2, d1 = tc.d1, d0 = tc.d0, s0 = tc.s0;
    if ((qp.active >> top) & 0x7fu) {
        const bool lastG = ta.last && j8 == 7;
        d5.x = dequant_layer(qp, d5.x, top + 5); d5.y = dequant_layer(qp, d5.y, top + 5);
        d5.z = dequant_layer(qp, d5.z, top + 5); d5.w = dequant_layer(qp, d5.w, top + (lastG ? 6 : 5));
        d4.x = dequant_layer(qp, d4.x, top + 4); d4.y = dequant_layer(qp, d4.y, top + (lastG ? 5 : 4));
        d3 = dequant_layer(qp, d3, top + (lastG ? 4 : 3));
        d2 = dequant_layer(qp, d2, top + ((ta.last && (j8 >> 1) == 3) ? 3 : 2));
        d1 = dequant_layer(qp, d1, top + ((ta.last && (j8 >> 2) == 1) ? 2 : 1));
        if (sub_bits == 0) {
            d0 = dequant_layer(qp, d0, 1);
            s0 = dequant_layer(qp, s0, 0);
        } else {
            d0 = dequant_layer(qp, d0, top + (ta.last ? 1 : 0));  // s0 was dequantized by the coarse kernel
        }
    }
    int l, r, s;
    unlift(s0, d0, l, r); s = (j8 & 4) ? r : l;
    unlift(s, d1, l, r);  s = (j8 & 2) ? r : l;
    unlift(s, d2, l, r);  s = (j8 & 1) ? r : l;
    int s4[2], s5[4];
    int4 x0, x1;
    unlift(s, d3, s4[0], s4[1]);
    unlift(s4[0], d4.x, s5[0], s5[1]);
    unlift(s4[1], d4.y, s5[2], s5[3]);
    unlift(s5[0], d5.x, x0.x, x0.y);
    unlift(s5[1], d5.y, x0.z, x0.w);
    unlift(s5[2], d5.z, x1.x, x1.y);
    unlift(s5[3], d5.w, x1.z, x1.w);
    if (live) {
        *reinterpret_cast<int4 *>(sp) = x0;
        *reinterpret_cast<int4 *>(sp + 4) = x1;
    }
}

// Register levels of one (base tile, channel): dequantize the lane's 2 x 128 + 2 x 64 + 2 x 32 bit
// coefficient runs, unfold levels 6..8 of its two depth-3 subtrees from the level-6 low-pass values sA, sB
// and scatter the 16 clamped leaves into the staged region (p0: shared-memory address of the lane's first
// leaf of this channel).
template <typename S, int PB, bool DEEP, int QS>
__device__ __forceinline__ void dec_register_levels(const Geometry &g, const QuantParams &qp, const LaneCoefs &c, int sA, int sB,
                                                    int lane, bool lastB, int top, uint32_t p0, int half)
{
    int4 a8 = c.a8, b8 = c.b8;
    int2 a7 = c.a7, b7 = c.b7;
    int a6 = c.a6, b6 = c.b6;
    if (QS == kQuantSmallest) {
        // only layers 8 and 9 are active and share one divisor (see enc_register_levels)
        int r7 = b7.y;
#define FRI_DQ8                                                                \
        a8.x = f(a8.x); a8.y = f(a8.y); a8.z = f(a8.z); a8.w = f(a8.w);             \
        b8.x = f(b8.x); b8.y = f(b8.y); b8.z = f(b8.z); b8.w = f(b8.w); r7 = f(r7);
        if (qp.multiply) {
            const unsigned q = (unsigned)qp.q[8];
            auto f = [q](int x) { return (int)((unsigned)x * q); };
            FRI_DQ8
        } else if (qp.pow2 & 0x100u) {
            const int k = qp.pow2_shift[8];
            auto f = [k](int x) { return trunc_div_pow2(x, k); };
            FRI_DQ8
        } else {
            const Div dv = qp.div(8);
            auto f = [dv](int x) { return trunc_div(x, dv); };
            FRI_DQ8
        }
#undef FRI_DQ8
        if (lane == 31) b7.y = r7;
    } else if (QS == kQuantGeneric && ((qp.active >> (top + 6)) & 0xfu)) {
        const int r8 = b8.w, r7 = b7.y, r6 = b6;  // raw values of the level-last nodes
        const int mul = qp.multiply;
#define FRI_DQLEVEL(L, STMTS)                                                       \
        if ((qp.active >> (top + (L))) & 1u) {                                              \
            if (mul) {                                                                      \
                const unsigned q = (unsigned)qp.q[top + (L)];                               \
                auto f = [q](int x) { return (int)((unsigned)x * q); };                     \
                STMTS                                                                       \
            } else if ((qp.pow2 >> (top + (L))) & 1u) {                                     \
                const int k = qp.pow2_shift[top + (L)];                                     \
                auto f = [k](int x) { return trunc_div_pow2(x, k); };                       \
                STMTS                                                                       \
            } else {                                                                        \
                const Div dv = qp.div(top + (L));                                           \
                auto f = [dv](int x) { return trunc_div(x, dv); };                          \
                STMTS                                                                       \
            }                                                                               \
        }
        FRI_DQLEVEL(8, a8.x = f(a8.x); a8.y = f(a8.y); a8.z = f(a8.z); a8.w = f(a8.w);
                       b8.x = f(b8.x); b8.y = f(b8.y); b8.z = f(b8.z); b8.w = f(b8.w);)
        FRI_DQLEVEL(7, a7.x = f(a7.x); a7.y = f(a7.y); b7.x = f(b7.x); b7.y = f(b7.y);)
        FRI_DQLEVEL(6, a6 = f(a6); b6 = f(b6);)
#undef FRI_DQLEVEL
        if (lastB && ((qp.fix >> (top + 6)) & 7u)) {  // only where the next layer's divisor differs
            if ((qp.fix >> (top + 8)) & 1u) b8.w = dequant_layer(qp, r8, top + 9);
            if ((qp.fix >> (top + 7)) & 1u) b7.y = dequant_layer(qp, r7, top + 8);
            if ((qp.fix >> (top + 6)) & 1u) b6 = dequant_layer(qp, r6, top + 7);
        }
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
    const uint32_t p1 = p0 + g.pitch, p2 = p1 + g.pitch;
#if FRI_WHATIF_SKIP_SCATTER
    {
        int x = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) x ^= v[i] ^ w[i];
        store_clamped<S>(p0, x);
        return;
    }
#endif
#define FRI_ST(ptr, dx, val) store_clamped<S>((uint32_t)((int)(ptr) + (dx) * PB), (val))
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

// Dequantization + inverse transform of the tiles of one group into the staged region; mirror image of
// encode_tiles: lane group lane / 8 first unfolds levels 0..5 of its channel (C == 3) or of its tile of a
// batch of up to four (C == 1) into the warp's scratch, then every lane unfolds its two depth-3 subtrees per
// (tile, channel) and scatters the 16 clamped leaves.  The register-level coefficient loads are
// software-pipelined one (tile, channel) ahead.
template <int C, typename S, bool DEEP, int QS, typename CT>
__device__ __forceinline__ void decode_tiles(const Geometry &g, const QuantParams &qp, const GroupDesc &gd, const RegionView &rv,
                                             const uint32_t *__restrict__ tile_unit, int frame, uint8_t *region,
                                             int32_t *scratch, const CT *__restrict__ coefs,
                                             const int32_t *__restrict__ dc_in)
{
    constexpr int SB = (int)sizeof(S);
    constexpr int PB = C * SB;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const int n_present = __popc(gd.tile_mask);
    const uint32_t lane_base = (uint32_t)__cvta_generic_to_shared(region) + rv.phi0 + lane_anchor_bytes(lane, g.pitch, PB);
    const int half = kHalfB.y * g.pitch + kHalfB.x * PB;
    const int sub_bits = DEEP ? g.sub_bits : 0, depth = DEEP ? g.depth : kBaseDepth;
    const int top = sub_bits;
    const bool sparse_group = (gd.tile_mask & (gd.tile_mask + 1u)) != 0;
    const int grp = min(lane >> 3, scratch_units(C) - 1), j8 = lane & 7;
    // The first channel's register-level coefficients of a tile are requested ahead of time: for the
    // warp's first tile before the top levels are unfolded (one round trip to L2/HBM covers both),
    // for every later tile while the previous tile's last channel is being unfolded.
    TaskAddr ta{};
    LaneCoefs cur{};
    if (warp < n_present) {
        ta = task_addr<C, DEEP>(g, tile_unit, frame, gd.tile_base + warp, 0);
        cur = load_lane_coefs(coefs + ta.block, ta.node, lane);
    }
    if (C == 1) {
        for (int e0 = warp; e0 < n_present; e0 += kTileBatch * n_warps) {
            {   // ---- levels 0..5 of up to four tiles, 8 lanes per tile
                const int eg = e0 + grp * n_warps;
                const bool live = eg < n_present;
                const TaskAddr tg = task_addr<C, DEEP>(g, tile_unit, frame, gd.tile_base + (live ? eg : e0), 0);
                const TopCoefs tc = load_top_coefs<DEEP, CT>(coefs, dc_in, tg, 0, j8, depth, sub_bits);
                dec_top_levels<DEEP>(qp, tc, tg, j8, top, sub_bits, live, scratch + grp * kScratchInts + 8 * j8);
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < kTileBatch; ++k) {
                const int e = e0 + k * n_warps;
                if (e >= n_present) break;
                const int slot = sparse_group ? (int)__fns(gd.tile_mask, 0, e + 1) : e;
                const bool has_next = e + n_warps < n_present;
                const TaskAddr tn = has_next ? task_addr<C, DEEP>(g, tile_unit, frame, gd.tile_base + e + n_warps, 0) : ta;
                const LaneCoefs c = cur;
                if (has_next) cur = load_lane_coefs(coefs + tn.block, tn.node, lane);
                const int sA = scratch[k * kScratchInts + lane], sB = scratch[k * kScratchInts + 32 + lane];
                dec_register_levels<S, PB, DEEP, QS>(g, qp, c, sA, sB, lane, ta.last && lane == 31, top,
                                                     lane_base + g.tile_off[slot], half);
                ta = tn;
            }
            __syncwarp();
        }
        return;
    }
    const bool grp_live = (lane >> 3) < C;
    for (int e = warp; e < n_present; e += n_warps) {
        const int slot = sparse_group ? (int)__fns(gd.tile_mask, 0, e + 1) : e;
        const bool has_next = e + n_warps < n_present;
        const TaskAddr tn = has_next ? task_addr<C, DEEP>(g, tile_unit, frame, gd.tile_base + e + n_warps, 0) : ta;
        const bool lastB = ta.last && lane == 31;
        const size_t node = ta.node;

        // ---- levels 0..5 of all channels, 8 lanes per channel
        {
            const TopCoefs tc = load_top_coefs<DEEP, CT>(coefs, dc_in, ta, grp, j8, depth, sub_bits);
            dec_top_levels<DEEP>(qp, tc, ta, j8, top, sub_bits, grp_live, scratch + grp * kScratchInts + 8 * j8);
        }
        __syncwarp();

        const uint32_t t0 = lane_base + g.tile_off[slot];
#pragma unroll
        for (int ch = 0; ch < C; ++ch) {
            const LaneCoefs c = cur;
            if (ch + 1 < C) cur = load_lane_coefs(coefs + ta.block + ((int64_t)(ch + 1) << depth), node, lane);
            else if (has_next) cur = load_lane_coefs(coefs + tn.block, tn.node, lane);
            const int sA = scratch[ch * kScratchInts + lane], sB = scratch[ch * kScratchInts + 32 + lane];
            dec_register_levels<S, PB, DEEP, QS>(g, qp, c, sA, sB, lane, lastB, top, t0 + ch * SB, half);
        }
        __syncwarp();
        ta = tn;
    }
}

// Zero-fills a region buffer (a lattice tile the reference's BFS never built — possible only
// next to the image border, e.g. 480x270 — still owns its pixels in the chunk masks: staging
// zeros for it reproduces from_wavelet's zero-initialised raster, wavelet_transform.rs:309-317).
__device__ __forceinline__ void zero_region(const Geometry &g, uint8_t *region)
{
    const int n16 = (g.region_h * g.pitch + 15) >> 4;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) reinterpret_cast<int4 *>(region)[i] = make_int4(0, 0, 0, 0);
}

// Write-out of a decoded region in 16-byte chunks aligned in global memory, one chunk per thread
// and iteration from the plan's chunk list.  Only bytes of pixels that belong to the group's
// tiles (chunk masks) and lie inside the image (set_pixel's bounds check, images.rs:104) are
// written: fully owned chunks as one 128-bit store, the chunks along the group's fractal outline
// byte-masked.  The first kWriteAhead list entries of every thread are fetched by
// write_out_preload() *before* the CTA barrier that completes the region, so their latency hides
// behind the barrier wait.
constexpr int kWriteAhead = 6;
// Fully owned chunks leave with a streaming (evict-first) store: decoded pixels are never re-read by the
// kernel and should not push the prefetched coefficient runs out of L2 (+1.7 % on 16 x 4K frames per launch;
// cache-global and write-through stores measured equal to the default).
#define FRI_PIXEL_STORE(ptr, val) __stcs((ptr), (val))

// Partially owned chunks (the group's fractal outline; 289 of a 4 x 4 RGB group's 1 682 chunks): an interior
// group writes them from the plan's edge list — first the fully owned 32-bit words (one LDS.32 + STG.32 each),
// then the owned samples of the words the outline passes through (one sample-sized load and store each) — so the
// loop has no per-word mask test and no divergence; the round's earlier form walked chunk masks word by word
// (skip / whole word / bytewise) and cost three times the instructions.  Groups that touch the image border keep
// the masked chunk form, which also clips.
constexpr int kWordAhead = 2;  // edge-list entries per thread fetched before the barrier: words ...
constexpr int kSampAhead = 4;  // ... and samples

struct WriteAhead {
    uint32_t e[kWriteAhead];   // fully owned chunks
    uint32_t w[kWordAhead];    // fully owned words of the partially owned chunks
    uint32_t s[kSampAhead];    // owned samples of the remaining words
};

__device__ __forceinline__ WriteAhead write_out_preload(const Geometry &g, const RegionView &rv,
                                                        const uint32_t *__restrict__ chunk_list,
                                                        const uint32_t *__restrict__ edge_list, uint64_t pol)
{
    WriteAhead w;
    const uint32_t *cl = chunk_list + (size_t)rv.phi0 * g.list_cap;
    const int n_full = g.list_full[rv.phi0];
    // Slots past the end of the list repeat its last chunk (a redundant store of the same bytes)
    // instead of being skipped: the unrolled stores then need no per-chunk test and branch.
    // (Skipping whole unused rounds with a CTA-uniform test — a 1-byte-per-pixel group has 883 full chunks
    // for 6 x 256 slots — measured equal for 1-channel images and cost the RGB kernel a spill.)
#pragma unroll
    for (int u = 0; u < kWriteAhead; ++u) {
        const int k = min((int)(threadIdx.x + u * blockDim.x), n_full - 1);
        w.e[u] = (rv.interior && n_full > 0) ? ld_table(cl + k, pol) : kNoChunk;
    }
    const uint32_t *el = edge_list + (size_t)rv.phi0 * g.edge_cap;
    const int n_words = g.edge_words[rv.phi0], n_samples = g.edge_samples[rv.phi0];
#pragma unroll
    for (int u = 0; u < kWordAhead; ++u) {
        const int k = (int)(threadIdx.x + u * blockDim.x);
        w.w[u] = (rv.interior && k < n_words) ? ld_table(el + k, pol) : kNoChunk;
    }
#pragma unroll
    for (int u = 0; u < kSampAhead; ++u) {
        const int k = (int)(threadIdx.x + u * blockDim.x);
        w.s[u] = (rv.interior && k < n_samples) ? ld_table(el + n_words + k, pol) : kNoChunk;
    }
    return w;
}

// One entry of the edge list: row << 20 | shared-memory byte offset.
template <typename T>
__device__ __forceinline__ void store_edge(const RegionView &rv, const uint8_t *region, uint32_t e)
{
    const int r = (int)(e >> 20), s = (int)(e & 0xfffffu);
    *reinterpret_cast<T *>(rv.gaddr(r, s)) = *reinterpret_cast<const T *>(region + s);
}

template <typename S>
__device__ __forceinline__ void write_out_group(const Geometry &g, const GroupDesc &gd, const RegionView &rv,
                                                const uint32_t *__restrict__ chunk_list,
                                                const uint16_t *__restrict__ chunk_mask,
                                                const uint32_t *__restrict__ edge_list, const uint8_t *region,
                                                const WriteAhead &ahead, uint64_t pol)
{
    const int n_threads = blockDim.x;
    const uint32_t *cl = chunk_list + (size_t)rv.phi0 * g.list_cap;
    const int n_full = g.list_full[rv.phi0], n_all = g.list_all[rv.phi0];
    if (rv.interior) {
        if (n_full > 0 && !FRI_WHATIF_SKIP_FULL) {
#pragma unroll
            for (int u = 0; u < kWriteAhead; ++u) {
                const int r = (int)(ahead.e[u] >> 16), s = (int)(ahead.e[u] & 0xffffu) << 4;
                FRI_PIXEL_STORE(reinterpret_cast<int4 *>(rv.gaddr(r, s)), *reinterpret_cast<const int4 *>(region + s));
            }
        }
#if FRI_TRACE
        if ((threadIdx.x == 0 || threadIdx.x == 255) && blockIdx.x < 16384) g_trace2[4 * blockIdx.x + (threadIdx.x == 0 ? 0 : 1)] = gtime();
#endif
#pragma unroll 1
        for (int k = threadIdx.x + kWriteAhead * n_threads; k < (FRI_WHATIF_SKIP_FULL ? 0 : n_full); k += n_threads) {
            const uint32_t e = ld_table(cl + k, pol);
            const int r = (int)(e >> 16), s = (int)(e & 0xffffu) << 4;
            FRI_PIXEL_STORE(reinterpret_cast<int4 *>(rv.gaddr(r, s)), *reinterpret_cast<const int4 *>(region + s));
        }
#if !FRI_WHATIF_SKIP_MIXED
        const uint32_t *el = edge_list + (size_t)rv.phi0 * g.edge_cap;
        const int n_words = g.edge_words[rv.phi0], n_samples = g.edge_samples[rv.phi0];
#pragma unroll
        for (int u = 0; u < kWordAhead; ++u)
            if (ahead.w[u] != kNoChunk) store_edge<uint32_t>(rv, region, ahead.w[u]);
#pragma unroll 1
        for (int k = threadIdx.x + kWordAhead * n_threads; k < n_words; k += n_threads) store_edge<uint32_t>(rv, region, ld_table(el + k, pol));
#pragma unroll
        for (int u = 0; u < kSampAhead; ++u)
            if (ahead.s[u] != kNoChunk) store_edge<S>(rv, region, ahead.s[u]);
#pragma unroll 1
        for (int k = threadIdx.x + kSampAhead * n_threads; k < n_samples; k += n_threads)
            store_edge<S>(rv, region, ld_table(el + n_words + k, pol));
#endif
#if FRI_TRACE
        if ((threadIdx.x == 0 || threadIdx.x == 255) && blockIdx.x < 16384) g_trace2[4 * blockIdx.x + (threadIdx.x == 0 ? 2 : 3)] = gtime();
#endif
    } else {
        const uint16_t *cmk = chunk_mask + (size_t)rv.phi0 * g.list_cap;
        const int stride32 = (int)g.row_stride;
#pragma unroll 1
        for (int k = threadIdx.x; k < n_all; k += n_threads) {
            const uint32_t e = ld_table(cl + k, pol);
            const int r = (int)(e >> 16), s = (int)(e & 0xffffu) << 4;
            if ((unsigned)(gd.y0 + r) >= (unsigned)g.height) continue;
            uint32_t m = k < n_full ? 0xffffu : ld_table(cmk + k, pol);
            const int xb = rv.xb0 + s - (r * g.pitch + rv.phi0);  // byte position of the chunk inside its image row
            const int lo = min(max(-xb, 0), 16), hi = min(max(stride32 - xb, 0), 16);
            m &= ((1u << hi) - 1u) & ~((1u << lo) - 1u);
            if (m) store_chunk_masked(rv.gaddr(r, s), region + s, m);
        }
    }
}

__device__ __forceinline__ size_t region_bytes(const Geometry &g) { return ((size_t)g.region_h * g.pitch + 15) & ~(size_t)15; }

// ------------------------------------------------------------------------------------------
// kernels: one CTA per (group, frame)
// ------------------------------------------------------------------------------------------
#ifndef FRI_TRACE
#define FRI_TRACE 0
#endif
#ifndef FRI_ENC_MINB
#define FRI_ENC_MINB 4
#endif
#ifndef FRI_DEC_MINB
#define FRI_DEC_MINB 4
#endif
template <int C, typename S, bool DEEP, int QS, typename CT>
__global__ void __launch_bounds__(kThreads, FRI_ENC_MINB)
fri_encode_kernel(const __grid_constant__ Geometry g, const __grid_constant__ QuantParams qp,
                  const GroupDesc *__restrict__ groups, const uint32_t *__restrict__ tile_unit,
                  const uint32_t *__restrict__ stage_list, const uint8_t *__restrict__ pixels,
                  CT *__restrict__ coefs, int32_t *__restrict__ dc_out, int lookahead, int group_offset, int bulk_staging,
                  int late_lookahead_ctas)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t *region = smem;
    int32_t *scratch = reinterpret_cast<int32_t *>(smem + region_bytes(g)) + (threadIdx.x >> 5) * (scratch_units(C) * kScratchInts);
    const uint64_t pol = table_policy();
    const GroupDesc gd = ld_group(groups + group_offset + blockIdx.x, pol);  // the grid covers groups [group_offset, + gridDim.x)
    pdl_launch_dependents();
    const int frame = blockIdx.y;
    const RegionView rv = region_view<C * (int)sizeof(S)>(g, gd, pixels + (int64_t)frame * g.frame_bytes);
    FRI_TRACE_MARK(0);
    // Two-stage copy: the chunks needed by every warp's first tile are committed first, so that
    // the first round of tiles runs while the rest of the footprint is still in flight.
    const int n_warps = blockDim.x >> 5;
    const int n_first = g.stage_first[rv.phi0], n_all = g.list_all[rv.phi0];
    // (full groups only: every warp then has a tile in both rounds and reaches the barrier between them)
    const bool two_stage = __popc(gd.tile_mask) == g.group_a * g.group_b && g.group_a * g.group_b == 2 * n_warps;
    // Interior groups are staged with one bulk copy (TMA) per row; groups that touch the image border
    // keep the chunk-list path, which clips and zero-fills.
    __shared__ __align__(8) uint64_t bars[2];
    const bool bulk = bulk_staging && rv.interior && g.n_rows_first > 0 && g.region_h <= (int)blockDim.x;
    if (bulk && threadIdx.x == 0) {
        mbar_init(&bars[0], two_stage ? g.n_rows_first : g.region_h);
        mbar_init(&bars[1], max(1, g.region_h - g.n_rows_first));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // the previous kernel of the stream may have written these pixels (or still read the coefficients) — unless
    // the caller has promised that consecutive calls are independent (fri_plan_set_independent_calls)
    if (DEEP || !g.independent_calls) pdl_wait();
    if (bulk) {
        // only the warps that issue copies (thread r copies row r) wait for the barrier initialisation; the
        // others go straight on to the look-ahead and meet them at the CTA barrier before the wait
        const int issuing = (g.region_h + 31) & ~31;
        if ((int)threadIdx.x < issuing) {
            asm volatile("bar.sync 1, %0;" ::"r"(issuing) : "memory");
            stage_rows_bulk(g, rv, region, bars, two_stage);
        }
    } else {
        stage_group(g, gd, rv, stage_list, region, 0, two_stage ? n_first : n_all, pol);
        cp_async_commit();
        if (two_stage) {
            stage_group(g, gd, rv, stage_list, region, n_first, n_all, pol);
            cp_async_commit();
            cp_async_wait_but_one();
        } else {
            cp_async_wait_all();
        }
    }
    auto look_ahead = [&]() {
    // Look-ahead: while this group's copies are in flight, pull the pixel rows of the group that
        // will run in this CTA slot one residency later towards L2, one 128-byte line per thread and
        // iteration, so that its copies are served by L2 instead of waiting in the DRAM queues.
        if (lookahead > 0) {
            const int64_t target = (int64_t)frame * gridDim.x + blockIdx.x + lookahead;  // linear CTA index in this launch
            if (target < (int64_t)gridDim.y * gridDim.x) {
                const int tf = (int)(target / gridDim.x);
                const GroupDesc tg = ld_group(groups + group_offset + (target - (int64_t)tf * gridDim.x), pol);
                const int lines_per_row = (g.row_bytes + 127) / 128 + 1;
                const int y_lo = max(tg.y0, 0), y_hi = min(tg.y0 + g.region_h, g.height);
                const int64_t xb = min(max((int64_t)tg.x0 * (C * (int)sizeof(S)), (int64_t)0), g.row_stride - 1);
                const char *base = reinterpret_cast<const char *>(pixels) + (int64_t)tf * g.frame_bytes + xb;
                const int total = (y_hi - y_lo) * lines_per_row;
                for (int i = threadIdx.x; i < total; i += blockDim.x) {
                    const int r = i / lines_per_row, c = i - r * lines_per_row;
                    if (xb + (int64_t)c * 128 < g.row_stride)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (int64_t)(y_lo + r) * g.row_stride + (int64_t)c * 128));
                }
            }
        }
    };
    // first-wave CTAs (all start together, all waiting on HBM) look ahead only once their own pixels are in
    const bool late_look = late_lookahead_ctas > 0 && (int64_t)frame * gridDim.x + blockIdx.x < late_lookahead_ctas;
    if (!late_look) look_ahead();
    __syncthreads();
    if (bulk) mbar_wait(&bars[0], 0);
    if (late_look) look_ahead();
    FRI_TRACE_MARK(1);
    encode_tiles<C, S, DEEP, QS, CT>(g, qp, gd, rv, tile_unit, frame, region, scratch, coefs, dc_out, two_stage,
                                     bulk ? bars : nullptr);
#if FRI_TRACE
    __syncthreads();
#endif
    FRI_TRACE_MARK(2);
}

template <int C, typename S, bool DEEP, int QS, typename CT>
__global__ void __launch_bounds__(kThreads, FRI_DEC_MINB)
fri_decode_kernel(const __grid_constant__ Geometry g, const __grid_constant__ QuantParams qp,
                  const GroupDesc *__restrict__ groups, const uint32_t *__restrict__ tile_unit,
                  const uint32_t *__restrict__ chunk_list, const uint16_t *__restrict__ chunk_mask,
                  const uint32_t *__restrict__ edge_list,
                  const CT *__restrict__ coefs, const int32_t *__restrict__ dc_in, uint8_t *__restrict__ pixels,
                  int group_offset, int no_prefetch_ctas)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t *region = smem;
    int32_t *scratch = reinterpret_cast<int32_t *>(smem + region_bytes(g)) + (threadIdx.x >> 5) * (scratch_units(C) * kScratchInts);
    const uint64_t pol = table_policy();
    const GroupDesc gd = ld_group(groups + group_offset + blockIdx.x, pol);
    pdl_launch_dependents();
    const int frame = blockIdx.y;
    const RegionView rv = region_view<C * (int)sizeof(S)>(g, gd, pixels + (int64_t)frame * g.frame_bytes);
    FRI_TRACE_MARK(0);
    const bool sparse = __popc(gd.tile_mask) != g.group_a * g.group_b;
    if (sparse) zero_region(g, region);
    // the previous kernel of the stream may have produced these coefficients (or still read the pixels)
    if (DEEP || !g.independent_calls) pdl_wait();
    if ((int64_t)blockIdx.y * gridDim.x + blockIdx.x >= no_prefetch_ctas) prefetch_group_coefs<C, DEEP, CT>(g, gd, frame, coefs, tile_unit);
    if (sparse) __syncthreads();
    decode_tiles<C, S, DEEP, QS, CT>(g, qp, gd, rv, tile_unit, frame, region, scratch, coefs, dc_in);
    const WriteAhead ahead = write_out_preload(g, rv, chunk_list, edge_list, pol);
    __syncthreads();
    FRI_TRACE_MARK(1);
    write_out_group<S>(g, gd, rv, chunk_list, chunk_mask, edge_list, region, ahead, pol);
#if FRI_TRACE
    __syncthreads();
#endif
    FRI_TRACE_MARK(2);
}

// ------------------------------------------------------------------------------------------
// depth > 9: base tiles of a retained fractal that lie entirely outside the image hold only `None`
// coefficients.  The transform kernels skip them (they are not in any group); the encoder writes the zeros the
// dense array holds for `None` with this pure store kernel: one warp per (absent base tile, channel), the nine
// level runs of the tile inside the fractal's heap.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fri_zero_absent_kernel(const uint32_t *__restrict__ absent_unit, uint32_t n_absent, int channels, int n_fractals, int sub_bits,
                       int depth, int32_t *__restrict__ coefs)
{
    const int lane = threadIdx.x & 31;
    const size_t task = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (task >= (size_t)n_absent * channels) return;
    const uint32_t u = __ldg(absent_unit + task / channels);
    const int ch = (int)(task % channels), frame = blockIdx.y;
    const uint32_t f = u >> sub_bits, node = (1u << sub_bits) + (u & ((1u << sub_bits) - 1u));
    int32_t *out = coefs + ((((int64_t)frame * n_fractals + f) * channels + ch) << depth);
    const int4 z4 = make_int4(0, 0, 0, 0);
    int4 *o8 = reinterpret_cast<int4 *>(out + ((size_t)node << 8));
    __stcs(o8 + lane, z4);
    __stcs(o8 + 32 + lane, z4);
    __stcs(reinterpret_cast<int4 *>(out + ((size_t)node << 7)) + lane, z4);
    __stcs(reinterpret_cast<int2 *>(out + ((size_t)node << 6)) + lane, make_int2(0, 0));
    __stcs(out + ((size_t)node << 5) + lane, 0);
#pragma unroll
    for (int L = 4; L >= 0; --L)
        if (lane < (1 << L)) __stcs(out + ((size_t)node << L) + lane, 0);
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
            out[pos] = quant_layer_enc(qp, d, layer_of(pos));
        }
        __syncthreads();
        int32_t *t = src; src = dst; dst = t;
    }
    if (threadIdx.x == 0) out[0] = quant_layer_enc(qp, src[0], 0);  // wavelet_transform.rs:221, layer 0
}

__global__ void __launch_bounds__(kCoarseThreads)
fri_coarse_inverse_kernel(const __grid_constant__ QuantParams qp, int sub_bits, int depth,
                          const int32_t *__restrict__ coefs, int32_t *__restrict__ dc)
{
    extern __shared__ __align__(16) int32_t cs[];
    const int n = 1 << sub_bits;
    // Level L reads 2^L values and writes 2^(L+1).  Two buffers, X = cs[0, n) and Y = cs[n, n + n/2): the
    // last level (sub_bits - 1) writes n values and must land in X, the one before n/2 values in Y, and so
    // on alternating — so level 0 writes into X when sub_bits is odd, into Y when it is even.
    int32_t *dst = (sub_bits & 1) ? cs : cs + n;
    int32_t *src = (sub_bits & 1) ? cs + n : cs;
    const int32_t *in = coefs + ((int64_t)blockIdx.x << depth);
    int32_t *out = dc + ((int64_t)blockIdx.x << sub_bits);
    if (threadIdx.x == 0) src[0] = dequant_layer(qp, in[0], 0);
    __syncthreads();
    for (int level = 0; level < sub_bits; ++level) {
        const int cnt = 1 << level;
        for (int j = threadIdx.x; j < cnt; j += kCoarseThreads) {
            const uint32_t pos = (uint32_t)(cnt + j);
            const int d = dequant_layer(qp, in[pos], layer_of(pos));
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

// ------------------------------------------------------------------------------------------
// emission order: gather the quantized coefficients of every channel
// into the order the reference's entropy coder consumes them, `None` slots dropped.
//
// Consecutive emitted coefficients come from different tiles (the scan walks one tree level across
// the whole image), so a flat gather reads one 32-byte sector per 4-byte coefficient.  Instead one
// CTA takes a group of lattice-adjacent tiles (the same groups the transform kernels use): the
// group's 2 KB coefficient blocks are read coalesced into shared memory, and its emission slots —
// which the plan lists per group in increasing emission index, so that the nodes a scan row picks
// up inside the group are adjacent — are written as runs.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
fri_emit_kernel(const GroupDesc *__restrict__ groups, const uint32_t *__restrict__ goff, const uint32_t *__restrict__ dst,
                const uint16_t *__restrict__ loc, unsigned long long stride, int channels, int n_tiles,
                const int32_t *__restrict__ coefs, T *__restrict__ out)
{
    extern __shared__ __align__(16) int32_t es[];
    const GroupDesc gd = groups[blockIdx.x];
    const int n_present = __popc(gd.tile_mask), frame = blockIdx.y;
    const uint32_t k0 = goff[blockIdx.x], k1 = goff[blockIdx.x + 1];
    for (int ch = 0; ch < channels; ++ch) {
        // tile t of the group, channel ch: 512 coefficients at ((frame * n_tiles + tile_base + t) * C + ch) << 9
        for (int idx = threadIdx.x; idx < n_present * (kTileLeaves / 4); idx += blockDim.x) {
            const int t = idx >> 7, v = idx & 127;
            const int4 *src = reinterpret_cast<const int4 *>(coefs + ((((size_t)frame * n_tiles + gd.tile_base + t) * channels + ch) << kBaseDepth));
            reinterpret_cast<int4 *>(es)[idx] = __ldcs(src + v);
        }
        __syncthreads();
        T *o = out + ((size_t)frame * channels + ch) * stride;  // stride: elements between the starts of two streams
        for (uint32_t k = k0 + threadIdx.x; k < k1; k += blockDim.x) o[__ldg(dst + k)] = (T)es[__ldg(loc + k)];
        __syncthreads();
    }
}

// The inverse gather (decoder side): emitted streams -> dense coefficient blocks, `None` slots 0.
// Same grouping: the group's emission slots are read as runs, placed in shared memory, and the 2 KB
// blocks are written coalesced.
template <typename T>
__global__ void __launch_bounds__(256)
fri_unemit_kernel(const GroupDesc *__restrict__ groups, const uint32_t *__restrict__ goff, const uint32_t *__restrict__ dst,
                  const uint16_t *__restrict__ loc, unsigned long long stride, int channels, int n_tiles,
                  const T *__restrict__ in, int32_t *__restrict__ coefs)
{
    extern __shared__ __align__(16) int32_t es[];
    const GroupDesc gd = groups[blockIdx.x];
    const int n_present = __popc(gd.tile_mask), frame = blockIdx.y;
    const uint32_t k0 = goff[blockIdx.x], k1 = goff[blockIdx.x + 1];
    const bool all_some = k1 - k0 == (uint32_t)n_present * kTileLeaves;  // no None slot in this group
    for (int ch = 0; ch < channels; ++ch) {
        if (!all_some) {
            for (int idx = threadIdx.x; idx < n_present * (kTileLeaves / 4); idx += blockDim.x)
                reinterpret_cast<int4 *>(es)[idx] = make_int4(0, 0, 0, 0);
            __syncthreads();
        }
        const T *src = in + ((size_t)frame * channels + ch) * stride;
        uint32_t k = k0 + threadIdx.x;
        for (; k + 3 * blockDim.x < k1; k += 4 * blockDim.x) {  // four independent gathers in flight per thread
            uint32_t l[4];
            int32_t v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                l[u] = __ldg(loc + k + u * blockDim.x);
                v[u] = (int32_t)__ldcs(src + __ldg(dst + k + u * blockDim.x));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) es[l[u]] = v[u];
        }
        for (; k < k1; k += blockDim.x) es[__ldg(loc + k)] = (int32_t)__ldcs(src + __ldg(dst + k));
        __syncthreads();
        for (int idx = threadIdx.x; idx < n_present * (kTileLeaves / 4); idx += blockDim.x) {
            const int t = idx >> 7, v = idx & 127;
            int4 *out = reinterpret_cast<int4 *>(coefs + ((((size_t)frame * n_tiles + gd.tile_base + t) * channels + ch) << kBaseDepth));
            __stcs(out + v, reinterpret_cast<const int4 *>(es)[idx]);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// 16-bit transport (host-buffer entry points fri_*_tq16): every coefficient an 8-bit image can
// produce fits an i16 (|d| <= 255, 0 <= s <= 255), so the copies over PCIe carry half the bytes.
// Pure streaming repack of the device-resident i32 array; eight coefficients per thread and
// iteration, saturating so that an out-of-range value can never wrap silently.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_sat16(int lo, int hi)
{
    uint32_t r;
    asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(r) : "r"(hi), "r"(lo));  // {sat(hi), sat(lo)}
    return r;
}

__global__ void __launch_bounds__(256)
fri_pack16_kernel(const int4 *__restrict__ src, int4 *__restrict__ dst, size_t n8)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        const int4 a = __ldcs(src + 2 * i), b = __ldcs(src + 2 * i + 1);
        __stcs(dst + i, make_int4((int)pack_sat16(a.x, a.y), (int)pack_sat16(a.z, a.w), (int)pack_sat16(b.x, b.y),
                                  (int)pack_sat16(b.z, b.w)));
    }
}

__global__ void __launch_bounds__(256)
fri_unpack16_kernel(const int4 *__restrict__ src, int4 *__restrict__ dst, size_t n8)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        const int4 v = __ldcs(src + i);
        __stcs(dst + 2 * i, make_int4((int)(short)v.x, v.x >> 16, (int)(short)v.y, v.y >> 16));
        __stcs(dst + 2 * i + 1, make_int4((int)(short)v.z, v.z >> 16, (int)(short)v.w, v.w >> 16));
    }
}

// ------------------------------------------------------------------------------------------
// 10-bit packed transport of the emission-ordered streams (host entry points fri_*_tq_emit10): a
// coefficient travels as the symbol the reference's entropy coder would see, pack_signed(k) = 2k for
// k >= 0, -2k - 1 for k < 0 (utils.rs:34-40), in the 1024-symbol alphabet (entropy_coding.rs:25) — every
// coefficient of an 8-bit image fits (|k| <= 255) and so does every value a decodable container can hold.
// Four symbols -> 40 bits -> 5 bytes, little-endian (symbol i of a block in bits [10 i, 10 i + 10)):
// 1.25 bytes per coefficient over PCIe instead of 2.  One thread packs 64 symbols (128 B in, 80 B out);
// streams are padded to a multiple of 64 symbols on both sides.  Encode saturates at +-511 / -512.
// A 9-bit form of the same layout (72 B per 64 symbols, saturating at -256 / +255) carries everything the
// transform of an 8-bit image can produce (|k| <= 255) in 1.125 bytes per coefficient.
// ------------------------------------------------------------------------------------------
template <int BITS>
__device__ __forceinline__ uint32_t zigzag_sat(int v)
{
    v = max(-(1 << (BITS - 1)), min((1 << (BITS - 1)) - 1, v));
    return (uint32_t)((v << 1) ^ (v >> 31));  // 2v for v >= 0, -2v - 1 for v < 0
}
__device__ __forceinline__ int unzigzag(uint32_t s) { return (int)(s >> 1) ^ -(int)(s & 1u); }  // utils.rs:42-48

// One thread packs 64 symbols of BITS bits (128 B in) into 8 * BITS bytes (symbol i of a block in bits
// [BITS i, BITS i + BITS), little-endian); all shifts are compile-time after unrolling.
template <int BITS>
__global__ void __launch_bounds__(256)
fri_pack_kernel(const int4 *__restrict__ src, int2 *__restrict__ dst, size_t n_blocks)
{
    for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < n_blocks; b += (size_t)gridDim.x * blockDim.x) {
        uint32_t w[2 * BITS];
        uint64_t acc = 0;
        int nb = 0, wi = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int4 v = __ldcs(src + 8 * b + j);
            const int x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int e = (t & 1) ? (x[t >> 1] >> 16) : (int)(short)x[t >> 1];
                acc |= (uint64_t)zigzag_sat<BITS>(e) << nb;
                nb += BITS;
                if (nb >= 32) {
                    w[wi++] = (uint32_t)acc;
                    acc >>= 32;
                    nb -= 32;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < BITS; ++k) __stcs(dst + (size_t)BITS * b + k, make_int2((int)w[2 * k], (int)w[2 * k + 1]));
    }
}

template <int BITS>
__global__ void __launch_bounds__(256)
fri_unpack_kernel(const int2 *__restrict__ src, int4 *__restrict__ dst, size_t n_blocks)
{
    for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < n_blocks; b += (size_t)gridDim.x * blockDim.x) {
        uint32_t w[2 * BITS + 1];
#pragma unroll
        for (int k = 0; k < BITS; ++k) {
            const int2 v = __ldcs(src + (size_t)BITS * b + k);
            w[2 * k] = (uint32_t)v.x;
            w[2 * k + 1] = (uint32_t)v.y;
        }
        w[2 * BITS] = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int out[4];
#pragma unroll
            for (int h = 0; h < 4; ++h) {  // symbols 8 j + 2 h, 8 j + 2 h + 1
                int v[2];
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const int bit = BITS * (8 * j + 2 * h + t), wi = bit >> 5, sh = bit & 31;
                    const uint32_t sym = (uint32_t)(((uint64_t)w[wi] | (uint64_t)w[wi + 1] << 32) >> sh) & ((1u << BITS) - 1u);
                    v[t] = unzigzag(sym);
                }
                out[h] = (v[0] & 0xffff) | (v[1] << 16);
            }
            __stcs(dst + 8 * b + j, make_int4(out[0], out[1], out[2], out[3]));
        }
    }
}

// Launch with the programmatic-stream-serialization attribute (see pdl_wait): FRI_PDL=0 turns it off.
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, int threads, size_t smem, cudaStream_t stream, Args &&...args)
{
    static const bool enabled = [] {
        const char *env = std::getenv("FRI_PDL");
        return !env || std::atoi(env) != 0;
    }();
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = enabled ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

template <typename K>
cudaError_t set_smem(K kernel, size_t bytes)
{
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

constexpr size_t kMaxSmem = 226 * 1024;  // dynamic part; leaves room for the kernels' few static bytes

}  // namespace

cudaError_t configure_kernels()
{
    cudaError_t e;
#define FRI_CFG(k) if ((e = set_smem(k, kMaxSmem)) != cudaSuccess) return e
#define FRI_CFG4(name)                                               \
    FRI_CFG((name<1, uint8_t, false, kQuantGeneric, int32_t>));      \
    FRI_CFG((name<3, uint8_t, false, kQuantGeneric, int32_t>));      \
    FRI_CFG((name<1, uint16_t, false, kQuantGeneric, int32_t>));     \
    FRI_CFG((name<3, uint16_t, false, kQuantGeneric, int32_t>));     \
    FRI_CFG((name<1, uint8_t, false, kQuantNone, int32_t>));         \
    FRI_CFG((name<3, uint8_t, false, kQuantNone, int32_t>));         \
    FRI_CFG((name<1, uint16_t, false, kQuantNone, int32_t>));        \
    FRI_CFG((name<3, uint16_t, false, kQuantNone, int32_t>));        \
    FRI_CFG((name<1, uint8_t, false, kQuantSmallest, int32_t>));     \
    FRI_CFG((name<3, uint8_t, false, kQuantSmallest, int32_t>));     \
    FRI_CFG((name<1, uint16_t, false, kQuantSmallest, int32_t>));    \
    FRI_CFG((name<3, uint16_t, false, kQuantSmallest, int32_t>));    \
    FRI_CFG((name<1, uint8_t, true, kQuantNone, int32_t>));          \
    FRI_CFG((name<3, uint8_t, true, kQuantNone, int32_t>));          \
    FRI_CFG((name<1, uint16_t, true, kQuantNone, int32_t>));         \
    FRI_CFG((name<3, uint16_t, true, kQuantNone, int32_t>));         \
    FRI_CFG((name<1, uint8_t, true, kQuantGeneric, int32_t>));       \
    FRI_CFG((name<3, uint8_t, true, kQuantGeneric, int32_t>));       \
    FRI_CFG((name<1, uint16_t, true, kQuantGeneric, int32_t>));      \
    FRI_CFG((name<3, uint16_t, true, kQuantGeneric, int32_t>));      \
    FRI_CFG((name<1, uint8_t, false, kQuantGeneric, int16_t>));      \
    FRI_CFG((name<3, uint8_t, false, kQuantGeneric, int16_t>));      \
    FRI_CFG((name<1, uint8_t, false, kQuantNone, int16_t>));         \
    FRI_CFG((name<3, uint8_t, false, kQuantNone, int16_t>));         \
    FRI_CFG((name<1, uint8_t, false, kQuantSmallest, int16_t>));     \
    FRI_CFG((name<3, uint8_t, false, kQuantSmallest, int16_t>))
    FRI_CFG4(fri_encode_kernel);
    FRI_CFG4(fri_decode_kernel);
    FRI_CFG(fri_coarse_forward_kernel);
    FRI_CFG(fri_coarse_inverse_kernel);
    FRI_CFG(fri_emit_kernel<int32_t>);
    FRI_CFG(fri_emit_kernel<int16_t>);
    FRI_CFG(fri_unemit_kernel<int32_t>);
    FRI_CFG(fri_unemit_kernel<int16_t>);
#undef FRI_CFG4
#undef FRI_CFG
    return cudaSuccess;
}

namespace {

// CTAs of the main kernels resident on the device at once.
int resident_ctas(const Geometry &g)
{
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    const int threads = cta_threads(g);
    const size_t smem = kernel_smem_bytes(g) + 1024;
    int per_sm = (int)std::min<size_t>((size_t)(228 * 1024) / smem, (size_t)(65536 / (64 * threads)));
    per_sm = std::max(1, std::min(per_sm, 2048 / threads));
    return sms * per_sm;
}

}  // namespace

cudaError_t launch_encode(const Geometry &g, const DeviceTables &t, const QuantParams &qp, const void *d_pixels,
                          uint32_t n_frames, void *d_coefs_any, bool half, int32_t *d_dc, cudaStream_t stream,
                          uint32_t *launches, int group_begin, int group_end)
{
    if (half && (g.sub_bits != 0 || g.sample_bytes != 1)) return cudaErrorInvalidValue;  // int16 arrays: depth 9, 8-bit samples
    int32_t *d_coefs = static_cast<int32_t *>(d_coefs_any);
    int16_t *d_coefs16 = static_cast<int16_t *>(d_coefs_any);
    if (group_end < 0) group_end = g.n_groups;
    const int n_groups = group_end - group_begin;
    const bool whole = group_begin == 0 && group_end == g.n_groups;
    if (n_frames == 0 || n_groups <= 0) return cudaSuccess;
    const size_t smem = kernel_smem_bytes(g);
    const int qclass = quant_class(qp, g);
    int lookahead = resident_ctas(g);  // prefetch distance: the group that will reuse this CTA's slot
    if (const char *env = std::getenv("FRI_LOOKAHEAD")) lookahead = std::atoi(env);  // tuning knob
    if (!whole) lookahead = 0;  // banded host pipeline: rows of later groups may not be on the device yet
    // Bulk-copy (TMA) staging of interior groups; FRI_STAGE_BULK=0 falls back to the chunk-list cp.async path
    // The CTAs of the first wave (they start together and all wait on HBM) issue their look-ahead only after
    // their own pixels have landed: +2.8 % on a single frame.  Tuning knob: FRI_ENC_LATE_LOOKAHEAD_CTAS.
    int late_look_ctas = resident_ctas(g);
    if (const char *env = std::getenv("FRI_ENC_LATE_LOOKAHEAD_CTAS")) late_look_ctas = std::atoi(env);
    int bulk_staging = 1;
    if (const char *env = std::getenv("FRI_STAGE_BULK")) bulk_staging = std::atoi(env) != 0;  // tuning knob
    const GroupDesc *gtab = whole && t.groups_launch ? t.groups_launch : t.groups;  // launch order (whole frames only)
    if (g.sub_bits > 0 && t.n_absent > 0 && group_begin == 0) {
        // base tiles outside the image: their low-pass roots are `None` (0 for the coarse kernel), their slots 0
        const size_t dc_bytes = (((size_t)n_frames * g.n_fractals * g.channels) << g.sub_bits) * sizeof(int32_t);
        cudaError_t e = cudaMemsetAsync(d_dc, 0, dc_bytes, stream);
        if (e != cudaSuccess) return e;
        const unsigned tasks = t.n_absent * (unsigned)g.channels;
        for (uint32_t f0 = 0; f0 < n_frames; f0 += 65535u) {
            const uint32_t nf = n_frames - f0 < 65535u ? n_frames - f0 : 65535u;
            fri_zero_absent_kernel<<<dim3((tasks + 7) / 8, nf), 256, 0, stream>>>(t.absent_unit, t.n_absent, g.channels, g.n_fractals,
                                                                                 g.sub_bits, g.depth,
                                                                                 d_coefs + (int64_t)f0 * g.coefs_per_frame);
            if (launches) ++*launches;
        }
    }
    const uint8_t *px = static_cast<const uint8_t *>(d_pixels);
    for (uint32_t f0 = 0; f0 < n_frames; f0 += 65535u) {  // gridDim.y limit
        const uint32_t nf = n_frames - f0 < 65535u ? n_frames - f0 : 65535u;
        const dim3 grid((unsigned)n_groups, nf);
        const uint8_t *p = px + (int64_t)f0 * g.frame_bytes;
        int32_t *c = d_coefs + (int64_t)f0 * g.coefs_per_frame;
        int16_t *c16 = d_coefs16 + (int64_t)f0 * g.coefs_per_frame;
        int32_t *dc = d_dc ? d_dc + (((int64_t)f0 * g.n_fractals * g.channels) << g.sub_bits) : nullptr;
        cudaError_t launch_err = cudaSuccess;
#define FRI_LAUNCH_Q(CC, SS, DD, QQ, TT, PTR) \
    launch_err = launch_pdl(fri_encode_kernel<CC, SS, DD, QQ, TT>, grid, cta_threads(g), smem, stream, g, qp, gtab, t.tile_unit, t.stage_list, p, PTR, dc, lookahead, group_begin, bulk_staging, late_look_ctas)
#define FRI_LAUNCH(CC, SS, TT, PTR)                                                           \
        do {                                                                                  \
            if (qclass == kQuantNone) FRI_LAUNCH_Q(CC, SS, false, kQuantNone, TT, PTR);       \
            else if (qclass == kQuantSmallest) FRI_LAUNCH_Q(CC, SS, false, kQuantSmallest, TT, PTR); \
            else FRI_LAUNCH_Q(CC, SS, false, kQuantGeneric, TT, PTR);                         \
        } while (0)
        if (half) {
            if (g.channels == 1) FRI_LAUNCH(1, uint8_t, int16_t, c16);
            else FRI_LAUNCH(3, uint8_t, int16_t, c16);
        } else if (g.sub_bits != 0) {
#define FRI_LAUNCH_DEEP(CC, SS)                                                               \
            do {                                                                              \
                if (qclass == kQuantNone) FRI_LAUNCH_Q(CC, SS, true, kQuantNone, int32_t, c); \
                else FRI_LAUNCH_Q(CC, SS, true, kQuantGeneric, int32_t, c);                   \
            } while (0)
            if (g.channels == 1 && g.sample_bytes == 1) FRI_LAUNCH_DEEP(1, uint8_t);
            else if (g.channels == 3 && g.sample_bytes == 1) FRI_LAUNCH_DEEP(3, uint8_t);
            else if (g.channels == 1 && g.sample_bytes == 2) FRI_LAUNCH_DEEP(1, uint16_t);
            else FRI_LAUNCH_DEEP(3, uint16_t);
#undef FRI_LAUNCH_DEEP
        } else if (g.channels == 1 && g.sample_bytes == 1) FRI_LAUNCH(1, uint8_t, int32_t, c);
        else if (g.channels == 3 && g.sample_bytes == 1) FRI_LAUNCH(3, uint8_t, int32_t, c);
        else if (g.channels == 1 && g.sample_bytes == 2) FRI_LAUNCH(1, uint16_t, int32_t, c);
        else FRI_LAUNCH(3, uint16_t, int32_t, c);
#undef FRI_LAUNCH
#undef FRI_LAUNCH_Q
        if (launch_err != cudaSuccess) return launch_err;
        if (launches) ++*launches;
    }
    if (g.sub_bits > 0 && group_end == g.n_groups) {  // after the last band
        const unsigned blocks = (unsigned)((int64_t)n_frames * g.n_fractals * g.channels);
        const size_t cs = ((size_t)3 << g.sub_bits) / 2 * sizeof(int32_t) + 16;
        fri_coarse_forward_kernel<<<blocks, kCoarseThreads, cs, stream>>>(qp, g.sub_bits, g.depth, d_dc, d_coefs);
        if (launches) ++*launches;
    }
    return cudaGetLastError();
}

cudaError_t launch_decode(const Geometry &g, const DeviceTables &t, const QuantParams &qp, const void *d_coefs_any, bool half,
                          uint32_t n_frames, void *d_pixels, int32_t *d_dc, cudaStream_t stream,
                          uint32_t *launches, int group_begin, int group_end)
{
    if (half && (g.sub_bits != 0 || g.sample_bytes != 1)) return cudaErrorInvalidValue;
    const int32_t *d_coefs = static_cast<const int32_t *>(d_coefs_any);
    const int16_t *d_coefs16 = static_cast<const int16_t *>(d_coefs_any);
    if (group_end < 0) group_end = g.n_groups;
    const int n_groups = group_end - group_begin;
    if (n_frames == 0 || n_groups <= 0) return cudaSuccess;
    const size_t smem = kernel_smem_bytes(g);
    const int qclass = quant_class(qp, g);
    if (g.sub_bits > 0 && group_begin == 0) {  // before the first band
        const unsigned blocks = (unsigned)((int64_t)n_frames * g.n_fractals * g.channels);
        const size_t cs = ((size_t)3 << g.sub_bits) / 2 * sizeof(int32_t) + 16;
        fri_coarse_inverse_kernel<<<blocks, kCoarseThreads, cs, stream>>>(qp, g.sub_bits, g.depth, d_coefs, d_dc);
        if (launches) ++*launches;
    }
    const GroupDesc *gtab = (group_begin == 0 && group_end == g.n_groups && t.groups_launch) ? t.groups_launch : t.groups;
    // The CTAs of the first wave start together and their demand loads alone keep HBM busy; an L2 prefetch of
    // the whole run on top of that only delays everybody's first tiles (+2-4 % without it on a single frame).
    // Every later CTA starts alone and prefetches its run.  Tuning knob: FRI_DEC_NO_PREFETCH_CTAS.
    int no_prefetch_ctas = resident_ctas(g);
    if (const char *env = std::getenv("FRI_DEC_NO_PREFETCH_CTAS")) no_prefetch_ctas = std::atoi(env);
    uint8_t *px = static_cast<uint8_t *>(d_pixels);
    for (uint32_t f0 = 0; f0 < n_frames; f0 += 65535u) {
        const uint32_t nf = n_frames - f0 < 65535u ? n_frames - f0 : 65535u;
        const dim3 grid((unsigned)n_groups, nf);
        uint8_t *p = px + (int64_t)f0 * g.frame_bytes;
        const int32_t *c = d_coefs + (int64_t)f0 * g.coefs_per_frame;
        const int16_t *c16 = d_coefs16 + (int64_t)f0 * g.coefs_per_frame;
        int32_t *dc = d_dc ? d_dc + (((int64_t)f0 * g.n_fractals * g.channels) << g.sub_bits) : nullptr;
        cudaError_t launch_err = cudaSuccess;
#define FRI_LAUNCH_Q(CC, SS, DD, QQ, TT, PTR) \
    launch_err = launch_pdl(fri_decode_kernel<CC, SS, DD, QQ, TT>, grid, cta_threads(g), smem, stream, g, qp, gtab, t.tile_unit, t.chunk_list, t.chunk_mask, t.edge_list, PTR, dc, p, group_begin, no_prefetch_ctas)
#define FRI_LAUNCH(CC, SS, TT, PTR)                                                           \
        do {                                                                                  \
            if (qclass == kQuantNone) FRI_LAUNCH_Q(CC, SS, false, kQuantNone, TT, PTR);       \
            else if (qclass == kQuantSmallest) FRI_LAUNCH_Q(CC, SS, false, kQuantSmallest, TT, PTR); \
            else FRI_LAUNCH_Q(CC, SS, false, kQuantGeneric, TT, PTR);                         \
        } while (0)
        if (half) {
            if (g.channels == 1) FRI_LAUNCH(1, uint8_t, int16_t, c16);
            else FRI_LAUNCH(3, uint8_t, int16_t, c16);
        } else if (g.sub_bits != 0) {
#define FRI_LAUNCH_DEEP(CC, SS)                                                               \
            do {                                                                              \
                if (qclass == kQuantNone) FRI_LAUNCH_Q(CC, SS, true, kQuantNone, int32_t, c); \
                else FRI_LAUNCH_Q(CC, SS, true, kQuantGeneric, int32_t, c);                   \
            } while (0)
            if (g.channels == 1 && g.sample_bytes == 1) FRI_LAUNCH_DEEP(1, uint8_t);
            else if (g.channels == 3 && g.sample_bytes == 1) FRI_LAUNCH_DEEP(3, uint8_t);
            else if (g.channels == 1 && g.sample_bytes == 2) FRI_LAUNCH_DEEP(1, uint16_t);
            else FRI_LAUNCH_DEEP(3, uint16_t);
#undef FRI_LAUNCH_DEEP
        } else if (g.channels == 1 && g.sample_bytes == 1) FRI_LAUNCH(1, uint8_t, int32_t, c);
        else if (g.channels == 3 && g.sample_bytes == 1) FRI_LAUNCH(3, uint8_t, int32_t, c);
        else if (g.channels == 1 && g.sample_bytes == 2) FRI_LAUNCH(1, uint16_t, int32_t, c);
        else FRI_LAUNCH(3, uint16_t, int32_t, c);
#undef FRI_LAUNCH
#undef FRI_LAUNCH_Q
        if (launch_err != cudaSuccess) return launch_err;
        if (launches) ++*launches;
    }
    return cudaGetLastError();
}

cudaError_t launch_unemit(const Geometry &g, const DeviceTables &t, const EmitTables &et, uint64_t count, const void *d_in,
                          bool half, uint32_t n_frames, int32_t *d_coefs, cudaStream_t stream, uint32_t *launches)
{   // `count` here is the stream stride in elements (>= the number of Some slots)
    if (n_frames == 0 || g.n_groups == 0) return cudaSuccess;
    const size_t smem = (size_t)g.group_a * g.group_b * kTileLeaves * sizeof(int32_t);
    const size_t esz = half ? sizeof(int16_t) : sizeof(int32_t);
    for (uint32_t f0 = 0; f0 < n_frames; f0 += 65535u) {  // gridDim.y limit
        const uint32_t nf = n_frames - f0 < 65535u ? n_frames - f0 : 65535u;
        const dim3 grid((unsigned)g.n_groups, nf);
        int32_t *c = d_coefs + (int64_t)f0 * g.coefs_per_frame;
        const uint8_t *in = static_cast<const uint8_t *>(d_in) + (size_t)f0 * g.channels * count * esz;
        if (half)
            fri_unemit_kernel<int16_t><<<grid, 256, smem, stream>>>(t.groups, et.goff, et.dst, et.loc, count, g.channels,
                                                                    g.n_fractals, reinterpret_cast<const int16_t *>(in), c);
        else
            fri_unemit_kernel<int32_t><<<grid, 256, smem, stream>>>(t.groups, et.goff, et.dst, et.loc, count, g.channels,
                                                                    g.n_fractals, reinterpret_cast<const int32_t *>(in), c);
        if (launches) ++*launches;
    }
    return cudaGetLastError();
}

cudaError_t launch_pack_bits(int bits, const int16_t *d_src, uint8_t *d_dst, size_t n_blocks, cudaStream_t stream, uint32_t *launches)
{
    if (n_blocks == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)std::min<size_t>((n_blocks + 255) / 256, (size_t)148 * 16);
    if (bits == 10)
        fri_pack_kernel<10><<<blocks, 256, 0, stream>>>(reinterpret_cast<const int4 *>(d_src), reinterpret_cast<int2 *>(d_dst), n_blocks);
    else if (bits == 9)
        fri_pack_kernel<9><<<blocks, 256, 0, stream>>>(reinterpret_cast<const int4 *>(d_src), reinterpret_cast<int2 *>(d_dst), n_blocks);
    else
        return cudaErrorInvalidValue;
    if (launches) ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_unpack_bits(int bits, const uint8_t *d_src, int16_t *d_dst, size_t n_blocks, cudaStream_t stream, uint32_t *launches)
{
    if (n_blocks == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)std::min<size_t>((n_blocks + 255) / 256, (size_t)148 * 16);
    if (bits == 10)
        fri_unpack_kernel<10><<<blocks, 256, 0, stream>>>(reinterpret_cast<const int2 *>(d_src), reinterpret_cast<int4 *>(d_dst), n_blocks);
    else if (bits == 9)
        fri_unpack_kernel<9><<<blocks, 256, 0, stream>>>(reinterpret_cast<const int2 *>(d_src), reinterpret_cast<int4 *>(d_dst), n_blocks);
    else
        return cudaErrorInvalidValue;
    if (launches) ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_pack16(const int32_t *d_src, int16_t *d_dst, size_t count, cudaStream_t stream, uint32_t *launches)
{
    if (count == 0) return cudaSuccess;
    const size_t n8 = count / 8;  // coefficient blocks are multiples of 512
    const unsigned blocks = (unsigned)std::min<size_t>((n8 + 255) / 256, (size_t)148 * 16);
    fri_pack16_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const int4 *>(d_src), reinterpret_cast<int4 *>(d_dst), n8);
    if (launches) ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_unpack16(const int16_t *d_src, int32_t *d_dst, size_t count, cudaStream_t stream, uint32_t *launches)
{
    if (count == 0) return cudaSuccess;
    const size_t n8 = count / 8;
    const unsigned blocks = (unsigned)std::min<size_t>((n8 + 255) / 256, (size_t)148 * 16);
    fri_unpack16_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const int4 *>(d_src), reinterpret_cast<int4 *>(d_dst), n8);
    if (launches) ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_emit(const Geometry &g, const DeviceTables &t, const EmitTables &et, uint64_t count, const int32_t *d_coefs,
                        uint32_t n_frames, void *d_out, bool half, cudaStream_t stream, uint32_t *launches)
{
    if (count == 0 || n_frames == 0) return cudaSuccess;
    const size_t smem = (size_t)g.group_a * g.group_b * kTileLeaves * sizeof(int32_t);
    const size_t esz = half ? sizeof(int16_t) : sizeof(int32_t);
    for (uint32_t f0 = 0; f0 < n_frames; f0 += 65535u) {  // gridDim.y limit
        const uint32_t nf = n_frames - f0 < 65535u ? n_frames - f0 : 65535u;
        const dim3 grid((unsigned)g.n_groups, nf);
        const int32_t *c = d_coefs + (int64_t)f0 * g.coefs_per_frame;
        uint8_t *o = static_cast<uint8_t *>(d_out) + (size_t)f0 * g.channels * count * esz;
        if (half)
            fri_emit_kernel<int16_t><<<grid, 256, smem, stream>>>(t.groups, et.goff, et.dst, et.loc, count, g.channels,
                                                                  g.n_fractals, c, reinterpret_cast<int16_t *>(o));
        else
            fri_emit_kernel<int32_t><<<grid, 256, smem, stream>>>(t.groups, et.goff, et.dst, et.loc, count, g.channels,
                                                                  g.n_fractals, c, reinterpret_cast<int32_t *>(o));
        if (launches) ++*launches;
    }
    return cudaGetLastError();
}

}  // namespace fri
