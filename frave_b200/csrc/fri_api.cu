// fri_api.cu — the C ABI of libfri_cuda (see include/fri_cuda.h for the contract).
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <mutex>
#include <thread>
#include <cstdlib>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/fri_cuda.h"
#include "fri_codec.h"
#include "fri_kernels.cuh"
#include "fri_plan.h"

using namespace fri;

namespace {

thread_local std::string g_last_error;

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

int cuda_fail(cudaError_t e, const char *what)
{
    return fail(FRI_E_CUDA, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}

#define FRI_CUDA(call)                                              \
    do {                                                            \
        cudaError_t e__ = (call);                                   \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call);       \
    } while (0)

constexpr int kSlots = 2;      // frames resident on the device in the host-buffer entry points
constexpr int kMaxBands = 8;   // a frame is streamed through the device in up to this many bands of groups

// Device buffers of one frame in flight.
struct Slot {
    void *d_pixels = nullptr;
    int32_t *d_coefs = nullptr;
    int16_t *d_coefs16 = nullptr;  // 16-bit transport staging (fri_*_tq16), allocated on first use
    int32_t *d_dc = nullptr;
    void *d_emit = nullptr;        // emission-ordered streams of the frame (fri_*_tq_emit*), sized for int32 streams
    void *d_pack = nullptr;        // the same streams in the 10-bit packed transport (fri_*_tq_emit10)
    cudaEvent_t compute_done = nullptr;  // last kernel of the frame that used the slot
    cudaEvent_t out_done = nullptr;      // last device-to-host copy of that frame
    bool used = false;
};

// The host-buffer entry points run a three-stage pipeline: copies in, kernels, copies out, each on
// its own in-order stream and chained by events, so host-to-device and device-to-host traffic
// overlap (PCIe is full duplex) — across the bands of one frame and across frames.
struct Pipeline {
    cudaStream_t in = nullptr, compute = nullptr, out = nullptr;
    cudaEvent_t in_ready[kMaxBands] = {};
    cudaEvent_t band_done[kMaxBands] = {};
};

}  // namespace

struct fri_plan {
    Plan plan;
    int device = -1;
    DeviceTables tables;
    void *d_groups_launch = nullptr;
    void *d_groups = nullptr, *d_tile_unit = nullptr, *d_chunk_mask = nullptr, *d_chunk_list = nullptr, *d_stage_list = nullptr, *d_edge_list = nullptr, *d_absent_unit = nullptr;
    Slot slots[kSlots];
    Pipeline pipe;
    bool slots_ready = false;
    uint32_t last_launches = 0;
    cudaMemPool_t pool = nullptr;  // stream-ordered scratch of the *_device entry points (created on first use)
    std::mutex pool_mutex;
    int bands = 0;  // fri_plan_set_bands: 0 = automatic
    bool async_mode = false;  // fri_plan_set_async
    bool independent_calls = false;  // fri_plan_set_independent_calls
    // emission order (computed on first use)
    int emit_state = 0;  // 0 = not computed, 1 = ready, -1 = failed (emit_error)
    std::string emit_error;
    std::vector<uint32_t> emit_order;  // [n_tiles * 512], None slots included
    std::vector<uint32_t> emit_src;    // Some slots only
    void *d_emit_goff = nullptr, *d_emit_dst = nullptr, *d_emit_loc = nullptr;  // the Some slots partitioned by group
    EmitTables emit_tables;
    bool emit_device_ready = false;     // all three tables uploaded
    // prediction tables (computed on first use)
    bool predict_ready = false;
    void *d_pred_adjacent = nullptr, *d_pred_steps = nullptr;
    void *h_dense = nullptr;    // pinned staging of fri_frv_decode: one frame of dense blocks
    void *h_symbols = nullptr;  // pinned staging of fri_frv_encode: [C][count] u16 symbols, then [C][count] u8 buckets
    PredictTables predict_tables;
    // host lattice index for the host predictor / entropy decoder (computed on first use)
    bool lattice_ready = false;
    LatticeIndex lattice;
    std::vector<uint8_t> some_flat;  // [n_tiles * 512] Some / None
    uint64_t fit_zero_rows[3] = {0, 0, 0};  // codec::fit_zero_rows of this plan
};

namespace {

std::once_flag g_cfg_once[16];
cudaError_t g_cfg_err[16];

int enter_device(const fri_plan *p)
{
    if (!p) return fail(FRI_E_INVALID, "plan is NULL");
    if (p->device < 0) return fail(FRI_E_CUDA, "plan was created without a device (device < 0): no CPU fallback exists");
    FRI_CUDA(cudaSetDevice(p->device));
    return FRI_OK;
}

int check_q(const int32_t *q)
{
    if (!q) return FRI_OK;
    for (int l = 0; l < 32; ++l)
        if (q[l] < 1) return fail(FRI_E_INVALID, "quantization matrix entry %d is %d; entries must be >= 1", l, q[l]);
    return FRI_OK;
}

// Stream-ordered scratch for the device-resident entry points (the depth > 9 low-pass roots, the int16 streams
// behind the packed transport): a private memory pool per plan that keeps its memory across calls (release
// threshold = never), so a call costs a sub-allocation, not a trip to the driver — the device's default pool
// hands memory back at every synchronisation and made such calls 0.2-0.7 ms slower.
int pool_alloc(const fri_plan *cp, void **out, size_t bytes, cudaStream_t st)
{
    fri_plan *p = const_cast<fri_plan *>(cp);
    {
        std::lock_guard<std::mutex> lock(p->pool_mutex);
        if (!p->pool) {
            cudaMemPoolProps props{};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = p->device;
            FRI_CUDA(cudaMemPoolCreate(&p->pool, &props));
            uint64_t never = UINT64_MAX;
            FRI_CUDA(cudaMemPoolSetAttribute(p->pool, cudaMemPoolAttrReleaseThreshold, &never));
        }
    }
    FRI_CUDA(cudaMallocFromPoolAsync(out, bytes, p->pool, st));
    return FRI_OK;
}

size_t dc_elems_per_frame(const Geometry &g)
{
    return g.sub_bits > 0 ? ((size_t)g.n_fractals * g.channels) << g.sub_bits : 0;
}

int ensure_slots(fri_plan *p)
{
    if (p->slots_ready) return FRI_OK;
    // Every resource is created at most once, so a call that failed half-way (out of device memory) can be
    // retried without leaking what the first attempt got.
    const Geometry &g = p->plan.geo;
    Pipeline &pl = p->pipe;
    for (cudaStream_t *st : {&pl.in, &pl.compute, &pl.out})
        if (!*st) FRI_CUDA(cudaStreamCreateWithFlags(st, cudaStreamNonBlocking));
    for (int k = 0; k < kMaxBands; ++k) {
        if (!pl.in_ready[k]) FRI_CUDA(cudaEventCreateWithFlags(&pl.in_ready[k], cudaEventDisableTiming));
        if (!pl.band_done[k]) FRI_CUDA(cudaEventCreateWithFlags(&pl.band_done[k], cudaEventDisableTiming));
    }
    for (auto &s : p->slots) {
        if (!s.d_pixels) FRI_CUDA(cudaMalloc(&s.d_pixels, (size_t)g.frame_bytes + 16));
        if (!s.d_coefs) FRI_CUDA(cudaMalloc(&s.d_coefs, (size_t)g.coefs_per_frame * sizeof(int32_t) + 16));
        if (g.sub_bits > 0 && !s.d_dc) FRI_CUDA(cudaMalloc(&s.d_dc, dc_elems_per_frame(g) * sizeof(int32_t)));
        if (!s.compute_done) FRI_CUDA(cudaEventCreateWithFlags(&s.compute_done, cudaEventDisableTiming));
        if (!s.out_done) FRI_CUDA(cudaEventCreateWithFlags(&s.out_done, cudaEventDisableTiming));
    }
    p->slots_ready = true;
    return FRI_OK;
}

// A slot is reused by the frame two calls later — possibly of the other direction or another element size
// (an encode and a decode may alternate on one handle, also in asynchronous mode).  Its buffers are free once
// the previous user's kernels AND device-to-host copies are done, and both the copy-in stream and the compute
// stream touch them first, so both wait for both events.
int acquire_slot(fri_plan *p, Slot &s)
{
    if (!s.used) return FRI_OK;
    Pipeline &pl = p->pipe;
    for (cudaStream_t st : {pl.in, pl.compute}) {
        FRI_CUDA(cudaStreamWaitEvent(st, s.compute_done, 0));
        FRI_CUDA(cudaStreamWaitEvent(st, s.out_done, 0));
    }
    return FRI_OK;
}

// Host buffers of the asynchronous mode must be page-locked: a pageable pointer would make cudaMemcpyAsync
// synchronous at best and — once the call has returned — let the caller free memory a copy still reads.
int check_pinned(const fri_plan *p, const void *ptr, const char *what)
{
    if (!p->async_mode) return FRI_OK;
    cudaPointerAttributes attr{};
    cudaError_t e = cudaPointerGetAttributes(&attr, ptr);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(FRI_E_INVALID, "%s: asynchronous mode needs page-locked host memory (fri_host_alloc)", what);
    }
    if (attr.type != cudaMemoryTypeHost && attr.type != cudaMemoryTypeManaged)
        return fail(FRI_E_INVALID, "%s: asynchronous mode needs page-locked host memory (fri_host_alloc); the buffer is pageable",
                    what);
    return FRI_OK;
}

// End of a host-buffer entry point: wait for the plan's three streams, unless the plan is in
// asynchronous mode (fri_plan_set_async), where the caller does that with fri_plan_sync.
int finish_host_call(fri_plan *p)
{
    if (p->async_mode) return FRI_OK;
    Pipeline &pl = p->pipe;
    FRI_CUDA(cudaStreamSynchronize(pl.out));
    FRI_CUDA(cudaStreamSynchronize(pl.compute));
    FRI_CUDA(cudaStreamSynchronize(pl.in));
    return FRI_OK;
}

int ensure_slots16(fri_plan *p)
{
    int rc = ensure_slots(p);
    if (rc) return rc;
    const Geometry &g = p->plan.geo;
    if (g.sample_bytes != 1)
        return fail(FRI_E_UNSUPPORTED, "16-bit coefficient transport needs 8-bit samples (residues of 16-bit samples need 18 bits)");
    for (auto &s : p->slots)
        if (!s.d_coefs16) FRI_CUDA(cudaMalloc(&s.d_coefs16, (size_t)g.coefs_per_frame * sizeof(int16_t) + 16));
    return FRI_OK;
}

// Splits the plan's groups (ordered top to bottom) into bands of consecutive groups.
struct Band {
    int g0, g1;        // groups [g0, g1)
    size_t t0, t1;     // base tiles [t0, t1) == fractals at depth 9
    int row_hi;        // encode: pixel rows [0, row_hi) must be on the device before the band runs
    int final_rows;    // decode: pixel rows [0, final_rows) are complete once the band has run
};

int make_bands(const fri_plan *p, Band bands[kMaxBands])
{
    const Plan &pl = p->plan;
    const Geometry &g = pl.geo;
    int n = g.sub_bits > 0 ? 1 : std::max(1, std::min(kMaxBands, g.n_groups / 256));
    if (const char *env = std::getenv("FRI_BANDS")) n = std::max(1, std::min(kMaxBands, std::atoi(env)));  // tuning knob
    if (p->bands > 0 && g.sub_bits == 0) n = std::min(kMaxBands, p->bands);
    n = std::min(n, std::max(1, g.n_groups));
    for (int k = 0; k < n; ++k) {
        Band &b = bands[k];
        b.g0 = (int)((int64_t)g.n_groups * k / n);
        b.g1 = (int)((int64_t)g.n_groups * (k + 1) / n);
        b.t0 = pl.groups[b.g0].tile_base;
        b.t1 = b.g1 < g.n_groups ? pl.groups[b.g1].tile_base : (size_t)g.n_base_tiles;
        int hi = 0;
        for (int i = b.g0; i < b.g1; ++i) hi = std::max(hi, pl.groups[i].y0 + g.region_h);
        b.row_hi = k == n - 1 ? g.height : std::max(0, std::min(g.height, hi));
    }
    for (int k = 0; k < n; ++k) {
        int lo = g.height;  // first row any later band still writes
        for (int i = bands[k].g1; i < g.n_groups; ++i) lo = std::min(lo, pl.groups[i].y0);
        bands[k].final_rows = k == n - 1 ? g.height : std::max(0, std::min(g.height, lo));
    }
    for (int k = 1; k < n; ++k) {  // monotone: a band never un-finishes rows / needs fewer rows
        bands[k].row_hi = std::max(bands[k].row_hi, bands[k - 1].row_hi);
        bands[k].final_rows = std::max(bands[k].final_rows, bands[k - 1].final_rows);
    }
    return n;
}

}  // namespace

#if FRI_TRACE
namespace fri { cudaError_t debug_trace(unsigned long long *out, size_t n); cudaError_t debug_trace2(unsigned long long *out, size_t n); }
#endif

extern "C" {

const char *fri_version(void) { return "libfri_cuda 0.1.0 sm_100a"; }
const char *fri_last_error(void) { return g_last_error.c_str(); }

int fri_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int fri_plan_create(fri_plan **out, int device, uint32_t width, uint32_t height, uint32_t channels, uint32_t depth,
                    uint32_t sample_bytes)
{
    if (!out) return fail(FRI_E_INVALID, "out is NULL");
    *out = nullptr;
    fri_plan *p = new (std::nothrow) fri_plan;
    if (!p) return fail(FRI_E_NOMEM, "out of host memory");
    std::string err;
    try {
        err = build_plan(p->plan, width, height, channels, depth, sample_bytes, 0, 0);
    } catch (const std::bad_alloc &) {
        delete p;
        return fail(FRI_E_NOMEM, "out of host memory while building the lattice");
    }
    if (!err.empty()) {
        delete p;
        return fail(FRI_E_INVALID, "%s", err.c_str());
    }
    p->device = device;
    if (device >= 0) {
        int n = fri_device_count();
        if (device >= n) {
            delete p;
            return fail(FRI_E_CUDA, "CUDA device %d requested but %d device(s) visible; there is no CPU fallback", device, n);
        }
        const Plan &pl = p->plan;
        const size_t smem = kernel_smem_bytes(pl.geo);
        if (smem > 227 * 1024) {
            delete p;
            return fail(FRI_E_UNSUPPORTED, "group needs %zu bytes of shared memory", smem);
        }
        auto upload = [&](void **dst, const void *src, size_t bytes) -> cudaError_t {
            if (bytes == 0) return cudaSuccess;
            cudaError_t e = cudaMalloc(dst, bytes);
            if (e != cudaSuccess) return e;
            return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
        };
        cudaError_t e = cudaSetDevice(device);
        if (e == cudaSuccess && device < 16) {
            std::call_once(g_cfg_once[device], [&] { g_cfg_err[device] = configure_kernels(); });
            e = g_cfg_err[device];
        } else if (e == cudaSuccess) {
            e = configure_kernels();
        }
        if (e == cudaSuccess) e = upload(&p->d_groups, pl.groups.data(), pl.groups.size() * sizeof(GroupDesc));
        // Launch order of whole-frame launches: groups whose staged region crosses the image border take the
        // clipping paths (slower staging and write-out) and, in plan order, the bottom border row would be
        // the last CTAs of the launch — they go first instead (FRI_ORDER=0 keeps plan order).
        int order = 1;
        if (const char *env = std::getenv("FRI_ORDER")) order = std::atoi(env);  // tuning knob
        if (e == cudaSuccess && order != 0) {
            std::vector<GroupDesc> lo(pl.groups);
            const Geometry &gg = pl.geo;
            auto border = [&](const GroupDesc &d) {
                return !(d.x0 >= 0 && d.y0 >= 0 && d.x0 + gg.region_w <= gg.width && d.y0 + gg.region_h <= gg.height);
            };
            if (order == 1) std::stable_partition(lo.begin(), lo.end(), border);
            else std::stable_partition(lo.begin(), lo.end(), [&](const GroupDesc &d) { return !border(d); });
            e = upload(&p->d_groups_launch, lo.data(), lo.size() * sizeof(GroupDesc));
        }
        if (e == cudaSuccess && pl.geo.sub_bits > 0)
            e = upload(&p->d_tile_unit, pl.tile_unit.data(), pl.tile_unit.size() * sizeof(uint32_t));
        if (e == cudaSuccess && !pl.absent_unit.empty())
            e = upload(&p->d_absent_unit, pl.absent_unit.data(), pl.absent_unit.size() * sizeof(uint32_t));
        if (e == cudaSuccess) e = upload(&p->d_chunk_mask, pl.chunk_mask.data(), pl.chunk_mask.size() * sizeof(uint16_t));
        if (e == cudaSuccess) e = upload(&p->d_stage_list, pl.stage_list.data(), pl.stage_list.size() * sizeof(uint32_t));
        if (e == cudaSuccess) e = upload(&p->d_chunk_list, pl.chunk_list.data(), pl.chunk_list.size() * sizeof(uint32_t));
        if (e == cudaSuccess) e = upload(&p->d_edge_list, pl.edge_list.data(), pl.edge_list.size() * sizeof(uint32_t));
        if (e != cudaSuccess) {
            fri_plan_destroy(p);
            return cuda_fail(e, "uploading the plan tables");
        }
        p->tables.groups = static_cast<const GroupDesc *>(p->d_groups);
        p->tables.groups_launch = static_cast<const GroupDesc *>(p->d_groups_launch);
        p->tables.tile_unit = static_cast<const uint32_t *>(p->d_tile_unit);
        p->tables.absent_unit = static_cast<const uint32_t *>(p->d_absent_unit);
        p->tables.n_absent = (uint32_t)pl.absent_unit.size();
        p->tables.chunk_mask = static_cast<const uint16_t *>(p->d_chunk_mask);
        p->tables.chunk_list = static_cast<const uint32_t *>(p->d_chunk_list);
        p->tables.edge_list = static_cast<const uint32_t *>(p->d_edge_list);
        p->tables.stage_list = static_cast<const uint32_t *>(p->d_stage_list);
    }
    *out = p;
    return FRI_OK;
}

void fri_plan_destroy(fri_plan *p)
{
    if (!p) return;
    if (p->device >= 0 && cudaSetDevice(p->device) == cudaSuccess) {
        Pipeline &pl = p->pipe;
        for (cudaStream_t st : {pl.in, pl.compute, pl.out})
            if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
        for (int k = 0; k < kMaxBands; ++k) {
            if (pl.in_ready[k]) cudaEventDestroy(pl.in_ready[k]);
            if (pl.band_done[k]) cudaEventDestroy(pl.band_done[k]);
        }
        for (auto &s : p->slots) {
            if (s.d_pixels) cudaFree(s.d_pixels);
            if (s.d_coefs) cudaFree(s.d_coefs);
            if (s.d_coefs16) cudaFree(s.d_coefs16);
            if (s.d_dc) cudaFree(s.d_dc);
            if (s.d_emit) cudaFree(s.d_emit);
            if (s.d_pack) cudaFree(s.d_pack);
            if (s.compute_done) cudaEventDestroy(s.compute_done);
            if (s.out_done) cudaEventDestroy(s.out_done);
        }
        if (p->pool) cudaMemPoolDestroy(p->pool);
        if (p->d_groups) cudaFree(p->d_groups);
        if (p->d_groups_launch) cudaFree(p->d_groups_launch);
        if (p->d_tile_unit) cudaFree(p->d_tile_unit);
        if (p->d_absent_unit) cudaFree(p->d_absent_unit);
        if (p->d_chunk_mask) cudaFree(p->d_chunk_mask);
        if (p->d_chunk_list) cudaFree(p->d_chunk_list);
        if (p->d_edge_list) cudaFree(p->d_edge_list);
        if (p->d_emit_goff) cudaFree(p->d_emit_goff);
        if (p->d_emit_dst) cudaFree(p->d_emit_dst);
        if (p->d_emit_loc) cudaFree(p->d_emit_loc);
        if (p->d_stage_list) cudaFree(p->d_stage_list);
        for (void *d : {p->d_pred_adjacent, p->d_pred_steps})
            if (d) cudaFree(d);
        if (p->h_symbols) cudaFreeHost(p->h_symbols);
        if (p->h_dense) cudaFreeHost(p->h_dense);
    }
    delete p;
}

uint32_t fri_plan_num_tiles(const fri_plan *p) { return p ? (uint32_t)p->plan.geo.n_fractals : 0; }
uint32_t fri_plan_num_built(const fri_plan *p) { return p ? p->plan.n_built : 0; }
uint32_t fri_plan_num_full_tiles(const fri_plan *p) { return p ? p->plan.n_full : 0; }
uint64_t fri_plan_coefs_per_frame(const fri_plan *p) { return p ? (uint64_t)p->plan.geo.coefs_per_frame : 0; }
uint64_t fri_plan_pixels_covered(const fri_plan *p) { return p ? p->plan.pixels_covered : 0; }

int fri_plan_centers(const fri_plan *p, int32_t *centers)
{
    if (!p || !centers) return fail(FRI_E_INVALID, "NULL argument");
    std::memcpy(centers, p->plan.centers.data(), p->plan.centers.size() * sizeof(int32_t));
    return FRI_OK;
}

int fri_plan_masks(const fri_plan *p, uint32_t *masks)
{
    if (!p || !masks) return fail(FRI_E_INVALID, "NULL argument");
    const Geometry &g = p->plan.geo;
    const size_t words = ((size_t)1 << g.depth) / 32;
    for (int32_t i = 0; i < g.n_fractals; ++i) {
        if (p->plan.full[i])
            std::memset(masks + i * words, 0xff, words * sizeof(uint32_t));
        else
            fractal_mask(g.depth, p->plan.centers[2 * i], p->plan.centers[2 * i + 1], g.width, g.height, masks + i * words);
    }
    return FRI_OK;
}

int fri_plan_launch_info(const fri_plan *p, int32_t info[16])
{
    if (!p || !info) return fail(FRI_E_INVALID, "NULL argument");
    const Geometry &g = p->plan.geo;
    const int32_t v[16] = {g.group_a, g.group_b, g.region_w, g.region_h, g.pitch, (int32_t)kernel_smem_bytes(g),
                           g.n_groups, g.n_base_tiles, cta_threads(g), g.chunks_per_row, g.depth, g.sub_bits,
                           g.list_full[0], g.list_all[0], 0, 0};
    std::memcpy(info, v, sizeof(v));
    return FRI_OK;
}

static int encode_device(const fri_plan *cp, const void *d_pixels, uint32_t n_frames, const int32_t *q, void *d_coefs, bool half,
                         void *stream)
{
    fri_plan *p = const_cast<fri_plan *>(cp);
    int rc = enter_device(p);
    if (rc) return rc;
    if ((rc = check_q(q))) return rc;
    if (n_frames == 0) return FRI_OK;
    if (!d_pixels || !d_coefs) return fail(FRI_E_INVALID, "NULL device buffer");
    const Geometry &g = p->plan.geo;
    if (half && (g.sample_bytes != 1 || g.sub_bits != 0))
        return fail(FRI_E_UNSUPPORTED, "int16 coefficient arrays need 8-bit samples and depth 9");
    if ((uintptr_t)d_coefs & 15) return fail(FRI_E_INVALID, "d_coefs must be 16-byte aligned");
    if ((uintptr_t)d_pixels & (uintptr_t)(g.sample_bytes - 1)) return fail(FRI_E_INVALID, "d_pixels must be aligned to the sample size");
    QuantParams qp;
    make_quant_params(qp, q, 0);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // depth > 9: the base tiles' low-pass roots go through a scratch array that lives for this call only —
    // stream-ordered allocation, so calls on different streams (or an encode and a decode in flight together)
    // never share it and the plan stays read-only
    int32_t *d_dc = nullptr;
    if (g.sub_bits > 0 && (rc = pool_alloc(p, reinterpret_cast<void **>(&d_dc), n_frames * dc_elems_per_frame(g) * sizeof(int32_t), st)))
        return rc;
    uint32_t launches = 0;
    Geometry gl = g;
    gl.independent_calls = p->independent_calls ? 1 : 0;
    const cudaError_t e = launch_encode(gl, p->tables, qp, d_pixels, n_frames, d_coefs, half, d_dc, st, &launches);
    if (d_dc) cudaFreeAsync(d_dc, st);
    p->last_launches = launches;
    if (e != cudaSuccess) return cuda_fail(e, "launch_encode");
    return FRI_OK;
}

static int decode_device(const fri_plan *cp, const void *d_coefs, bool half, uint32_t n_frames, const int32_t *q, int dequant_mode,
                         void *d_pixels, void *stream)
{
    fri_plan *p = const_cast<fri_plan *>(cp);
    int rc = enter_device(p);
    if (rc) return rc;
    if ((rc = check_q(q))) return rc;
    if (dequant_mode != FRI_DEQUANT_DIVIDE && dequant_mode != FRI_DEQUANT_MULTIPLY)
        return fail(FRI_E_INVALID, "dequant_mode must be FRI_DEQUANT_DIVIDE or FRI_DEQUANT_MULTIPLY");
    if (n_frames == 0) return FRI_OK;
    if (!d_pixels || !d_coefs) return fail(FRI_E_INVALID, "NULL device buffer");
    const Geometry &g = p->plan.geo;
    if (half && (g.sample_bytes != 1 || g.sub_bits != 0))
        return fail(FRI_E_UNSUPPORTED, "int16 coefficient arrays need 8-bit samples and depth 9");
    if ((uintptr_t)d_coefs & 15) return fail(FRI_E_INVALID, "d_coefs must be 16-byte aligned");
    if ((uintptr_t)d_pixels & (uintptr_t)(g.sample_bytes - 1)) return fail(FRI_E_INVALID, "d_pixels must be aligned to the sample size");
    QuantParams qp;
    make_quant_params(qp, q, dequant_mode == FRI_DEQUANT_MULTIPLY);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // from_wavelet zero-initialises the raster (wavelet_transform.rs:309-317); only needed when
    // the retained fractals do not cover every pixel.
    if (p->plan.pixels_covered != (uint64_t)g.width * g.height)
        FRI_CUDA(cudaMemsetAsync(d_pixels, 0, (size_t)g.frame_bytes * n_frames, st));
    int32_t *d_dc = nullptr;  // per-call low-pass scratch (depth > 9), see encode_device
    if (g.sub_bits > 0 && (rc = pool_alloc(p, reinterpret_cast<void **>(&d_dc), n_frames * dc_elems_per_frame(g) * sizeof(int32_t), st)))
        return rc;
    uint32_t launches = 0;
    Geometry gl = g;
    gl.independent_calls = p->independent_calls && p->plan.pixels_covered == (uint64_t)g.width * g.height ? 1 : 0;  // (the memset orders itself)
    const cudaError_t e = launch_decode(gl, p->tables, qp, d_coefs, half, n_frames, d_pixels, d_dc, st, &launches);
    if (d_dc) cudaFreeAsync(d_dc, st);
    p->last_launches = launches;
    if (e != cudaSuccess) return cuda_fail(e, "launch_decode");
    return FRI_OK;
}

/* ---- one image split over several GPUs by ranges of tile groups -------------------------- */
struct Part {
    int g0 = 0, g1 = 0;       // groups [g0, g1)
    int64_t t0 = 0, t1 = 0;   // tiles [t0, t1) in plan order: the coefficient blocks the part produces / consumes
    int row0 = 0, row1 = 0;   // pixel rows [row0, row1) the part's groups read (encode) or write into (decode)
};

static int plan_part(const fri_plan *p, uint32_t part, uint32_t n_parts, Part &out)
{
    if (!p) return fail(FRI_E_INVALID, "plan is NULL");
    const Plan &pl = p->plan;
    const Geometry &g = pl.geo;
    if (g.sub_bits != 0) return fail(FRI_E_UNSUPPORTED, "an image is split by tile groups at depth 9 only (deeper trees fold every base tile's root in one pass)");
    if (n_parts == 0 || part >= n_parts || n_parts > (uint32_t)std::max(1, g.n_groups))
        return fail(FRI_E_INVALID, "part %u of %u: need part < n_parts <= %d groups", part, n_parts, g.n_groups);
    out.g0 = (int)((int64_t)g.n_groups * part / n_parts);
    out.g1 = (int)((int64_t)g.n_groups * (part + 1) / n_parts);
    out.t0 = pl.groups[out.g0].tile_base;
    out.t1 = out.g1 < g.n_groups ? (int64_t)pl.groups[out.g1].tile_base : (int64_t)g.n_base_tiles;
    int lo = g.height, hi = 0;
    for (int i = out.g0; i < out.g1; ++i) {
        lo = std::min(lo, pl.groups[i].y0);
        hi = std::max(hi, pl.groups[i].y0 + g.region_h);
    }
    out.row0 = std::max(0, std::min(lo, g.height));
    out.row1 = std::max(out.row0, std::min(hi, g.height));
    return FRI_OK;
}

int fri_plan_part(const fri_plan *p, uint32_t part, uint32_t n_parts, uint32_t *group_begin, uint32_t *group_end,
                  uint32_t *tile_begin, uint32_t *tile_end, uint32_t *row_begin, uint32_t *row_end)
{
    Part pt;
    int rc = plan_part(p, part, n_parts, pt);
    if (rc) return rc;
    if (group_begin) *group_begin = (uint32_t)pt.g0;
    if (group_end) *group_end = (uint32_t)pt.g1;
    if (tile_begin) *tile_begin = (uint32_t)pt.t0;
    if (tile_end) *tile_end = (uint32_t)pt.t1;
    if (row_begin) *row_begin = (uint32_t)pt.row0;
    if (row_end) *row_end = (uint32_t)pt.row1;
    return FRI_OK;
}

// The kernels address a whole frame; a range of groups touches some rows and tiles only, so the caller's band
// buffers are handed over as frame bases shifted back by the band's first row / first tile.
static int range_call(const fri_plan *cp, bool encode, const void *d_in, void *d_out, const int32_t *q, int dequant_mode,
                      int g0, int g1, int64_t tile_first, int row_first, void *stream)
{
    fri_plan *p = const_cast<fri_plan *>(cp);
    int rc = enter_device(p);
    if (rc) return rc;
    if ((rc = check_q(q))) return rc;
    if (!encode && dequant_mode != FRI_DEQUANT_DIVIDE && dequant_mode != FRI_DEQUANT_MULTIPLY)
        return fail(FRI_E_INVALID, "dequant_mode must be FRI_DEQUANT_DIVIDE or FRI_DEQUANT_MULTIPLY");
    if (!d_in || !d_out) return fail(FRI_E_INVALID, "NULL device buffer");
    const Geometry &g = p->plan.geo;
    const void *d_pixels_rows = encode ? d_in : d_out;
    const void *d_coefs_tiles = encode ? d_out : d_in;
    if ((uintptr_t)d_coefs_tiles & 15) return fail(FRI_E_INVALID, "the coefficient buffer must be 16-byte aligned");
    if ((uintptr_t)d_pixels_rows & (uintptr_t)(g.sample_bytes - 1)) return fail(FRI_E_INVALID, "the pixel band must be aligned to the sample size");
    const size_t block = (size_t)g.channels << g.depth;
    const uintptr_t px_base = (uintptr_t)d_pixels_rows - (uintptr_t)row_first * (uintptr_t)g.row_stride;
    const uintptr_t co_base = (uintptr_t)d_coefs_tiles - (uintptr_t)tile_first * block * sizeof(int32_t);
    QuantParams qp;
    make_quant_params(qp, q, !encode && dequant_mode == FRI_DEQUANT_MULTIPLY);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint32_t launches = 0;
    const cudaError_t e = encode
        ? launch_encode(g, p->tables, qp, reinterpret_cast<const void *>(px_base), 1, reinterpret_cast<void *>(co_base), false, nullptr, st,
                        &launches, g0, g1)
        : launch_decode(g, p->tables, qp, reinterpret_cast<const void *>(co_base), false, 1, reinterpret_cast<void *>(px_base), nullptr, st,
                        &launches, g0, g1);
    p->last_launches = launches;
    if (e != cudaSuccess) return cuda_fail(e, encode ? "launch_encode (group range)" : "launch_decode (group range)");
    return FRI_OK;
}

int fri_encode_tq_device_part(const fri_plan *p, const void *d_pixel_rows, const int32_t *q, int32_t *d_coef_tiles, uint32_t part,
                              uint32_t n_parts, void *stream)
{
    Part pt;
    int rc = plan_part(p, part, n_parts, pt);
    if (rc) return rc;
    return range_call(p, true, d_pixel_rows, d_coef_tiles, q, FRI_DEQUANT_DIVIDE, pt.g0, pt.g1, pt.t0, pt.row0, stream);
}

int fri_decode_tq_device_part(const fri_plan *p, const int32_t *d_coef_tiles, const int32_t *q, int dequant_mode, void *d_pixel_rows,
                              uint32_t part, uint32_t n_parts, void *stream)
{
    Part pt;
    int rc = plan_part(p, part, n_parts, pt);
    if (rc) return rc;
    return range_call(p, false, d_coef_tiles, d_pixel_rows, q, dequant_mode, pt.g0, pt.g1, pt.t0, pt.row0, stream);
}

int fri_plan_groups_in_rows(const fri_plan *p, uint32_t group_begin, uint32_t group_end, uint32_t row_begin, uint32_t row_end,
                            uint32_t *first, uint32_t *last, uint32_t *span_begin, uint32_t *span_end)
{
    if (!p) return fail(FRI_E_INVALID, "plan is NULL");
    const Plan &pl = p->plan;
    const Geometry &g = pl.geo;
    if (g.sub_bits != 0) return fail(FRI_E_UNSUPPORTED, "group ranges exist at depth 9 only");
    if (group_begin > group_end || group_end > (uint32_t)g.n_groups) return fail(FRI_E_INVALID, "bad group range");
    int lo = -1, hi = -1;
    for (int i = (int)group_begin; i < (int)group_end; ++i) {
        const int y0 = std::max(0, pl.groups[i].y0), y1 = std::min(g.height, pl.groups[i].y0 + g.region_h);
        if (y0 < (int)row_end && y1 > (int)row_begin) {
            if (lo < 0) lo = i;
            hi = i + 1;
        }
    }
    if (lo < 0) lo = hi = (int)group_begin;
    int s0 = g.height, s1 = 0;
    for (int i = lo; i < hi; ++i) {
        s0 = std::min(s0, std::max(0, pl.groups[i].y0));
        s1 = std::max(s1, std::min(g.height, pl.groups[i].y0 + g.region_h));
    }
    if (lo == hi) s0 = s1 = 0;
    if (first) *first = (uint32_t)lo;
    if (last) *last = (uint32_t)hi;
    if (span_begin) *span_begin = (uint32_t)s0;
    if (span_end) *span_end = (uint32_t)s1;
    return FRI_OK;
}

int fri_decode_tq_device_groups(const fri_plan *p, const int32_t *d_coef_tiles, uint32_t tile_first, const int32_t *q, int dequant_mode,
                                void *d_pixel_rows, int32_t row_first, uint32_t group_begin, uint32_t group_end, void *stream)
{
    if (!p) return fail(FRI_E_INVALID, "plan is NULL");
    const Geometry &g = p->plan.geo;
    if (g.sub_bits != 0) return fail(FRI_E_UNSUPPORTED, "group ranges exist at depth 9 only");
    if (group_begin > group_end || group_end > (uint32_t)g.n_groups) return fail(FRI_E_INVALID, "bad group range");
    if (group_begin == group_end) return FRI_OK;
    if ((int64_t)tile_first > (int64_t)p->plan.groups[group_begin].tile_base) return fail(FRI_E_INVALID, "tile_first lies behind the first group's tiles");
    return range_call(p, false, d_coef_tiles, d_pixel_rows, q, dequant_mode, (int)group_begin, (int)group_end, tile_first, row_first, stream);
}

int fri_encode_tq_device(const fri_plan *p, const void *d_pixels, uint32_t n_frames, const int32_t *q, int32_t *d_coefs,
                         void *stream)
{
    return encode_device(p, d_pixels, n_frames, q, d_coefs, false, stream);
}

int fri_encode_tq_device16(const fri_plan *p, const void *d_pixels, uint32_t n_frames, const int32_t *q, int16_t *d_coefs,
                           void *stream)
{
    return encode_device(p, d_pixels, n_frames, q, d_coefs, true, stream);
}

int fri_decode_tq_device(const fri_plan *p, const int32_t *d_coefs, uint32_t n_frames, const int32_t *q, int dequant_mode,
                         void *d_pixels, void *stream)
{
    return decode_device(p, d_coefs, false, n_frames, q, dequant_mode, d_pixels, stream);
}

int fri_decode_tq_device16(const fri_plan *p, const int16_t *d_coefs, uint32_t n_frames, const int32_t *q, int dequant_mode,
                           void *d_pixels, void *stream)
{
    return decode_device(p, d_coefs, true, n_frames, q, dequant_mode, d_pixels, stream);
}

static int encode_host(fri_plan *p, const void *pixels, uint32_t n_frames, const int32_t *q, void *coefs, bool half)
{
    int rc = enter_device(p);
    if (rc) return rc;
    if ((rc = check_q(q))) return rc;
    if (n_frames == 0) return FRI_OK;
    if (!pixels || !coefs) return fail(FRI_E_INVALID, "NULL host buffer");
    if ((rc = check_pinned(p, pixels, "pixels")) || (rc = check_pinned(p, coefs, "coefs"))) return rc;
    if ((rc = half ? ensure_slots16(p) : ensure_slots(p))) return rc;
    const Geometry &g = p->plan.geo;
    Pipeline &pl = p->pipe;
    QuantParams qp;
    make_quant_params(qp, q, 0);
    p->last_launches = 0;
    Band bands[kMaxBands];
    const int n_bands = make_bands(p, bands);
    const size_t block = (size_t)g.channels << g.depth;  // coefficients per fractal
    for (uint32_t f = 0; f < n_frames; ++f) {
        Slot &s = p->slots[f % kSlots];
        const uint8_t *src = static_cast<const uint8_t *>(pixels) + (size_t)f * g.frame_bytes;
        const size_t esz = half ? sizeof(int16_t) : sizeof(int32_t);  // bytes per coefficient on the host side
        uint8_t *dst = static_cast<uint8_t *>(coefs) + (size_t)f * g.coefs_per_frame * esz;
        if ((rc = acquire_slot(p, s))) return rc;
        int rows_up = 0;
        for (int k = 0; k < n_bands; ++k) {
            const Band &b = bands[k];
            if (b.row_hi > rows_up) {
                FRI_CUDA(cudaMemcpyAsync(static_cast<uint8_t *>(s.d_pixels) + (size_t)rows_up * g.row_stride,
                                         src + (size_t)rows_up * g.row_stride, (size_t)(b.row_hi - rows_up) * g.row_stride,
                                         cudaMemcpyHostToDevice, pl.in));
                rows_up = b.row_hi;
            }
            FRI_CUDA(cudaEventRecord(pl.in_ready[k], pl.in));
            FRI_CUDA(cudaStreamWaitEvent(pl.compute, pl.in_ready[k], 0));
            // 16-bit transport: at depth 9 the kernel writes int16 itself; deep trees (coarse levels work on
            // int32) are repacked on the device
            const bool native16 = half && g.sub_bits == 0;
            void *d_out = native16 ? static_cast<void *>(s.d_coefs16) : static_cast<void *>(s.d_coefs);
            FRI_CUDA(launch_encode(g, p->tables, qp, s.d_pixels, 1, d_out, native16, s.d_dc, pl.compute, &p->last_launches, b.g0, b.g1));
            // the band's coefficient range (deep trees: one band, the coarse kernel has touched every fractal's top levels)
            const size_t c0 = g.sub_bits == 0 ? b.t0 * block : 0;
            const size_t cn = g.sub_bits == 0 ? (b.t1 - b.t0) * block : (size_t)g.coefs_per_frame;
            if (half && !native16) FRI_CUDA(launch_pack16(s.d_coefs + c0, s.d_coefs16 + c0, cn, pl.compute, &p->last_launches));
            FRI_CUDA(cudaEventRecord(pl.band_done[k], pl.compute));
            FRI_CUDA(cudaStreamWaitEvent(pl.out, pl.band_done[k], 0));
            const void *d_src = half ? static_cast<const void *>(s.d_coefs16 + c0) : static_cast<const void *>(s.d_coefs + c0);
            FRI_CUDA(cudaMemcpyAsync(dst + c0 * esz, d_src, cn * esz, cudaMemcpyDeviceToHost, pl.out));
        }
        FRI_CUDA(cudaEventRecord(s.compute_done, pl.compute));
        FRI_CUDA(cudaEventRecord(s.out_done, pl.out));
        s.used = true;
    }
    return finish_host_call(p);
}

int fri_encode_tq(fri_plan *p, const void *pixels, uint32_t n_frames, const int32_t *q, int32_t *coefs)
{
    return encode_host(p, pixels, n_frames, q, coefs, false);
}

int fri_encode_tq16(fri_plan *p, const void *pixels, uint32_t n_frames, const int32_t *q, int16_t *coefs)
{
    return encode_host(p, pixels, n_frames, q, coefs, true);
}

static int decode_host(fri_plan *p, const void *coefs, uint32_t n_frames, const int32_t *q, int dequant_mode, void *pixels,
                       bool half)
{
    int rc = enter_device(p);
    if (rc) return rc;
    if ((rc = check_q(q))) return rc;
    if (dequant_mode != FRI_DEQUANT_DIVIDE && dequant_mode != FRI_DEQUANT_MULTIPLY)
        return fail(FRI_E_INVALID, "dequant_mode must be FRI_DEQUANT_DIVIDE or FRI_DEQUANT_MULTIPLY");
    if (n_frames == 0) return FRI_OK;
    if (!pixels || !coefs) return fail(FRI_E_INVALID, "NULL host buffer");
    if ((rc = check_pinned(p, pixels, "pixels")) || (rc = check_pinned(p, coefs, "coefs"))) return rc;
    if ((rc = half ? ensure_slots16(p) : ensure_slots(p))) return rc;
    const Geometry &g = p->plan.geo;
    Pipeline &pl = p->pipe;
    QuantParams qp;
    make_quant_params(qp, q, dequant_mode == FRI_DEQUANT_MULTIPLY);
    p->last_launches = 0;
    Band bands[kMaxBands];
    const int n_bands = make_bands(p, bands);
    const size_t block = (size_t)g.channels << g.depth;
    const bool need_zero = p->plan.pixels_covered != (uint64_t)g.width * g.height;
    for (uint32_t f = 0; f < n_frames; ++f) {
        Slot &s = p->slots[f % kSlots];
        const size_t esz = half ? sizeof(int16_t) : sizeof(int32_t);
        const uint8_t *src = static_cast<const uint8_t *>(coefs) + (size_t)f * g.coefs_per_frame * esz;
        uint8_t *dst = static_cast<uint8_t *>(pixels) + (size_t)f * g.frame_bytes;
        if ((rc = acquire_slot(p, s))) return rc;
        if (need_zero) FRI_CUDA(cudaMemsetAsync(s.d_pixels, 0, (size_t)g.frame_bytes, pl.compute));
        int rows_out = 0;
        for (int k = 0; k < n_bands; ++k) {
            const Band &b = bands[k];
            const size_t c0 = g.sub_bits == 0 ? b.t0 * block : 0;
            const size_t cn = g.sub_bits == 0 ? (b.t1 - b.t0) * block : (size_t)g.coefs_per_frame;
            void *d_dst = half ? static_cast<void *>(s.d_coefs16 + c0) : static_cast<void *>(s.d_coefs + c0);
            FRI_CUDA(cudaMemcpyAsync(d_dst, src + c0 * esz, cn * esz, cudaMemcpyHostToDevice, pl.in));
            FRI_CUDA(cudaEventRecord(pl.in_ready[k], pl.in));
            FRI_CUDA(cudaStreamWaitEvent(pl.compute, pl.in_ready[k], 0));
            const bool native16 = half && g.sub_bits == 0;  // the kernel reads int16 itself at depth 9
            if (half && !native16) FRI_CUDA(launch_unpack16(s.d_coefs16 + c0, s.d_coefs + c0, cn, pl.compute, &p->last_launches));
            const void *d_in = native16 ? static_cast<const void *>(s.d_coefs16) : static_cast<const void *>(s.d_coefs);
            FRI_CUDA(launch_decode(g, p->tables, qp, d_in, native16, 1, s.d_pixels, s.d_dc, pl.compute, &p->last_launches, b.g0, b.g1));
            FRI_CUDA(cudaEventRecord(pl.band_done[k], pl.compute));
            FRI_CUDA(cudaStreamWaitEvent(pl.out, pl.band_done[k], 0));
            if (b.final_rows > rows_out) {  // rows no later band writes
                FRI_CUDA(cudaMemcpyAsync(dst + (size_t)rows_out * g.row_stride,
                                         static_cast<uint8_t *>(s.d_pixels) + (size_t)rows_out * g.row_stride,
                                         (size_t)(b.final_rows - rows_out) * g.row_stride, cudaMemcpyDeviceToHost, pl.out));
                rows_out = b.final_rows;
            }
        }
        FRI_CUDA(cudaEventRecord(s.compute_done, pl.compute));
        FRI_CUDA(cudaEventRecord(s.out_done, pl.out));
        s.used = true;
    }
    return finish_host_call(p);
}

int fri_decode_tq(fri_plan *p, const int32_t *coefs, uint32_t n_frames, const int32_t *q, int dequant_mode, void *pixels)
{
    return decode_host(p, coefs, n_frames, q, dequant_mode, pixels, false);
}

int fri_decode_tq16(fri_plan *p, const int16_t *coefs, uint32_t n_frames, const int32_t *q, int dequant_mode, void *pixels)
{
    return decode_host(p, coefs, n_frames, q, dequant_mode, pixels, true);
}

/* ---- emission order ------------------------------------------------ */
static int ensure_emission(fri_plan *p)
{
    if (!p) return fail(FRI_E_INVALID, "plan is NULL");
    if (p->emit_state == 0) {
        std::string err;
        try {
            err = build_emission_order(p->plan, p->emit_order);
        } catch (const std::bad_alloc &) {
            err = "out of host memory";
        }
        if (!err.empty()) {
            p->emit_state = -1;
            p->emit_error = err;
        } else {
            // drop the slots the reference skips (`if let Some(value)`, entropy_coding.rs:287,298,314)
            const Geometry &g = p->plan.geo;
            std::vector<uint32_t> mask(kTileLeaves / 32);
            std::vector<uint8_t> some((size_t)g.n_fractals * kTileLeaves);
            for (int32_t t = 0; t < g.n_fractals; ++t) {
                if (p->plan.full[t]) {
                    std::fill(some.begin() + (size_t)t * kTileLeaves, some.begin() + (size_t)(t + 1) * kTileLeaves, 1);
                    continue;
                }
                fractal_mask(g.depth, p->plan.centers[2 * t], p->plan.centers[2 * t + 1], g.width, g.height, mask.data());
                for (int i = 0; i < kTileLeaves; ++i) some[(size_t)t * kTileLeaves + i] = (mask[i >> 5] >> (i & 31)) & 1u;
            }
            p->emit_src.clear();
            p->emit_src.reserve(p->emit_order.size());
            for (uint32_t v : p->emit_order)
                if (some[v]) p->emit_src.push_back(v);
            p->emit_state = 1;
        }
    }
    if (p->emit_state < 0) return fail(FRI_E_UNSUPPORTED, "%s", p->emit_error.c_str());
    return FRI_OK;
}

static int ensure_emission_device(fri_plan *p)
{
    int rc = ensure_emission(p);
    if (rc) return rc;
    if (p->emit_device_ready || p->emit_src.empty()) return FRI_OK;
    // Partition the Some slots by the group that owns their tile (a group's tiles are consecutive in
    // plan order), keeping increasing emission index inside every group.
    const Plan &pl = p->plan;
    const Geometry &g = pl.geo;
    const size_t count = p->emit_src.size();
    std::vector<uint32_t> goff((size_t)g.n_groups + 1, 0), dst(count);
    std::vector<uint16_t> loc(count);
    try {
        std::vector<uint32_t> group_of((size_t)g.n_fractals);
        for (int32_t gi = 0; gi < g.n_groups; ++gi) {
            const uint32_t t0 = pl.groups[gi].tile_base, n = (uint32_t)__builtin_popcount(pl.groups[gi].tile_mask);
            for (uint32_t t = t0; t < t0 + n; ++t) group_of[t] = (uint32_t)gi;
        }
        for (uint32_t v : p->emit_src) ++goff[group_of[v >> kBaseDepth] + 1];
        for (int32_t gi = 0; gi < g.n_groups; ++gi) goff[gi + 1] += goff[gi];
        std::vector<uint32_t> fill(goff.begin(), goff.end() - 1);
        for (size_t i = 0; i < count; ++i) {
            const uint32_t v = p->emit_src[i], tile = v >> kBaseDepth, gi = group_of[tile];
            const uint32_t k = fill[gi]++;
            dst[k] = (uint32_t)i;
            loc[k] = (uint16_t)(((tile - pl.groups[gi].tile_base) << kBaseDepth) | (v & (kTileLeaves - 1)));
        }
    } catch (const std::bad_alloc &) {
        return fail(FRI_E_NOMEM, "out of host memory while building the emission tables");
    }
    auto upload = [&](void **d, const void *h, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc(d, bytes);
        return e != cudaSuccess ? e : cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice);
    };
    cudaError_t e = upload(&p->d_emit_goff, goff.data(), goff.size() * sizeof(uint32_t));
    if (e == cudaSuccess) e = upload(&p->d_emit_dst, dst.data(), count * sizeof(uint32_t));
    if (e == cudaSuccess) e = upload(&p->d_emit_loc, loc.data(), count * sizeof(uint16_t));
    if (e != cudaSuccess) {  // all three or none: a later call starts over
        for (void **d : {&p->d_emit_goff, &p->d_emit_dst, &p->d_emit_loc}) {
            if (*d) cudaFree(*d);
            *d = nullptr;
        }
        return cuda_fail(e, "uploading the emission tables");
    }
    p->emit_device_ready = true;
    p->emit_tables.goff = static_cast<const uint32_t *>(p->d_emit_goff);
    p->emit_tables.dst = static_cast<const uint32_t *>(p->d_emit_dst);
    p->emit_tables.loc = static_cast<const uint16_t *>(p->d_emit_loc);
    return FRI_OK;
}

uint64_t fri_plan_emission_count(fri_plan *p)
{
    if (ensure_emission(p)) return 0;
    return (uint64_t)p->emit_src.size();
}

int fri_plan_emission_order(fri_plan *p, uint32_t *order)
{
    if (!order) return fail(FRI_E_INVALID, "NULL argument");
    int rc = ensure_emission(p);
    if (rc) return rc;
    std::memcpy(order, p->emit_order.data(), p->emit_order.size() * sizeof(uint32_t));
    return FRI_OK;
}

// Element type of an emission-ordered stream at the boundary: int32, int16, or zig-zag symbols packed at
// `bits` (10 or 9) bits each.
struct StreamFmt {
    int elem_bytes;  // 4 or 2 (dense streams); 0 = packed
    int bits;        // packed only
    bool packed() const { return elem_bytes == 0; }
    bool half() const { return elem_bytes != 4; }  // the device-side staging of this format is int16
};
static const StreamFmt kFmtI32{4, 0}, kFmtI16{2, 0};
static StreamFmt fmt_packed(int bits) { return StreamFmt{0, bits}; }

// Padded length of one channel's stream inside the device staging of the packed transport: a multiple of the
// 64-symbol packing block.
static size_t padded_count(size_t count) { return (count + kPackBlock - 1) / kPackBlock * kPackBlock; }
static size_t packed_bytes(size_t count, int bits) { return padded_count(count) / kPackBlock * (size_t)pack_block_bytes(bits); }

// Zeroes elements [count, stride) of every one of n_streams int16 streams of padded stride.
static cudaError_t zero_stream_padding(int16_t *streams, size_t count, size_t stride, size_t n_streams, cudaStream_t st)
{
    if (stride == count || n_streams == 0) return cudaSuccess;
    return cudaMemset2DAsync(streams + count, stride * sizeof(int16_t), 0, (stride - count) * sizeof(int16_t), n_streams, st);
}

static int check_stream_fmt(const fri_plan *p, StreamFmt fmt)
{
    if (fmt.packed() && fmt.bits != 9 && fmt.bits != 10)
        return fail(FRI_E_INVALID, "packed streams carry 9 or 10 bits per symbol, not %d", fmt.bits);
    if (fmt.half() && p->plan.geo.sample_bytes != 1)
        return fail(FRI_E_UNSUPPORTED, "16-bit / packed emission needs 8-bit samples (residues of 16-bit samples need 18 bits)");
    return FRI_OK;
}

static int emit_device(fri_plan *p, const int32_t *d_coefs, uint32_t n_frames, void *d_out, StreamFmt fmt, void *stream)
{
    int rc = enter_device(p);
    if (rc) return rc;
    if ((rc = check_stream_fmt(p, fmt))) return rc;
    if ((rc = ensure_emission_device(p))) return rc;
    if (n_frames == 0) return FRI_OK;
    if (!d_coefs || !d_out) return fail(FRI_E_INVALID, "NULL device buffer");
    if ((uintptr_t)d_coefs & 15) return fail(FRI_E_INVALID, "d_coefs must be 16-byte aligned");
    const Geometry &g = p->plan.geo;
    const size_t count = p->emit_src.size();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint32_t launches = 0;
    cudaError_t e;
    if (fmt.packed()) {
        if ((uintptr_t)d_out & 15) return fail(FRI_E_INVALID, "packed streams must be 16-byte aligned");
        // gather into int16 streams of padded stride (stream-ordered scratch), then pack
        const size_t stride = padded_count(count), total = (size_t)n_frames * g.channels * stride;
        int16_t *tmp = nullptr;
        if ((rc = pool_alloc(p, reinterpret_cast<void **>(&tmp), total * sizeof(int16_t), st))) return rc;
        e = zero_stream_padding(tmp, count, stride, (size_t)n_frames * g.channels, st);  // the padding packs as symbol 0
        if (e == cudaSuccess) e = launch_emit(g, p->tables, p->emit_tables, stride, d_coefs, n_frames, tmp, true, st, &launches);
        if (e == cudaSuccess) e = launch_pack_bits(fmt.bits, tmp, static_cast<uint8_t *>(d_out), total / kPackBlock, st, &launches);
        cudaFreeAsync(tmp, st);
    } else {
        e = launch_emit(g, p->tables, p->emit_tables, count, d_coefs, n_frames, d_out, fmt.half(), st, &launches);
    }
    p->last_launches = launches;
    if (e != cudaSuccess) return cuda_fail(e, "emission gather");
    return FRI_OK;
}

int fri_emit_device(fri_plan *p, const int32_t *d_coefs, uint32_t n_frames, int32_t *d_out, void *stream)
{
    return emit_device(p, d_coefs, n_frames, d_out, kFmtI32, stream);
}

int fri_emit_device16(fri_plan *p, const int32_t *d_coefs, uint32_t n_frames, int16_t *d_out, void *stream)
{
    return emit_device(p, d_coefs, n_frames, d_out, kFmtI16, stream);
}

int fri_emit_device10(fri_plan *p, const int32_t *d_coefs, uint32_t n_frames, uint8_t *d_out, void *stream)
{
    return emit_device(p, d_coefs, n_frames, d_out, fmt_packed(10), stream);
}

int fri_emit_device_packed(fri_plan *p, const int32_t *d_coefs, uint32_t n_frames, int bits, uint8_t *d_out, void *stream)
{
    return emit_device(p, d_coefs, n_frames, d_out, fmt_packed(bits), stream);
}

uint64_t fri_plan_emission_packed_bytes(fri_plan *p)
{
    if (ensure_emission(p)) return 0;
    return (uint64_t)packed_bytes(p->emit_src.size(), 10);
}

uint64_t fri_plan_emission_packed_size(fri_plan *p, int bits)
{
    if (bits != 9 && bits != 10) {
        fail(FRI_E_INVALID, "packed streams carry 9 or 10 bits per symbol, not %d", bits);
        return 0;
    }
    if (ensure_emission(p)) return 0;
    return (uint64_t)packed_bytes(p->emit_src.size(), bits);
}

// Device staging of the emission-ordered streams of one frame per slot: sized for int32 streams of padded
// stride (the largest of the formats), plus the packed image of the same streams.
static int ensure_emit_slots(fri_plan *p)
{
    int rc = ensure_slots(p);
    if (rc) return rc;
    const Geometry &g = p->plan.geo;
    const size_t count = p->emit_src.size();
    for (auto &s : p->slots) {
        if (!s.d_emit) FRI_CUDA(cudaMalloc(&s.d_emit, (size_t)g.channels * padded_count(count) * sizeof(int32_t) + 16));
        if (!s.d_pack && g.sample_bytes == 1) FRI_CUDA(cudaMalloc(&s.d_pack, (size_t)g.channels * packed_bytes(count, 10) + 16));
    }
    return FRI_OK;
}

static size_t stream_frame_bytes(const Geometry &g, size_t count, StreamFmt fmt)
{
    return fmt.packed() ? (size_t)g.channels * packed_bytes(count, fmt.bits) : (size_t)g.channels * count * (size_t)fmt.elem_bytes;
}

static int encode_emit_host(fri_plan *p, const void *pixels, uint32_t n_frames, const int32_t *q, void *out, StreamFmt fmt)
{
    int rc = enter_device(p);
    if (rc) return rc;
    if ((rc = check_q(q))) return rc;
    if ((rc = check_stream_fmt(p, fmt))) return rc;
    if ((rc = ensure_emission_device(p))) return rc;
    if (n_frames == 0) return FRI_OK;
    if (!pixels || !out) return fail(FRI_E_INVALID, "NULL host buffer");
    if ((rc = check_pinned(p, pixels, "pixels")) || (rc = check_pinned(p, out, "streams"))) return rc;
    if ((rc = ensure_emit_slots(p))) return rc;
    const Geometry &g = p->plan.geo;
    const size_t count = p->emit_src.size();
    const size_t stride = fmt.packed() ? padded_count(count) : count;  // elements between two streams on the device
    const size_t per_frame = stream_frame_bytes(g, count, fmt);
    QuantParams qp;
    make_quant_params(qp, q, 0);
    p->last_launches = 0;
    Pipeline &pl = p->pipe;
    for (uint32_t f = 0; f < n_frames; ++f) {
        Slot &s = p->slots[f % kSlots];
        if ((rc = acquire_slot(p, s))) return rc;
        FRI_CUDA(cudaMemcpyAsync(s.d_pixels, static_cast<const uint8_t *>(pixels) + (size_t)f * g.frame_bytes,
                                 (size_t)g.frame_bytes, cudaMemcpyHostToDevice, pl.in));
        FRI_CUDA(cudaEventRecord(pl.in_ready[0], pl.in));
        FRI_CUDA(cudaStreamWaitEvent(pl.compute, pl.in_ready[0], 0));
        FRI_CUDA(launch_encode(g, p->tables, qp, s.d_pixels, 1, s.d_coefs, false, s.d_dc, pl.compute, &p->last_launches));
        if (fmt.packed())  // the padding of every stream packs as symbol 0
            FRI_CUDA(zero_stream_padding(static_cast<int16_t *>(s.d_emit), count, stride, (size_t)g.channels, pl.compute));
        FRI_CUDA(launch_emit(g, p->tables, p->emit_tables, stride, s.d_coefs, 1, s.d_emit, fmt.half(), pl.compute,
                             &p->last_launches));
        if (fmt.packed())
            FRI_CUDA(launch_pack_bits(fmt.bits, static_cast<const int16_t *>(s.d_emit), static_cast<uint8_t *>(s.d_pack),
                                      (size_t)g.channels * stride / kPackBlock, pl.compute, &p->last_launches));
        FRI_CUDA(cudaEventRecord(pl.band_done[0], pl.compute));
        FRI_CUDA(cudaEventRecord(s.compute_done, pl.compute));
        FRI_CUDA(cudaStreamWaitEvent(pl.out, pl.band_done[0], 0));
        FRI_CUDA(cudaMemcpyAsync(static_cast<uint8_t *>(out) + (size_t)f * per_frame, fmt.packed() ? s.d_pack : s.d_emit, per_frame,
                                 cudaMemcpyDeviceToHost, pl.out));
        FRI_CUDA(cudaEventRecord(s.out_done, pl.out));
        s.used = true;
    }
    return finish_host_call(p);
}

int fri_encode_tq_emit(fri_plan *p, const void *pixels, uint32_t n_frames, const int32_t *q, int32_t *out)
{
    return encode_emit_host(p, pixels, n_frames, q, out, kFmtI32);
}

int fri_encode_tq_emit16(fri_plan *p, const void *pixels, uint32_t n_frames, const int32_t *q, int16_t *out)
{
    return encode_emit_host(p, pixels, n_frames, q, out, kFmtI16);
}

int fri_encode_tq_emit10(fri_plan *p, const void *pixels, uint32_t n_frames, const int32_t *q, uint8_t *out)
{
    return encode_emit_host(p, pixels, n_frames, q, out, fmt_packed(10));
}

int fri_encode_tq_emit_packed(fri_plan *p, const void *pixels, uint32_t n_frames, const int32_t *q, int bits, uint8_t *out)
{
    return encode_emit_host(p, pixels, n_frames, q, out, fmt_packed(bits));
}

static int unemit_device(fri_plan *p, const void *d_streams, StreamFmt fmt, uint32_t n_frames, int32_t *d_coefs, void *stream)
{
    int rc = enter_device(p);
    if (rc) return rc;
    if ((rc = check_stream_fmt(p, fmt))) return rc;
    if ((rc = ensure_emission_device(p))) return rc;
    if (n_frames == 0) return FRI_OK;
    if (!d_coefs || !d_streams) return fail(FRI_E_INVALID, "NULL device buffer");
    if ((uintptr_t)d_coefs & 15) return fail(FRI_E_INVALID, "d_coefs must be 16-byte aligned");
    const Geometry &g = p->plan.geo;
    const size_t count = p->emit_src.size();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint32_t launches = 0;
    cudaError_t e;
    if (fmt.packed()) {
        if ((uintptr_t)d_streams & 15) return fail(FRI_E_INVALID, "packed streams must be 16-byte aligned");
        const size_t stride = padded_count(count), total = (size_t)n_frames * g.channels * stride;
        int16_t *tmp = nullptr;
        if ((rc = pool_alloc(p, reinterpret_cast<void **>(&tmp), total * sizeof(int16_t), st))) return rc;
        e = launch_unpack_bits(fmt.bits, static_cast<const uint8_t *>(d_streams), tmp, total / kPackBlock, st, &launches);
        if (e == cudaSuccess) e = launch_unemit(g, p->tables, p->emit_tables, stride, tmp, true, n_frames, d_coefs, st, &launches);
        cudaFreeAsync(tmp, st);
    } else {
        e = launch_unemit(g, p->tables, p->emit_tables, count, d_streams, fmt.half(), n_frames, d_coefs, st, &launches);
    }
    p->last_launches = launches;
    if (e != cudaSuccess) return cuda_fail(e, "emission un-gather");
    return FRI_OK;
}

int fri_unemit_device(fri_plan *p, const int32_t *d_streams, uint32_t n_frames, int32_t *d_coefs, void *stream)
{
    return unemit_device(p, d_streams, kFmtI32, n_frames, d_coefs, stream);
}

int fri_unemit_device16(fri_plan *p, const int16_t *d_streams, uint32_t n_frames, int32_t *d_coefs, void *stream)
{
    return unemit_device(p, d_streams, kFmtI16, n_frames, d_coefs, stream);
}

int fri_unemit_device10(fri_plan *p, const uint8_t *d_streams, uint32_t n_frames, int32_t *d_coefs, void *stream)
{
    return unemit_device(p, d_streams, fmt_packed(10), n_frames, d_coefs, stream);
}

int fri_unemit_device_packed(fri_plan *p, const uint8_t *d_streams, uint32_t n_frames, int bits, int32_t *d_coefs, void *stream)
{
    return unemit_device(p, d_streams, fmt_packed(bits), n_frames, d_coefs, stream);
}

static int decode_emit_host(fri_plan *p, const void *streams, StreamFmt fmt, uint32_t n_frames, const int32_t *q, int dequant_mode,
                            void *pixels)
{
    int rc = enter_device(p);
    if (rc) return rc;
    if ((rc = check_q(q))) return rc;
    if (dequant_mode != FRI_DEQUANT_DIVIDE && dequant_mode != FRI_DEQUANT_MULTIPLY)
        return fail(FRI_E_INVALID, "dequant_mode must be FRI_DEQUANT_DIVIDE or FRI_DEQUANT_MULTIPLY");
    if ((rc = check_stream_fmt(p, fmt))) return rc;
    if ((rc = ensure_emission_device(p))) return rc;
    if (n_frames == 0) return FRI_OK;
    if (!pixels || !streams) return fail(FRI_E_INVALID, "NULL host buffer");
    if ((rc = check_pinned(p, pixels, "pixels")) || (rc = check_pinned(p, streams, "streams"))) return rc;
    if ((rc = ensure_emit_slots(p))) return rc;
    const Geometry &g = p->plan.geo;
    const size_t count = p->emit_src.size();
    const size_t stride = fmt.packed() ? padded_count(count) : count;
    const size_t per_frame = stream_frame_bytes(g, count, fmt);
    QuantParams qp;
    make_quant_params(qp, q, dequant_mode == FRI_DEQUANT_MULTIPLY);
    p->last_launches = 0;
    const bool need_zero = p->plan.pixels_covered != (uint64_t)g.width * g.height;
    Pipeline &pl = p->pipe;
    for (uint32_t f = 0; f < n_frames; ++f) {
        Slot &s = p->slots[f % kSlots];
        if ((rc = acquire_slot(p, s))) return rc;
        FRI_CUDA(cudaMemcpyAsync(fmt.packed() ? s.d_pack : s.d_emit, static_cast<const uint8_t *>(streams) + (size_t)f * per_frame,
                                 per_frame, cudaMemcpyHostToDevice, pl.in));
        FRI_CUDA(cudaEventRecord(pl.in_ready[0], pl.in));
        FRI_CUDA(cudaStreamWaitEvent(pl.compute, pl.in_ready[0], 0));
        if (need_zero) FRI_CUDA(cudaMemsetAsync(s.d_pixels, 0, (size_t)g.frame_bytes, pl.compute));
        if (fmt.packed())
            FRI_CUDA(launch_unpack_bits(fmt.bits, static_cast<const uint8_t *>(s.d_pack), static_cast<int16_t *>(s.d_emit),
                                        (size_t)g.channels * stride / kPackBlock, pl.compute, &p->last_launches));
        FRI_CUDA(launch_unemit(g, p->tables, p->emit_tables, stride, s.d_emit, fmt.half(), 1, s.d_coefs, pl.compute,
                               &p->last_launches));
        FRI_CUDA(launch_decode(g, p->tables, qp, s.d_coefs, false, 1, s.d_pixels, s.d_dc, pl.compute, &p->last_launches));
        FRI_CUDA(cudaEventRecord(pl.band_done[0], pl.compute));
        FRI_CUDA(cudaEventRecord(s.compute_done, pl.compute));
        FRI_CUDA(cudaStreamWaitEvent(pl.out, pl.band_done[0], 0));
        FRI_CUDA(cudaMemcpyAsync(static_cast<uint8_t *>(pixels) + (size_t)f * g.frame_bytes, s.d_pixels, (size_t)g.frame_bytes,
                                 cudaMemcpyDeviceToHost, pl.out));
        FRI_CUDA(cudaEventRecord(s.out_done, pl.out));
        s.used = true;
    }
    return finish_host_call(p);
}

int fri_decode_tq_emit(fri_plan *p, const int32_t *streams, uint32_t n_frames, const int32_t *q, int dequant_mode, void *pixels)
{
    return decode_emit_host(p, streams, kFmtI32, n_frames, q, dequant_mode, pixels);
}

int fri_decode_tq_emit16(fri_plan *p, const int16_t *streams, uint32_t n_frames, const int32_t *q, int dequant_mode, void *pixels)
{
    return decode_emit_host(p, streams, kFmtI16, n_frames, q, dequant_mode, pixels);
}

int fri_decode_tq_emit10(fri_plan *p, const uint8_t *streams, uint32_t n_frames, const int32_t *q, int dequant_mode, void *pixels)
{
    return decode_emit_host(p, streams, fmt_packed(10), n_frames, q, dequant_mode, pixels);
}

int fri_decode_tq_emit_packed(fri_plan *p, const uint8_t *streams, uint32_t n_frames, int bits, const int32_t *q, int dequant_mode,
                              void *pixels)
{
    return decode_emit_host(p, streams, fmt_packed(bits), n_frames, q, dequant_mode, pixels);
}

/* ---- prediction + context bucketing, encode side ------------------------ */
static int ensure_lattice(fri_plan *p);

static int ensure_predict_device(fri_plan *p)
{
    int rc = ensure_emission_device(p);
    if (rc) return rc;
    if (p->predict_ready) return FRI_OK;
    if ((rc = ensure_lattice(p))) return rc;
    // the device image of the host predictor's neighbour tables (codec::Predictor, fri_codec.cpp)
    std::vector<uint32_t> steps((size_t)kTileLeaves * 14);
    PredictTables &t = p->predict_tables;
    std::vector<int32_t> adjacent;
    try {
        const codec::Predictor pr(p->lattice, p->plan.centers.data(), p->plan.geo.channels);
        auto pack = [](const codec::Predictor::NodeStep &n) { return n.heap < 0 ? 0xffffu : ((uint32_t)n.heap | (uint32_t)n.cell << 16); };
        for (int h = 0; h < kTileLeaves; ++h) {
            const codec::Predictor::HeapSteps &hs = pr.steps[h];
            uint32_t *o = steps.data() + (size_t)h * 14;
            for (int j = 0; j < 6; ++j) o[j] = pack(hs.regular[j]);
            for (int j = 0; j < 4; ++j) o[6 + j] = pack(hs.alt[j]);
            for (int j = 0; j < 4; ++j) o[10 + j] = pack(hs.probe[j]);
        }
        adjacent = pr.adjacent;
        adjacent.resize((size_t)p->plan.geo.n_fractals * 9, -1);
        for (int j = 0; j < 3; ++j) t.lf_cell[j] = pr.lf_cell[j];
    } catch (const std::bad_alloc &) {
        return fail(FRI_E_NOMEM, "out of host memory while building the prediction tables");
    } catch (const std::logic_error &e) {
        return fail(FRI_E_UNSUPPORTED, "%s", e.what());
    }
    auto upload = [&](void **d, const void *h, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc(d, bytes ? bytes : 4);
        return e != cudaSuccess || bytes == 0 ? e : cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice);
    };
    cudaError_t e = upload(&p->d_pred_adjacent, adjacent.data(), adjacent.size() * sizeof(int32_t));
    if (e == cudaSuccess) e = upload(&p->d_pred_steps, steps.data(), steps.size() * sizeof(uint32_t));
    if (e == cudaSuccess) e = configure_predict_kernel();
    if (e != cudaSuccess) {
        for (void **d : {&p->d_pred_adjacent, &p->d_pred_steps}) {
            if (*d) cudaFree(*d);
            *d = nullptr;
        }
        return cuda_fail(e, "uploading the prediction tables");
    }
    t.adjacent = static_cast<const int32_t *>(p->d_pred_adjacent);
    t.steps = static_cast<const uint32_t *>(p->d_pred_steps);
    p->predict_ready = true;
    return FRI_OK;
}

int fri_predict_device(fri_plan *p, const int32_t *d_coefs, uint32_t n_frames, const float *value_params,
                       const float *width_params, uint8_t *d_bucket, int32_t *d_pred, uint16_t *d_sym, uint32_t *d_hist,
                       uint32_t *d_overflow, void *stream)
{
    int rc = enter_device(p);
    if (rc) return rc;
    if ((rc = ensure_predict_device(p))) return rc;
    if (n_frames == 0) return FRI_OK;
    if (!d_coefs || !d_bucket || !d_pred || !d_sym || !d_hist) return fail(FRI_E_INVALID, "NULL device buffer");
    if (!value_params || !width_params) return fail(FRI_E_INVALID, "the predictor parameters are inputs: pass [C][3][6] floats each");
    const Geometry &g = p->plan.geo;
    PredictParams prm{};
    std::memcpy(prm.value, value_params, sizeof(float) * 18 * g.channels);
    std::memcpy(prm.width, width_params, sizeof(float) * 18 * g.channels);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FRI_CUDA(cudaMemsetAsync(d_hist, 0, (size_t)n_frames * g.channels * 10 * 1024 * sizeof(uint32_t), st));
    if (d_overflow) FRI_CUDA(cudaMemsetAsync(d_overflow, 0, sizeof(uint32_t), st));
    uint32_t launches = 0;
    const cudaError_t e = launch_predict(g, p->tables, p->emit_tables, p->predict_tables, prm, p->emit_src.size(), d_coefs, n_frames,
                                         d_bucket, d_pred, d_sym, d_hist, d_overflow, st, &launches);
    p->last_launches = launches;
    if (e != cudaSuccess) return cuda_fail(e, "launch_predict");
    return FRI_OK;
}

/* ---- host codec behind the transform: parameter fit, context model, rANS, `frif` container
 *      (fri_codec.cpp) ------------------------------------------- */
static int ensure_lattice(fri_plan *p)
{
    int rc = ensure_emission(p);
    if (rc) return rc;
    if (p->lattice_ready) return FRI_OK;
    try {
        build_lattice_index(p->plan, p->lattice);
        const Geometry &g = p->plan.geo;
        p->some_flat.assign((size_t)g.n_fractals * kTileLeaves, 1);
        std::vector<uint32_t> mask(kTileLeaves / 32);
        for (int32_t t = 0; t < g.n_fractals; ++t) {
            if (p->plan.full[t]) continue;
            fractal_mask(g.depth, p->plan.centers[2 * t], p->plan.centers[2 * t + 1], g.width, g.height, mask.data());
            for (int i = 0; i < kTileLeaves; ++i) p->some_flat[(size_t)t * kTileLeaves + i] = (mask[i >> 5] >> (i & 31)) & 1u;
        }
        codec::fit_zero_rows(p->plan, p->some_flat, p->fit_zero_rows);
    } catch (const std::bad_alloc &) {
        return fail(FRI_E_NOMEM, "out of host memory while building the lattice index");
    }
    p->lattice_ready = true;
    return FRI_OK;
}

static int host_threads_hint()
{
    const unsigned n = std::thread::hardware_concurrency();
    return (int)std::max(1u, std::min(n, 16u));
}

/* Predictor parameters of one device-resident frame: the normal equations are summed on the device (two passes
 * of fri_fit_kernel, exact integers), the 6 x 6 systems solved on the host between and after them. */
static int fit_on_device(fri_plan *p, const int32_t *d_coefs, float *value_params, float *width_params, cudaStream_t st)
{
    const Geometry &g = p->plan.geo;
    const int C = g.channels;
    constexpr int kTerms = 27;
    const size_t n = (size_t)C * 3 * kTerms;
    unsigned long long *d_sums = nullptr;
    int rc = pool_alloc(p, reinterpret_cast<void **>(&d_sums), n * sizeof(unsigned long long), st);
    if (rc) return rc;
    std::vector<unsigned long long> sums(n);
    PredictParams prm{};
    auto pass = [&](bool width) -> int {
        FRI_CUDA(cudaMemsetAsync(d_sums, 0, n * sizeof(unsigned long long), st));
        FRI_CUDA(launch_fit(g, p->tables, p->emit_tables, p->predict_tables, prm, width, d_coefs, d_sums, st, &p->last_launches));
        FRI_CUDA(cudaMemcpyAsync(sums.data(), d_sums, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        FRI_CUDA(cudaStreamSynchronize(st));
        return FRI_OK;
    };
    auto solve = [&](bool width, float *out) {
        for (int cs = 0; cs < C * 3; ++cs) {
            codec::FitSums f;
            for (int i = 0; i < 21; ++i) f.a[i] = sums[(size_t)cs * kTerms + i];
            for (int i = 0; i < 6; ++i) f.b[i] = sums[(size_t)cs * kTerms + 21 + i];
            if (width) f.a[0] += p->fit_zero_rows[cs % 3];
            codec::solve_fit(f, width ? (double)codec::kFitScale : 1.0, out + (size_t)cs * 6);
        }
    };
    rc = pass(false);
    if (!rc) {
        solve(false, value_params);
        std::memcpy(prm.value, value_params, sizeof(float) * 18 * C);
        rc = pass(true);
    }
    if (!rc) solve(true, width_params);
    cudaFreeAsync(d_sums, st);
    return rc;
}

int fri_fit_device(fri_plan *p, const int32_t *d_coefs, float *value_params, float *width_params, void *stream)
{
    int rc = enter_device(p);
    if (rc) return rc;
    if (!d_coefs || !value_params || !width_params) return fail(FRI_E_INVALID, "NULL argument");
    if ((rc = ensure_predict_device(p))) return rc;
    if ((rc = ensure_lattice(p))) return rc;
    p->last_launches = 0;
    return fit_on_device(p, d_coefs, value_params, width_params, static_cast<cudaStream_t>(stream));
}

int fri_fit_parameters(fri_plan *p, const int32_t *coefs, float *value_params, float *width_params)
{
    if (!p || !coefs || !value_params || !width_params) return fail(FRI_E_INVALID, "NULL argument");
    int rc = ensure_lattice(p);
    if (rc) return rc;
    try {
        codec::fit_parameters(p->plan, p->lattice, p->some_flat, coefs, value_params, width_params, host_threads_hint());
    } catch (const std::bad_alloc &) {
        return fail(FRI_E_NOMEM, "out of host memory");
    }
    return FRI_OK;
}

int fri_predict_host(fri_plan *p, const int32_t *coefs, const float *value_params, const float *width_params, uint8_t *bucket,
                     int32_t *pred, uint16_t *sym, uint32_t *hist, uint32_t *overflow)
{
    if (!p || !coefs || !value_params || !width_params || !bucket || !pred || !sym || !hist) return fail(FRI_E_INVALID, "NULL argument");
    int rc = ensure_lattice(p);
    if (rc) return rc;
    const Geometry &g = p->plan.geo;
    const int C = g.channels;
    const size_t count = p->emit_src.size();
    const codec::Predictor pr(p->lattice, p->plan.centers.data(), C);
    std::memset(hist, 0, sizeof(uint32_t) * (size_t)C * codec::kContexts * codec::kAlphabet);
    uint32_t over = 0;
    for (int ch = 0; ch < C; ++ch) {
        float vp[3][6], wp[3][6];
        std::memcpy(vp, value_params + (size_t)ch * 18, sizeof(vp));
        std::memcpy(wp, width_params + (size_t)ch * 18, sizeof(wp));
        for (size_t k = 0; k < count; ++k) {
            const uint32_t src = p->emit_src[k];
            const int tile = (int)(src >> kBaseDepth), heap = (int)(src & (kTileLeaves - 1));
            int b;
            int32_t pv;
            if (heap < 2) pr.lf(coefs, tile, heap, ch, b, pv);
            else pr.hf(coefs, tile, heap, ch, vp, wp, b, pv);
            const int32_t value = coefs[(((size_t)tile * C + ch) << kBaseDepth) + heap];
            const uint32_t s = codec::pack_signed((int32_t)((uint32_t)value - (uint32_t)pv));
            bucket[ch * count + k] = (uint8_t)b;
            pred[ch * count + k] = pv;
            sym[ch * count + k] = (uint16_t)std::min<uint32_t>(s, 0xffffu);
            if (s < (uint32_t)codec::kAlphabet) ++hist[((size_t)ch * codec::kContexts + b) * codec::kAlphabet + s];
            else ++over;
        }
    }
    if (overflow) *overflow = over;
    return FRI_OK;
}

static int colorspace_code(const Geometry &g, int colorspace)
{
    if (colorspace == 0) return g.channels == 1 ? 1 : 2;  // Luma / RGB (images.rs:23-29)
    return colorspace;
}

int fri_frv_pack(fri_plan *p, int colorspace, const float *value_params, const float *width_params, const uint8_t *bucket,
                 const uint16_t *sym, const uint32_t *hist, uint8_t **out, size_t *out_len)
{
    if (!p || !value_params || !width_params || !bucket || !sym || !hist || !out || !out_len) return fail(FRI_E_INVALID, "NULL argument");
    *out = nullptr;
    *out_len = 0;
    int rc = ensure_emission(p);
    if (rc) return rc;
    const Geometry &g = p->plan.geo;
    if (g.sample_bytes != 1) return fail(FRI_E_UNSUPPORTED, "the frif container holds 8-bit images (images.rs:84)");
    colorspace = colorspace_code(g, colorspace);
    if (colorspace < 1 || colorspace > 3 || (colorspace == 1) != (g.channels == 1))
        return fail(FRI_E_INVALID, "colorspace must be 1 (Luma, 1 channel), 2 (RGB) or 3 (YCbCr), or 0 for the default");
    const int C = g.channels;
    const size_t count = p->emit_src.size();
    try {
        std::vector<codec::ChannelPayload> payload((size_t)C);
        std::vector<std::string> err((size_t)C);
        std::vector<std::thread> th;
        for (int ch = 0; ch < C; ++ch)
            th.emplace_back([&, ch] {
                std::memcpy(payload[ch].value_params, value_params + (size_t)ch * 18, sizeof(float) * 18);
                std::memcpy(payload[ch].width_params, width_params + (size_t)ch * 18, sizeof(float) * 18);
                err[ch] = codec::entropy_encode_channel(sym + (size_t)ch * count, bucket + (size_t)ch * count, count,
                                                        hist + (size_t)ch * codec::kContexts * codec::kAlphabet, payload[ch]);
            });
        for (auto &t : th) t.join();
        for (int ch = 0; ch < C; ++ch)
            if (!err[ch].empty()) return fail(FRI_E_UNSUPPORTED, "channel %d: %s", ch, err[ch].c_str());
        const std::vector<uint8_t> bytes = codec::serialize((uint32_t)g.height, (uint32_t)g.width, colorspace, payload);
        uint8_t *buf = static_cast<uint8_t *>(std::malloc(bytes.size() ? bytes.size() : 1));
        if (!buf) return fail(FRI_E_NOMEM, "out of host memory");
        std::memcpy(buf, bytes.data(), bytes.size());
        *out = buf;
        *out_len = bytes.size();
    } catch (const std::bad_alloc &) {
        return fail(FRI_E_NOMEM, "out of host memory");
    }
    return FRI_OK;
}

void fri_frv_free(uint8_t *bytes) { std::free(bytes); }

int fri_frv_info(const uint8_t *bytes, size_t len, uint32_t *width, uint32_t *height, uint32_t *channels)
{
    if (!bytes || len < 16 || std::memcmp(bytes, "frif", 4) != 0) return fail(FRI_E_INVALID, "Invalid signature for FRIF image.");
    auto u32 = [&](size_t o) { return (uint32_t)bytes[o] | (uint32_t)bytes[o + 1] << 8 | (uint32_t)bytes[o + 2] << 16 | (uint32_t)bytes[o + 3] << 24; };
    const uint32_t cs = u32(12) >> 30 & 3;
    if (cs == 0) return fail(FRI_E_INVALID, "Invalid metadata");
    if (height) *height = u32(4);
    if (width) *width = u32(8);
    if (channels) *channels = cs == 1 ? 1 : 3;
    return FRI_OK;
}

int fri_frv_unpack(fri_plan *p, const uint8_t *bytes, size_t len, int32_t *coefs)
{
    if (!p || !bytes || !coefs) return fail(FRI_E_INVALID, "NULL argument");
    int rc = ensure_lattice(p);
    if (rc) return rc;
    const Geometry &g = p->plan.geo;
    try {
        uint32_t h = 0, w = 0;
        int cs = 0;
        std::vector<codec::ChannelPayload> payload;
        const std::string perr = codec::deserialize(bytes, len, h, w, cs, payload);
        if (!perr.empty()) return fail(FRI_E_INVALID, "%s", perr.c_str());
        const int C = cs == 1 ? 1 : 3;
        if ((int32_t)h != g.height || (int32_t)w != g.width || C != g.channels)
            return fail(FRI_E_INVALID, "container holds a %ux%ux%d image, the plan is for %dx%dx%d", w, h, C, g.width, g.height, g.channels);
        if ((int)payload.size() != C) return fail(FRI_E_INVALID, "container holds %zu channel(s), expected %d", payload.size(), C);
        std::memset(coefs, 0, sizeof(int32_t) * (size_t)g.coefs_per_frame);  // from_metadata: every covered coefficient Some(0)
        const codec::Predictor pr(p->lattice, p->plan.centers.data(), C);
        std::vector<std::string> err((size_t)C);
        std::vector<std::thread> th;  // channels are independent: a channel's predictor reads its own channel only
        for (int ch = 0; ch < C; ++ch)
            th.emplace_back([&, ch] { err[ch] = codec::entropy_decode_channel(payload[ch], pr, p->emit_src, ch, coefs); });
        for (auto &t : th) t.join();
        for (int ch = 0; ch < C; ++ch)
            if (!err[ch].empty()) return fail(FRI_E_INVALID, "channel %d: %s", ch, err[ch].c_str());
    } catch (const std::bad_alloc &) {
        return fail(FRI_E_NOMEM, "out of host memory");
    }
    return FRI_OK;
}

int fri_frv_encode(fri_plan *p, const void *pixels, const int32_t *q, int colorspace, uint8_t **out, size_t *out_len)
{
    if (!out || !out_len) return fail(FRI_E_INVALID, "NULL argument");
    *out = nullptr;
    *out_len = 0;
    int rc = enter_device(p);
    if (rc) return rc;
    if (!pixels) return fail(FRI_E_INVALID, "NULL host buffer");
    if ((rc = check_q(q))) return rc;
    if ((rc = ensure_predict_device(p))) return rc;
    if ((rc = ensure_lattice(p))) return rc;
    if ((rc = ensure_slots(p))) return rc;
    const Geometry &g = p->plan.geo;
    if (g.sample_bytes != 1) return fail(FRI_E_UNSUPPORTED, "the frif container holds 8-bit images (images.rs:84)");
    const int C = g.channels;
    const size_t count = p->emit_src.size(), n_hist = (size_t)C * codec::kContexts * codec::kAlphabet;
    QuantParams qp;
    make_quant_params(qp, q, 0);
    Pipeline &pl = p->pipe;
    Slot &s = p->slots[0];
    cudaStream_t st = pl.compute;
    if ((rc = acquire_slot(p, s))) return rc;
    p->last_launches = 0;
    try {
        // 1. transform + quantization on the device
        FRI_CUDA(cudaMemcpyAsync(s.d_pixels, pixels, (size_t)g.frame_bytes, cudaMemcpyHostToDevice, st));
        FRI_CUDA(launch_encode(g, p->tables, qp, s.d_pixels, 1, s.d_coefs, false, s.d_dc, st, &p->last_launches));
        // 2. predictor parameters: normal equations summed on the device, solved on the host
        //    (context_modeling.rs:204-214; bit-identical to fri_fit_parameters on the same coefficients)
        std::vector<float> vp((size_t)C * 18), wp((size_t)C * 18);
        if ((rc = fit_on_device(p, s.d_coefs, vp.data(), wp.data(), st))) return rc;
        // 3. prediction + context buckets + histograms on the device
        PredictParams prm{};
        std::memcpy(prm.value, vp.data(), sizeof(float) * 18 * C);
        std::memcpy(prm.width, wp.data(), sizeof(float) * 18 * C);
        // buckets + symbols come back through pinned memory kept with the plan (150 MB for 4096 x 4096 x 3:
        // pageable destinations cost 20x the copy time)
        if (!p->h_symbols) FRI_CUDA(cudaHostAlloc(&p->h_symbols, (size_t)C * count * 3 + 16, cudaHostAllocDefault));
        uint8_t *d_bucket = nullptr;
        int32_t *d_pred = nullptr;  // the predictions themselves are not needed: the symbols carry the residuals
        uint16_t *d_sym = nullptr;
        uint32_t *d_hist = nullptr;
        struct Scratch {  // stream-ordered frees on every way out of this scope
            void **slots[3];
            cudaStream_t st;
            ~Scratch()
            {
                for (void **d : slots)
                    if (*d) cudaFreeAsync(*d, st);
            }
        } scratch{{reinterpret_cast<void **>(&d_bucket), reinterpret_cast<void **>(&d_sym), reinterpret_cast<void **>(&d_hist)}, st};
        if ((rc = pool_alloc(p, reinterpret_cast<void **>(&d_bucket), (size_t)C * count, st))) return rc;
        if ((rc = pool_alloc(p, reinterpret_cast<void **>(&d_sym), (size_t)C * count * sizeof(uint16_t), st))) return rc;
        if ((rc = pool_alloc(p, reinterpret_cast<void **>(&d_hist), (n_hist + 1) * sizeof(uint32_t), st))) return rc;
        uint16_t *sym = static_cast<uint16_t *>(p->h_symbols);
        uint8_t *bucket = reinterpret_cast<uint8_t *>(sym + (size_t)C * count);
        std::vector<uint32_t> hist(n_hist + 1);
        FRI_CUDA(cudaMemsetAsync(d_hist, 0, (n_hist + 1) * sizeof(uint32_t), st));
        FRI_CUDA(launch_predict(g, p->tables, p->emit_tables, p->predict_tables, prm, count, s.d_coefs, 1, d_bucket, d_pred, d_sym,
                                d_hist, d_hist + n_hist, st, &p->last_launches));
        FRI_CUDA(cudaMemcpyAsync(bucket, d_bucket, (size_t)C * count, cudaMemcpyDeviceToHost, st));
        FRI_CUDA(cudaMemcpyAsync(sym, d_sym, (size_t)C * count * sizeof(uint16_t), cudaMemcpyDeviceToHost, st));
        FRI_CUDA(cudaMemcpyAsync(hist.data(), d_hist, hist.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        FRI_CUDA(cudaEventRecord(s.compute_done, st));
        FRI_CUDA(cudaEventRecord(s.out_done, st));
        s.used = true;
        FRI_CUDA(cudaStreamSynchronize(st));
        if (hist[n_hist] != 0)
            return fail(FRI_E_UNSUPPORTED, "%u residual(s) fall outside the 1024-symbol alphabet (the reference panics at entropy_coding.rs:99)",
                        hist[n_hist]);
        // 4. rANS + container on the host
        return fri_frv_pack(p, colorspace, vp.data(), wp.data(), bucket, sym, hist.data(), out, out_len);
    } catch (const std::bad_alloc &) {
        return fail(FRI_E_NOMEM, "out of host memory");
    }
}

int fri_frv_decode(fri_plan *p, const uint8_t *bytes, size_t len, const int32_t *q, int dequant_mode, void *pixels)
{
    int rc = enter_device(p);
    if (rc) return rc;
    if (!bytes || !pixels) return fail(FRI_E_INVALID, "NULL argument");
    // the entropy decoder writes into pinned memory kept with the plan: the upload then runs at link speed
    if (!p->h_dense) FRI_CUDA(cudaHostAlloc(&p->h_dense, (size_t)p->plan.geo.coefs_per_frame * sizeof(int32_t) + 16, cudaHostAllocDefault));
    int32_t *coefs = static_cast<int32_t *>(p->h_dense);
    if ((rc = fri_frv_unpack(p, bytes, len, coefs))) return rc;  // entropy decoding: serial, host (entropy_coding.rs:354-449)
    const bool was_async = p->async_mode;
    p->async_mode = false;  // the staging buffer is reused by the next call: the copies must have finished on return
    rc = fri_decode_tq(p, coefs, 1, q, dequant_mode, pixels);
    p->async_mode = was_async;
    return rc;
}

int fri_host_alloc(void **out, size_t bytes)
{
    if (!out) return fail(FRI_E_INVALID, "out is NULL");
    *out = nullptr;
    FRI_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return FRI_OK;
}

void fri_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

uint32_t fri_plan_last_launches(const fri_plan *p) { return p ? p->last_launches : 0; }

int fri_plan_set_async(fri_plan *p, int on)
{
    if (!p) return fail(FRI_E_INVALID, "plan is NULL");
    p->async_mode = on != 0;
    return FRI_OK;
}

int fri_plan_set_independent_calls(fri_plan *p, int on)
{
    if (!p) return fail(FRI_E_INVALID, "plan is NULL");
    p->independent_calls = on != 0;
    return FRI_OK;
}

int fri_plan_sync(fri_plan *p)
{
    int rc = enter_device(p);
    if (rc) return rc;
    if (!p->slots_ready) return FRI_OK;  // nothing was ever enqueued
    Pipeline &pl = p->pipe;
    FRI_CUDA(cudaStreamSynchronize(pl.out));
    FRI_CUDA(cudaStreamSynchronize(pl.compute));
    FRI_CUDA(cudaStreamSynchronize(pl.in));
    return FRI_OK;
}

int fri_plan_set_bands(fri_plan *p, int bands)
{
    if (!p) return fail(FRI_E_INVALID, "plan is NULL");
    if (bands < 0 || bands > kMaxBands) return fail(FRI_E_INVALID, "bands must be in [0, %d] (0 = automatic)", kMaxBands);
    p->bands = bands;
    return FRI_OK;
}

int32_t fri_quant_divide(int32_t value, int32_t q)
{
    if (q <= 1) return value;
    if ((q & (q - 1)) == 0) {  // the kernels' power-of-two path
        int k = 0;
        while ((1 << k) < q) ++k;
        return trunc_div_pow2(value, k);
    }
    return trunc_div(value, make_div(q));
}

/* the multiply-high path for any q >= 2, powers of two included (for tests) */
int32_t fri_quant_divide_magic(int32_t value, int32_t q) { return q <= 1 ? value : trunc_div(value, make_div(q)); }

int32_t fri_quant_divide_small(int32_t value, int32_t q)
{
    SmallDiv sd;
    if (q <= 1) return value;
    return make_small_div(q, sd) ? trunc_div_small(value, sd) : trunc_div(value, make_div(q));
}

#if FRI_TRACE
int fri_debug_trace(unsigned long long *out, size_t n) { return fri::debug_trace(out, n) == cudaSuccess ? 0 : -2; }
int fri_debug_trace2(unsigned long long *out, size_t n) { return fri::debug_trace2(out, n) == cudaSuccess ? 0 : -2; }
#endif

}  // extern "C"
