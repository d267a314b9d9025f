// fri_kernels.cuh — launch interface between the C ABI (fri_api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>

#include "fri_plan.h"

namespace fri {

// Truncating division by a run-time constant (Rust `i32 / i32` with a positive divisor,
// quantization.rs:19 and :37) as a multiply-high + shifts, exact for every i32 numerator.
// Granlund & Montgomery / libdivide style: magic == 0 -> divisor is a power of two.
struct Div {
    uint32_t magic;
    uint32_t more;  // bits 0..4: shift; bit 6: the "add" fix-up step is needed
};
constexpr uint32_t kDivAdd = 0x40u;

#if defined(__CUDACC__)
#define FRI_HDI __host__ __device__ __forceinline__
#else
#define FRI_HDI inline
#endif

FRI_HDI uint32_t mulhi_u32(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

FRI_HDI int32_t trunc_div(int32_t d, Div dv)
{
    const uint32_t n = d < 0 ? 0u - (uint32_t)d : (uint32_t)d;
    uint32_t r;
    if (dv.magic == 0) {
        r = n >> (dv.more & 31u);
    } else {
        uint32_t t = mulhi_u32(n, dv.magic);
        if (dv.more & kDivAdd) t = ((n - t) >> 1) + t;
        r = t >> (dv.more & 31u);
    }
    return d < 0 ? (int32_t)(0u - r) : (int32_t)r;
}

// Quantization matrix prepared on the host (quantization.rs:3-25).
struct QuantParams {
    uint32_t magic[32];
    uint8_t more[32];
    int32_t q[32];
    uint32_t active;   // bit l set <=> q[l] != 1
    int32_t multiply;  // decode only: 1 = multiply (true dequantizer), 0 = divide (reference)
    FRI_HDI Div div(int l) const { return Div{magic[l], more[l]}; }
};

Div make_div(int32_t q);
void make_quant_params(QuantParams &qp, const int32_t *q, int multiply);

constexpr int kThreads = 256;       // 8 warps per CTA
constexpr int kWarps = kThreads / 32;
constexpr int kScratchInts = 64;    // per-warp exchange buffer for the top 64 coefficients

size_t kernel_smem_bytes(const Geometry &g);

// Device-side tables of a plan.
struct DeviceTables {
    const GroupDesc *groups = nullptr;
    const uint32_t *tile_unit = nullptr;  // nullptr at depth 9
    const uint32_t *ownership = nullptr;
};

// One-time per-process kernel attribute setup (max dynamic shared memory).
cudaError_t configure_kernels();

// Enqueue the fused forward transform + quantization for n_frames frames.
//   d_dc: scratch for the base tiles' low-pass roots, [n_frames][n_fractals][C][2^sub_bits]
//         (depth > 9 only; finished by launch_coarse_forward).
cudaError_t launch_encode(const Geometry &g, const DeviceTables &t, const QuantParams &qp, const void *d_pixels,
                          uint32_t n_frames, int32_t *d_coefs, int32_t *d_dc, cudaStream_t stream,
                          uint32_t *launches);
// Enqueue the fused dequantization + inverse transform + scatter.
cudaError_t launch_decode(const Geometry &g, const DeviceTables &t, const QuantParams &qp, const int32_t *d_coefs,
                          uint32_t n_frames, void *d_pixels, int32_t *d_dc, cudaStream_t stream,
                          uint32_t *launches);

}  // namespace fri
