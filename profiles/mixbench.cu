// mixbench.cu — ceiling for the path's byte mix with NO transform: a pure streaming kernel that
// reads 1 B and writes 4 B per sample (encode mix) or reads 4 B and writes 1 B (decode mix),
// fully coalesced 128-bit accesses.  Context for the roofline fractions in DESIGN.md.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mixbench mixbench.cu && ./mixbench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int U>
__global__ void widen(const uint32_t *__restrict__ in, int4 *__restrict__ out, size_t n4)
{
    size_t i = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x;
    uint32_t v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = i + (size_t)u * blockDim.x < n4 ? __ldcs(in + i + (size_t)u * blockDim.x) : 0;
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (i + (size_t)u * blockDim.x < n4)
            __stcs(out + i + (size_t)u * blockDim.x, make_int4(v[u] & 255, (v[u] >> 8) & 255, (v[u] >> 16) & 255, v[u] >> 24));
}

template <int U>
__global__ void narrow(const int4 *__restrict__ in, uint32_t *__restrict__ out, size_t n4)
{
    size_t i = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x;
    int4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = i + (size_t)u * blockDim.x < n4 ? __ldcs(in + i + (size_t)u * blockDim.x) : make_int4(0, 0, 0, 0);
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (i + (size_t)u * blockDim.x < n4)
            __stcs(out + i + (size_t)u * blockDim.x, (uint32_t)(v[u].x & 255) | (uint32_t)(v[u].y & 255) << 8 | (uint32_t)(v[u].z & 255) << 16 | (uint32_t)v[u].w << 24);
}

int main()
{
    const size_t sizes[2] = {50331648, (size_t)8 * 3840 * 2160 * 3};  // samples: 4096^2 RGB, 8 x 4K RGB
    for (size_t n : sizes) {
        const size_t n4 = n / 4;
        const int sets = 6;
        uint32_t *b8[sets]; int4 *b32[sets];
        for (int s = 0; s < sets; ++s) { cudaMalloc(&b8[s], n); cudaMalloc(&b32[s], n * 4); cudaMemset(b8[s], 1, n); cudaMemset(b32[s], 1, n * 4); }
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        constexpr int U = 4;
        const int threads = 256;
        const unsigned blocks = (unsigned)((n4 + threads * U - 1) / (threads * U));
        for (int mode = 0; mode < 2; ++mode) {
            float best = 1e9f, sum = 0;
            const int reps = 30;
            for (int r = -3; r < reps; ++r) {
                const int s = (r + 3) % sets;
                cudaEventRecord(a);
                if (mode == 0) widen<U><<<blocks, threads>>>(b8[s], b32[s], n4);
                else narrow<U><<<blocks, threads>>>(b32[s], b8[s], n4);
                cudaEventRecord(b);
                cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b);
                if (r >= 0) { best = ms < best ? ms : best; sum += ms; }
            }
            printf("%s mix, %zu samples (%.0f MB): avg %.1f us = %.0f GB/s, best %.1f us = %.0f GB/s\n", mode == 0 ? "encode (1B rd + 4B wr)" : "decode (4B rd + 1B wr)",
                   n, 5.0 * n / 1e6, 1e3 * sum / reps, 5.0 * n / (sum / reps * 1e-3) / 1e9, 1e3 * best, 5.0 * n / (best * 1e-3) / 1e9);
        }
        for (int s = 0; s < sets; ++s) { cudaFree(b8[s]); cudaFree(b32[s]); }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
