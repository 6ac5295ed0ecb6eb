// mw_layout.cuh -- layout of the intermediate between the two transform passes (shared by the FFTMesh-convention
// kernels, mw_ocean_kernels.cuh, and the OceanRenderer-convention kernels, mw_renderer_kernels.cuh).
#pragma once
#include "mw_fft.cuh"

namespace mwk {

// packed points per thread of the FFT engine.  32 (radix 32 x 32: ONE shared-memory exchange per 1024-point transform,
// one warp per line, no named barriers) is implemented and parity-green, but measured equal-to-slower than 16
// (radix 16 x 16 x 4, two exchanges) at N = 1024 on B200: 427 vs 423 us per 16-tile frame (profiles/r01_summary.md).
// Build with -DMW_PTS_1024=32 to select it.
#ifndef MW_PTS_1024
#define MW_PTS_1024 16
#endif
__host__ __device__ constexpr int fft_pts(int N) { return N == 1024 ? MW_PTS_1024 : 16; }

#ifndef MW_SLABW_1024
#define MW_SLABW_1024 8
#endif
__host__ __device__ constexpr int slab_w(int N) { return N < 1024 ? 8 : (N == 1024 ? MW_SLABW_1024 : 4); }
__host__ __device__ constexpr size_t xab_index(int N, int n, int b)
{
    return ((size_t)(b / slab_w(N)) * N + n) * slab_w(N) + (b % slab_w(N));
}
__host__ __device__ constexpr size_t xc_index(int N, int n, int b)
{
    return ((size_t)(b / (2 * slab_w(N))) * N + n) * (2 * slab_w(N)) + (b % (2 * slab_w(N)));
}
__host__ __device__ constexpr size_t xab_tile_elems(int N) { return (size_t)N * N; }
// FFTMesh-convention C field (float2 units): slabs of 4 W columns; inside a row the columns are ordered so that the 16-byte
// pair (X[4c], X[4c+1]) of line c sits at float4 position c and (X[4c+2], X[4c+3]) at W + c -- what pass 2 loads with two
// coalesced instructions -- while pass 1's 32 consecutive columns still fill one contiguous row.
__host__ __device__ constexpr unsigned xc4_pos(int N, int jj)  // jj = column within the slab
{
    return (unsigned)(((jj % 4) / 2) * (2 * slab_w(N)) + (jj / 4) * 2 + (jj % 2));
}
__host__ __device__ constexpr size_t xc4_index(int N, int n, int b)
{
    return ((size_t)(b / (4 * slab_w(N))) * N + n) * (4 * slab_w(N)) + xc4_pos(N, b % (4 * slab_w(N)));
}

}  // namespace mwk
