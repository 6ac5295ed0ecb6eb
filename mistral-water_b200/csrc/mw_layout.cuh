// mw_layout.cuh -- layout of the intermediate between the two transform passes (shared by the FFTMesh-convention
// kernels, mw_ocean_kernels.cuh, and the OceanRenderer-convention kernels, mw_renderer_kernels.cuh).
#pragma once
#include "mw_fft.cuh"

namespace mwk {

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

}  // namespace mwk
