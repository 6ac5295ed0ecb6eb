// Microbenchmark: how fast can the column pass's output pattern be written at all?  Developer tool.
// Emulates k_cols_extract's stores (N=1024, 16 tiles): CTA = 4 columns x 1024 rows; thread (g = tid>>2, c = tid&3)
// writes rows g + 64b + 256rr; per row group of 4 lanes: hds 32 B, normal 48 B, whitecap 16 B.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int N = 1024;
template <int W, int MODE>   // W = columns per CTA (4 or 8 or 16); MODE 0: slab pattern, 1: fully contiguous
__global__ void k(float2* disp, float* normal, float* white, int tiles)
{
    const int tid = threadIdx.x;
    constexpr int LOGW = W == 4 ? 2 : (W == 8 ? 3 : 4);
    const int c = tid & (W - 1), g = tid >> LOGW;          // blockDim = W * 64
    const size_t obase = (size_t)blockIdx.y * N * N;
    const int b0 = blockIdx.x * W;
    const float v = tid * 0.001f;
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        const int ar = g + 64 * (s & 3) + 256 * (s >> 2);
        size_t o = obase + (size_t)ar * N + b0 + c;
        if (MODE == 1) o = obase + (size_t)blockIdx.x * (W * N) + s * (W * 64) + tid;
        disp[o] = make_float2(v, v + s);
        white[o] = v;
        if (MODE == 0) {
            float4* dst = reinterpret_cast<float4*>(normal + 3 * (o - (c & 3))) + (c & 3);
            if ((c & 3) < 3) *dst = make_float4(v, v, v, v);
        } else {
            float4* dst = reinterpret_cast<float4*>(normal + 3 * (o - tid)) + tid;
            if (tid < (W * 64 * 3) / 4) *dst = make_float4(v, v, v, v);
        }
    }
}
template <int W, int MODE>
void run(const char* name, float2* d, float* n, float* w, int tiles)
{
    dim3 grid(N / W, tiles);
    k<W, MODE><<<grid, W * 64>>>(d, n, w, tiles);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) k<W, MODE><<<grid, W * 64>>>(d, n, w, tiles);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    const double bytes = (double)tiles * N * N * 24;
    printf("%-44s %7.1f us  %7.1f GB/s   (%s)\n", name, ms * 1e3, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    const int tiles = 16;
    float2* d; float *n, *w;
    cudaMalloc(&d, (size_t)tiles * N * N * 8); cudaMalloc(&n, (size_t)tiles * N * N * 12); cudaMalloc(&w, (size_t)tiles * N * N * 4);
    run<4, 0>("4-column slabs (as k_cols_extract)", d, n, w, tiles);
    run<8, 0>("8-column slabs", d, n, w, tiles);
    run<16, 0>("16-column slabs", d, n, w, tiles);
    run<4, 1>("fully contiguous (wrong place), 256 thr", d, n, w, tiles);
    run<16, 1>("fully contiguous (wrong place), 1024 thr", d, n, w, tiles);
    return 0;
}
