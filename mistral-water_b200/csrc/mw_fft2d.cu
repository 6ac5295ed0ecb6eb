// mw_fft2d.cu -- the engine's 2-D transform exposed on caller data (mw_fft2d in mistral_ocean.h).
//
// Functionally this is the reference's Stockham blit chain (Shaders/FFT/Stockham.shader:31-57,
// Scripts/OceanRenderer.cs:229-262: log2 N horizontal radix-2 stages, then log2 N vertical ones) on
// one complex field: sign = -1 reproduces it; sign = +1 is the conjugate transform the FFTMesh
// synthesis uses.  It shares mwfft::fft_line with the ocean kernels, so a parity check of this entry
// point against numpy / the literal stage-by-stage restatement pins the FFT core itself.
#include <vector>
#include "mw_fft.cuh"

namespace {

using mwfft::Plan;
using mwfft::pad_idx;

// LR packed lines per CTA; packed line l holds rows 2l and 2l+1 (contiguous along the transform direction).
template <int N, int LR, int SIGN>
__global__ void __launch_bounds__(LR * (N / 16)) k_fft_rows(const float2* __restrict__ in, float2* __restrict__ out,
                                                           const float2* __restrict__ gtw, int lines_total)
{
    using P = Plan<N>;
    constexpr int T = P::T;
    constexpr int LP = mwfft::line_pitch(N, 8);
    extern __shared__ float4 smem4[];
    float4* tw2 = smem4;
    float2* tw3 = reinterpret_cast<float2*>(smem4 + P::TW2_F4);
    const int lr = threadIdx.x / T, g = threadIdx.x % T;
    const int ln = blockIdx.x * LR + lr;
    const bool active = ln < lines_total;
    float4* line = smem4 + P::TW_BYTES / 16 + lr * LP;
    mwfft::load_twiddles<N, SIGN>(tw2, tw3, gtw);
    const float2* s0 = in + (size_t)(2 * ln) * N;
    const float2* s1 = s0 + N;
    if (active) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const int i = g + T * c;
            const float2 a = s0[i], b = s1[i];
            line[pad_idx(i)] = make_float4(a.x, b.x, a.y, b.y);
        }
    }
    __syncthreads();
    float2* d0 = out + (size_t)(2 * ln) * N;
    float2* d1 = d0 + N;
    mwfft::fft_line<N, SIGN>(line, g, lr, active, tw2, tw3, [&](int idx, int, mwfft::cpk v) {
        d0[idx] = make_float2(v.re.x, v.im.x);
        d1[idx] = make_float2(v.re.y, v.im.y);
    });
}

// Slab of 8 columns per CTA (4 packed lines): transposing load, FFT along the strided direction,
// transposing store.
template <int N, int SIGN>
__global__ void __launch_bounds__(4 * (N / 16)) k_fft_cols(const float2* __restrict__ in, float2* __restrict__ out,
                                                          const float2* __restrict__ gtw)
{
    using P = Plan<N>;
    constexpr int T = P::T;
    constexpr int MAIN = 4 * T;
    constexpr int LP = mwfft::line_pitch(N, 4);
    extern __shared__ float4 smem4[];
    float4* tw2 = smem4;
    float2* tw3 = reinterpret_cast<float2*>(smem4 + P::TW2_F4);
    float4* lines = smem4 + P::TW_BYTES / 16;
    const int tid = threadIdx.x;
    const int q = tid / T, g = tid % T;
    const int b0 = blockIdx.x * 8;
    const size_t base = (size_t)blockIdx.y * N * N;
    mwfft::load_twiddles<N, SIGN>(tw2, tw3, gtw);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int e = tid + k * MAIN;
        const int n = e >> 2, c2 = e & 3;
        const float4 v = *reinterpret_cast<const float4*>(in + base + (size_t)n * N + b0 + 2 * c2);
        lines[c2 * LP + pad_idx(n)] = make_float4(v.x, v.z, v.y, v.w);
    }
    __syncthreads();
    float4* line = lines + q * LP;
    mwfft::fft_line<N, SIGN>(line, g, q, true, tw2, tw3, [&](int, int pidx, mwfft::cpk v) {
        line[pidx] = make_float4(v.re.x, v.im.x, v.re.y, v.im.y);
    });
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int e = tid + k * MAIN;
        const int n = e >> 2, c2 = e & 3;
        *reinterpret_cast<float4*>(out + base + (size_t)n * N + b0 + 2 * c2) = lines[c2 * LP + pad_idx(n)];
    }
}

template <int N, int SIGN>
int run2d(int batch, const float2* d_in, float2* d_tmp, float2* d_out, const float2* d_tw, cudaStream_t st)
{
    constexpr int T = N / 16;
    constexpr int LR = (T >= 64) ? 2 : (128 / T);
    constexpr size_t smem_r = Plan<N>::TW_BYTES + (size_t)LR * mwfft::line_pitch(N, 8) * sizeof(float4);
    constexpr size_t smem_c = Plan<N>::TW_BYTES + (size_t)4 * mwfft::line_pitch(N, 4) * sizeof(float4);
    MW_CUDA(cudaFuncSetAttribute(k_fft_rows<N, LR, SIGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r));
    MW_CUDA(cudaFuncSetAttribute(k_fft_cols<N, SIGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
    const int lines_total = batch * N / 2;
    k_fft_rows<N, LR, SIGN><<<(lines_total + LR - 1) / LR, LR * T, smem_r, st>>>(d_in, d_tmp, d_tw, lines_total);
    MW_LAUNCH_CHECK();
    k_fft_cols<N, SIGN><<<dim3(N / 8, batch), 4 * T, smem_c, st>>>(d_tmp, d_out, d_tw);
    MW_LAUNCH_CHECK();
    return MW_OK;
}

template <int SIGN>
int dispatch(int n, int batch, const float2* d_in, float2* d_tmp, float2* d_out, const float2* d_tw, cudaStream_t st)
{
    switch (n) {
        case 32: return run2d<32, SIGN>(batch, d_in, d_tmp, d_out, d_tw, st);
        case 64: return run2d<64, SIGN>(batch, d_in, d_tmp, d_out, d_tw, st);
        case 128: return run2d<128, SIGN>(batch, d_in, d_tmp, d_out, d_tw, st);
        case 256: return run2d<256, SIGN>(batch, d_in, d_tmp, d_out, d_tw, st);
        case 512: return run2d<512, SIGN>(batch, d_in, d_tmp, d_out, d_tw, st);
        case 1024: return run2d<1024, SIGN>(batch, d_in, d_tmp, d_out, d_tw, st);
        case 2048: return run2d<2048, SIGN>(batch, d_in, d_tmp, d_out, d_tw, st);
    }
    mw_set_error("mw_fft2d: n must be a power of two in [32, 2048], got %d", n);
    return MW_E_INVALID_ARG;
}

}  // namespace

extern "C" int mw_fft2d(int device, int32_t n, int32_t batch, int sign, const float* in, float* out)
{
    if (!in || !out || batch < 1 || (sign != 1 && sign != -1)) {
        mw_set_error("mw_fft2d: bad argument (null buffer, batch < 1 or sign not +-1)");
        return MW_E_INVALID_ARG;
    }
    if (n < 32 || n > 2048 || (n & (n - 1))) {
        mw_set_error("mw_fft2d: n must be a power of two in [32, 2048], got %d", n);
        return MW_E_INVALID_ARG;
    }
    MW_CUDA(cudaSetDevice(device));
    const size_t total = (size_t)batch * n * n;
    float2 *d_a = nullptr, *d_b = nullptr, *d_tw = nullptr;
    std::vector<float2> tw(n);
    const double PI_D = 3.14159265358979323846;
    for (int x = 0; x < n; ++x) tw[x] = make_float2((float)cos(2.0 * PI_D * x / n), (float)sin(2.0 * PI_D * x / n));
    int rc = MW_OK;
    cudaStream_t st = nullptr;
    auto cleanup = [&]() {
        if (d_a) cudaFree(d_a);
        if (d_b) cudaFree(d_b);
        if (d_tw) cudaFree(d_tw);
        if (st) cudaStreamDestroy(st);
    };
#define MW_TRY(expr)                                                                                   \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess) {                                                                       \
            mw_set_error("%s failed: %s", #expr, cudaGetErrorString(_e));                              \
            cleanup();                                                                                 \
            return _e == cudaErrorMemoryAllocation ? MW_E_OOM : MW_E_CUDA;                             \
        }                                                                                              \
    } while (0)
    MW_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    MW_TRY(cudaMalloc((void**)&d_a, total * sizeof(float2)));
    MW_TRY(cudaMalloc((void**)&d_b, total * sizeof(float2)));
    MW_TRY(cudaMalloc((void**)&d_tw, n * sizeof(float2)));
    MW_TRY(cudaMemcpyAsync(d_tw, tw.data(), n * sizeof(float2), cudaMemcpyHostToDevice, st));
    MW_TRY(cudaMemcpyAsync(d_a, in, total * sizeof(float2), cudaMemcpyHostToDevice, st));
    // rows: a -> b ; columns: b -> a
    rc = sign > 0 ? dispatch<+1>(n, batch, d_a, d_b, d_a, d_tw, st) : dispatch<-1>(n, batch, d_a, d_b, d_a, d_tw, st);
    if (rc == MW_OK) {
        MW_TRY(cudaMemcpyAsync(out, d_a, total * sizeof(float2), cudaMemcpyDeviceToHost, st));
        MW_TRY(cudaStreamSynchronize(st));
    }
    cleanup();
    return rc;
}
