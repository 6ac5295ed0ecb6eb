// mw_direct_kernels.cuh -- the direct-sum frame for grids the transform identity does not cover (sm_100a).
//
// FFTMesh.Displacement (Scripts/FFTMesh.cs:192-220) evaluates S(x) = sum_{n,m} htilde(t,n,m) e^{i k.x} literally.  That sum
// is a DFT only on periodic, even, power-of-two grids (length == resolution * unitWidth) -- and the
// reference's own FFT Mesh demo scene is not one (resolution 12, length 12.39: Demo/FFT Mesh.unity:147,150).  So that the
// shipped scene runs through the engine, small grids of ANY resolution and length take this path: the same O(N^2)-per-vertex
// sum, on the GPU, one thread block per vertex.  It is a GPU kernel, not a fallback to the host: nothing here runs on the CPU.
//
//   k_direct_htilde      htilde(t, n, m) for the whole grid            (FFTMesh.cs:178-190), once per frame instead of N^2 times
//   k_direct_displace    Displacement(x, t, out nor) per vertex        (:192-220) + the vertex update of EvaluateWaves (:243-247)
//   k_direct_whitecap    forward-difference Jacobian + smoothstep      (:253-276)
//
// Arithmetic: fp32 storage and operation order as in the source; Mathf.Cos / Sin are "double libm, then round", which is
// what the double-precision sincos below gives.  The only liberty is the summation order (a block-wide tree instead of the
// reference's sequential loop): the terms are the same fp32 numbers.
#pragma once
#include "mw_ocean_kernels.cuh"

namespace mwk {

constexpr int DIRECT_THREADS = 128;

// wave-vector component of Displacement: 2 * PI * (i - resolution / 2.0f) / length   (FFTMesh.cs:201, :204), source order
__device__ __forceinline__ float direct_k(int i, int N, float length)
{
    return __fdiv_rn(__fmul_rn(__fmul_rn(2.0f, MW_PI_F), __fsub_rn((float)i, __fdiv_rn((float)N, 2.0f))), length);
}

__global__ void k_direct_htilde(const float4* __restrict__ spec, const float* __restrict__ omega, float2* __restrict__ H,
                                int64_t n2, int tiles, float t)
{
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n2 * tiles) return;
    const float4 s = spec[gid];
    const float omegat = __fmul_rn(omega[gid % n2], t);                 // :183
    const float cs = (float)cos((double)omegat), sn = (float)sin((double)omegat);
    // :188, literal operation order: c0 = (cos, sin), c1 = (cos, -sin)
    const float rx = __fsub_rn(__fadd_rn(__fsub_rn(__fmul_rn(s.x, cs), __fmul_rn(s.y, sn)), __fmul_rn(s.z, cs)), __fmul_rn(s.w, -sn));
    const float ry = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(s.x, sn), __fmul_rn(s.y, cs)), __fmul_rn(s.z, -sn)), __fmul_rn(s.w, cs));
    H[gid] = make_float2(rx, ry);
}

// one block per vertex; blockIdx.y = tile
__global__ void __launch_bounds__(DIRECT_THREADS) k_direct_displace(const float2* __restrict__ H, float* __restrict__ height,
                                                                    float2* __restrict__ disp, float* __restrict__ normal, int N,
                                                                    float unit_width, float length)
{
    const int v = blockIdx.x, tile = blockIdx.y;
    const int n2 = N * N;
    // rest position of the vertex (GenerateMesh, FFTMesh.cs:104-112): x = vertices[index].xz
    const int vi = v / N, vj = v % N, half = N / 2;
    const float off = (N % 2 == 0) ? __fdiv_rn(unit_width, 2.0f) : 0.0f;
    const float px = __fadd_rn(__fmul_rn((float)(vi - half), unit_width), off);
    const float pz = __fadd_rn(__fmul_rn((float)(vj - half), unit_width), off);
    const float2* Ht = H + (size_t)tile * n2;
    float hx = 0.f, nx = 0.f, nz = 0.f, dx = 0.f, dz = 0.f;
    for (int idx = threadIdx.x; idx < n2; idx += DIRECT_THREADS) {
        const int i = idx / N, j = idx % N;
        const float kx = direct_k(i, N, length), kz = direct_k(j, N, length);
        const float k_length = __fsqrt_rn(__fadd_rn(__fmul_rn(kx, kx), __fmul_rn(kz, kz)));       // Vector2.magnitude
        const float kDotX = __fadd_rn(__fmul_rn(kx, px), __fmul_rn(kz, pz));                       // Vector2.Dot (:206)
        double sd, cd;
        sincos((double)kDotX, &sd, &cd);
        const float c = (float)cd, s = (float)sd;                                                  // :207
        const float2 h = Ht[idx];
        const float hcx = __fsub_rn(__fmul_rn(h.x, c), __fmul_rn(h.y, s));                         // :209
        const float hcy = __fadd_rn(__fmul_rn(h.x, s), __fmul_rn(h.y, c));
        hx += hcx;                                                                                 // :210 (h.x is the height)
        nx += __fmul_rn(-kx, hcy);                                                                 // :211
        nz += __fmul_rn(-kz, hcy);
        if (k_length < MW_EPSILON_F) continue;                                                     // :212-213
        dx += __fmul_rn(__fdiv_rn(kx, k_length), hcy);                                             // :214
        dz += __fmul_rn(__fdiv_rn(-kz, k_length), hcy);
    }
    // block-wide sums
    __shared__ float red[5][DIRECT_THREADS / 32];
    float vals[5] = {hx, nx, nz, dx, dz};
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        float x = vals[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = x;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            float x = 0.f;
            for (int w = 0; w < DIRECT_THREADS / 32; ++w) x += red[q][w];
            tot[q] = x;
        }
        const size_t o = (size_t)tile * n2 + v;
        if (height) height[o] = tot[0];
        if (disp) disp[o] = make_float2(tot[3], tot[4]);                                           // hds (:247)
        if (normal) {
            // nor = Vector3.Normalize(Vector3.up - n)   (:218); Normalize returns zero below 1e-5
            const float ax = -tot[1], ay = 1.0f, az = -tot[2];
            const float len = sqrtf(ax * ax + ay * ay + az * az);
            const float inv = len > 1e-5f ? 1.0f / len : 0.0f;
            normal[3 * o + 0] = ax * inv; normal[3 * o + 1] = ay * inv; normal[3 * o + 2] = az * inv;
        }
    }
}

// FFTMesh.cs:253-276 for one tile set: Jacobian of the forward differences, noise from the normal, smoothstep
__global__ void k_direct_whitecap(const float2* __restrict__ disp, const float* __restrict__ normal, float* __restrict__ whitecap,
                                  float* __restrict__ jacobian, int N, int tiles)
{
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n2 = (int64_t)N * N;
    if (gid >= n2 * tiles) return;
    const int idx = (int)(gid % n2);
    const int i = idx / N, j = idx % N;
    const float2 d = disp[gid];
    float2 ddx = make_float2(0.f, 0.f), ddy = make_float2(0.f, 0.f);
    if (i != N - 1) { const float2 e = disp[gid + N]; ddx = make_float2(0.5f * (d.x - e.x), 0.5f * (d.y - e.y)); }   // :260-263
    if (j != N - 1) { const float2 e = disp[gid + 1]; ddy = make_float2(0.5f * (d.x - e.x), 0.5f * (d.y - e.y)); }   // :264-267
    const float jac = __fsub_rn(__fmul_rn(1.0f + ddx.x, 1.0f + ddy.y), __fmul_rn(ddx.y, ddy.x));                     // :268
    if (jacobian) jacobian[gid] = jac;
    if (whitecap) {
        const float ax = fabsf(normal[3 * gid + 0]) * 0.3f, az = fabsf(normal[3 * gid + 2]) * 0.3f;                  // :269
        float turb = fmaxf(1.0f - jac + sqrtf(ax * ax + az * az), 0.0f);                                             // :270
        turb = fminf(turb, 1.0f);                                                                                     // SmoothStep clamps its argument
        whitecap[gid] = turb * turb * (3.0f - 2.0f * turb);                                                           // :273
    }
}

}  // namespace mwk
