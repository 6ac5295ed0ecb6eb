// mw_gerstner.cu -- pond renderer: Gerstner sum-of-waves vertex displacement (sm_100a).
//
// Replaces the per-vertex work of Shaders/MistralWaterLib.cginc Gerstner (:71-99) /
// GerstnerLevelOne (:101-125) as dispatched by Displacement (:154-180), generalised to a W-wave
// table (both reference variants are special cases; see mw_gerstner_from_material /
// mw_gerstner_append_level_one).
//
// Roofline note: 24 B/vertex of traffic against 2 transcendentals + ~12 flops per
// wave; at W = 32 this kernel is bound by the MUFU/FP32 pipes, not by HBM.
#include <string.h>
#include "mw_common.cuh"

namespace {

struct GerstnerTable {
    int n_waves;
    mw_gerstner_wave w[MW_GERSTNER_MAX_WAVES];
};

// sin/cos of an fp32 phase: explicit 3-term Cody-Waite reduction to [-pi, pi] followed by the
// MUFU approximations (abs. error ~5e-7 on the reduced range).  |theta| stays far below 2^17 here.
__device__ __forceinline__ void sincos_reduced(float th, float* s, float* c)
{
    // round-to-nearest of theta / 2pi by the add-magic trick (|theta / 2pi| < 2^22): two FADDs on the FMA pipe
    // instead of an FRND on the XU pipe, which sin and cos (MUFU) already saturate
    const float k = __fadd_rn(__fadd_rn(th * 0.15915494309189535f, 12582912.0f), -12582912.0f);
    float r = fmaf(k, -6.28125f, th);                  // 2pi = 6.28125 + 1.9350051879882812e-3 + 3.019916050561733e-7
    r = fmaf(k, -1.9350051879882812e-3f, r);
    r = fmaf(k, -3.019916050561733e-7f, r);
    *s = __sinf(r);
    *c = __cosf(r);
}

constexpr int VPT = 4;  // vertices per thread: 3 float4 in, 3 float4 out
constexpr int THREADS = 256;

// NRM: 0 = no normal output, 1 = (0, 1, 0) (what the reference ships, :98 / :121), 2 = analytic normal of the displaced
// surface, 3 = the value Gerstner() computes at :92-97 before discarding it
template <int NRM>
__global__ void __launch_bounds__(THREADS) k_gerstner(const __grid_constant__ GerstnerTable tab, const float* __restrict__ pos,
                                                      float* __restrict__ out, float* __restrict__ nrm, int64_t n, float t,
                                                      float smoothing)
{
    const int64_t v0 = ((int64_t)blockIdx.x * THREADS + threadIdx.x) * VPT;
    if (v0 >= n) return;
    float p[VPT * 3];
    const bool full = v0 + VPT <= n;
    if (full) {
        const float4* src = reinterpret_cast<const float4*>(pos + 3 * v0);  // 48-byte aligned: v0 % 4 == 0
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float4 q = ldg_stream4(src + k);
            p[4 * k + 0] = q.x; p[4 * k + 1] = q.y; p[4 * k + 2] = q.z; p[4 * k + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < VPT * 3; ++k) p[k] = (3 * v0 + k < 3 * n) ? pos[3 * v0 + k] : 0.f;
    }
    float ox[VPT], oy[VPT], oz[VPT];
    // analytic normal: partial derivatives of P = (x + offs.x, offs.y, z + offs.z) with respect to the rest position
    //   dP/dx = (1 - jxx, hx, -jxz), dP/dz = (-jxz, hz, 1 - jzz),  jab = sum amp_xz f Da Db sin, ha = sum amp_y f Da cos
    float jxx[VPT], jxz[VPT], jzz[VPT], hx[VPT], hz[VPT];
#pragma unroll
    for (int v = 0; v < VPT; ++v) { ox[v] = oy[v] = oz[v] = 0.f; jxx[v] = jxz[v] = jzz[v] = hx[v] = hz[v] = 0.f; }
    const int nw = tab.n_waves;
#pragma unroll 2
    for (int w = 0; w < nw; ++w) {
        const mw_gerstner_wave W = tab.w[w];
        const float ph = __fmul_rn(W.rate, t);  // speeds * t   (MistralWaterLib.cginc:81, :114)
        const float ax = __fmul_rn(W.amp_xz, W.dir_x), az = __fmul_rn(W.amp_xz, W.dir_y);
#pragma unroll
        for (int v = 0; v < VPT; ++v) {
            // theta = freq * dot(dir, sVertex.xz) + rate * t   (:80-84, :114-116).  |theta| reaches
            // ~1e3 on a 1024-wide pond, where one fp32 ulp of theta is ~1e-4 rad: the phase is
            // therefore formed with the source's own roundings (no FMA contraction) so that it is
            // the same fp32 number the reference forms; only sin/cos are approximated.
            const float d = __fadd_rn(__fmul_rn(W.dir_x, p[3 * v + 0]), __fmul_rn(W.dir_y, p[3 * v + 2]));
            const float th = __fadd_rn(__fmul_rn(W.freq, d), ph);
            float s, c;
            sincos_reduced(th, &s, &c);
            ox[v] = fmaf(ax, c, ox[v]);       // :86 / :114
            oz[v] = fmaf(az, c, oz[v]);       // :87 / :115
            oy[v] = fmaf(W.amp_y, s, oy[v]);  // :88 / :116
            if (NRM == 2) {
                const float fs = W.freq * s, fc = W.freq * c;
                jxx[v] = fmaf(ax * W.dir_x, fs, jxx[v]);
                jxz[v] = fmaf(ax * W.dir_y, fs, jxz[v]);
                jzz[v] = fmaf(az * W.dir_y, fs, jzz[v]);
                hx[v] = fmaf(W.amp_y * W.dir_x, fc, hx[v]);
                hz[v] = fmaf(W.amp_y * W.dir_y, fc, hz[v]);
            }
        }
    }
    float nv[VPT * 3];
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
        float nx = 0.f, ny = 1.f, nz = 0.f;
        if (NRM == 2) {
            // dP/dz x dP/dx (y up): a = dP/dz = (-jxz, hz, 1 - jzz), b = dP/dx = (1 - jxx, hx, -jxz)
            const float a0 = -jxz[v], a1 = hz[v], a2 = 1.0f - jzz[v], b0 = 1.0f - jxx[v], b1 = hx[v], b2 = -jxz[v];
            nx = a1 * b2 - a2 * b1; ny = a2 * b0 - a0 * b2; nz = a0 * b1 - a1 * b0;
            const float inv = rsqrtf(nx * nx + ny * ny + nz * nz);
            nx *= inv; ny *= inv; nz *= inv;
        } else if (NRM == 3) {
            nx = (0.0f - ox[v]) * smoothing; ny = 2.0f - oz[v]; nz = 0.0f * smoothing;  // :92-96
            const float inv = rsqrtf(nx * nx + ny * ny + nz * nz);                       // :97 normalize
            nx *= inv; ny *= inv; nz *= inv;
        }
        nv[3 * v + 0] = nx; nv[3 * v + 1] = ny; nv[3 * v + 2] = nz;
    }
#pragma unroll
    for (int v = 0; v < VPT; ++v) { p[3 * v + 0] += ox[v]; p[3 * v + 1] += oy[v]; p[3 * v + 2] += oz[v]; }  // :176
    if (full) {
        float4* dst = reinterpret_cast<float4*>(out + 3 * v0);
#pragma unroll
        for (int k = 0; k < 3; ++k) dst[k] = make_float4(p[4 * k + 0], p[4 * k + 1], p[4 * k + 2], p[4 * k + 3]);
        if (NRM) {  // mode 1: (0,1,0) x4 = 0 1 0 0 | 1 0 0 1 | 0 0 1 0   (:98, :121)
            float4* dn = reinterpret_cast<float4*>(nrm + 3 * v0);
#pragma unroll
            for (int k = 0; k < 3; ++k) dn[k] = make_float4(nv[4 * k + 0], nv[4 * k + 1], nv[4 * k + 2], nv[4 * k + 3]);
        }
    } else {
        for (int k = 0; k < VPT * 3; ++k)
            if (3 * v0 + k < 3 * n) {
                out[3 * v0 + k] = p[k];
                if (NRM) nrm[3 * v0 + k] = nv[k];
            }
    }
}

}  // namespace

extern "C" int mw_gerstner_from_material(mw_gerstner_params* p, float amplitude, float frequency, float steepness,
                                         const float w_speed[4], const float w_direction_ab[4], const float w_direction_cd[4])
{
    if (!p || !w_speed || !w_direction_ab || !w_direction_cd) { mw_set_error("null argument"); return MW_E_INVALID_ARG; }
    const float amp = amplitude * 0.01f;  // MistralWaterLib.cginc:172
    const float dirs[4][2] = {{w_direction_ab[0], w_direction_ab[1]}, {w_direction_ab[2], w_direction_ab[3]},
                              {w_direction_cd[0], w_direction_cd[1]}, {w_direction_cd[2], w_direction_cd[3]}};
    p->n_waves = 4;
    for (int k = 0; k < 4; ++k) {
        mw_gerstner_wave& w = p->waves[k];
        w.dir_x = dirs[k][0]; w.dir_y = dirs[k][1];
        w.freq = frequency;          // :80
        w.rate = w_speed[k];         // :81
        w.amp_xz = steepness * amp;  // :77-78
        w.amp_y = amp;               // :88
    }
    return MW_OK;
}

extern "C" int mw_gerstner_append_level_one(mw_gerstner_params* p, float amplitude, float frequency, float steepness)
{
    if (!p) { mw_set_error("null argument"); return MW_E_INVALID_ARG; }
    if (p->n_waves < 0 || p->n_waves + 5 > MW_GERSTNER_MAX_WAVES) { mw_set_error("wave table full"); return MW_E_INVALID_ARG; }
    // MistralWaterLib.cginc:105-109
    static const float amps[5] = {0.7f, 0.6f, 0.6f, 0.7f, 0.9f};
    static const float steeps[5] = {0.95f, 0.615f, 0.821f, 0.462f, 0.611f};
    static const float speeds[5] = {-2.112f, 0.6124f, -0.878f, -3.6234f, 1.f};
    static const float dir[5][2] = {{1.f, -0.2f}, {-0.9f, 1.f}, {0.2f, 0.2f}, {-1.0f, 0.77f}, {0.99f, -1.145f}};
    static const float fs[5] = {0.954f, 1.52f, 0.44f, 0.21f, 0.8f};
    for (int i = 0; i < 5; ++i) {
        mw_gerstner_wave& w = p->waves[p->n_waves++];
        w.dir_x = dir[i][0]; w.dir_y = dir[i][1];
        w.freq = frequency * fs[i];              // :114
        w.rate = speeds[i] * frequency * fs[i];  // :114
        w.amp_xz = steepness * amplitude * steeps[i] * amps[i];
        w.amp_y = amplitude * amps[i];           // :116
    }
    return MW_OK;
}

extern "C" int mw_gerstner_displace(const mw_gerstner_params* p, const float* pos_xyz, float* out_xyz, float* out_nrm,
                                    int64_t n, float t, void* cuda_stream)
{
    if (!p || !pos_xyz || !out_xyz || n < 0) { mw_set_error("mw_gerstner_displace: bad argument"); return MW_E_INVALID_ARG; }
    if (p->n_waves < 0 || p->n_waves > MW_GERSTNER_MAX_WAVES) { mw_set_error("n_waves out of range"); return MW_E_INVALID_ARG; }
    if (n == 0) return MW_OK;
    MW_CUDA(cudaSetDevice(p->device));
    GerstnerTable tab;
    tab.n_waves = p->n_waves;
    memcpy(tab.w, p->waves, sizeof tab.w);
    const bool dev = (p->flags & MW_DEVICE_PTRS) != 0;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const float* d_pos = pos_xyz;
    float* d_out = out_xyz;
    float* d_nrm = out_nrm;
    float* scratch = nullptr;
    const size_t bytes = (size_t)n * 3 * sizeof(float);
    if (!dev) {
        const size_t pitch = ((size_t)3 * n + 3) & ~(size_t)3;  // keep every sub-buffer 16-byte aligned (float4 I/O)
        MW_CUDA(cudaMalloc((void**)&scratch, pitch * sizeof(float) * (out_nrm ? 3 : 2)));
        d_pos = scratch; d_out = scratch + pitch; d_nrm = out_nrm ? scratch + 2 * pitch : nullptr;
        cudaError_t e = cudaMemcpyAsync(scratch, pos_xyz, bytes, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) { cudaFree(scratch); mw_set_error("H2D failed: %s", cudaGetErrorString(e)); return MW_E_CUDA; }
    } else if ((reinterpret_cast<uintptr_t>(pos_xyz) | reinterpret_cast<uintptr_t>(out_xyz) |
                reinterpret_cast<uintptr_t>(out_nrm)) & 15) {
        mw_set_error("device buffers must be 16-byte aligned");
        return MW_E_INVALID_ARG;
    }
    const int64_t threads_needed = (n + VPT - 1) / VPT;
    const unsigned grid = (unsigned)((threads_needed + THREADS - 1) / THREADS);
    if (!d_nrm) k_gerstner<0><<<grid, THREADS, 0, st>>>(tab, d_pos, d_out, nullptr, n, t, 0.f);
    else if (p->flags & MW_GERSTNER_NORMAL_ANALYTIC) k_gerstner<2><<<grid, THREADS, 0, st>>>(tab, d_pos, d_out, d_nrm, n, t, 0.f);
    else if (p->flags & MW_GERSTNER_NORMAL_DISCARDED) k_gerstner<3><<<grid, THREADS, 0, st>>>(tab, d_pos, d_out, d_nrm, n, t, p->smoothing);
    else k_gerstner<1><<<grid, THREADS, 0, st>>>(tab, d_pos, d_out, d_nrm, n, t, 0.f);
    g_mw_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && !dev) {
        e = cudaMemcpyAsync(out_xyz, d_out, bytes, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && out_nrm) e = cudaMemcpyAsync(out_nrm, d_nrm, bytes, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    if (scratch) cudaFree(scratch);
    if (e != cudaSuccess) { mw_set_error("mw_gerstner_displace failed: %s", cudaGetErrorString(e)); return MW_E_CUDA; }
    return MW_OK;
}


// ---------------------------------------------------------------------------------------------
// `Wave` displacement mode (MistralWaterLib.cginc:127-152 through Displacement :160-164)
// ---------------------------------------------------------------------------------------------
namespace {
// sin / cos of speed + x * frequency: the phase is formed with the source's own fp32 roundings, reduced with a
// three-term Cody-Waite step and evaluated on the MUFU unit (same scheme as k_gerstner)
__device__ __forceinline__ void wave_sincos(float theta, float* s, float* c)
{
    const float k = rintf(theta * 0.15915494309189535f);
    float r = fmaf(k, -6.28318548202514648f, theta);
    r = fmaf(k, 1.74845553146e-7f, r);
    r = fmaf(k, 1.2e-14f, r) ;
    *s = __sinf(r);
    *c = __cosf(r);
}
__global__ void __launch_bounds__(256) k_wave(const float* __restrict__ pos, float* __restrict__ out, float* __restrict__ nrm,
                                              int64_t n, float speed, float amp, float frequency, float smoothing)
{
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const float x = pos[3 * v], y = pos[3 * v + 1], z = pos[3 * v + 2];
    float s0, c0, s1, c1, s2, c2, t;
    wave_sincos(__fadd_rn(speed, __fmul_rn(x, frequency)), &s0, &t);
    wave_sincos(__fadd_rn(speed, __fmul_rn(__fadd_rn(x, 0.05f), frequency)), &s1, &t);
    wave_sincos(__fadd_rn(speed, __fmul_rn(z, frequency)), &t, &c0);
    wave_sincos(__fadd_rn(speed, __fmul_rn(__fadd_rn(z, 0.05f), frequency)), &t, &c2);
    s2 = s0;  // v2 shares x with v0 (:131), v1 shares z with v0 (:130)
    c1 = c0;
    const float y0 = y + s0 * amp - c0 * amp;
    float y1 = y + s1 * amp - c1 * amp;
    float y2 = y + s2 * amp - c2 * amp;
    y1 -= (y1 - y0) * (1.0f - smoothing);  // :144-145
    y2 -= (y2 - y0) * (1.0f - smoothing);
    out[3 * v] = x;
    out[3 * v + 1] = y + y0;               // v.vertex.y += offsets.y (:162), offsets = the displaced v0
    out[3 * v + 2] = z;
    if (nrm) {
        // cross(v2 - v0, v1 - v0) with v2 - v0 = (0, y2 - y0, 0.05), v1 - v0 = (0.05, y1 - y0, 0)   (:147)
        const float ax = 0.0f, ay = y2 - y0, az = __fsub_rn(__fadd_rn(z, 0.05f), z);
        const float bx = __fsub_rn(__fadd_rn(x, 0.05f), x), by = y1 - y0, bz = 0.0f;
        const float cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
        const float inv = rsqrtf(cx * cx + cy * cy + cz * cz);
        nrm[3 * v] = cx * inv; nrm[3 * v + 1] = cy * inv; nrm[3 * v + 2] = cz * inv;
    }
}
}  // namespace

extern "C" int mw_wave_displace(const mw_wave_params* p, const float* pos_xyz, float* out_xyz, float* out_nrm, int64_t n,
                                float t, void* cuda_stream)
{
    if (!p || !pos_xyz || !out_xyz || n < 0) { mw_set_error("mw_wave_displace: bad argument"); return MW_E_INVALID_ARG; }
    if (n == 0) return MW_OK;
    MW_CUDA(cudaSetDevice(p->device));
    const bool dev = (p->flags & MW_DEVICE_PTRS) != 0;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const float* d_pos = pos_xyz;
    float *d_out = out_xyz, *d_nrm = out_nrm, *scratch = nullptr;
    const size_t bytes = (size_t)n * 3 * sizeof(float);
    if (!dev) {
        MW_CUDA(cudaMalloc((void**)&scratch, bytes * (out_nrm ? 3 : 2)));
        d_pos = scratch; d_out = scratch + 3 * n; d_nrm = out_nrm ? scratch + 6 * n : nullptr;
        cudaError_t e = cudaMemcpyAsync(scratch, pos_xyz, bytes, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) { cudaFree(scratch); mw_set_error("H2D failed: %s", cudaGetErrorString(e)); return MW_E_CUDA; }
    }
    volatile float speed = p->speed * t;          // :133
    volatile float amp = p->amplitude * 0.01f;    // :134
    k_wave<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_pos, d_out, d_nrm, n, speed, amp, p->frequency, p->smoothing);
    g_mw_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && !dev) {
        e = cudaMemcpyAsync(out_xyz, d_out, bytes, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && out_nrm) e = cudaMemcpyAsync(out_nrm, d_nrm, bytes, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    if (scratch) cudaFree(scratch);
    if (e != cudaSuccess) { mw_set_error("mw_wave_displace failed: %s", cudaGetErrorString(e)); return MW_E_CUDA; }
    return MW_OK;
}
