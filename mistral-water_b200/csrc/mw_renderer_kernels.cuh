// mw_renderer_kernels.cuh -- device code of the OceanRenderer (GPU-shader convention) path, sm_100a.
//
// What the reference does per frame (Scripts/OceanRenderer.cs:216-316 GenerateTexture) is a chain of 3 + 4 log2 R + 2
// full-screen blits on R x R float textures (R = 8 * resolution, :136): Dispersion -> Spectrum -> 2 log2 R Stockham
// passes -> SpectrumHeight -> 2 log2 R Stockham passes -> OceanNormal -> WhiteCap.  Here it is three kernels:
//
//   k_r_rows    phase += rate * dt (Dispersion.shader:32-41), h = h0 e^{i phase} + h0conj e^{-i phase}
//               (Spectrum.shader:40-45), the chop spectra hx, hz (:47-50) and the height spectrum
//               (SpectrumHeight.shader:46), then the horizontal half of both Stockham chains        (pass 1)
//   k_r_cols    the vertical half of both chains -> the displacement and height images              (pass 2)
//   k_r_maps    OceanNormal.shader:32-56 and WhiteCap.shader:33-45 on those two images              (pass 3)
//
// The two complex fields of the reference's RGBA texel ARE a packed pair of the FFT engine (mw_fft.cuh): (hx, hz)
// ride one packed line, the height spectra of two adjacent rows / columns ride another.
// Images are [y][x] RGBAFloat, x contiguous; "horizontal" = along x.
#pragma once
#include "mw_layout.cuh"

namespace mwr {

using mwfft::Plan;
using mwfft::pad_idx;
using mwk::slab_w;
using mwk::xab_index;
using mwk::xab_tile_elems;

#define MWR_PI_F 3.1415926536f  // FFTCommon.cginc:7
#define MWR_G_F 9.81f           // :9
#define MWR_EPS_F 0.0001f       // :8

// GetWave (FFTCommon.cginc:58-67) component for texel index i (n = i + .5 enters, `n -= 0.5` recovers i)
__device__ __forceinline__ float wave_rn(int i, int R, float length)
{
    const float n = (float)(i < R / 2 ? i : i - R);
    return __fdiv_rn(__fmul_rn(__fmul_rn(2.0f, MWR_PI_F), n), length);
}

// =============================================================================================
// init-time kernels
// =============================================================================================
// Phillips (FFTCommon.cginc:69-85); (i, j) are already GetWave's integer indices
__device__ __forceinline__ float phillips_r(int i, int j, int R, float amp, float wx, float wy, float length)
{
    const float kx = wave_rn(i, R, length), kz = wave_rn(j, R, length);
    const float klen = __fsqrt_rn(__fadd_rn(__fmul_rn(kx, kx), __fmul_rn(kz, kz)));
    if (klen < MWR_EPS_F) return 0.0f;
    const float klen2 = __fmul_rn(klen, klen), klen4 = __fmul_rn(klen2, klen2);
    const float wlen = __fsqrt_rn(__fadd_rn(__fmul_rn(wx, wx), __fmul_rn(wy, wy)));
    const float kdw = __fadd_rn(__fmul_rn(__fdiv_rn(kx, klen), __fdiv_rn(wx, wlen)), __fmul_rn(__fdiv_rn(kz, klen), __fdiv_rn(wy, wlen)));
    const float kdw2 = __fmul_rn(kdw, kdw);
    const float l = __fdiv_rn(__fmul_rn(wlen, wlen), MWR_G_F);
    const float l2 = __fmul_rn(l, l);
    const float L2 = __fmul_rn(__fmul_rn(l2, 0.01f), 0.01f);  // damping 0.01 (:82)
    const float e1 = (float)exp((double)__fdiv_rn(-1.0f, __fmul_rn(klen2, l2)));
    const float e2 = (float)exp((double)__fmul_rn(-klen2, L2));
    return __fmul_rn(__fmul_rn(__fdiv_rn(__fmul_rn(amp, e1), klen4), kdw2), e2);
}

// UVRandom (FFTCommon.cginc:37-41) with a correctly rounded sin: GPU sin is implementation-defined and the hash
// amplifies its error by 4e4, so the reference's own noise differs from GPU to GPU; hosts that want
// theirs upload it (mw_renderer_set_initial)
__device__ __forceinline__ float uv_random(float u, float v, float salt, float rnd)
{
    const float d = __fadd_rn(__fmul_rn(__fadd_rn(u, salt), 12.9898f), __fmul_rn(__fadd_rn(v, rnd), 78.233f));
    const float s = __fmul_rn((float)sin((double)d), 43758.5453f);
    return __fsub_rn(s, floorf(s));
}

// hTilde0 (FFTCommon.cginc:87-99)
__device__ __forceinline__ float2 htilde0_r(float u, float v, float r1, float r2, float phi)
{
    const float rand1 = fminf(fmaxf(uv_random(u, v, 10.612f, r1), 0.01f), 1.0f);
    const float rand2 = fminf(fmaxf(uv_random(u, v, 11.899f, r2), 0.01f), 1.0f);
    const float x = (float)sqrt((double)__fmul_rn(-2.0f, (float)log((double)rand1)));
    const float y = __fmul_rn(__fmul_rn(2.0f, MWR_PI_F), rand2);
    const float s = (float)sqrt((double)__fdiv_rn(phi, 2.0f));
    return make_float2(__fmul_rn(__fmul_rn(x, (float)cos((double)y)), s), __fmul_rn(__fmul_rn(x, (float)sin((double)y)), s));
}

// InitialSpectrum.shader:42-54 -> initial[tile][y][x] = (h0, h0conj); tile t uses seeds + t
__global__ void k_r_initial(float4* __restrict__ initial, int R, int tiles, float length, float amp, float wx, float wy,
                            float seed1, float seed2)
{
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n2 = (int64_t)R * R;
    if (gid >= n2 * tiles) return;
    const int tile = (int)(gid / n2), idx = (int)(gid % n2);
    const int x = idx % R, y = idx / R;
    const float u = __fdiv_rn((float)x + 0.5f, (float)R), v = __fdiv_rn((float)y + 0.5f, (float)R);
    const float phi1 = phillips_r(x, y, R, amp, wx, wy, length);
    // :47  Phillips(_Resolution - n, _Resolution - m) with n = x + .5: GetWave sees R - 1 - x
    const float phi2 = phillips_r(R - 1 - x, R - 1 - y, R, amp, wx, wy, length);
    const float s1 = __fadd_rn(seed1, (float)tile), s2 = __fadd_rn(seed2, (float)tile);
    const float2 a = htilde0_r(u, v, __fdiv_rn(s1, 2.0f), __fmul_rn(s2, 2.0f), phi1);  // :49
    const float2 b = htilde0_r(u, v, s1, s2, phi2);                                    // :50
    initial[gid] = make_float4(a.x, a.y, b.x, -b.y);
}

// CalcDispersion without the `* dt` (FFTCommon.cginc:106-114): sqrt(G |k| (1 + |k|^2 / 370 / 370))
__global__ void k_r_rate(float* __restrict__ rate, int R, float length)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= R * R) return;
    const float kx = wave_rn(idx % R, R, length), kz = wave_rn(idx / R, R, length);
    const float wlen = __fsqrt_rn(__fadd_rn(__fmul_rn(kx, kx), __fmul_rn(kz, kz)));
    const float cap = __fadd_rn(1.0f, __fdiv_rn(__fdiv_rn(__fmul_rn(wlen, wlen), 370.0f), 370.0f));
    rate[idx] = __fsqrt_rn(__fmul_rn(__fmul_rn(MWR_G_F, wlen), cap));
}

// =============================================================================================
// pass 1: dispersion + spectrum + spectrum-height + horizontal transform
// =============================================================================================
struct RRowArgs {
    const float4* initial;  // [tiles][R][R] (h0, h0conj)
    float* phase;           // [tiles][R][R], updated in place (the ping/pong phase textures)
    const float* rate;      // [R][R]
    const float* kw;        // [R] GetWave component per texel index
    const float4* twimg;    // twiddle image, sign -1
    float4* XAB;            // [tiles][R/W][R][W]   (hx, hz) after the horizontal transform
    float2* XC;             // [tiles][R/2W][R][2W] h after the horizontal transform
    float dt;               // deltaTime * mult (OceanRenderer.cs:223)
    float choppiness;
    int tile0;
};

// One CTA = two image rows y0, y0 + 1 = 3 packed lines: (hx, hz) of y0, (hx, hz) of y1, (h of y0, h of y1).
#ifndef MW_RROWS_UNROLL
#define MW_RROWS_UNROLL 2   // texel batches of the evolve loop in flight (developer switch)
#endif
template <int N>
__global__ void __launch_bounds__(3 * (N / 16)) k_r_rows(const RRowArgs a)
{
    using P = Plan<N>;
    constexpr int T = P::T;
    constexpr int NT = 3 * T;
    constexpr int LP = mwfft::line_pitch(N, 8);
    constexpr int W = slab_w(N);
    extern __shared__ float4 smem4[];
    float4* tw2 = smem4;
    float2* tw3 = reinterpret_cast<float2*>(smem4 + P::TW2_F4);
    float4* lines = smem4 + P::TW_BYTES / 16;

    const int tile = a.tile0 + blockIdx.y, xt = blockIdx.y;
    const int y0 = 2 * blockIdx.x, y1 = y0 + 1;
    const float4* ini = a.initial + (size_t)tile * N * N;
    float* ph = a.phase + (size_t)tile * N * N;
    const float ky0 = __ldg(a.kw + y0), ky1 = __ldg(a.kw + y1);
    const float two_pi = __fmul_rn(2.0f, MWR_PI_F);
    const uint64_t pol = evict_first_policy();  // the initial spectrum is read once per frame
    mwfft::load_twiddle_image<N, NT>(smem4, a.twimg);
    // launched as a programmatic dependent: the phase texture (updated in place below) and the intermediate belong to the
    // stream's earlier kernels until the predecessor has completed
    pdl_trigger();
    pdl_wait();

    auto texel = [&](float4 s, float phase_old, float rate, float kx, float ky, unsigned o, float4& F, float2& H) {
        // Dispersion.shader:37-40, GetDispersion: fmod(phase + rate * dt, 2 PI)
        const float sum = __fadd_rn(phase_old, __fmul_rn(rate, a.dt));
        const float phase = __fsub_rn(sum, __fmul_rn(two_pi, truncf(__fdiv_rn(sum, two_pi))));
        ph[o] = phase;
        float sn, cs;
        sincosf(phase, &sn, &cs);
        // Spectrum.shader:45  h = h0 * pv + h0conj * conj(pv)
        H = make_float2((s.x + s.z) * cs - (s.y - s.w) * sn, (s.x - s.z) * sn + (s.y + s.w) * cs);
        // :47-49  hx = -MultByI(h * wave.x / w) * choppiness = (h.y, -h.x) * wave.x / w * choppiness
        const float w = fmaxf(MWR_EPS_F, sqrtf(kx * kx + ky * ky));
        const float sc = __fdividef(a.choppiness, w);
        const float fx = kx * sc, fz = ky * sc;
        F = make_float4(H.y * fx, H.y * fz, -H.x * fx, -H.x * fz);  // (hx.re, hz.re, hx.im, hz.im)
    };
    constexpr int UNR = MW_RROWS_UNROLL;
#pragma unroll UNR
    for (int x = threadIdx.x; x < N; x += NT) {
        const unsigned o0 = (unsigned)y0 * N + x, o1 = (unsigned)y1 * N + x;
        const float4 s0 = ldg_once4(ini + o0, pol), s1 = ldg_once4(ini + o1, pol);
        const float p0 = ph[o0], p1 = ph[o1];
        const float r0 = __ldg(a.rate + o0), r1 = __ldg(a.rate + o1);
        const float kx = __ldg(a.kw + x);
        float4 F0, F1;
        float2 H0, H1;
        texel(s0, p0, r0, kx, ky0, o0, F0, H0);
        texel(s1, p1, r1, kx, ky1, o1, F1, H1);
        const int px = pad_idx(x);
        lines[px] = F0;
        lines[LP + px] = F1;
        lines[2 * LP + px] = make_float4(H0.x, H1.x, H0.y, H1.y);
    }
    __syncthreads();

    const int q = threadIdx.x / T, g = threadIdx.x % T;
    float4* line = lines + q * LP;
    mwfft::cpk v[16];
    mwfft::load_line_regs<N>(v, line, g);
    auto line_sync = [&] { mwfft::group_sync<T>(q); };
    line_sync();
    mwfft::fft_line_inreg<N, -1>(v, line, g, tw2, tw3, line_sync);
    if (q < 2) {
        const int row = q ? y1 : y0;
        float4* dst = a.XAB + (size_t)xt * xab_tile_elems(N);
#pragma unroll
        for (int sl = 0; sl < 16; ++sl)
            dst[xab_index(N, row, g + mwfft::final_off<N>(sl))] = make_float4(v[sl].re.x, v[sl].re.y, v[sl].im.x, v[sl].im.y);
    } else {
        float2* dst = a.XC + (size_t)xt * N * N;
#pragma unroll
        for (int sl = 0; sl < 16; ++sl) {
            const int idx = g + mwfft::final_off<N>(sl);
            dst[mwk::xc_index(N, y0, idx)] = make_float2(v[sl].re.x, v[sl].im.x);
            dst[mwk::xc_index(N, y1, idx)] = make_float2(v[sl].re.y, v[sl].im.y);
        }
    }
    (void)W;
}

// =============================================================================================
// pass 2: vertical transform -> displacement / height images
// =============================================================================================
struct RColArgs {
    const float4* XAB;
    const float2* XC;
    const float4* twimg;
    float4* displacement;  // [tiles][R][R] (Re hx, Im hx, Re hz, Im hz)   = displacementTexture
    float4* height;        // [tiles][R][R] (Re h, Im h, Re h, Im h)       = heightTexture
    int tile0;
    int ab_blocks;         // blockIdx.x < ab_blocks: (hx, hz) slab of W columns; else: h slab of 2 W columns
};

#ifndef MW_RCOLS_TMA
#define MW_RCOLS_TMA 1   // (hx, hz) slab by one bulk copy; 0 = 16 per-thread LDG.128 (developer switch, tools/variant builds)
#endif
template <int N>
__global__ void __launch_bounds__(slab_w(N) * (N / 16)) k_r_cols(const RColArgs a)
{
    using P = Plan<N>;
    constexpr int T = P::T;
    constexpr int W = slab_w(N);
    constexpr int LOGW = mwfft::ilog2(W);
    constexpr int LP = mwfft::line_pitch(N, W);
    extern __shared__ float4 smem4[];
    float4* tw2 = smem4;
    float2* tw3 = reinterpret_cast<float2*>(smem4 + P::TW2_F4);
    float4* lines = smem4 + P::TW_BYTES / 16;
    __shared__ uint64_t slab_bar;

    const int tile = a.tile0 + blockIdx.y, xt = blockIdx.y;
    const int tid = threadIdx.x;
    const int c = tid & (W - 1), g = tid >> LOGW;
    const size_t plane = (size_t)N * N;
    float4* line = lines + c * LP;
    const bool is_ab = (int)blockIdx.x < a.ab_blocks;
    const int b0 = is_ab ? blockIdx.x * W : ((int)blockIdx.x - a.ab_blocks) * (2 * W);
    mwfft::cpk v[16];
    pdl_trigger();
    pdl_wait();  // the intermediate is k_r_rows' (the predecessor in this stream)
    if (is_ab && MW_RCOLS_TMA) {
        // the slab is one contiguous block of N * W * 16 bytes in the slab-major intermediate: ONE bulk copy (TMA 1-D form) into
        // the line buffers, which are free until the first stage writes them -- as in mwk::k_cols_seam; 16 x 512 per-thread
        // LDG.128 kept the load/store queue full instead (stall reason lg_throttle 9.8, profiles/r02_summary.md)
        if (tid == 0) mbar_init(&slab_bar, 1);
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&slab_bar, (unsigned)(N * W * sizeof(float4)));
            bulk_g2s(lines, a.XAB + (size_t)xt * xab_tile_elems(N) + (size_t)blockIdx.x * N * W, (unsigned)(N * W * sizeof(float4)), &slab_bar);
        }
        mwfft::load_twiddle_image<N, W * T>(smem4, a.twimg);
        mbar_wait(&slab_bar, 0);
        const float4* src = lines + g * W + c;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float4 e = src[(T * k) * W];
            v[k].re = make_float2(e.x, e.y);
            v[k].im = make_float2(e.z, e.w);
        }
        __syncthreads();  // everyone has its inputs before the first stage overwrites the raw slab with the lines
    } else if (is_ab) {
        const float4* src = a.XAB + (size_t)xt * xab_tile_elems(N) + ((size_t)blockIdx.x * N + g) * W + c;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float4 e = ldg_fresh4(src + (size_t)(T * k) * W);
            v[k].re = make_float2(e.x, e.y);
            v[k].im = make_float2(e.z, e.w);
        }
        mwfft::load_twiddle_image<N, W * T>(smem4, a.twimg);
    } else {
        // 16 contiguous bytes = columns b0 + 2c, b0 + 2c + 1 of one row: (re0, im0, re1, im1)
        const float4* src = reinterpret_cast<const float4*>(a.XC + (size_t)xt * plane + ((size_t)(b0 / (2 * W)) * N + g) * (2 * W) + 2 * c);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float4 e = ldg_fresh4(src + (size_t)(T * k) * W);
            v[k].re = make_float2(e.x, e.z);
            v[k].im = make_float2(e.y, e.w);
        }
        mwfft::load_twiddle_image<N, W * T>(smem4, a.twimg);
    }
    auto cta_sync = [] { __syncthreads(); };
    mwfft::fft_line_inreg<N, -1>(v, line, g, tw2, tw3, cta_sync);
    if (is_ab) {
        float4* dst = a.displacement + (size_t)tile * plane + (size_t)g * N + b0 + c;
#pragma unroll
        for (int s = 0; s < 16; ++s)
            dst[(size_t)mwfft::final_off<N>(s) * N] = make_float4(v[s].re.x, v[s].im.x, v[s].re.y, v[s].im.y);
    } else {
        float4* dst = a.height + (size_t)tile * plane + (size_t)g * N + b0 + 2 * c;
#pragma unroll
        for (int s = 0; s < 16; ++s) {
            float4* d = dst + (size_t)mwfft::final_off<N>(s) * N;
            d[0] = make_float4(v[s].re.x, v[s].im.x, v[s].re.x, v[s].im.x);
            d[1] = make_float4(v[s].re.y, v[s].im.y, v[s].re.y, v[s].im.y);
        }
    }
}

// =============================================================================================
// pass 3: OceanNormal.shader + WhiteCap.shader
// =============================================================================================
struct RMapArgs {
    const float4* displacement;
    const float4* height;
    float4* normal;      // [tiles][R][R] (n, 1)                   = normalTexture        or NULL
    float* white;        // [tiles][R][R] the R channel (ColorMask R)                      or NULL
    float4* white_rgba;  // [tiles][R][R] (xx, xx, xx, 1) as the fragment returns it        or NULL
    float* jacobian;     // [tiles][R][R] (developer / test output)                         or NULL
    int R;
    int tile0;           // first image of this launch (blockIdx.z counts from it)
    int step;            // WhiteCap tap distance in texels = R / mesh resolution (8)
    int repeat;          // 0: clamp at the border (RenderTexture default), 1: repeat
    float texel_size;    // _Length / _Resolution (OceanNormal.shader:42)
};

__global__ void __launch_bounds__(256) k_r_maps(const RMapArgs a)
{
    const int R = a.R;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    pdl_trigger();
    pdl_wait();  // the two images are k_r_cols'
    if (x >= R || y >= R) return;
    const size_t base = (size_t)(a.tile0 + blockIdx.z) * R * R;
    const float4* D = a.displacement + base;
    const float4* H = a.height + base;
    auto wrap = [&](int i) { return a.repeat ? (i + R) & (R - 1) : min(max(i, 0), R - 1); };
    const int xm = wrap(x - 1), xp = wrap(x + 1), ym = wrap(y - 1), yp = wrap(y + 1);
    const float ts = a.texel_size;
    const float4 dc = ld_plain4(D + (size_t)y * R + x);
    // GetVec (OceanNormal.shader:32-37) = (disp.r, height.r, disp.b); center = disp.rgb as written (:44)
    auto vec = [&](int xx, int yy) {
        const float4 d = ld_plain4(D + (size_t)yy * R + xx);
        const float h = ld_plain1(reinterpret_cast<const float*>(H + (size_t)yy * R + xx));
        return make_float3(d.x, h, d.z);
    };
    const float3 vr = vec(xp, y), vl = vec(xm, y), vt = vec(x, ym), vb = vec(x, yp);
    const float3 right = make_float3(ts + vr.x - dc.x, vr.y - dc.y, vr.z - dc.z);
    const float3 left = make_float3(-ts + vl.x - dc.x, vl.y - dc.y, vl.z - dc.z);
    const float3 top = make_float3(vt.x - dc.x, vt.y - dc.y, -ts + vt.z - dc.z);
    const float3 bottom = make_float3(vb.x - dc.x, vb.y - dc.y, ts + vb.z - dc.z);
    auto cross = [](float3 p, float3 q) { return make_float3(p.y * q.z - p.z * q.y, p.z * q.x - p.x * q.z, p.x * q.y - p.y * q.x); };
    const float3 c1 = cross(right, top), c2 = cross(top, left), c3 = cross(left, bottom), c4 = cross(bottom, right);
    const float sx = c1.x + c2.x + c3.x + c4.x, sy = c1.y + c2.y + c3.y + c4.y, sz = c1.z + c2.z + c3.z + c4.z;
    const float inv = rsqrtf(sx * sx + sy * sy + sz * sz);
    const float nx = sx * inv, ny = sy * inv, nz = sz * inv;
    const size_t o = base + (size_t)y * R + x;
    const uint64_t pol = evict_first_policy();  // the maps leave the engine here
    if (a.normal) st_once(a.normal + o, make_float4(nx, ny, nz, 1.0f), pol);
    if (a.white || a.white_rgba || a.jacobian) {
        // WhiteCap.shader:35-36: +-step texel central differences of disp.rb, / 8
        const int st = a.step;
        const float4 dN = ld_plain4(D + (size_t)wrap(y - st) * R + x), dS = ld_plain4(D + (size_t)wrap(y + st) * R + x);
        const float4 dW = ld_plain4(D + (size_t)y * R + wrap(x - st)), dE = ld_plain4(D + (size_t)y * R + wrap(x + st));
        const float dDdy_x = -0.5f * (dN.x - dS.x) / 8.0f, dDdy_y = -0.5f * (dN.z - dS.z) / 8.0f;
        const float dDdx_x = -0.5f * (dW.x - dE.x) / 8.0f, dDdx_y = -0.5f * (dW.z - dE.z) / 8.0f;
        const float ax = 0.3f * nx, az = 0.3f * nz;                                        // :37
        const float jac = (1.0f + dDdx_x) * (1.0f + dDdy_y) - dDdx_y * dDdy_x;              // :38
        const float turb = fmaxf(0.0f, 1.0f - jac + sqrtf(ax * ax + az * az));             // :39
        const float s = fminf(turb, 1.0f);
        const float xx = s * s * (3.0f - 2.0f * s);                                         // :42 smoothstep(0, 1, turb)
        if (a.white) st_once(a.white + o, xx, pol);
        if (a.white_rgba) st_once(a.white_rgba + o, make_float4(xx, xx, xx, 1.0f), pol);
        if (a.jacobian) st_once(a.jacobian + o, jac, pol);
    }
}

// =============================================================================================
// OceanRenderer.GenerateMesh (OceanRenderer.cs:172-207, = FFTMesh.cs:101-139 without the spectrum)
// =============================================================================================
// vertices / normals [N*N] float3, uvs [N*N] float2, indices [(N-1)^2 * 6] int32, in the reference's emission order.
__global__ void k_mesh_generate(float* __restrict__ vertices, float* __restrict__ normals, float* __restrict__ uvs,
                                int* __restrict__ indices, int N, float unit_width)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * N) return;
    const int i = idx / N, j = idx % N;
    const int half = N / 2;
    const float off = (N % 2 == 0) ? __fdiv_rn(unit_width, 2.0f) : 0.0f;
    if (vertices) {
        vertices[3 * idx + 0] = __fadd_rn(__fmul_rn((float)(i - half), unit_width), off);
        vertices[3 * idx + 1] = 0.0f;
        vertices[3 * idx + 2] = __fadd_rn(__fmul_rn((float)(j - half), unit_width), off);
    }
    if (normals) { normals[3 * idx + 0] = 0.f; normals[3 * idx + 1] = 1.f; normals[3 * idx + 2] = 0.f; }
    if (uvs) {
        uvs[2 * idx + 0] = __fdiv_rn(__fmul_rn((float)i, 1.0f), (float)(N - 1));
        uvs[2 * idx + 1] = __fdiv_rn(__fmul_rn((float)j, 1.0f), (float)(N - 1));
    }
    if (indices && j != N - 1) {
        // the loop emits, for (i, j) in row-major order with j < N-1: 3 indices if i != N-1, then 3 more if i != 0.
        // rows before i: row 0 contributes 3 (N-1), rows 1..N-2 contribute 6 (N-1) each, row N-1 contributes 3 (N-1)
        const int per_cell_here = (i != N - 1 ? 3 : 0) + (i != 0 ? 3 : 0);
        const int before_rows = i == 0 ? 0 : 3 * (N - 1) + (i - 1) * 6 * (N - 1);
        int* o = indices + before_rows + j * per_cell_here;
        if (i != N - 1) { o[0] = idx; o[1] = idx + 1; o[2] = idx + N; o += 3; }
        if (i != 0) { o[0] = idx; o[1] = idx - N + 1; o[2] = idx + 1; }
    }
}

}  // namespace mwr
