// mw_ocean_kernels.cuh -- device code of the Tessendorf hot path (sm_100a).
//
// Frame pipeline (replaces FFTMesh.EvaluateWaves, Scripts/FFTMesh.cs:224-280):
//
//   k_spectrum_rows   h0,h0conj --evolve--> h(k,t) --phase ramp + Hermitian packing--> 3 complex
//                     fields --row FFT (along m)--> intermediate X[f][n][b]          (pass 1)
//   k_cols_extract    X --column FFT (along n)--> height / hds / normal / whitecap   (pass 2)
//
// Why this equals the reference's O(N^4) direct sum: SURVEY.md section 3.4 / DESIGN.md.
#pragma once
#include "mw_fft.cuh"

namespace mwk {

using mwfft::Plan;
using mwfft::pad_idx;

// =============================================================================================
// init-time kernels
// =============================================================================================

// Dispersion(n, m), FFTMesh.cs:141-147, bit-exact: every operation is the fp32 round-to-nearest
// one the C# expression performs, in the same order, with no FMA contraction.
__device__ __forceinline__ float dispersion_rn(int n, int m, int N, float length)
{
    const float w = __fdiv_rn(__fmul_rn(2.0f, MW_PI_F), length);
    const float kx = __fdiv_rn(__fmul_rn(MW_PI_F, (float)(2 * n - N)), length);
    const float kz = __fdiv_rn(__fmul_rn(MW_PI_F, (float)(2 * m - N)), length);
    const float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(kx, kx), __fmul_rn(kz, kz)));
    return __fmul_rn(floorf(__fdiv_rn(__fsqrt_rn(__fmul_rn(MW_G_F, mag)), w)), w);
}

__global__ void k_dispersion(float* __restrict__ omega, int N, float length)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * N) return;
    omega[idx] = dispersion_rn(idx / N, idx % N, N, length);
}

// Phillips(n, m), FFTMesh.cs:149-166.  fp32 storage and operation order as in the source;
// Mathf.Exp/Sqrt are "double libm, then round", which is what the double intrinsics give here.
__device__ __forceinline__ float phillips_rn(int n, int m, int N, float length, float amplitude, float wx, float wy)
{
    const float kx = __fmul_rn(__fdiv_rn((float)(2 * n - N), length), MW_PI_F);
    const float kz = __fmul_rn(__fdiv_rn((float)(2 * m - N), length), MW_PI_F);
    const float k_length = __fsqrt_rn(__fadd_rn(__fmul_rn(kx, kx), __fmul_rn(kz, kz)));
    if (k_length < MW_EPSILON_F) return 0.0f;
    const float k2 = __fmul_rn(k_length, k_length);
    const float k4 = __fmul_rn(k2, k2);
    // k.normalized, wind.normalized (zero when magnitude <= 1e-5)
    float nkx = 0.f, nkz = 0.f, nwx = 0.f, nwy = 0.f;
    if (k_length > 1e-5f) { nkx = __fdiv_rn(kx, k_length); nkz = __fdiv_rn(kz, k_length); }
    const float w_length = __fsqrt_rn(__fadd_rn(__fmul_rn(wx, wx), __fmul_rn(wy, wy)));
    if (w_length > 1e-5f) { nwx = __fdiv_rn(wx, w_length); nwy = __fdiv_rn(wy, w_length); }
    const float kDotW = __fadd_rn(__fmul_rn(nkx, nwx), __fmul_rn(nkz, nwy));
    const float kDotW2 = __fmul_rn(kDotW, kDotW);
    const float l = __fdiv_rn(__fmul_rn(w_length, w_length), MW_G_F);
    const float l2 = __fmul_rn(l, l);
    const float damping = 0.001f;
    const float L2 = __fmul_rn(__fmul_rn(l2, damping), damping);
    const float e1 = (float)exp((double)__fdiv_rn(-1.0f, __fmul_rn(k2, l2)));
    const float e2 = (float)exp((double)__fmul_rn(-k2, L2));
    return __fmul_rn(__fmul_rn(__fdiv_rn(__fmul_rn(amplitude, e1), k4), kDotW2), e2);
}

// htilde0(n, m), FFTMesh.cs:168-176
__device__ __forceinline__ float2 htilde0_rn(int n, int m, float z1, float z2, int N, float length, float amplitude,
                                             float wx, float wy)
{
    const float lg = (float)log((double)z1);
    const float rad = (float)sqrt((double)__fmul_rn(-2.0f, lg));
    const float ang = __fmul_rn(__fmul_rn(2.0f, MW_PI_F), z2);
    const float rx = __fmul_rn(rad, (float)cos((double)ang));
    const float ry = __fmul_rn(rad, (float)sin((double)ang));
    const float s = (float)sqrt((double)__fdiv_rn(phillips_rn(n, m, N, length, amplitude, wx, wy), 2.0f));
    return make_float2(__fmul_rn(rx, s), __fmul_rn(ry, s));
}

// GenerateMesh's spectrum part, FFTMesh.cs:114-116: spec[idx] = (h0.x, h0.y, h0conj.x, h0conj.y)
__global__ void k_init_spectrum(float4* __restrict__ spec, int N, int tiles, float length, float amplitude, float wx,
                                float wy, uint64_t seed)
{
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n2 = (int64_t)N * N;
    if (gid >= n2 * tiles) return;
    const int tile = (int)(gid / n2);
    const int idx = (int)(gid % n2);
    const int i = idx / N, j = idx % N;
    const uint64_t key = seed + (uint64_t)tile;
    uint32_t r[4];
    philox4x32_10((uint32_t)idx, 0u, 0u, 0u, (uint32_t)key, (uint32_t)(key >> 32), r);
    const float2 a = htilde0_rn(i, j, u32_to_unit_open0(r[0]), u32_to_unit_open0(r[1]), N, length, amplitude, wx, wy);
    const float2 b = htilde0_rn(N - i, N - j, u32_to_unit_open0(r[2]), u32_to_unit_open0(r[3]), N, length, amplitude, wx, wy);
    spec[gid] = make_float4(a.x, a.y, b.x, -b.y);
}

__global__ void k_pack_h0(float4* __restrict__ spec, const float2* __restrict__ h0, const float2* __restrict__ h0c, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2 a = h0[i], b = h0c[i];
    spec[i] = make_float4(a.x, a.y, b.x, b.y);
}
__global__ void k_unpack_h0(const float4* __restrict__ spec, float2* __restrict__ h0, float2* __restrict__ h0c, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 s = spec[i];
    h0[i] = make_float2(s.x, s.y);
    h0c[i] = make_float2(s.z, s.w);
}

// =============================================================================================
// per-frame: h(k,t)
// =============================================================================================

// htilde(t, n, m), FFTMesh.cs:178-190, from the packed spectrum and the precomputed omega.
__device__ __forceinline__ float2 htilde_eval(float4 s, float cs, float sn)
{
    // res.x = h0.x*c - h0.y*s + h0c.x*c + h0c.y*s ; res.y = h0.x*s + h0.y*c - h0c.x*s + h0c.y*c
    return make_float2((s.x + s.z) * cs - (s.y - s.w) * sn, (s.x - s.z) * sn + (s.y + s.w) * cs);
}

__global__ void k_evolve(const float4* __restrict__ spec, const float* __restrict__ omega, float2* __restrict__ out,
                         int64_t n2, int tiles, float t)
{
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n2 * tiles) return;
    const float4 s = spec[gid];
    const float omegat = __fmul_rn(omega[gid % n2], t);
    float sn, cs;
    sincosf(omegat, &sn, &cs);
    // literal operation order of :188 (no regrouping) for the debug/parity entry point
    const float rx = s.x * cs - s.y * sn + s.z * cs - s.w * (-sn);
    const float ry = s.x * sn + s.y * cs + s.z * (-sn) + s.w * cs;
    out[gid] = make_float2(rx, ry);
}

// =============================================================================================
// pass 1: evolve + pack + row FFT
// =============================================================================================
struct RowArgs {
    const float4* spec;    // [tiles][N][N]  (h0, h0conj)
    const float* omega;    // [N][N]
    const float2* ramp;    // [2N]  exp(i pi s (1-N)/N), s = n + m
    const float* kd;       // [N]   2 pi (i - N/2) / L, fp32 as FFTMesh.cs:201
    const float2* tw;      // [N]   exp(+2 pi i x / N)
    float2* X;             // [tiles][3][N][N]
    float t;
};

// RP row pairs per CTA; 6 lines per pair (2 rows x 3 fields); N/32 threads per line.
template <int N, int RP>
__global__ void __launch_bounds__(RP * 6 * (N / 32)) k_spectrum_rows(const RowArgs a)
{
    using P = Plan<N>;
    constexpr int T = P::T;
    constexpr int PAIR_THREADS = 6 * T;
    extern __shared__ float2 smem[];

    const int tile = blockIdx.y;
    const int rp = threadIdx.x / PAIR_THREADS;
    const int lt = threadIdx.x % PAIR_THREADS;
    const int pair = blockIdx.x * RP + rp;  // < N/2
    float2* lines = smem + rp * 6 * P::PITCH;

    const int rA = pair == 0 ? 0 : pair;
    const int rB = pair == 0 ? N / 2 : N - pair;
    const float4* spec = a.spec + (size_t)tile * N * N;

    // ---- evolve + pack: one task = a grid point and its mirror (-k) ----
    const int ntask = pair == 0 ? N + 2 : N;
    for (int task = lt; task < ntask; task += PAIR_THREADS) {
        int n1, m1, n2, m2, sel1, sel2;
        if (pair != 0) {
            n1 = rA; m1 = task; n2 = rB; m2 = (N - task) & (N - 1); sel1 = 0; sel2 = 1;
        } else {  // rows 0 and N/2 mirror onto themselves
            const int half = task >= N / 2 + 1;
            n1 = n2 = half ? N / 2 : 0;
            m1 = task - half * (N / 2 + 1);
            m2 = (N - m1) & (N - 1);
            sel1 = sel2 = half;
        }
        const float4 s1 = ldg_stream4(spec + n1 * N + m1);
        const float4 s2 = ldg_stream4(spec + n2 * N + m2);
        const float omegat = __fmul_rn(__ldg(a.omega + n1 * N + m1), a.t);  // FFTMesh.cs:183
        float sn, cs;
        sincosf(omegat, &sn, &cs);
        const float2 E1 = cmul(htilde_eval(s1, cs, sn), __ldg(a.ramp + n1 + m1));
        const float2 E2 = cmul(htilde_eval(s2, cs, sn), __ldg(a.ramp + n2 + m2));
        const float kx1 = __ldg(a.kd + n1), kz1 = __ldg(a.kd + m1);
        const float kx2 = __ldg(a.kd + n2), kz2 = __ldg(a.kd + m2);
        const float k2 = kx1 * kx1 + kz1 * kz1;
        const float inv = k2 < MW_EPSILON_F * MW_EPSILON_F ? 0.0f : rsqrtf(k2);  // FFTMesh.cs:213-214
        const float2 E1c = cconj(E1), E2c = cconj(E2);
        // F[n,m] = ((p1) E1 - (p2) conj(E2)) * (-i/2), p = (kx + i kz) or (kx + i kz)/|k|
        auto pack = [](float2 p1, float2 e1, float2 p2, float2 e2c) {
            const float2 d = csub(cmul(p1, e1), cmul(p2, e2c));
            return make_float2(0.5f * d.y, -0.5f * d.x);
        };
        const float2 pk1 = make_float2(kx1, kz1), pk2 = make_float2(kx2, kz2);
        const float2 pu1 = make_float2(kx1 * inv, kz1 * inv), pu2 = make_float2(kx2 * inv, kz2 * inv);
        float2* l1 = lines + sel1 * 3 * P::PITCH + pad_idx(m1);
        float2* l2 = lines + sel2 * 3 * P::PITCH + pad_idx(m2);
        l1[0] = pack(pu1, E1, pu2, E2c);            // field 0: chop displacement (Dx, Dz)
        l1[P::PITCH] = pack(pk1, E1, pk2, E2c);     // field 1: slopes (sx, sz)
        l1[2 * P::PITCH] = E1;                      // field 2: height
        l2[0] = pack(pu2, E2, pu1, E1c);
        l2[P::PITCH] = pack(pk2, E2, pk1, E1c);
        l2[2 * P::PITCH] = E2;
    }
    __syncthreads();

    // ---- row FFT: line q = sel * 3 + field ----
    const int q = lt / T, g = lt % T;
    const int sel = q / 3, f = q % 3;
    const int row = sel ? rB : rA;
    float2* dst = a.X + (((size_t)tile * 3 + f) * N + row) * N;
    mwfft::fft_line<N, +1>(lines + q * P::PITCH, g, true, a.tw, [&](int idx, float2 v) { dst[idx] = v; });
}

// =============================================================================================
// pass 2: column FFT + extraction (+ Jacobian whitecap)
// =============================================================================================
struct ColArgs {
    const float2* X;    // [tiles][3][N][N]
    const float2* tw;   // [N]
    float* height;      // [tiles][N*N]     or NULL
    float2* disp;       // [tiles][N*N]     or NULL   (hds)
    float* normal;      // [tiles][N*N][3]  or NULL
    float* whitecap;    // [tiles][N*N]     or NULL
    float* jacobian;    // [tiles][N*N]     or NULL
};

// Slab of W columns per CTA; W + 1 thread groups (the last one transforms the halo column b0 + W of
// the displacement field so that hds[index + 1] of FFTMesh.cs:266 is on chip).
template <int N, int W>
__global__ void __launch_bounds__((W + 1) * (N / 32)) k_cols_extract(const ColArgs a)
{
    using P = Plan<N>;
    constexpr int T = P::T;
    constexpr int MAIN = W * T;  // threads that own the W real columns
    extern __shared__ float2 smem[];
    float2* lines = smem;                                               // [W + 1][PITCH]
    float* noise = reinterpret_cast<float*>(smem + (W + 1) * P::PITCH); // [N][W]

    const int tile = blockIdx.y;
    const int b0 = blockIdx.x * W;
    const int tid = threadIdx.x;
    const int q = tid / T, g = tid % T;
    const bool is_halo = q == W;
    const bool halo_live = b0 + W < N;
    const size_t plane = (size_t)N * N;
    const size_t obase = (size_t)tile * plane;
    const bool want_white = a.whitecap != nullptr || a.jacobian != nullptr;

    // field order: 2 (height), 1 (slopes -> normal, noise), 0 (displacement -> hds, Jacobian)
#pragma unroll 1
    for (int f = 2; f >= 0; --f) {
        const float2* Xf = a.X + ((size_t)tile * 3 + f) * plane;
        const bool skip = (f == 2 && !a.height) || (f == 1 && !a.normal && !want_white) ||
                          (f == 0 && !a.disp && !want_white);
        if (skip) continue;  // uniform across the CTA
        // ---- transposing load: rows of W float2 from global -> line c at position n ----
        if (!is_halo) {
            float2 v[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const int e = tid + k * MAIN;
                v[k] = __ldg(Xf + (size_t)(e / W) * N + b0 + (e % W));
            }
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const int e = tid + k * MAIN;
                lines[(e % W) * P::PITCH + pad_idx(e / W)] = v[k];
            }
        } else if (f == 0 && halo_live && want_white) {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const int n = g + k * T;
                lines[W * P::PITCH + pad_idx(n)] = __ldg(Xf + (size_t)n * N + b0 + W);
            }
        }
        __syncthreads();
        // ---- column FFT; results (sign-fixed, real pairs) go back into the line ----
        {
            const bool active = !is_halo || (f == 0 && halo_live && want_white);
            float2* line = lines + q * P::PITCH;
            const int b = b0 + q;
            // sigma[a,b] = -(-1)^(a+b); displacement field also carries Dz's extra minus (FFTMesh.cs:215)
            mwfft::fft_line<N, +1>(line, g, active, a.tw, [&](int idx, float2 v) {
                const float s = ((idx + b) & 1) ? 1.0f : -1.0f;
                line[pad_idx(idx)] = make_float2(s * v.x, f == 0 ? -s * v.y : s * v.y);
            });
        }
        __syncthreads();
        // ---- extraction: thread <-> (row a = e / W, column c = e % W), c fastest ----
        if (!is_halo) {
#pragma unroll 4
            for (int k = 0; k < 32; ++k) {
                const int e = tid + k * MAIN;
                const int ar = e / W, c = e % W;
                const size_t o = obase + (size_t)ar * N + b0 + c;
                const float2 val = lines[c * P::PITCH + pad_idx(ar)];
                if (f == 2) {
                    a.height[o] = val.x;  // FFTMesh.cs:219 h.x
                } else if (f == 1) {
                    // nor = normalize(up - n) = (sx, 1, sz) / |.|   (FFTMesh.cs:212, 218)
                    const float inv = rsqrtf(val.x * val.x + 1.0f + val.y * val.y);
                    const float nx = val.x * inv, ny = inv, nz = val.y * inv;
                    if (a.normal) {
                        a.normal[3 * o + 0] = nx;
                        a.normal[3 * o + 1] = ny;
                        a.normal[3 * o + 2] = nz;
                    }
                    // noise = |(|n.x|, |n.z|) * 0.3|   (FFTMesh.cs:269-270)
                    const float ax = fabsf(nx) * 0.3f, az = fabsf(nz) * 0.3f;
                    noise[ar * W + c] = sqrtf(ax * ax + az * az);
                } else {
                    if (a.disp) a.disp[o] = val;  // hds (FFTMesh.cs:247)
                    if (want_white) {
                        float2 dDdx = make_float2(0.f, 0.f), dDdy = make_float2(0.f, 0.f);
                        if (ar != N - 1) {  // hds[index + resolution]  (:260-263)
                            const float2 nb = lines[c * P::PITCH + pad_idx(ar + 1)];
                            dDdx = make_float2(0.5f * (val.x - nb.x), 0.5f * (val.y - nb.y));
                        }
                        if (b0 + c != N - 1) {  // hds[index + 1]  (:264-267)
                            const float2 nb = lines[(c + 1) * P::PITCH + pad_idx(ar)];
                            dDdy = make_float2(0.5f * (val.x - nb.x), 0.5f * (val.y - nb.y));
                        }
                        const float jac = (1.0f + dDdx.x) * (1.0f + dDdy.y) - dDdx.y * dDdy.x;  // :268
                        if (a.jacobian) a.jacobian[o] = jac;
                        if (a.whitecap) {
                            float turb = fmaxf(1.0f - jac + noise[ar * W + c], 0.0f);  // :270
                            turb = fminf(turb, 1.0f);                                  // SmoothStep clamps
                            a.whitecap[o] = -2.0f * turb * turb * turb + 3.0f * turb * turb;  // :273
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
}

// =============================================================================================
// optional mesh-facing outputs (FFTMesh.cs:243-245, 274): displaced vertices and Color[]
// =============================================================================================
__global__ void k_mesh_outputs(const float* __restrict__ height, const float2* __restrict__ disp,
                               const float* __restrict__ whitecap, float* __restrict__ vertices,
                               float4* __restrict__ colors, int N, int tiles, float unit_width, float choppiness)
{
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n2 = (int64_t)N * N;
    if (gid >= n2 * tiles) return;
    const int idx = (int)(gid % n2);
    const int i = idx / N, j = idx % N;
    if (vertices) {
        // rest position, FFTMesh.cs:107-112 (N is even): (i - N/2) * uw + uw / 2
        const float off = __fdiv_rn(unit_width, 2.0f);
        const float px = __fadd_rn(__fmul_rn((float)(i - N / 2), unit_width), off);
        const float pz = __fadd_rn(__fmul_rn((float)(j - N / 2), unit_width), off);
        const float2 d = disp[gid];
        vertices[3 * gid + 0] = __fsub_rn(px, __fmul_rn(d.x, choppiness));  // :245
        vertices[3 * gid + 1] = height[gid];                                 // :243
        vertices[3 * gid + 2] = __fsub_rn(pz, __fmul_rn(d.y, choppiness));  // :244
    }
    if (colors) {
        const float w = whitecap[gid];
        colors[gid] = make_float4(w, w, w, w);  // :274
    }
}

}  // namespace mwk
