// mw_ocean_kernels.cuh -- device code of the Tessendorf hot path (sm_100a).
//
// Frame pipeline (replaces FFTMesh.EvaluateWaves, Scripts/FFTMesh.cs:224-280):
//
//   k_spectrum_rows   h0,h0conj --evolve--> h(k,t) --phase ramp + Hermitian packing--> 3 complex
//                     fields --row FFT (along m)--> intermediate XAB / XC            (pass 1)
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
// Intermediate layout (ours to choose; 24 B per grid point):
//   XAB[tile][n][b] : float4 (A.re, B.re, A.im, B.im)   A = chop-displacement field, B = slope field
//   XC [tile][n][b] : float2 (C.re, C.im)               C = height field
struct RowArgs {
    const float4* spec;    // [tiles][N][N]  (h0, h0conj)
    const float* omega;    // [N][N]
    const float2* ramp;    // [2N]  exp(i pi s (1-N)/N), s = n + m
    const float* kd;       // [N]   2 pi (i - N/2) / L, fp32 as FFTMesh.cs:201
    const float2* tw;      // [N]   exp(+2 pi i x / N)
    float4* XAB;           // [tiles][N][N]
    float2* XC;            // [tiles][N][N]
    float t;
};

// F[n,m] = (p1 * e1 - p2 * conj(e2)) * (-i/2): the Hermitian "imaginary part" packing of SURVEY 3.4
__device__ __forceinline__ float2 herm_pack(float2 p1, float2 e1, float2 p2, float2 e2)
{
    const float2 d = csub(cmul(p1, e1), cmul(p2, cconj(e2)));
    return make_float2(0.5f * d.y, -0.5f * d.x);
}

// RP row pairs per CTA; 3 packed lines per pair: (A,B) of row rA, (A,B) of row rB, (C of rA, C of rB).
template <int N, int RP, int MINB>
__global__ void __launch_bounds__(RP * 3 * (N / 16), MINB) k_spectrum_rows(const RowArgs a)
{
    using P = Plan<N>;
    constexpr int T = P::T;
    constexpr int PAIR_THREADS = 3 * T;
    constexpr int PITCH = mwfft::plane_pitch(N, 8);
    extern __shared__ float2 smem[];

    const int tile = blockIdx.y;
    const int rp = threadIdx.x / PAIR_THREADS;
    const int lt = threadIdx.x % PAIR_THREADS;
    const int pair = blockIdx.x * RP + rp;  // < N/2
    float2* lines = smem + rp * 6 * PITCH;  // line q: re plane at (2q) * PITCH, im plane at (2q+1) * PITCH

    const bool special = pair == 0;         // rows 0 and N/2 mirror onto themselves
    const int rA = special ? 0 : pair;
    const int rB = special ? N / 2 : N - pair;
    const float4* spec = a.spec + (size_t)tile * N * N;
    const float kxA = __ldg(a.kd + rA), kxB = __ldg(a.kd + rB);

    // ---- evolve + pack.  One task = the four grid points (rA|rB, m|m') with m' = -m mod N; the set is
    //      closed under k -> -k, so every Hermitian partner is on hand and every point is read once. ----
    for (int m = lt; m <= N / 2; m += PAIR_THREADS) {
        const int mm = (N - m) & (N - 1);
        const float4 s1 = ldg_stream4(spec + rA * N + m);    // P1 = (rA, m)
        const float4 s2 = ldg_stream4(spec + rB * N + mm);   // P2 = (rB, m')
        const float4 s3 = ldg_stream4(spec + rA * N + mm);   // P3 = (rA, m')
        const float4 s4 = ldg_stream4(spec + rB * N + m);    // P4 = (rB, m)
        // omega depends on |k| only: general rows  w(P1) = w(P2), w(P3) = w(P4);
        //                            special rows  w(P1) = w(P3), w(P4) = w(P2)
        const float w1 = __ldg(a.omega + rA * N + m);
        const float w2 = __ldg(a.omega + (special ? rB * N + m : rA * N + mm));
        float sn1, cs1, sn2, cs2;
        sincosf(__fmul_rn(w1, a.t), &sn1, &cs1);  // FFTMesh.cs:183
        sincosf(__fmul_rn(w2, a.t), &sn2, &cs2);
        const float2 E1 = cmul(htilde_eval(s1, cs1, sn1), __ldg(a.ramp + rA + m));
        const float2 E2 = cmul(special ? htilde_eval(s2, cs2, sn2) : htilde_eval(s2, cs1, sn1), __ldg(a.ramp + rB + mm));
        const float2 E3 = cmul(special ? htilde_eval(s3, cs1, sn1) : htilde_eval(s3, cs2, sn2), __ldg(a.ramp + rA + mm));
        const float2 E4 = cmul(htilde_eval(s4, cs2, sn2), __ldg(a.ramp + rB + m));
        const float kzm = __ldg(a.kd + m), kzmm = __ldg(a.kd + mm);
        // |k| is shared inside a mirror pair; it differs between rows A and B only for the special pair
        const float k2A = kxA * kxA + kzm * kzm, k2B = kxB * kxB + kzm * kzm;
        const float invA = k2A < MW_EPSILON_F * MW_EPSILON_F ? 0.0f : rsqrtf(k2A);  // FFTMesh.cs:213-214
        const float invB = k2B < MW_EPSILON_F * MW_EPSILON_F ? 0.0f : rsqrtf(k2B);
        const float2 k1 = make_float2(kxA, kzm), k2 = make_float2(kxB, kzmm), k3 = make_float2(kxA, kzmm), k4 = make_float2(kxB, kzm);
        const float2 u1 = make_float2(k1.x * invA, k1.y * invA), u3 = make_float2(k3.x * invA, k3.y * invA);
        const float2 u2 = make_float2(k2.x * invB, k2.y * invB), u4 = make_float2(k4.x * invB, k4.y * invB);
        float2 A1, A2, A3, A4, B1, B2, B3, B4;
        if (!special) {  // partners: P1 <-> P2, P3 <-> P4
            A1 = herm_pack(u1, E1, u2, E2); A2 = herm_pack(u2, E2, u1, E1);
            A3 = herm_pack(u3, E3, u4, E4); A4 = herm_pack(u4, E4, u3, E3);
            B1 = herm_pack(k1, E1, k2, E2); B2 = herm_pack(k2, E2, k1, E1);
            B3 = herm_pack(k3, E3, k4, E4); B4 = herm_pack(k4, E4, k3, E3);
        } else {         // partners: P1 <-> P3, P4 <-> P2
            A1 = herm_pack(u1, E1, u3, E3); A3 = herm_pack(u3, E3, u1, E1);
            A4 = herm_pack(u4, E4, u2, E2); A2 = herm_pack(u2, E2, u4, E4);
            B1 = herm_pack(k1, E1, k3, E3); B3 = herm_pack(k3, E3, k1, E1);
            B4 = herm_pack(k4, E4, k2, E2); B2 = herm_pack(k2, E2, k4, E4);
        }
        const int pm = pad_idx(m), pmm = pad_idx(mm);
        // line 0 = (A,B) of row rA ; line 1 = (A,B) of row rB ; line 2 = (C of rA, C of rB)
        lines[0 * PITCH + pm] = make_float2(A1.x, B1.x);  lines[1 * PITCH + pm] = make_float2(A1.y, B1.y);
        lines[0 * PITCH + pmm] = make_float2(A3.x, B3.x); lines[1 * PITCH + pmm] = make_float2(A3.y, B3.y);
        lines[2 * PITCH + pm] = make_float2(A4.x, B4.x);  lines[3 * PITCH + pm] = make_float2(A4.y, B4.y);
        lines[2 * PITCH + pmm] = make_float2(A2.x, B2.x); lines[3 * PITCH + pmm] = make_float2(A2.y, B2.y);
        lines[4 * PITCH + pm] = make_float2(E1.x, E4.x);  lines[5 * PITCH + pm] = make_float2(E1.y, E4.y);
        lines[4 * PITCH + pmm] = make_float2(E3.x, E2.x); lines[5 * PITCH + pmm] = make_float2(E3.y, E2.y);
    }
    __syncthreads();

    // ---- row FFT of the three packed lines ----
    const int q = lt / T, g = lt % T;
    float2* pre = lines + 2 * q * PITCH;
    float2* pim = pre + PITCH;
    const size_t tbase = (size_t)tile * N * N;
    if (q < 2) {
        float4* dst = a.XAB + tbase + (size_t)(q ? rB : rA) * N;
        mwfft::fft_line<N, +1>(pre, pim, g, rp * 3 + q, true, a.tw,
                               [&](int idx, mwfft::cpk v) { dst[idx] = make_float4(v.re.x, v.re.y, v.im.x, v.im.y); });
    } else {
        float2* dA = a.XC + tbase + (size_t)rA * N;
        float2* dB = a.XC + tbase + (size_t)rB * N;
        mwfft::fft_line<N, +1>(pre, pim, g, rp * 3 + q, true, a.tw, [&](int idx, mwfft::cpk v) {
            dA[idx] = make_float2(v.re.x, v.im.x);
            dB[idx] = make_float2(v.re.y, v.im.y);
        });
    }
}

// =============================================================================================
// pass 2: column FFT + extraction (+ Jacobian whitecap)
// =============================================================================================
struct ColArgs {
    const float4* XAB;  // [tiles][N][N]
    const float2* XC;   // [tiles][N][N]
    const float2* tw;   // [N]
    float* height;      // [tiles][N*N]     or NULL
    float2* disp;       // [tiles][N*N]     or NULL   (hds)
    float* normal;      // [tiles][N*N][3]  or NULL
    float* whitecap;    // [tiles][N*N]     or NULL
    float* jacobian;    // [tiles][N*N]     or NULL
};

// Slab of W columns per CTA, W + 1 packed-line thread groups.
//   phase 1: the (A,B) pairs of the W columns + the halo column b0 + W (so that hds[index + 1] of
//            FFTMesh.cs:266 is on chip)  -> hds, normal, Jacobian, whitecap
//   phase 2: the C field of the W columns, two columns per packed line  -> height
template <int N, int W, int MINB>
__global__ void __launch_bounds__((W + 1) * (N / 16), MINB) k_cols_extract(const ColArgs a)
{
    using P = Plan<N>;
    constexpr int T = P::T;
    constexpr int MAIN = W * T;  // threads that own the W real columns
    constexpr int PITCH = mwfft::plane_pitch(N, W);
    constexpr int LOGW = mwfft::ilog2(W);
    extern __shared__ float2 smem[];  // line q: re plane at (2q) * PITCH, im plane at (2q+1) * PITCH

    const int tile = blockIdx.y;
    const int b0 = blockIdx.x * W;
    const int tid = threadIdx.x;
    const int q = tid / T, g = tid % T;
    const bool is_halo = q == W;
    const size_t plane = (size_t)N * N;
    const size_t obase = (size_t)tile * plane;
    const bool want_white = a.whitecap != nullptr || a.jacobian != nullptr;
    const bool halo_live = want_white && b0 + W < N;

    // ------------------------------------------------------------------ phase 1: (A, B)
    if (a.disp || a.normal || want_white) {
        const float4* X = a.XAB + obase;
        if (!is_halo) {
            float4 v[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int e = tid + k * MAIN;
                v[k] = __ldg(X + (size_t)(e >> LOGW) * N + b0 + (e & (W - 1)));
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int e = tid + k * MAIN;
                const int o = 2 * (e & (W - 1)) * PITCH + pad_idx(e >> LOGW);
                smem[o] = make_float2(v[k].x, v[k].y);
                smem[o + PITCH] = make_float2(v[k].z, v[k].w);
            }
        } else if (halo_live) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int n = g + k * T;
                const float4 h = __ldg(X + (size_t)n * N + b0 + W);
                smem[2 * W * PITCH + pad_idx(n)] = make_float2(h.x, h.y);
                smem[(2 * W + 1) * PITCH + pad_idx(n)] = make_float2(h.z, h.w);
            }
        }
        __syncthreads();
        {
            float2* pre = smem + 2 * q * PITCH;
            float2* pim = pre + PITCH;
            const int b = b0 + q;
            // sigma[a,b] = -(-1)^(a+b); lane x = field A -> (dx, dz) with Dz's extra minus (FFTMesh.cs:215),
            // lane y = field B -> (sx, sz).  Stored back as re plane = (dx, sx), im plane = (dz, sz).
            mwfft::fft_line<N, +1>(pre, pim, g, q, !is_halo || halo_live, a.tw, [&](int idx, mwfft::cpk v) {
                const float s = ((idx + b) & 1) ? 1.0f : -1.0f;
                const int p = pad_idx(idx);
                pre[p] = make_float2(s * v.re.x, s * v.re.y);
                pim[p] = make_float2(-s * v.im.x, s * v.im.y);
            });
        }
        __syncthreads();
        if (!is_halo) {
#pragma unroll 4
            for (int k = 0; k < 16; ++k) {
                const int e = tid + k * MAIN;
                const int ar = e >> LOGW, c = e & (W - 1);
                const size_t o = obase + (size_t)ar * N + b0 + c;
                const float2* pre = smem + 2 * c * PITCH;
                const float2* pim = pre + PITCH;
                const float2 vr = pre[pad_idx(ar)], vi = pim[pad_idx(ar)];  // (dx, sx), (dz, sz)
                // nor = normalize(up - n) = (sx, 1, sz) / |.|   (FFTMesh.cs:212, 218)
                const float inv = rsqrtf(vr.y * vr.y + 1.0f + vi.y * vi.y);
                const float nx = vr.y * inv, nz = vi.y * inv;
                if (a.normal) {
                    a.normal[3 * o + 0] = nx;
                    a.normal[3 * o + 1] = inv;
                    a.normal[3 * o + 2] = nz;
                }
                if (a.disp) a.disp[o] = make_float2(vr.x, vi.x);  // hds (FFTMesh.cs:247)
                if (want_white) {
                    float2 dDdx = make_float2(0.f, 0.f), dDdy = make_float2(0.f, 0.f);
                    if (ar != N - 1) {  // hds[index + resolution]  (:260-263)
                        const float nbx = pre[pad_idx(ar + 1)].x, nbz = pim[pad_idx(ar + 1)].x;
                        dDdx = make_float2(0.5f * (vr.x - nbx), 0.5f * (vi.x - nbz));
                    }
                    if (b0 + c != N - 1) {  // hds[index + 1]  (:264-267)
                        const float nbx = pre[2 * PITCH + pad_idx(ar)].x, nbz = pim[2 * PITCH + pad_idx(ar)].x;
                        dDdy = make_float2(0.5f * (vr.x - nbx), 0.5f * (vi.x - nbz));
                    }
                    const float jac = (1.0f + dDdx.x) * (1.0f + dDdy.y) - dDdx.y * dDdy.x;  // :268
                    if (a.jacobian) a.jacobian[o] = jac;
                    if (a.whitecap) {
                        // noise = |(|n.x|, |n.z|) * 0.3|   (:269-270)
                        const float ax = fabsf(nx) * 0.3f, az = fabsf(nz) * 0.3f;
                        float turb = fmaxf(1.0f - jac + sqrtf(ax * ax + az * az), 0.0f);  // :270
                        turb = fminf(turb, 1.0f);                                          // SmoothStep clamps
                        a.whitecap[o] = -2.0f * turb * turb * turb + 3.0f * turb * turb;   // :273
                    }
                }
            }
        }
        __syncthreads();
    }

    // ------------------------------------------------------------------ phase 2: C (height), W/2 packed lines
    if (a.height) {
        constexpr int HL = W / 2;       // packed lines
        constexpr int HMAIN = HL * T;   // threads at work
        const float2* X = a.XC + obase;
        const bool on = tid < HMAIN;
        if (on) {
            float4 v[16];  // (C[n][b].re, C[n][b].im, C[n][b+1].re, C[n][b+1].im)
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int e = tid + k * HMAIN;
                const int n = e / HL, c2 = e % HL;
                v[k] = __ldg(reinterpret_cast<const float4*>(X + (size_t)n * N + b0 + 2 * c2));
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int e = tid + k * HMAIN;
                const int n = e / HL, c2 = e % HL;
                const int o = 2 * c2 * PITCH + pad_idx(n);
                smem[o] = make_float2(v[k].x, v[k].z);
                smem[o + PITCH] = make_float2(v[k].y, v[k].w);
            }
        }
        __syncthreads();
        {
            float2* pre = smem + 2 * q * PITCH;
            float2* pim = pre + PITCH;
            const int b = b0 + 2 * q;  // lane x = column b, lane y = column b + 1 (opposite sigma)
            mwfft::fft_line<N, +1>(pre, pim, g, q, on, a.tw, [&](int idx, mwfft::cpk v) {
                const float s = ((idx + b) & 1) ? 1.0f : -1.0f;
                pre[pad_idx(idx)] = make_float2(s * v.re.x, -s * v.re.y);  // height = sigma * Re (FFTMesh.cs:219)
            });
        }
        __syncthreads();
        if (on) {
#pragma unroll 4
            for (int k = 0; k < 16; ++k) {
                const int e = tid + k * HMAIN;
                const int n = e / HL, c2 = e % HL;
                const float2 h = smem[2 * c2 * PITCH + pad_idx(n)];
                *reinterpret_cast<float2*>(a.height + obase + (size_t)n * N + b0 + 2 * c2) = h;
            }
        }
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
