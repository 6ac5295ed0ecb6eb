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
#include "mw_layout.cuh"

namespace mwk {

using mwfft::Plan;
using mwfft::pad_idx;

// =============================================================================================
// init-time kernels
// =============================================================================================

// Dispersion(n, m), FFTMesh.cs:141-147, bit-exact: every operation is the fp32 round-to-nearest
// one the C# expression performs, in the same order, with no FMA contraction.
// The quantisation makes omega an integer multiple q of w0 = 2 pi / L: omega = fl(q * w0).  q is kept next to omega
// so that the per-frame e^{i omega t} can come from a table with one entry per q (k_phase_table) -- the entry is
// computed from the same fp32 omega with the same fp32 product omega * t, so the values are the ones a per-point
// evaluation gives, bit for bit.
__device__ __forceinline__ float dispersion_w0(float length) { return __fdiv_rn(__fmul_rn(2.0f, MW_PI_F), length); }
__device__ __forceinline__ float dispersion_q(int n, int m, int N, float length)
{
    const float w = dispersion_w0(length);
    const float kx = __fdiv_rn(__fmul_rn(MW_PI_F, (float)(2 * n - N)), length);
    const float kz = __fdiv_rn(__fmul_rn(MW_PI_F, (float)(2 * m - N)), length);
    const float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(kx, kx), __fmul_rn(kz, kz)));
    return floorf(__fdiv_rn(__fsqrt_rn(__fmul_rn(MW_G_F, mag)), w));
}
__device__ __forceinline__ float dispersion_rn(int n, int m, int N, float length)
{
    return __fmul_rn(dispersion_q(n, m, N, length), dispersion_w0(length));
}

__global__ void k_dispersion(float* __restrict__ omega, int* __restrict__ qidx, int N, float length)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * N) return;
    const float q = dispersion_q(idx / N, idx % N, N, length);
    omega[idx] = __fmul_rn(q, dispersion_w0(length));
    qidx[idx] = (int)q;
}

// Per frame: ptab[q] = (cos, sin)(fl(fl(q * w0) * t)), FFTMesh.cs:183-185 for every distinct omega of the grid.
__global__ void k_phase_table(float2* __restrict__ ptab, int entries, float length, float t)
{
    pdl_trigger();  // pass 1 of the first tile group may become resident and fetch its spectrum meanwhile
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= entries) return;
    const float omegat = __fmul_rn(__fmul_rn((float)q, dispersion_w0(length)), t);
    float sn, cs;
    sincosf(omegat, &sn, &cs);
    ptab[q] = make_float2(cs, sn);
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

// What pass 1 reads every frame: -(h0, h0conj) * ramp[n + m].  The phase ramp r[n] r[m] and the constant sign of the
// height field (C' = -H r r, see "Signs" below) are constant in time and commute with the evolution
// h0 e^{iwt} + h0conj e^{-iwt}, so they are applied once here instead of once per point per frame.
__global__ void k_ramp_spectrum(const float4* __restrict__ spec, const float2* __restrict__ ramp, float4* __restrict__ spec_r,
                                int N, int64_t total)
{
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int idx = (int)(gid % ((int64_t)N * N));
    const float2 r = ramp[idx / N + idx % N];
    const float4 s = spec[gid];
    const float2 a = cmul(make_float2(s.x, s.y), r), b = cmul(make_float2(s.z, s.w), r);
    spec_r[gid] = make_float4(-a.x, -a.y, -b.x, -b.y);
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
// Intermediate layout (ours to choose), SLAB-MAJOR so that what one pass-2 CTA reads is one contiguous
// block of memory (a row-major intermediate makes pass 2 read 64-byte pieces 16 KB apart, which runs the
// HBM at a fraction of its bandwidth -- measured):
//   XAB[tile][b / 8][n][8]   : float4 (A.re, B.re, A.im, B.im), A = chop-displacement field, B = slope field
//                              (columns 8s..8s+7 of row n: one 128-byte line)          16 B per grid point
//   XC [tile][b / 16][n][16] : float2 (C.re, C.im), C = height field                     8 B per grid point
// Pass 1 therefore scatters whole 128-byte lines (writes: merged in the 126 MB L2 before they reach HBM).  The
// halo column a pass-2 CTA needs (column 8s+8, for the Jacobian's forward difference) is read from the next slab:
// a 16-byte read out of every 128-byte row there -- reads tolerate that, writes of partial sectors do not.
// Slabs are 8 (16) columns wide because pass 2 must WRITE its outputs in rows of that many columns: measured
// (tools/ubench/store_pattern.cu), 4-column output rows leave half-filled 32-byte sectors (whitecap 16 B,
// normal 48 B per row) and the same bytes take 224 us instead of 71 us (8 columns) per 16 tiles.
// (N = 2048: 4-column slabs -- nine 2048-point packed lines do not fit in shared memory.)
struct RowArgs {
    const float4* spec;    // [tiles][N][N]  -(h0, h0conj) * ramp[n + m]   (k_ramp_spectrum)
    const int* qidx;       // [N][N]  omega / w0  (k_dispersion)
    const float2* ptab;    // [q_max + 1]  (cos, sin)(omega_q t) of this frame  (k_phase_table)
    const float* kd;       // [N]   2 pi (i - N/2) / L, fp32 as FFTMesh.cs:201
    const float4* twimg;   // shared-memory twiddle tables as a ready-made image (mwfft::twiddle_image_host)
    float4* XAB;           // [tiles][N/8][N][8]
    float2* XC;            // [tiles][N/32][N][32] (mw_layout.cuh: xc4_index)
    int tile0;             // first tile of this launch (blockIdx.y counts from it); X is indexed by blockIdx.y
    long long* dbg;        // developer phase-timing buffer (NULL in production)
    int dbg_flags;         // developer experiments: 1 = no output stores, 2 = no FFT, 4 = no spectrum loads
    int pdl;               // programmatic dependent launch: 1 = release the successor at CTA start, 2 = before the result stores
    float length, t;       // INLINE_PHASE kernels only: e^{i omega t} is evaluated here instead of being read from ptab
    unsigned* seam_flags;  // [tiles][seam_nab] "first column published" flags of k_cols_seam (mw_cols_seam.cuh), cleared here; or NULL
    int seam_nab;
};

// Signs.  The direct sum equals sigma[a,b] * T[a,b] with sigma = -(-1)^(a+b), and Dz carries an
// extra minus (FFTMesh.cs:215).  None of that costs an instruction here:
//   * (-1)^b : pass 1 stores spectrum element m at line position (m + N/2) mod N   (shift theorem);
//   * (-1)^a : pass 2 stores intermediate row n at line position (n + N/2) mod N;
//   * the remaining constant signs go into the packing multipliers below, chosen so that the finished
//     transforms ARE the outputs:  A' -> (dx + i dz),  B' -> (sx + i sz),  C' -> height (real part).
//       A' = -Im-part(ux Hc) + i Im-part(uz Hc)   => multiplier (-ux + i uz)
//       B' = -Im-part(kx H)  - i Im-part(kz H)    => multiplier (-kx - i kz)
//       C' = -H
//
// Hermitian "imaginary part" packing: with E = H r r at a grid point P, E~ = the same at the mirror
// point -P and M the multiplier above,   F(P) = (-i/2) (M(P) E(P) - M(-P) conj(E(-P))).
// The spectrum is stored as Eh = -E (k_ramp_spectrum), and M(-P) = -M(P) exactly (kd[N - i] == -kd[i] in fp32)
// everywhere except where an index mirrors onto itself (rows / columns 0 and N/2), so for a general row pair
//       F(P)  = Mh(P) (Eh(P) + conj(Eh(-P))),   F(-P) = -Mh(P) conj(Eh(P) + conj(Eh(-P))),   Mh = (i/2) M
// -- one packed complex product serves both points of a mirror pair.  (Column 0 is self-mirrored too: general form.)

// e^{i omega t} applied to the stored pair: a e + b conj(e)   (htilde, FFTMesh.cs:178-190)
__device__ __forceinline__ float2 htilde_tab(float4 s, float2 e)
{
    return make_float2((s.x + s.z) * e.x - (s.y - s.w) * e.y, (s.x - s.z) * e.y + (s.y + s.w) * e.x);
}

// RP row pairs per CTA; 3 packed lines per pair: (A,B) of row rA, (A,B) of row rB, (C of rA, C of rB).
// INLINE_PHASE (small single frames, where a frame is a chain of launch latencies): the two e^{i omega t} of a task are
// evaluated here -- sincosf(fl(fl(q w0) t)), the very expression k_phase_table tabulates, hence bit-identical -- so that the
// frame is two kernels instead of three.
template <int N, int RP, int MINB, bool INLINE_PHASE = false>
__global__ void __launch_bounds__(RP * 3 * (N / fft_pts(N)), MINB) k_spectrum_rows(const RowArgs a)
{
    constexpr int PTS = fft_pts(N);
    using P = Plan<N, PTS>;
    constexpr int T = P::T;
    constexpr int PAIR_THREADS = 3 * T;
    constexpr int LP = mwfft::line_pitch(N, 8);
    constexpr int NIT = (N / 2 + 1 + PAIR_THREADS - 1) / PAIR_THREADS;  // tasks per thread (3)
    extern __shared__ float4 smem4[];
    float4* tw2 = smem4;
    float2* tw3 = reinterpret_cast<float2*>(smem4 + P::TW2_F4);
    float4* all_lines = smem4 + P::TW_BYTES / 16;

    const int tile = a.tile0 + blockIdx.y;  // spectrum / global tile index
    const int xt = blockIdx.y;              // slot inside the intermediate buffer of this launch
    const int rp = threadIdx.x / PAIR_THREADS;
    const int lt = threadIdx.x % PAIR_THREADS;
    const int pair = blockIdx.x * RP + rp;  // < N/2
    float4* lines = all_lines + rp * 3 * LP;

#if MW_DEVHOOKS
#define MW_RSTAMP(i) do { if (a.dbg && threadIdx.x == 0) a.dbg[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 + (i)] = clock64(); } while (0)
#else
#define MW_RSTAMP(i) do { } while (0)
#endif
    MW_RSTAMP(0);
    if (a.pdl == 1) pdl_trigger();

    const bool special = pair == 0;         // rows 0 and N/2 mirror onto themselves
    const int rA = special ? 0 : pair;
    const int rB = special ? N / 2 : N - pair;
    const float4* specT = a.spec + (size_t)tile * N * N;   // CTA-uniform base; everything below is a 32-bit offset
    const unsigned oA = (unsigned)rA * N, oB = (unsigned)rB * N;

    // ---- evolve + pack.  One task = the four grid points (rA|rB, m|m') with m' = -m mod N; the set is
    //      closed under k -> -k, so every Hermitian partner is on hand and every point is read once.
    //      All loads of a thread's (up to) three tasks are issued before the first use. ----
    if (MW_DBG(a, 512)) return;
    if (!MW_DBG(a, 32)) {
        const uint64_t pol = evict_first_policy();
        float4 s1[NIT], s2[NIT], s3[NIT], s4[NIT];
        int q1[NIT], q2[NIT];
        float kz[NIT];
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int m0 = lt + it * PAIR_THREADS;
            const int m = m0 <= N / 2 ? m0 : N / 2;  // surplus threads: harmless repeat of the last task's loads
            const int mm = (N - m) & (N - 1);
            s1[it] = ldg_once4(specT + (oA + m), pol);    // P1 = (rA, m)
            s2[it] = ldg_once4(specT + (oB + mm), pol);   // P2 = (rB, m')
            s3[it] = ldg_once4(specT + (oA + mm), pol);   // P3 = (rA, m')
            s4[it] = ldg_once4(specT + (oB + m), pol);    // P4 = (rB, m)
            // omega depends on |k| only: general rows  w(P1) = w(P2), w(P3) = w(P4);
            //                            special rows  w(P1) = w(P3), w(P4) = w(P2)
            q1[it] = __ldg(a.qidx + (oA + m));
            q2[it] = __ldg(a.qidx + (special ? oB + m : oA + mm));
            kz[it] = __ldg(a.kd + m);
        }
        // Everything fetched so far is constant across frames.  The phase table is this frame's (k_phase_table), and the
        // intermediate written below may still be being read by the previous pass 2 of this stream: wait for the predecessor.
        pdl_wait();
        if (a.seam_flags && blockIdx.x == 0)   // (the previous pass 2 of this tile has completed: its flags can go)
            for (int i = threadIdx.x; i < a.seam_nab; i += RP * PAIR_THREADS) a.seam_flags[(size_t)tile * a.seam_nab + i] = 0u;
        float2 e1[NIT], e2[NIT];
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            if constexpr (INLINE_PHASE) {
                const float w0 = dispersion_w0(a.length);
                float sn, cs;
                sincosf(__fmul_rn(__fmul_rn((float)q1[it], w0), a.t), &sn, &cs);
                e1[it] = make_float2(cs, sn);
                sincosf(__fmul_rn(__fmul_rn((float)q2[it], w0), a.t), &sn, &cs);
                e2[it] = make_float2(cs, sn);
            } else {
                e1[it] = ldg_fresh2(a.ptab + q1[it]);
                e2[it] = ldg_fresh2(a.ptab + q2[it]);
            }
        }
        // the twiddle tables are fetched while the spectrum loads are in flight
        mwfft::load_twiddle_image<N, RP * PAIR_THREADS, PTS>(smem4, a.twimg);
        const float kxA = __ldg(a.kd + rA), kxB = __ldg(a.kd + rB);
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int m = lt + it * PAIR_THREADS;
            if (m > N / 2) break;                       // (last iteration only)
            const int mm = (N - m) & (N - 1);
            const float kzm = kz[it];
            const float kzmm = (m == 0) ? kzm : -kzm;   // kd[N - m] == -kd[m] exactly; kd[0] mirrors onto itself
            // Eh = -(H r r) at the four points (the spectrum is stored ramped and negated)
            const float2 E1 = htilde_tab(s1[it], e1[it]);
            const float2 E2 = htilde_tab(s2[it], special ? e2[it] : e1[it]);
            const float2 E3 = htilde_tab(s3[it], special ? e1[it] : e2[it]);
            const float2 E4 = htilde_tab(s4[it], e2[it]);
            // line positions shifted by N/2: the transform then carries the (-1)^b of sigma
            const int pm = pad_idx((m + N / 2) & (N - 1)), pmm = pad_idx((mm + N / 2) & (N - 1));
            float4 F1, F2, F3, F4;
            if (!special && m != 0) {
                // |k| is shared by the four points; Mh'' = -(i/2) M ... with Eh = -E:  F = Mh (Eh + conj(Eh~)),
                // Mh(P) = (i/2)(-M(P))... written out: re = (-hz inv, hz), im = (-hx inv, -hx), hx = kx/2, hz = kz/2
                // (lane x = the A' field, lane y = the B' field)
                const float k2 = kxA * kxA + kzm * kzm;
                const float inv = k2 < MW_EPSILON_F * MW_EPSILON_F ? 0.0f : rsqrtf(k2);  // FFTMesh.cs:213-214
                const float hx = 0.5f * kxA, hz = 0.5f * kzm, hzz = 0.5f * kzmm;
                const float2 are1 = make_float2(-hz * inv, hz), are3 = make_float2(-hzz * inv, hzz);
                const float2 aim = make_float2(-hx * inv, -hx);
                auto pack_pair = [](float2 are, float2 aim, float2 es, float2 ep, float4& Fs, float4& Fp) {
                    const float x = es.x + ep.x, y = es.y - ep.y;      // S = Eh(P) + conj(Eh(-P))
                    const float2 xx = make_float2(x, x), yy = make_float2(y, y);
                    const float2 q = __fmul2_rn(aim, yy), u = __fmul2_rn(aim, xx);
                    const float2 sre = __ffma2_rn(are, xx, mwfft::neg2(q));           // a x - b y
                    const float2 sim = __ffma2_rn(are, yy, u);                        // a y + b x
                    const float2 pre = __ffma2_rn(are, mwfft::neg2(xx), mwfft::neg2(q));  // -a x - b y
                    const float2 pim = __ffma2_rn(are, yy, mwfft::neg2(u));           // a y - b x
                    Fs = make_float4(sre.x, sre.y, sim.x, sim.y);
                    Fp = make_float4(pre.x, pre.y, pim.x, pim.y);
                };
                pack_pair(are1, aim, E1, E2, F1, F2);   // P1 <-> P2
                pack_pair(are3, aim, E3, E4, F3, F4);   // P3 <-> P4
            } else {
                // rows 0 and N/2 (partners P1 <-> P3 and P4 <-> P2) and column 0 of any row: the Nyquist index mirrors
                // onto itself with the same kd, so M(-P) != -M(P) there and the packing is done in its general form
                //   F = (i/2) (M_s Eh_s - M_p conj(Eh_p))
                const float k2A = kxA * kxA + kzm * kzm, k2B = kxB * kxB + kzm * kzm;
                const float invA = k2A < MW_EPSILON_F * MW_EPSILON_F ? 0.0f : rsqrtf(k2A);
                const float invB = k2B < MW_EPSILON_F * MW_EPSILON_F ? 0.0f : rsqrtf(k2B);
                const float uxA = -kxA * invA, uxB = -kxB * invB;
                const mwfft::cpk M1 = {make_float2(uxA, -kxA), make_float2(kzm * invA, -kzm)};    // P1 = (rA, m)
                const mwfft::cpk M3 = {make_float2(uxA, -kxA), make_float2(kzmm * invA, -kzmm)};  // P3 = (rA, m')
                const mwfft::cpk M2 = {make_float2(uxB, -kxB), make_float2(kzmm * invB, -kzmm)};  // P2 = (rB, m')
                const mwfft::cpk M4 = {make_float2(uxB, -kxB), make_float2(kzm * invB, -kzm)};    // P4 = (rB, m)
                auto pack2 = [](const mwfft::cpk& ms, float2 es, const mwfft::cpk& mp, float2 ep) {
                    const mwfft::cpk d = mwfft::psub(mwfft::pmul(ms, es.x, es.y), mwfft::pmul(mp, ep.x, -ep.y));
                    return make_float4(-0.5f * d.im.x, -0.5f * d.im.y, 0.5f * d.re.x, 0.5f * d.re.y);
                };
                if (special) {
                    F1 = pack2(M1, E1, M3, E3); F3 = pack2(M3, E3, M1, E1);
                    F4 = pack2(M4, E4, M2, E2); F2 = pack2(M2, E2, M4, E4);
                } else {
                    F1 = pack2(M1, E1, M2, E2); F2 = pack2(M2, E2, M1, E1);
                    F3 = pack2(M3, E3, M4, E4); F4 = pack2(M4, E4, M3, E3);
                }
            }
            // line 0 = (A,B) of row rA ; line 1 = (A,B) of row rB ; line 2 = (C of rA, C of rB), C' = Eh
            lines[pm] = F1;
            lines[pmm] = F3;
            lines[LP + pm] = F4;
            lines[LP + pmm] = F2;
            // C' is stored Hermitian-symmetrised, Cs(P) = (Eh(P) + conj(Eh(-P))) / 2 (= S / 2 above), Cs(-P) = conj(Cs(P)):
            // its finished transform is REAL (= height, the real part of the transform of C'), so pass 2 transforms two
            // columns per complex line (z = column j + i column j') and a C slab covers twice the columns.
            const float2 Sa = special ? make_float2(0.5f * (E1.x + E3.x), 0.5f * (E1.y - E3.y))    // P1 <-> P3
                                      : make_float2(0.5f * (E1.x + E2.x), 0.5f * (E1.y - E2.y));   // P1 <-> P2
            const float2 Sb = special ? make_float2(0.5f * (E4.x + E2.x), 0.5f * (E4.y - E2.y))    // P4 <-> P2
                                      : make_float2(0.5f * (E3.x + E4.x), 0.5f * (E3.y - E4.y));   // P3 <-> P4
            // special: Cs(P1) = Sa, Cs(P3) = conj(Sa), Cs(P4) = Sb, Cs(P2) = conj(Sb)
            // general: Cs(P1) = Sa, Cs(P2) = conj(Sa), Cs(P3) = Sb, Cs(P4) = conj(Sb)
            const float2 C1 = Sa;
            const float2 C2 = special ? make_float2(Sb.x, -Sb.y) : make_float2(Sa.x, -Sa.y);
            const float2 C3 = special ? make_float2(Sa.x, -Sa.y) : Sb;
            const float2 C4 = special ? Sb : make_float2(Sb.x, -Sb.y);
            lines[2 * LP + pm] = make_float4(C1.x, C4.x, C1.y, C4.y);
            lines[2 * LP + pmm] = make_float4(C3.x, C2.x, C3.y, C2.y);
        }
    } else {
        mwfft::load_twiddle_image<N, RP * PAIR_THREADS, PTS>(smem4, a.twimg);
        pdl_wait();
        if (a.seam_flags && blockIdx.x == 0)
            for (int i = threadIdx.x; i < a.seam_nab; i += RP * PAIR_THREADS) a.seam_flags[(size_t)tile * a.seam_nab + i] = 0u;
    }
    MW_RSTAMP(1);
    __syncthreads();
    MW_RSTAMP(2);

    // ---- row FFT of the three packed lines ----
    const int q = lt / T, g = lt % T;
    float4* line = lines + q * LP;
    if (MW_DBG(a, 64)) return;
    {
        // one instance of the transform for all three lines (code size: the kernel has to stay resident in
        // the instruction cache while several CTAs run different phases); results come back in registers
        mwfft::cpk v[PTS];
        mwfft::load_line_regs<N, PTS>(v, line, g);
        const int bar_id = rp * 3 + q;
        auto line_sync = [&] { mwfft::group_sync<T>(bar_id); };
        line_sync();  // everyone has read before anyone overwrites (in-place exchange)
        mwfft::fft_line_inreg<N, +1, PTS>(v, line, g, tw2, tw3, line_sync);
        if (a.pdl == 2) pdl_trigger();
        constexpr int W = slab_w(N);
        // (an evict-last priority on these intermediate stores was measured: no effect once the once-touched traffic is evict-first)
#define MW_XST(ptr, val) (*(ptr) = (val))
        if (q < 2) {
            const int row = q ? rB : rA;
            float4* dst = a.XAB + (size_t)xt * xab_tile_elems(N);
            if constexpr (T % W == 0) {
                // result index = g + (multiple of T): slab and column-in-slab of g, then compile-time slab steps
                const unsigned base = ((unsigned)(g / W) * N + row) * W + (g % W);
#pragma unroll
                for (int sl = 0; sl < PTS; ++sl)
                    MW_XST(dst + (base + (unsigned)(mwfft::final_off<N, PTS>(sl) / W) * (N * W)),
                           make_float4(v[sl].re.x, v[sl].re.y, v[sl].im.x, v[sl].im.y));
            } else {
#pragma unroll
                for (int sl = 0; sl < PTS; ++sl)
                    dst[xab_index(N, row, g + mwfft::final_off<N, PTS>(sl))] = make_float4(v[sl].re.x, v[sl].re.y, v[sl].im.x, v[sl].im.y);
            }
        } else {
            float2* dst = a.XC + (size_t)xt * N * N;
            if constexpr (T % (4 * W) == 0) {
                // result column = g + (multiple of T): slab and in-row position of g, then compile-time slab steps
                const unsigned base = (unsigned)(g / (4 * W)) * N * (4 * W) + xc4_pos(N, g % (4 * W));
                const unsigned bA = base + (unsigned)rA * (4 * W), bB = base + (unsigned)rB * (4 * W);
#pragma unroll
                for (int sl = 0; sl < PTS; ++sl) {
                    const unsigned off = (unsigned)(mwfft::final_off<N, PTS>(sl) / (4 * W)) * (N * 4 * W);
                    MW_XST(dst + (bA + off), make_float2(v[sl].re.x, v[sl].im.x));
                    MW_XST(dst + (bB + off), make_float2(v[sl].re.y, v[sl].im.y));
                }
            } else {
#pragma unroll
                for (int sl = 0; sl < PTS; ++sl) {
                    const int idx = g + mwfft::final_off<N, PTS>(sl);
                    dst[xc4_index(N, rA, idx)] = make_float2(v[sl].re.x, v[sl].im.x);
                    dst[xc4_index(N, rB, idx)] = make_float2(v[sl].re.y, v[sl].im.y);
                }
            }
        }
    }
    MW_RSTAMP(3);
}

// =============================================================================================
// pass 2: column FFT + extraction (+ Jacobian whitecap)
// =============================================================================================
#ifndef MW_NSTAGE
#define MW_NSTAGE 4
#endif
struct ColArgs {
    const float4* XAB;  // [tiles][N/8][N][8]
    const float2* XC;   // [tiles][N/32][N][32] Hermitian-symmetrised C' after the row transform (xc4_index)
    const float4* twimg;  // twiddle-table image
    float* height;      // [tiles][N*N]     or NULL
    float2* disp;       // [tiles][N*N]     or NULL   (hds)
    float* normal;      // [tiles][N*N][3]  or NULL
    float* whitecap;    // [tiles][N*N]     or NULL
    float* jacobian;    // [tiles][N*N]     or NULL
    long long* dbg;     // developer phase-timing buffer (NULL in production): 8 clock64 stamps per CTA
    int dbg_flags;      // developer experiments (tools/phase_timing.py)
    int tile0;          // first tile of this launch (outputs are indexed by tile0 + blockIdx.y, X by blockIdx.y)
    int pdl;            // programmatic dependent launch: 1 = release the successor at CTA start, 2 = before the extraction
    int ab_blocks;      // blockIdx.x <  ab_blocks : (A,B) slab of 8 columns  (0 if no A/B output is wanted)
                        // blockIdx.x >= ab_blocks : C slab of 32 columns
    float2* seam;       // k_cols_seam only: [tiles][N / W][N] (dx, dz) / 2 of every slab's first column
    unsigned* seam_flags;  //               [tiles][N / W] "published" flags (cleared by pass 1)
    unsigned* seam_timeouts;  //            one counter: hand-overs that gave up waiting (never expected; mw_ocean_sync reports it)
    // TMA descriptors of the whitecap / hds / normal planes as 2-D float tensors [tiles * N rows][N * {1,2,3} floats]
    // (only read by the kernel variants that store through TMA, see cols_tma_store)
    alignas(64) CUtensorMap tm_white, tm_disp, tm_normal;
};
// Build option (-DMW_COLS_TMA_STORE=1): the (A,B) slab's outputs leave as TMA tensor stores, one 256-row x 8-column box per
// output plane per chunk, from a shared-memory staging area instead of ~45 per-thread 4..16-byte stores per thread.
#ifndef MW_COLS_TMA_STORE
#define MW_COLS_TMA_STORE 0
#endif
__host__ __device__ constexpr bool cols_tma_store(int N, int outs) { return MW_COLS_TMA_STORE && N == 1024 && outs == 7 && fft_pts(N) == 16 && slab_w(N) == 8; }
__host__ __device__ constexpr size_t cols_stage_bytes(int N, int outs, int threads)
{
    return cols_tma_store(N, outs) ? (size_t)256 * slab_w(N) * (4 + 8 + 12) + 128
                                   : (size_t)((threads + 31) / 32) * 96 * (N >= 256 ? MW_NSTAGE : 1) * sizeof(float);
}

__device__ __forceinline__ float sqrt_approx(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));  // MUFU.SQRT, ~1 ulp: the whitecap tolerance is 1e-5 of the Jacobian scale
    return r;
}
__device__ __forceinline__ float rsqrt_ftz(float x)  // argument >= 1 here: no denormal fix-up wanted
{
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// 9 packed lines per CTA, two kinds of CTA in one launch:
//   (A,B) CTA: the (A,B) pairs of W = 8 columns + the halo column b0 + W (so that hds[index + 1] of
//              FFTMesh.cs:266 is on chip)  -> hds, normal, Jacobian, whitecap
//   C CTA    : the Hermitian-symmetrised C field of 4 W = 32 columns, FOUR columns per packed line (two real columns per
//              complex transform; 8 lines busy)  -> height
//
// Thread <-> data: thread tid < 8T owns line c = tid & 7 and residue g = tid >> 3 of the line (T = N/16
// residues), so a warp is 4 consecutive rows x 8 columns.  With that mapping
//   * the first-stage inputs {row g + T k} x {8 columns} are loaded from global memory straight into
//     registers as 128-byte row segments of one contiguous slab (no staging copy, no transposition pass);
//   * after the last stage a warp holds 4 consecutive rows x 8 columns of finished values in registers,
//     which is the shape of a store of whole 32-byte sectors: hds 64 B, normal 96 B, whitecap 32 B per row.
//     Normals, hds and the whitecap are computed from registers; only the (dx, dz) pairs go through shared
//     memory once more, for the forward differences of the Jacobian (FFTMesh.cs:260-267), and the normals
//     take a per-warp 384-byte staging hop to leave as 16-byte stores.
// Threads tid >= 8T (one more group of T) run the halo line.
#ifndef MW_COLS_MAXREG
#define MW_COLS_MAXREG 128
#endif
// OUTS: which (A,B)-slab outputs exist, as a compile-time set (bit 0 hds, 1 normal, 2 whitecap, 3 Jacobian) so that the
// extraction is straight-line code; OUTS = -1 decides per pointer at run time (any other combination).
__host__ __device__ constexpr int nstage_slots(int N) { return N >= 256 ? MW_NSTAGE : 1; }
// Register cap per resolution = what decides the CTAs per SM of this kernel ((W + 1) N / 16 threads per CTA):
// N = 1024: 576 threads, one CTA per SM either way; N = 512: 288 threads, 112 registers let two CTAs share an SM
// (128 leaves one); N = 256: 144 threads, 96 registers -> four CTAs (the shared-memory limit) instead of three.
// (measured, profiles/r02_occ_sweep.jsonl: 112 / 96 registers here + 3 / 3 / 2 pass-1 CTAs per SM at N = 256 / 512 / 2048
//  take 7-8 % off the frame at those resolutions: 256^2 x 256 301 -> 278 us, 512^2 x 64 431 -> 403 us, one 2048^2 tile 154 -> 142 us)
#ifndef MW_COLS_MAXREG_512
#define MW_COLS_MAXREG_512 112
#endif
#ifndef MW_COLS_MAXREG_256
#define MW_COLS_MAXREG_256 96
#endif
__host__ __device__ constexpr int cols_maxreg(int N)
{
    return fft_pts(N) == 32 ? 168
         : N == 1024 ? (MW_SLABW_1024 == 8 ? MW_COLS_MAXREG : 96)
         : N == 512 ? MW_COLS_MAXREG_512
         : N == 256 ? MW_COLS_MAXREG_256
         : 128;
}
template <int N, int MINB, int OUTS>
__global__ void __launch_bounds__((slab_w(N) + 1) * (N / fft_pts(N)), MINB)
__maxnreg__(cols_maxreg(N)) k_cols_extract(const __grid_constant__ ColArgs a)
{
    constexpr int NS = nstage_slots(N);
    constexpr int PTS = fft_pts(N);
    using P = Plan<N, PTS>;
    constexpr int T = P::T;
    constexpr int W = slab_w(N);
    constexpr int LOGW = mwfft::ilog2(W);
    constexpr int LP = mwfft::line_pitch(N, W);
    constexpr bool LINEAR = (T % 16 == 0);
    extern __shared__ float4 smem4[];
    float4* tw2 = smem4;
    float2* tw3 = reinterpret_cast<float2*>(smem4 + P::TW2_F4);
    float4* lines = smem4 + P::TW_BYTES / 16;                              // [W + 1][LP]
    float* nstage = reinterpret_cast<float*>(lines + (W + 1) * LP);         // [warps][NS][96] normals of 32/W rows x W columns, NS slots

    const int tile = a.tile0 + blockIdx.y;
    const int xt = blockIdx.y;
    const int tid = threadIdx.x;
    const bool is_halo = tid >= W * T;
    const int c = is_halo ? W : (tid & (W - 1));
    const int g = is_halo ? tid - W * T : (tid >> LOGW);
    const size_t plane = (size_t)N * N;
    const size_t obase = (size_t)tile * plane;
    float4* line = lines + c * LP;
    auto cta_sync = [] { __syncthreads(); };

#if MW_DEVHOOKS
#define MW_STAMP(i) do { if (a.dbg && tid == 0) a.dbg[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 + (i)] = clock64(); } while (0)
#else
#define MW_STAMP(i) do { } while (0)
#endif
    MW_STAMP(0);

    const uint64_t pol = evict_first_policy();  // outputs are written once and not read again by this engine
    mwfft::cpk v[PTS];
    const bool is_ab = (int)blockIdx.x < a.ab_blocks;
    const bool want_white = OUTS < 0 ? (a.whitecap != nullptr || a.jacobian != nullptr) : (OUTS & 12) != 0;
    const int b0 = is_ab ? blockIdx.x * W : ((int)blockIdx.x - a.ab_blocks) * (4 * W);
    // a group without a live line (halo group of a C slab, of the last slab, or when no whitecap is wanted)
    // transforms zeros: same instruction stream for every thread, no divergent barrier
    const bool active = is_ab ? (!is_halo || (want_white && b0 + W < N)) : !is_halo;
    if (a.pdl == 1) pdl_trigger();
    if (MW_DBG(a, 8) && !is_ab) return;
    if (T >= 32 && !active) {  // whole warps with nothing to transform: help with the tables, then leave
        mwfft::load_twiddle_image<N, (W + 1) * T, PTS>(smem4, a.twimg);
        return;                // (exited threads are not waited for by barriers)
    }

    pdl_wait();  // the intermediate is pass 1's (the predecessor in this stream)
    // ---- first-stage inputs straight from global memory: line position p = g + T k holds intermediate row
    //      (p + N/2) mod N = g + T ((k + 8) mod 16)  (the (-1)^a of sigma); slab-major layout => contiguous ----
#ifndef MW_COLS_TMA
#define MW_COLS_TMA 1
#endif
    __shared__ uint64_t slab_bar;
    if (is_ab && MW_COLS_TMA) {
        // The (A,B) slab is one contiguous block of N * W * 16 bytes (slab-major layout): ONE thread asks the copy engine for
        // it (cp.async.bulk into the line buffers, which are free until the first stage writes them; SASS: UBLKCP + SYNCS),
        // everybody waits on the mbarrier and picks its first-stage inputs out of shared memory.  The per-thread alternative
        // (-DMW_COLS_TMA=0: 16 x 576 LDG.128) keeps only ~24 KB in flight per SM -- the request queue sets the pace
        // (`lg_throttle`), 3.6 TB/s aggregate against 5.8 TB/s for the bulk copy.  Measured on B200 (profiles/
        // r02_switch_sweep.jsonl): 16 x 1024^2 frame 406.5 -> 403.6 us with the L2-resident intermediate, 445 -> 437 us with
        // the intermediate in HBM (one launch for all tiles), 256^2 x 256 301 -> 295 us.
        float4* raw = lines;  // [N][W] float4, aliases the line buffers
        if (tid == 0) mbar_init(&slab_bar, 1);
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&slab_bar, (unsigned)(N * W * sizeof(float4)));
            bulk_g2s(raw, a.XAB + (size_t)xt * xab_tile_elems(N) + (size_t)blockIdx.x * N * W, (unsigned)(N * W * sizeof(float4)), &slab_bar);
        }
        if (is_halo) {
            // the halo column is entry 0 of every row of the NEXT slab: 16 bytes out of each 128-byte row, by plain loads
            const float4* src = a.XAB + (size_t)xt * xab_tile_elems(N) + ((size_t)blockIdx.x * N + g) * W + (size_t)N * W;
#pragma unroll
            for (int k = 0; k < PTS; ++k) {
                float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
                if (active) e = ldg_fresh4(src + (size_t)(T * ((k + PTS / 2) & (PTS - 1))) * W);
                v[k].re = make_float2(e.x, e.y);
                v[k].im = make_float2(e.z, e.w);
            }
        }
        mwfft::load_twiddle_image<N, (W + 1) * T, PTS>(smem4, a.twimg);
        mbar_wait(&slab_bar, 0);
        if (!is_halo) {
            const float4* src = raw + g * W + c;
#pragma unroll
            for (int k = 0; k < PTS; ++k) {
                const float4 e = src[(T * ((k + PTS / 2) & (PTS - 1))) * W];
                v[k].re = make_float2(e.x, e.y);
                v[k].im = make_float2(e.z, e.w);
            }
        }
        __syncthreads();  // everyone has its inputs before the first stage overwrites the raw slab with the lines
    } else if (is_ab) {
        // (the halo group, c == W, reads entry 0 of the same row of the next slab, N * W elements further on)
        const float4* src = a.XAB + (size_t)xt * xab_tile_elems(N) + ((size_t)blockIdx.x * N + g) * W +
                            (is_halo ? (size_t)N * W : (size_t)c);
#pragma unroll
        for (int k = 0; k < PTS; ++k) {
            float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
            if (active && !MW_DBG(a, 4)) e = ldg_fresh4(src + (size_t)(T * ((k + PTS / 2) & (PTS - 1))) * W);
            v[k].re = make_float2(e.x, e.y);
            v[k].im = make_float2(e.z, e.w);
        }
    } else {
        // C slab: 4 W columns, line c = columns b0 + 4c .. 4c + 3 as two complex lines z = X[4c] + i X[4c+1] (lane x) and
        // z = X[4c+2] + i X[4c+3] (lane y): the columns are Hermitian along n (pass 1 stored C' symmetrised), so each
        // transform is real and Re / Im of the result are the heights of the two columns.  A row of the slab is
        // [(X0, X1) of the W lines | (X2, X3) of the W lines], 16 bytes each: two fully coalesced loads.
        const float4* src = reinterpret_cast<const float4*>(a.XC + (size_t)xt * plane + ((size_t)(b0 / (4 * W)) * N + g) * (4 * W)) +
                            (active ? c : 0);
#pragma unroll
        for (int k = 0; k < PTS; ++k) {
            float4 e0 = make_float4(0.f, 0.f, 0.f, 0.f), e1 = e0;
            if (active) {
                const float4* p = src + (size_t)(T * ((k + PTS / 2) & (PTS - 1))) * (2 * W);
                e0 = ldg_fresh4(p);
                e1 = ldg_fresh4(p + W);
            }
            v[k].re = make_float2(e0.x - e0.w, e1.x - e1.w);
            v[k].im = make_float2(e0.y + e0.z, e1.y + e1.z);
        }
    }
    // the twiddle tables are fetched while the slab loads above are in flight
    if (!(is_ab && MW_COLS_TMA)) mwfft::load_twiddle_image<N, (W + 1) * T, PTS>(smem4, a.twimg);
    MW_STAMP(1);
    if (!MW_DBG(a, 2)) mwfft::fft_line_inreg<N, +1, PTS>(v, line, g, tw2, tw3, cta_sync);  // one instance for both kinds
    MW_STAMP(2);
    if (a.pdl == 2) pdl_trigger();

    if (!is_ab) {
        if (active) {
            // height (FFTMesh.cs:219) of four consecutive columns: (Re, Im) of lane x, (Re, Im) of lane y
            float* dst = a.height + obase + (size_t)g * N + b0 + 4 * c;
#pragma unroll
            for (int s = 0; s < PTS; ++s)
                st_once(reinterpret_cast<float4*>(dst + (size_t)mwfft::final_off<N, PTS>(s) * N),
                        make_float4(v[s].re.x, v[s].im.x, v[s].re.y, v[s].im.y), pol);
        }
        return;
    }

    // ------------------------------------------------------------------ (A, B) slab
    // finished transform: lane x = dx + i dz, lane y = sx + i sz.  (dx, dz) / 2 go to shared memory (in place of the
    // line, 8 bytes per row) for the neighbours' forward differences  0.5 (hds[idx] - hds[idx + N]),
    // 0.5 (hds[idx] - hds[idx + 1])  (FFTMesh.cs:260-267).  The "no neighbour => derivative 0" edges (:260, :264)
    // are data, not branches: row N (one past the end) holds a copy of row N - 1, and in the last slab the halo line
    // holds a copy of column N - 1.
    const bool has_disp = OUTS < 0 ? a.disp != nullptr : (OUTS & 1) != 0;
    const bool has_normal = OUTS < 0 ? a.normal != nullptr : (OUTS & 2) != 0;
    const bool has_white = OUTS < 0 ? a.whitecap != nullptr : (OUTS & 4) != 0;
    const bool has_jac = OUTS < 0 ? a.jacobian != nullptr : (OUTS & 8) != 0;
    const bool need_d = has_white || has_jac;
    const bool last_slab = b0 + W >= N;  // CTA-uniform
    float2* D = reinterpret_cast<float2*>(line);
    const int pg = pad_idx(g);
    auto dpos = [&](int s) { return LINEAR ? pg + mwfft::pad_step(mwfft::final_off<N, PTS>(s)) : pad_idx(g + mwfft::final_off<N, PTS>(s)); };
    if (need_d) {
        if (!(is_halo && last_slab)) {
#pragma unroll
            for (int s = 0; s < PTS; ++s) D[dpos(s)] = make_float2(0.5f * v[s].re.x, 0.5f * v[s].im.x);
            if (g == T - 1) D[pad_idx(N)] = make_float2(0.5f * v[PTS - 1].re.x, 0.5f * v[PTS - 1].im.x);  // the last slot is row g + N - T
        }
        if (last_slab && !is_halo && c == W - 1) {
            float2* Dh = reinterpret_cast<float2*>(line + LP);
#pragma unroll
            for (int s = 0; s < PTS; ++s) Dh[dpos(s)] = make_float2(0.5f * v[s].re.x, 0.5f * v[s].im.x);
        }
    }
    __syncthreads();
    MW_STAMP(3);
    if (T >= 32 && is_halo) return;  // whole warps: done (for T < 32 they share a warp with owners and stay for the warp syncs)
    if constexpr (cols_tma_store(N, OUTS)) {
        // Result slots [ch B, ch B + B) of every thread are the rows g + T b of ONE 256-row chunk of the slab (final_off):
        // the CTA stages the chunk's whitecap / hds / normal values in the layout of the output rows and one thread hands
        // the three boxes to the copy engine; the next chunk is computed while the engine reads the staging area.
        constexpr int FB = mwfft::Final<N, PTS>::B, CHUNK = FB * T;
        static_assert(CHUNK == 256 && mwfft::Final<N, PTS>::S == 256, "one chunk = 256 consecutive rows");
        float* stg = nstage + ((128u - (smem_u32(nstage) & 127u)) & 127u) / 4;  // tensor copies want 128-byte aligned shared memory
        float* st_w = stg;                                                // [256][W]     floats
        float2* st_d = reinterpret_cast<float2*>(stg + CHUNK * W);        // [256][W]     float2
        float* st_n = stg + CHUNK * W * 3;                                // [256][3 W]   floats
        const int dn = pad_idx(g + 1) - pg;
        const float2* De = reinterpret_cast<const float2*>(line + LP);
        const int row0 = (a.tile0 + (int)blockIdx.y) * N;
#pragma unroll
        for (int ch = 0; ch < PTS / FB; ++ch) {
            float o_nx[FB], o_ny[FB], o_nz[FB], o_wc[FB];
#pragma unroll
            for (int b = 0; b < FB; ++b) {
                const int s = ch * FB + b;
                const float dx = v[s].re.x, sx = v[s].re.y, dz = v[s].im.x, sz = v[s].im.y;
                const float r2 = fmaf(sx, sx, sz * sz);
                const float inv = rsqrt_ftz(r2 + 1.0f);
                o_nx[b] = sx * inv; o_ny[b] = inv; o_nz[b] = sz * inv;
                const int pa = dpos(s);
                const float2 nbs = D[pa + dn], nbe = De[pa];
                const float hx = 0.5f * dx, hz = 0.5f * dz;
                const float ddx_x = hx - nbs.x, ddx_y = hz - nbs.y, ddy_x = hx - nbe.x, ddy_y = hz - nbe.y;
                const float jac = fmaf(1.0f + ddx_x, 1.0f + ddy_y, -(ddx_y * ddy_x));
                const float noise = 0.3f * inv * sqrt_approx(r2);
                float turb = fminf(fmaxf(1.0f - jac + noise, 0.0f), 1.0f);
                o_wc[b] = turb * turb * fmaf(-2.0f, turb, 3.0f);
            }
            if (ch > 0) {  // the engine must have read the previous chunk out of the staging area
                if (tid == 0) bulk_wait_read();
                __syncthreads();
            }
#pragma unroll
            for (int b = 0; b < FB; ++b) {
                const int s = ch * FB + b;
                const int lr = g + T * b;  // row within the chunk
                st_w[lr * W + c] = o_wc[b];
                st_d[lr * W + c] = make_float2(v[s].re.x, v[s].im.x);
                st_n[lr * 3 * W + 3 * c + 0] = o_nx[b];
                st_n[lr * 3 * W + 3 * c + 1] = o_ny[b];
                st_n[lr * 3 * W + 3 * c + 2] = o_nz[b];
            }
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
                const int y = row0 + mwfft::final_off<N, PTS>(ch * FB);
                tma_store_2d(&a.tm_white, st_w, b0, y);
                tma_store_2d(&a.tm_disp, st_d, 2 * b0, y);
                tma_store_2d(&a.tm_normal, st_n, 3 * b0, y);
                bulk_commit();
            }
        }
        if (tid == 0) bulk_wait_read();  // the staging area lives as long as this thread does
        return;
    }
    {
        const bool own = !is_halo && !MW_DBG(a, 1);  // halo threads run the same code with every memory access predicated off
        if (MW_DBG(a, 16)) return;
        const int dn = pad_idx(g + 1) - pg;  // padded distance to the next row (1 or 2)
        // east neighbour's line ((small N) halo threads run along with every access predicated off: keep their reads in bounds)
        const float2* De = reinterpret_cast<const float2*>(is_halo ? line : line + LP);
        const int lane = tid & 31;
        // per-thread output bases; a slot then adds a compile-time multiple of N
        const size_t o0 = obase + (size_t)g * N + b0 + c;
        float2* p_disp = has_disp ? a.disp + o0 : nullptr;
        float* p_white = has_white ? a.whitecap + o0 : nullptr;
        float* p_jac = has_jac ? a.jacobian + o0 : nullptr;
        // normals leave as 16-byte stores: lane l = (row l >> LOGW, column l & (W-1)) puts (nx, ny, nz) at floats
        // [3l, 3l+3) of the warp's 96-float block of a slot; lanes 0..23 then store the block as 24 float4
        // (3W/4 per row = the slab's 12W contiguous bytes of an output row).  NS slots are staged per warp sync pair.
        constexpr int QR = 3 * W / 4;
        const int rr = lane / QR, qq = lane - QR * rr;  // row of the warp, float4 within the row
        float* wst = nstage + (tid >> 5) * (96 * NS);
        // column b0 of output row (first row of the warp + rr), as float4 index into the normal plane
        float4* p_nrm = has_normal ? reinterpret_cast<float4*>(a.normal) + (3 * (obase + (size_t)(g - (lane >> LOGW) + rr) * N + b0)) / 4 + qq
                                   : nullptr;
        const bool nrm_lane = lane < 24 && !MW_DBG(a, 1) && (T >= 32 || tid - lane + W * rr < W * T);  // (small N: rows of halo lanes do not exist)
#pragma unroll
        for (int s0 = 0; s0 < PTS; s0 += NS) {
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                const int s = s0 + j;
                const int off = mwfft::final_off<N, PTS>(s);       // output row = g + off
                const float dx = v[s].re.x, sx = v[s].re.y, dz = v[s].im.x, sz = v[s].im.y;
                // nor = normalize(up - n) = (sx, 1, sz) / |.|   (FFTMesh.cs:212, 218)
                const float r2 = fmaf(sx, sx, sz * sz);
                const float inv = rsqrt_ftz(r2 + 1.0f);
                const float nx = sx * inv, nz = sz * inv;
                if (has_normal) {
                    wst[96 * j + 3 * lane + 0] = nx;
                    wst[96 * j + 3 * lane + 1] = inv;
                    wst[96 * j + 3 * lane + 2] = nz;
                }
                if (has_disp && own) st_once(p_disp + (size_t)off * N, make_float2(dx, dz), pol);  // hds (FFTMesh.cs:247)
                if (need_d) {
                    const int pa = dpos(s);
                    const float2 nbs = D[LINEAR ? pa + dn : pad_idx(g + off + 1)];  // hds[index + resolution] / 2  (:260-263)
                    const float2 nbe = De[pa];                                       // hds[index + 1] / 2           (:264-267)
                    const float hx = 0.5f * dx, hz = 0.5f * dz;
                    const float ddx_x = hx - nbs.x, ddx_y = hz - nbs.y, ddy_x = hx - nbe.x, ddy_y = hz - nbe.y;
                    const float jac = fmaf(1.0f + ddx_x, 1.0f + ddy_y, -(ddx_y * ddy_x));  // :268
                    if (has_jac && own) st_once(p_jac + (size_t)off * N, jac, pol);
                    if (has_white && own) {
                        // noise = |(|n.x|, |n.z|) * 0.3| = 0.3 sqrt(sx^2 + sz^2) / |(sx, 1, sz)|   (:269-270)
                        const float noise = 0.3f * inv * sqrt_approx(r2);
                        float turb = fmaxf(1.0f - jac + noise, 0.0f);                           // :270
                        turb = fminf(turb, 1.0f);                                               // SmoothStep clamps
                        st_once(p_white + (size_t)off * N, turb * turb * fmaf(-2.0f, turb, 3.0f), pol);  // :273
                    }
                }
            }
            if (has_normal) {
                __syncwarp();
                if (nrm_lane) {
#pragma unroll
                    for (int j = 0; j < NS; ++j) {
                        const float4 q = *reinterpret_cast<const float4*>(wst + 96 * j + 4 * lane);
                        st_once(p_nrm + (3 * (size_t)mwfft::final_off<N, PTS>(s0 + j) * N) / 4, q, pol);
                    }
                }
                __syncwarp();
            }
        }
    }
    MW_STAMP(4);
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
        // rest position, FFTMesh.cs:107-112: (i - N/2) * uw (+ uw / 2 on even grids)
        const float off = (N % 2 == 0) ? __fdiv_rn(unit_width, 2.0f) : 0.0f;
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
