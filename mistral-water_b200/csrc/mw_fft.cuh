// mw_fft.cuh -- the in-CTA Stockham FFT engine (sm_100a), packed-pair edition.
//
// What it computes is what log2(N) blits of the reference's radix-2 Stockham fragment shader
// compute (Shaders/FFT/Stockham.shader:31-57, scheduled by Scripts/OceanRenderer.cs:229-262):
// an un-normalised, natural-order-in / natural-order-out DFT of one line.  How it computes it is
// different:
//   * a line lives in shared memory and is transformed in 2-3 autosort stages of radix 16 (radix
//     16 x 16 x N/256), each stage done entirely in registers, with ONE shared-memory exchange
//     between stages -- instead of log2(N) round trips through memory;
//   * every "line" is a PAIR of lines transformed together: Blackwell's packed fp32 instructions
//     (FADD2 / FMUL2 / FFMA2, PTX add/mul/fma.rn.f32x2) operate on two fp32 lanes per register pair
//     with scalar / immediate twiddles broadcast to both lanes.  The FP32 lane throughput is the
//     same as scalar code, but the issue slots are halved, which is what these kernels are short of.
//     The two lanes carry two fields of the same grid line (same twiddles), never the re/im of one
//     number.
//
// Thread layout: a packed line of N points is served by a group of T = N / 16 threads; thread g
// always reads the 16 elements {g + T*c : c < 16} (the Stockham read pattern "stride N/R" has this
// form for every radix R when each thread owns 16/R butterflies), runs 16/R radix-R butterflies in
// registers, and scatters the results to their autosorted positions.
#pragma once
#include "mw_common.cuh"

namespace mwfft {

constexpr int PTS = 16;  // packed points per thread

// two complex numbers side by side: (re.x + i im.x) and (re.y + i im.y)
struct cpk {
    float2 re, im;
};

__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ cpk padd(cpk a, cpk b) { return {__fadd2_rn(a.re, b.re), __fadd2_rn(a.im, b.im)}; }
__device__ __forceinline__ cpk psub(cpk a, cpk b) { return {__fadd2_rn(a.re, neg2(b.re)), __fadd2_rn(a.im, neg2(b.im))}; }
// multiply both numbers by the same scalar twiddle (c + i s)
__device__ __forceinline__ cpk pmul(cpk d, float c, float s)
{
    const float2 cc = make_float2(c, c), ss = make_float2(s, s);
    return {__ffma2_rn(d.re, cc, __fmul2_rn(d.im, neg2(ss))), __ffma2_rn(d.re, ss, __fmul2_rn(d.im, cc))};
}

// cos(2 pi q / 32), sin(2 pi q / 32) for q in [0, 16): compile-time constants once unrolled.
__host__ __device__ constexpr float cos32(int q)
{
    return q == 0 ? 1.0f
         : q == 1 ? 0.98078528040323043f
         : q == 2 ? 0.92387953251128674f
         : q == 3 ? 0.83146961230254524f
         : q == 4 ? 0.70710678118654752f
         : q == 5 ? 0.55557023301960218f
         : q == 6 ? 0.38268343236508978f
         : q == 7 ? 0.19509032201612825f
         : q == 8 ? 0.0f
         : -cos32(16 - q);
}
__host__ __device__ constexpr float sin32(int q) { return q <= 8 ? cos32(8 - q) : cos32(q - 8); }

__host__ __device__ constexpr int ilog2(int x) { return x <= 1 ? 0 : 1 + ilog2(x >> 1); }
__host__ __device__ constexpr int bitrev(int i, int bits)
{
    int r = 0;
    for (int b = 0; b < bits; ++b) r |= ((i >> b) & 1) << (bits - 1 - b);
    return r;
}

// d * W_32^Q with W_32 = exp(SIGN * 2 pi i / 32), Q in [0, 16)
template <int SIGN, int Q>
__device__ __forceinline__ cpk mul_w32(cpk d)
{
    if constexpr (Q == 0) {
        return d;
    } else if constexpr (Q == 8) {  // * (SIGN i): folded into the consumers' operand signs by ptxas
        if constexpr (SIGN > 0) return {neg2(d.im), d.re};
        else return {d.im, neg2(d.re)};
    } else if constexpr (Q == 4) {  // * (1 + SIGN i) / sqrt2
        const float2 h = make_float2(0.70710678118654752f, 0.70710678118654752f);
        if constexpr (SIGN > 0) return {__fmul2_rn(__fadd2_rn(d.re, neg2(d.im)), h), __fmul2_rn(__fadd2_rn(d.re, d.im), h)};
        else return {__fmul2_rn(__fadd2_rn(d.re, d.im), h), __fmul2_rn(__fadd2_rn(d.im, neg2(d.re)), h)};
    } else if constexpr (Q == 12) {  // * (-1 + SIGN i) / sqrt2
        const float2 h = make_float2(0.70710678118654752f, 0.70710678118654752f);
        const float2 nh = make_float2(-0.70710678118654752f, -0.70710678118654752f);
        if constexpr (SIGN > 0) return {__fmul2_rn(__fadd2_rn(d.re, d.im), nh), __fmul2_rn(__fadd2_rn(d.re, neg2(d.im)), h)};
        else return {__fmul2_rn(__fadd2_rn(d.im, neg2(d.re)), h), __fmul2_rn(__fadd2_rn(d.re, d.im), nh)};
    } else {
        return pmul(d, cos32(Q), SIGN > 0 ? sin32(Q) : -sin32(Q));
    }
}

// One decimation-in-frequency level over the R registers v[BASE + STRIDE * i], i < R.
template <int SIGN, int LEN, int BASE, int STRIDE, int BLK, int K>
__device__ __forceinline__ void dif_pair(cpk (&v)[PTS])
{
    constexpr int i0 = BASE + STRIDE * (BLK + K);
    constexpr int i1 = BASE + STRIDE * (BLK + K + LEN / 2);
    const cpk a = v[i0], b = v[i1];
    v[i0] = padd(a, b);
    v[i1] = mul_w32<SIGN, K * (32 / LEN)>(psub(a, b));
}
template <int SIGN, int R, int LEN, int BASE, int STRIDE, int BLK, int K>
struct DifK {
    static __device__ __forceinline__ void run(cpk (&v)[PTS])
    {
        dif_pair<SIGN, LEN, BASE, STRIDE, BLK, K>(v);
        if constexpr (K + 1 < LEN / 2) DifK<SIGN, R, LEN, BASE, STRIDE, BLK, K + 1>::run(v);
    }
};
template <int SIGN, int R, int LEN, int BASE, int STRIDE, int BLK>
struct DifBlk {
    static __device__ __forceinline__ void run(cpk (&v)[PTS])
    {
        DifK<SIGN, R, LEN, BASE, STRIDE, BLK, 0>::run(v);
        if constexpr (BLK + LEN < R) DifBlk<SIGN, R, LEN, BASE, STRIDE, BLK + LEN>::run(v);
    }
};
template <int SIGN, int R, int LEN, int BASE, int STRIDE>
struct DifLevel {
    static __device__ __forceinline__ void run(cpk (&v)[PTS])
    {
        DifBlk<SIGN, R, LEN, BASE, STRIDE, 0>::run(v);
        if constexpr (LEN > 2) DifLevel<SIGN, R, LEN / 2, BASE, STRIDE>::run(v);
    }
};
// In-register radix-R DFT of v[BASE + STRIDE * i]; output X[bitrev(i)] is left in slot i.
template <int SIGN, int R, int BASE, int STRIDE>
__device__ __forceinline__ void dft_regs(cpk (&v)[PTS])
{
    if constexpr (R >= 2) DifLevel<SIGN, R, R, BASE, STRIDE>::run(v);
}
template <int SIGN, int R, int B, int BASE = 0>
__device__ __forceinline__ void dft_all(cpk (&v)[PTS])
{
    dft_regs<SIGN, R, BASE, B>(v);
    if constexpr (BASE + 1 < B) dft_all<SIGN, R, B, BASE + 1>(v);
}

// ---------------------------------------------------------------------------------------------
// shared-memory layout of a packed line: two planes (re pairs, im pairs) of float2, 64-bit accesses
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr int pad_idx(int i) { return i + (i >> 4); }
// plane pitch in float2: room for pad_idx(N-1); a line is two planes, so the line stride is 2 * pitch,
// which must be == 16/W (mod 16) for the transposing accesses of a W-column slab (lane = column + W * row)
// to touch 16 distinct 8-byte bank groups per half warp  =>  pitch == 8/W (mod 8)
__host__ __device__ constexpr int plane_pitch(int n, int w) { return ((n + n / 16 + 15) / 16) * 16 + 8 / w; }

template <int N>
struct Plan {
    static_assert(N >= 32 && N <= 4096 && (N & (N - 1)) == 0, "N must be a power of two in [32, 4096]");
    static constexpr int T = N / PTS;                           // threads per packed line
    static constexpr int R1 = 16;                               // first radix
    static constexpr int R2 = N / 16 < 16 ? N / 16 : 16;        // second radix
    static constexpr int R3 = N / (16 * R2);                    // third radix (1 => two stages)
};

// all T threads of a line group (T <= 32: within one warp; else whole warps on a named barrier)
template <int T>
__device__ __forceinline__ void group_sync(int line_id)
{
    if constexpr (T <= 32) {
        __syncwarp();
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(line_id + 1), "n"(T) : "memory");
    }
}

// twiddle table lookup: tw[x] = exp(+2 pi i x / N); SIGN < 0 conjugates.
template <int SIGN>
__device__ __forceinline__ float2 tw_get(const float2* __restrict__ tw, int x)
{
    float2 w = __ldg(tw + x);
    if constexpr (SIGN < 0) w.y = -w.y;
    return w;
}

// A Stockham stage (radix R, S = product of the radices before it) on registers that already hold
// {line[g + T*c]}.  Applies the stage twiddles, runs the 16/R butterflies and hands every result to
// `emit(dest_index, value)`.
template <int N, int SIGN, int R, int S, class Emit>
__device__ __forceinline__ void stage_regs(cpk (&v)[PTS], int g, const float2* __restrict__ tw, Emit&& emit)
{
    constexpr int T = N / PTS;
    constexpr int B = PTS / R;        // butterflies per thread
    constexpr int TWS = N / (S * R);  // table stride: W_{S*R} = tw[TWS]
    if constexpr (S > 1) {
#pragma unroll
        for (int b = 0; b < B; ++b) {
            const int k = (g + b * T) & (S - 1);
#pragma unroll
            for (int r = 1; r < R; ++r) {
                const float2 w = tw_get<SIGN>(tw, r * k * TWS);
                v[b + r * B] = pmul(v[b + r * B], w.x, w.y);
            }
        }
    }
    dft_all<SIGN, R, B>(v);
    constexpr int LOGR = ilog2(R);
#pragma unroll
    for (int b = 0; b < B; ++b) {
        const int j = g + b * T;
        const int k = j & (S - 1);
        const int base = (j - k) * R + k;  // (j / S) * S * R + k
#pragma unroll
        for (int i = 0; i < R; ++i) emit(base + bitrev(i, LOGR) * S, v[b + i * B]);
    }
}

template <int N>
__device__ __forceinline__ void load_line_regs(cpk (&v)[PTS], const float2* pre, const float2* pim, int g)
{
    constexpr int T = N / PTS;
#pragma unroll
    for (int c = 0; c < PTS; ++c) {
        const int p = pad_idx(g + T * c);
        v[c].re = pre[p];
        v[c].im = pim[p];
    }
}

// Full transform of one packed line held in shared memory (planes `pre`, `pim`, padded with pad_idx),
// by the T threads of its group (g = index within the group, line_id = barrier id of the group).
// `active` = this group has a real line; inactive groups still arrive at their barrier.  The final
// stage's results go to `emit(index, value)` in natural order; emit may write into the line itself
// (every read of the group is complete before the first emit).
template <int N, int SIGN, class Emit>
__device__ __forceinline__ void fft_line(float2* pre, float2* pim, int g, int line_id, bool active,
                                         const float2* __restrict__ tw, Emit&& emit)
{
    using P = Plan<N>;
    constexpr int T = P::T;
    cpk v[PTS];
    auto to_smem = [&](int idx, cpk val) {
        const int p = pad_idx(idx);
        pre[p] = val.re;
        pim[p] = val.im;
    };
    if (active) load_line_regs<N>(v, pre, pim, g);
    group_sync<T>(line_id);  // everyone has read before anyone overwrites (in-place exchange)
    if (active) stage_regs<N, SIGN, P::R1, 1>(v, g, tw, to_smem);
    group_sync<T>(line_id);
    if (active) load_line_regs<N>(v, pre, pim, g);
    group_sync<T>(line_id);
    if constexpr (P::R3 == 1) {
        if (active) stage_regs<N, SIGN, P::R2, P::R1>(v, g, tw, emit);
    } else {
        if (active) stage_regs<N, SIGN, P::R2, P::R1>(v, g, tw, to_smem);
        group_sync<T>(line_id);
        if (active) load_line_regs<N>(v, pre, pim, g);
        group_sync<T>(line_id);
        if (active) stage_regs<N, SIGN, P::R3, P::R1 * P::R2>(v, g, tw, emit);
    }
}

}  // namespace mwfft
