// mw_fft.cuh -- the in-CTA Stockham FFT engine (sm_100a).
//
// What it computes is what log2(N) blits of the reference's radix-2 Stockham fragment shader
// compute (Shaders/FFT/Stockham.shader:31-57, scheduled by Scripts/OceanRenderer.cs:229-262):
// an un-normalised, natural-order-in / natural-order-out DFT of one line.  How it computes it
// is different: one line lives in shared memory, each thread owns 32 points in registers, and a
// line of N = 32 * M points is done in two autosort stages (radix 32 then radix M; three stages
// for N = 2048) with a single shared-memory exchange between stages, instead of log2(N)
// round trips through memory.
//
// Thread layout: a line is served by a "group" of T = N / 32 threads; thread g of the group
// always reads the 32 elements {g + T*c : c < 32} (the Stockham read pattern "stride N/R" has
// this form for every radix R when each thread owns 32/R butterflies), runs 32/R radix-R
// butterflies in registers, and scatters the results to their autosorted positions.
#pragma once
#include "mw_common.cuh"

namespace mwfft {

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

// (a - b) * W_len^k with W_len = exp(SIGN * 2 pi i / len); q = k * 32 / len in [0, 16).
template <int SIGN, int Q>
__device__ __forceinline__ float2 mul_w32(float2 d)
{
    if constexpr (Q == 0) {
        return d;
    } else if constexpr (Q == 8) {  // * (SIGN i)
        return SIGN > 0 ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
    } else if constexpr (Q == 4) {  // * (1 + SIGN i) / sqrt2
        constexpr float h = 0.70710678118654752f;
        return SIGN > 0 ? make_float2((d.x - d.y) * h, (d.x + d.y) * h) : make_float2((d.x + d.y) * h, (d.y - d.x) * h);
    } else if constexpr (Q == 12) {  // * (-1 + SIGN i) / sqrt2
        constexpr float h = 0.70710678118654752f;
        return SIGN > 0 ? make_float2((-d.x - d.y) * h, (d.x - d.y) * h) : make_float2((d.y - d.x) * h, (-d.x - d.y) * h);
    } else {
        constexpr float c = cos32(Q);
        constexpr float s = SIGN > 0 ? sin32(Q) : -sin32(Q);
        return make_float2(fmaf(d.x, c, -d.y * s), fmaf(d.x, s, d.y * c));
    }
}

// One DIF level: blocks of length LEN over the R registers v[BASE + STRIDE * i], i < R.
template <int SIGN, int R, int LEN, int BASE, int STRIDE, int BLK, int K>
__device__ __forceinline__ void dif_pair(float2 (&v)[32])
{
    constexpr int i0 = BASE + STRIDE * (BLK + K);
    constexpr int i1 = BASE + STRIDE * (BLK + K + LEN / 2);
    const float2 a = v[i0], b = v[i1];
    v[i0] = cadd(a, b);
    v[i1] = mul_w32<SIGN, K * (32 / LEN)>(csub(a, b));
}
template <int SIGN, int R, int LEN, int BASE, int STRIDE, int BLK, int K>
struct DifK {
    static __device__ __forceinline__ void run(float2 (&v)[32])
    {
        dif_pair<SIGN, R, LEN, BASE, STRIDE, BLK, K>(v);
        if constexpr (K + 1 < LEN / 2) DifK<SIGN, R, LEN, BASE, STRIDE, BLK, K + 1>::run(v);
    }
};
template <int SIGN, int R, int LEN, int BASE, int STRIDE, int BLK>
struct DifBlk {
    static __device__ __forceinline__ void run(float2 (&v)[32])
    {
        DifK<SIGN, R, LEN, BASE, STRIDE, BLK, 0>::run(v);
        if constexpr (BLK + LEN < R) DifBlk<SIGN, R, LEN, BASE, STRIDE, BLK + LEN>::run(v);
    }
};
template <int SIGN, int R, int LEN, int BASE, int STRIDE>
struct DifLevel {
    static __device__ __forceinline__ void run(float2 (&v)[32])
    {
        DifBlk<SIGN, R, LEN, BASE, STRIDE, 0>::run(v);
        if constexpr (LEN > 2) DifLevel<SIGN, R, LEN / 2, BASE, STRIDE>::run(v);
    }
};
// In-register radix-R DFT of v[BASE + STRIDE * i]; output X[bitrev(i)] is left in slot i.
template <int SIGN, int R, int BASE, int STRIDE>
__device__ __forceinline__ void dft_regs(float2 (&v)[32])
{
    if constexpr (R >= 2) DifLevel<SIGN, R, R, BASE, STRIDE>::run(v);
}

// ---------------------------------------------------------------------------------------------
// shared-memory line layout
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr int pad_idx(int i) { return i + (i >> 5); }
// line pitch in float2: room for pad_idx(N-1), and == 2 (mod 16) so that lanes walking across lines
// (the transposing loads / stores of the column pass) spread over the banks
__host__ __device__ constexpr int line_pitch(int n) { return ((n + n / 32 + 15) / 16) * 16 + 2; }

template <int N>
struct Plan {
    static_assert(N >= 32 && N <= 4096 && (N & (N - 1)) == 0, "N must be a power of two in [32, 4096]");
    static constexpr int T = N / 32;                       // threads per line
    static constexpr int R1 = 32;                          // first radix
    static constexpr int R2 = N <= 1024 ? N / 32 : 32;     // second radix (1 => no second stage)
    static constexpr int R3 = N <= 1024 ? 1 : N / 1024;    // third radix
    static constexpr int PITCH = line_pitch(N);
};

template <int T>
__device__ __forceinline__ void group_sync()
{
    if constexpr (T <= 32) __syncwarp();
    else __syncthreads();
}

// twiddle table lookup: tw[x] = exp(+2 pi i x / N); SIGN < 0 conjugates.
template <int SIGN>
__device__ __forceinline__ float2 tw_get(const float2* __restrict__ tw, int x)
{
    float2 w = __ldg(tw + x);
    if constexpr (SIGN < 0) w.y = -w.y;
    return w;
}

// A later Stockham stage (radix R, S = product of the radices before it) on registers that
// already hold {line[g + T*c]}.  Applies the stage twiddles, runs the 32/R butterflies and hands
// every result to `emit(dest_index, value)`.
template <int N, int SIGN, int R, int S, class Emit>
__device__ __forceinline__ void stage_regs(float2 (&v)[32], int g, const float2* __restrict__ tw, Emit&& emit)
{
    constexpr int T = N / 32;
    constexpr int B = 32 / R;          // butterflies per thread
    constexpr int TWS = N / (S * R);   // table stride: W_{S*R} = tw[TWS]
#pragma unroll
    for (int b = 0; b < B; ++b) {
        const int j = g + b * T;
        const int k = j & (S - 1);
        if constexpr (S > 1) {
#pragma unroll
            for (int r = 1; r < R; ++r) v[b + r * B] = cmul(v[b + r * B], tw_get<SIGN>(tw, r * k * TWS));
        }
    }
    // butterflies: BASE = b, STRIDE = B (unrolled by hand over b through templates)
    if constexpr (B == 1) {
        dft_regs<SIGN, R, 0, 1>(v);
    } else if constexpr (B == 2) {
        dft_regs<SIGN, R, 0, 2>(v); dft_regs<SIGN, R, 1, 2>(v);
    } else if constexpr (B == 4) {
        dft_regs<SIGN, R, 0, 4>(v); dft_regs<SIGN, R, 1, 4>(v); dft_regs<SIGN, R, 2, 4>(v); dft_regs<SIGN, R, 3, 4>(v);
    } else if constexpr (B == 8) {
        dft_regs<SIGN, R, 0, 8>(v); dft_regs<SIGN, R, 1, 8>(v); dft_regs<SIGN, R, 2, 8>(v); dft_regs<SIGN, R, 3, 8>(v);
        dft_regs<SIGN, R, 4, 8>(v); dft_regs<SIGN, R, 5, 8>(v); dft_regs<SIGN, R, 6, 8>(v); dft_regs<SIGN, R, 7, 8>(v);
    } else if constexpr (B == 16) {
        dft_regs<SIGN, R, 0, 16>(v); dft_regs<SIGN, R, 1, 16>(v); dft_regs<SIGN, R, 2, 16>(v); dft_regs<SIGN, R, 3, 16>(v);
        dft_regs<SIGN, R, 4, 16>(v); dft_regs<SIGN, R, 5, 16>(v); dft_regs<SIGN, R, 6, 16>(v); dft_regs<SIGN, R, 7, 16>(v);
        dft_regs<SIGN, R, 8, 16>(v); dft_regs<SIGN, R, 9, 16>(v); dft_regs<SIGN, R, 10, 16>(v); dft_regs<SIGN, R, 11, 16>(v);
        dft_regs<SIGN, R, 12, 16>(v); dft_regs<SIGN, R, 13, 16>(v); dft_regs<SIGN, R, 14, 16>(v); dft_regs<SIGN, R, 15, 16>(v);
    }
    constexpr int LOGR = ilog2(R);
#pragma unroll
    for (int b = 0; b < B; ++b) {
        const int j = g + b * T;
        const int k = j & (S - 1);
        const int base = (j - k) * R + k;   // (j / S) * S * R + k
#pragma unroll
        for (int i = 0; i < R; ++i) emit(base + bitrev(i, LOGR) * S, v[b + i * B]);
    }
}

template <int N>
__device__ __forceinline__ void load_line_regs(float2 (&v)[32], const float2* line, int g)
{
    constexpr int T = N / 32;
#pragma unroll
    for (int c = 0; c < 32; ++c) v[c] = line[pad_idx(g + T * c)];
}

// Full transform of one line held in shared memory (`line`, padded with pad_idx), by the T
// threads of its group (g = index within the group).  `active` = this group has a real line;
// inactive groups still take part in CTA-wide barriers when T > 32.  The final stage's results go
// to `emit(index, value)` (natural order); emit may write into the line buffer itself (every read
// of the group is complete before the first emit).
template <int N, int SIGN, class Emit>
__device__ __forceinline__ void fft_line(float2* line, int g, bool active, const float2* __restrict__ tw, Emit&& emit)
{
    using P = Plan<N>;
    constexpr int T = P::T;
    float2 v[32];
    auto to_smem = [&](int idx, float2 val) { line[pad_idx(idx)] = val; };

    if (active) load_line_regs<N>(v, line, g);
    group_sync<T>();  // everyone has read before anyone overwrites (in-place exchange)
    if constexpr (P::R2 == 1) {
        if (active) stage_regs<N, SIGN, P::R1, 1>(v, g, tw, emit);
    } else {
        if (active) stage_regs<N, SIGN, P::R1, 1>(v, g, tw, to_smem);
        group_sync<T>();
        if (active) load_line_regs<N>(v, line, g);
        group_sync<T>();  // emit / the next stage may overwrite the line from here on
        if constexpr (P::R3 == 1) {
            if (active) stage_regs<N, SIGN, P::R2, P::R1>(v, g, tw, emit);
        } else {
            if (active) stage_regs<N, SIGN, P::R2, P::R1>(v, g, tw, to_smem);
            group_sync<T>();
            if (active) load_line_regs<N>(v, line, g);
            group_sync<T>();
            if (active) stage_regs<N, SIGN, P::R3, P::R1 * P::R2>(v, g, tw, emit);
        }
    }
}

}  // namespace mwfft
