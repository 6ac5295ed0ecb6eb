// mw_fft.cuh -- the in-CTA Stockham FFT engine (sm_100a), packed-pair edition.
//
// What it computes is what log2(N) blits of the reference's radix-2 Stockham fragment shader
// compute (Shaders/FFT/Stockham.shader:31-57, scheduled by Scripts/OceanRenderer.cs:229-262):
// an un-normalised, natural-order-in / natural-order-out DFT of one line.  How it computes it is
// different:
//   * a line lives in shared memory and is transformed in 2-3 autosort stages of radix 16 (radix
//     16 x 16 x N/256), each stage done entirely in registers, with ONE shared-memory exchange
//     (128-bit accesses) between stages -- instead of log2(N) round trips through memory; the
//     inter-stage twiddles come from small shared-memory tables;
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

constexpr int PTS16 = 16;  // default packed points per thread (the PTS template parameter below: 16 or 32)

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
__device__ __forceinline__ void dif_pair(cpk* v)
{
    constexpr int i0 = BASE + STRIDE * (BLK + K);
    constexpr int i1 = BASE + STRIDE * (BLK + K + LEN / 2);
    const cpk a = v[i0], b = v[i1];
    v[i0] = padd(a, b);
    v[i1] = mul_w32<SIGN, K * (32 / LEN)>(psub(a, b));
}
template <int SIGN, int R, int LEN, int BASE, int STRIDE, int BLK, int K>
struct DifK {
    static __device__ __forceinline__ void run(cpk* v)
    {
        dif_pair<SIGN, LEN, BASE, STRIDE, BLK, K>(v);
        if constexpr (K + 1 < LEN / 2) DifK<SIGN, R, LEN, BASE, STRIDE, BLK, K + 1>::run(v);
    }
};
template <int SIGN, int R, int LEN, int BASE, int STRIDE, int BLK>
struct DifBlk {
    static __device__ __forceinline__ void run(cpk* v)
    {
        DifK<SIGN, R, LEN, BASE, STRIDE, BLK, 0>::run(v);
        if constexpr (BLK + LEN < R) DifBlk<SIGN, R, LEN, BASE, STRIDE, BLK + LEN>::run(v);
    }
};
template <int SIGN, int R, int LEN, int BASE, int STRIDE>
struct DifLevel {
    static __device__ __forceinline__ void run(cpk* v)
    {
        DifBlk<SIGN, R, LEN, BASE, STRIDE, 0>::run(v);
        if constexpr (LEN > 2) DifLevel<SIGN, R, LEN / 2, BASE, STRIDE>::run(v);
    }
};
// In-register radix-R DFT of v[BASE + STRIDE * i]; output X[bitrev(i)] is left in slot i.
template <int SIGN, int R, int BASE, int STRIDE>
__device__ __forceinline__ void dft_regs(cpk* v)
{
    if constexpr (R >= 2) DifLevel<SIGN, R, R, BASE, STRIDE>::run(v);
}
template <int SIGN, int R, int B, int BASE = 0>
__device__ __forceinline__ void dft_all(cpk* v)
{
    dft_regs<SIGN, R, BASE, B>(v);
    if constexpr (BASE + 1 < B) dft_all<SIGN, R, B, BASE + 1>(v);
}

// ---------------------------------------------------------------------------------------------
// shared-memory layout of a packed line: float4 elements (re.x, re.y, im.x, im.y), 128-bit accesses
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr int pad_idx(int i) { return i + (i >> 4); }
// line pitch in float4: room for pad_idx(N-1), and == 8/W (mod 8) so that the transposing accesses of a
// W-column slab (lane = column + W * row) touch 8 distinct 16-byte bank groups per quarter warp
__host__ __device__ constexpr int line_pitch(int n, int w) { return ((n + n / 16 + 7) / 8) * 8 + 8 / w; }

template <int N, int PTS = PTS16>
struct Plan {
    static_assert(N >= 32 && N <= 4096 && (N & (N - 1)) == 0, "N must be a power of two in [32, 4096]");
    static_assert((PTS == 16 || PTS == 32) && N >= 2 * PTS, "16 or 32 packed points per thread");
    static constexpr int T = N / PTS;                           // threads per packed line
    static constexpr int R1 = PTS;                              // first radix: all of a thread's points in one butterfly
    static constexpr int R2 = N / R1 < PTS ? N / R1 : PTS;      // second radix
    static constexpr int R3 = N / (R1 * R2);                    // third radix (1 => two stages, ONE shared-memory exchange)
    // twiddle tables kept in shared memory (filled once per CTA by load_twiddles / load_twiddle_image):
    //   tw2[k][r] = W_{R1 R2}^{r k}, k < R1, r < R2: rows of R2 float2 (+ 16 B pad so that the LDS.128 of
    //               eight lanes with consecutive k are conflict free)
    //   tw3[k]    = W_N^k, k < R1 R2 (third stage: the powers r = 2.. are formed by multiplication)
    static constexpr int TW2_ROW = R2 / 2 + 1;                  // float4 per row
    static constexpr int TW2_F4 = R1 * TW2_ROW;
    static constexpr int TW3_F2 = R3 > 1 ? R1 * R2 : 0;
    static constexpr int TW_BYTES = TW2_F4 * 16 + TW3_F2 * 8;
};

// all T threads of a line group (T <= 32: within one warp; else whole warps on a named barrier)
template <int T>
__device__ __forceinline__ void group_sync(int line_id)
{
    if constexpr (T <= 32) {
        __syncwarp();
    } else {
#ifndef MW_CTA_BARRIERS
        asm volatile("bar.sync %0, %1;" ::"r"(line_id + 1), "n"(T) : "memory");
#else
        (void)line_id;
        __syncthreads();  // every group of the CTA runs the same stage sequence, so a CTA-wide barrier is valid
#endif
    }
}

// Fill the shared twiddle tables from the global table gtw[x] = exp(+2 pi i x / N) (SIGN < 0 conjugates).
// Every thread of the CTA takes part; the caller's next __syncthreads publishes the tables.
template <int N, int SIGN, int PTS = PTS16>
__device__ __forceinline__ void load_twiddles(float4* tw2, float2* tw3, const float2* __restrict__ gtw)
{
    using P = Plan<N, PTS>;
    float2* t2 = reinterpret_cast<float2*>(tw2);
    constexpr int TWS2 = N / (P::R1 * P::R2);
    for (int i = threadIdx.x; i < P::R1 * P::R2; i += blockDim.x) {
        const int k = i / P::R2, r = i % P::R2;
        float2 w = __ldg(gtw + r * k * TWS2);
        if (SIGN < 0) w.y = -w.y;
        t2[k * (2 * P::TW2_ROW) + r] = w;
    }
    if constexpr (P::R3 > 1) {
        for (int i = threadIdx.x; i < P::R1 * P::R2; i += blockDim.x) {
            float2 w = __ldg(gtw + i);
            if (SIGN < 0) w.y = -w.y;
            tw3[i] = w;
        }
    }
}

// The same tables as a ready-made image in global memory (host-built once per handle, twiddle_image below): the
// per-CTA fill is then a straight 16-byte copy by NT threads (compile-time trip count, no index arithmetic).
template <int N, int NT, int PTS = PTS16>
__device__ __forceinline__ void load_twiddle_image(float4* smem_tw, const float4* __restrict__ img)
{
    constexpr int F4 = Plan<N, PTS>::TW_BYTES / 16;
#pragma unroll
    for (int i = 0; i < (F4 + NT - 1) / NT; ++i) {
        const int e = threadIdx.x + i * NT;
        if (F4 % NT == 0 || e < F4) smem_tw[e] = __ldg(img + e);
    }
}
// Host side: fills img (Plan<N>::TW_BYTES / 16 float4) with the layout load_twiddles produces.
template <class F2>
inline void twiddle_image_host(int n, int sign, float* img /* TW_BYTES / 4 floats */, F2 gtw /* gtw(x) -> (cos, sin)(2 pi x / n) */,
                               int pts = PTS16)
{
    const int r1 = pts;
    const int r2 = n / r1 < pts ? n / r1 : pts;
    const int r3 = n / (r1 * r2);
    const int row = r2 / 2 + 1;                 // float4 per tw2 row
    const int tw2_f4 = r1 * row;
    const int tws2 = n / (r1 * r2);
    for (int i = 0; i < tw2_f4 * 4; ++i) img[i] = 0.f;
    for (int k = 0; k < r1; ++k)
        for (int r = 0; r < r2; ++r) {
            float c, s;
            gtw((r * k * tws2) % n, c, s);
            img[(k * 2 * row + r) * 2 + 0] = c;
            img[(k * 2 * row + r) * 2 + 1] = sign < 0 ? -s : s;
        }
    if (r3 > 1)
        for (int i = 0; i < r1 * r2; ++i) {
            float c, s;
            gtw(i, c, s);
            img[tw2_f4 * 4 + 2 * i + 0] = c;
            img[tw2_f4 * 4 + 2 * i + 1] = sign < 0 ? -s : s;
        }
}
inline int twiddle_image_bytes(int n, int pts = PTS16)
{
    const int r1 = pts;
    const int r2 = n / r1 < pts ? n / r1 : pts;
    const int r3 = n / (r1 * r2);
    return r1 * (r2 / 2 + 1) * 16 + (r3 > 1 ? r1 * r2 * 8 : 0);
}

__device__ __forceinline__ float2 cmul_s(float2 a, float2 b)
{
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// Padded-index arithmetic without per-element shifts: for a multiple of 16, pad_idx(x + m) = pad_idx(x) +
// m + m / 16, so every access of a thread is "one base + compile-time offsets" (immediate-offset LDS/STS).
__host__ __device__ constexpr int pad_step(int m) { return m + m / 16; }

// Hands the R outputs of each of the B butterflies to emit(idx, pidx, value): idx = natural index of the
// output, pidx = pad_idx(idx) (for emitters that write into a line).
template <int N, int R, int S, int PTS, class Emit>
__device__ __forceinline__ void emit_all(cpk* v, int g, Emit&& emit)
{
    constexpr int T = N / PTS, B = PTS / R, LOGR = ilog2(R);
#pragma unroll
    for (int b = 0; b < B; ++b) {
        const int j = g + b * T;
        const int k = j & (S - 1);
        const int base = (j - k) * R + k;  // (j / S) * S * R + k
        if constexpr (S % 16 == 0) {
            const int pbase = pad_idx(base);
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const int rr = bitrev(i, LOGR);
                emit(base + rr * S, pbase + rr * pad_step(S), v[b + i * B]);
            }
        } else {  // S == 1 (first stage, R = PTS, B = 1): idx = R j + r, so pad_idx(idx) = (R + R / 16) j + r + r / 16
            static_assert(S == 1 && R == PTS && R % 16 == 0, "first stage is radix PTS");
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const int rr = bitrev(i, LOGR);
                emit(base + rr, (R + R / 16) * j + rr + rr / 16, v[b + i * B]);
            }
        }
    }
}

// Stage 1 (radix PTS, no twiddles)
template <int N, int SIGN, int PTS, class Emit>
__device__ __forceinline__ void stage1(cpk* v, int g, Emit&& emit)
{
    dft_all<SIGN, PTS, 1>(v);
    emit_all<N, PTS, 1, PTS>(v, g, emit);
}
// Stage 2 (radix R2, S = R1): twiddles W_{R1 R2}^{r k}, k = j mod R1, one table row per butterfly
template <int N, int SIGN, int PTS>
__device__ __forceinline__ void stage2_compute(cpk* v, int g, const float4* tw2)
{
    using P = Plan<N, PTS>;
    constexpr int R = P::R2, B = PTS / R, T = P::T;
#pragma unroll
    for (int b = 0; b < B; ++b) {
        const int k = (g + b * T) & (P::R1 - 1);
        const float4* row = tw2 + k * P::TW2_ROW;
#pragma unroll
        for (int r2 = 0; r2 < R / 2; ++r2) {
            const float4 w = row[r2];  // twiddles 2 r2 and 2 r2 + 1
            if (r2 > 0) v[b + (2 * r2) * B] = pmul(v[b + (2 * r2) * B], w.x, w.y);
            v[b + (2 * r2 + 1) * B] = pmul(v[b + (2 * r2 + 1) * B], w.z, w.w);
        }
    }
    dft_all<SIGN, R, B>(v);
}
template <int N, int SIGN, int PTS, class Emit>
__device__ __forceinline__ void stage2(cpk* v, int g, const float4* tw2, Emit&& emit)
{
    stage2_compute<N, SIGN, PTS>(v, g, tw2);
    emit_all<N, Plan<N, PTS>::R2, Plan<N, PTS>::R1, PTS>(v, g, emit);
}
// Stage 3 (radix R3, S = R1 R2): twiddles W_N^{r k}; w^1 from the table, higher powers by multiplication
template <int N, int SIGN, int PTS>
__device__ __forceinline__ void stage3_compute(cpk* v, int g, const float2* tw3)
{
    using P = Plan<N, PTS>;
    constexpr int R = P::R3, B = PTS / R, T = P::T, S = P::R1 * P::R2;
#pragma unroll
    for (int b = 0; b < B; ++b) {
        const int k = (g + b * T) & (S - 1);
        float2 pw[R];
        pw[1] = tw3[k];
#pragma unroll
        for (int r = 2; r < R; ++r) pw[r] = cmul_s(pw[r / 2], pw[r - r / 2]);
#pragma unroll
        for (int r = 1; r < R; ++r) v[b + r * B] = pmul(v[b + r * B], pw[r].x, pw[r].y);
    }
    dft_all<SIGN, R, B>(v);
}
template <int N, int SIGN, int PTS, class Emit>
__device__ __forceinline__ void stage3(cpk* v, int g, const float2* tw3, Emit&& emit)
{
    stage3_compute<N, SIGN, PTS>(v, g, tw3);
    emit_all<N, Plan<N, PTS>::R3, Plan<N, PTS>::R1 * Plan<N, PTS>::R2, PTS>(v, g, emit);
}

template <int N, int PTS = PTS16>
__device__ __forceinline__ void load_line_regs(cpk* v, const float4* line, int g)
{
    constexpr int T = N / PTS;
    if constexpr (T % 16 == 0) {
        const float4* p = line + pad_idx(g);
#pragma unroll
        for (int c = 0; c < PTS; ++c) {
            const float4 e = p[c * pad_step(T)];
            v[c].re = make_float2(e.x, e.y);
            v[c].im = make_float2(e.z, e.w);
        }
    } else {
#pragma unroll
        for (int c = 0; c < PTS; ++c) {
            const float4 e = line[pad_idx(g + T * c)];
            v[c].re = make_float2(e.x, e.y);
            v[c].im = make_float2(e.z, e.w);
        }
    }
}

// Full transform of one packed line held in shared memory (`line`, padded with pad_idx), by the T
// threads of its group (g = index within the group, line_id = barrier id of the group).  `active` =
// this group has a real line; inactive groups still arrive at their barrier.  The final stage's
// results go to `emit(idx, pidx, value)` in natural order (pidx = pad_idx(idx)); emit may write into
// the line itself (every read of the group is complete before the first emit).
template <int N, int SIGN, class Emit>
__device__ __forceinline__ void fft_line(float4* line, int g, int line_id, bool active, const float4* tw2,
                                         const float2* tw3, Emit&& emit)
{
    constexpr int PTS = PTS16;
    using P = Plan<N>;
    constexpr int T = P::T;
    constexpr int NSYNC = P::R3 == 1 ? 3 : 5;
    if (!active) {  // keep the barrier protocol, touch nothing (straight-line code for the active path)
#pragma unroll
        for (int i = 0; i < NSYNC; ++i) group_sync<T>(line_id);
        return;
    }
    cpk v[PTS16];
    auto to_smem = [&](int, int pidx, cpk val) { line[pidx] = make_float4(val.re.x, val.re.y, val.im.x, val.im.y); };
    load_line_regs<N>(v, line, g);
    group_sync<T>(line_id);  // everyone has read before anyone overwrites (in-place exchange)
    stage1<N, SIGN, PTS>(v, g, to_smem);
    group_sync<T>(line_id);
    load_line_regs<N>(v, line, g);
    group_sync<T>(line_id);
    if constexpr (P::R3 == 1) {
        stage2<N, SIGN, PTS>(v, g, tw2, emit);
    } else {
        stage2<N, SIGN, PTS>(v, g, tw2, to_smem);
        group_sync<T>(line_id);
        load_line_regs<N>(v, line, g);
        group_sync<T>(line_id);
        stage3<N, SIGN, PTS>(v, g, tw3, emit);
    }
}

// Register-to-register variant for callers whose threads already hold their first-stage inputs
// {line position g + T c} in v (loaded straight from global memory) and want the results left in
// registers: on return slot s of v holds output index final_idx<N>(g, s).  `sync` must synchronise all
// threads that share `line` (the caller chooses the thread <-> line mapping).  There is no "inactive"
// path on purpose: a group without a real line transforms zeros in its own line buffer, so that every
// thread of the CTA executes the same barrier instructions (no divergent __syncthreads).
template <int N, int PTS = PTS16>
struct Final {
    using P = Plan<N, PTS>;
    static constexpr int R = P::R3 == 1 ? P::R2 : P::R3;
    static constexpr int S = P::R3 == 1 ? P::R1 : P::R1 * P::R2;
    static constexpr int B = PTS / R;
    static constexpr int NSYNC = P::R3 == 1 ? 2 : 4;
};
template <int N, int PTS = PTS16>
__device__ __forceinline__ int final_idx(int g, int slot)
{
    using F = Final<N, PTS>;
    const int b = slot % F::B, i = slot / F::B;
    const int j = g + b * (N / PTS);
    const int k = j & (F::S - 1);
    return (j - k) * F::R + k + bitrev(i, ilog2(F::R)) * F::S;
}
// final_idx<N>(g, slot) == g + final_off<N>(slot): the group index only enters additively, because a thread's
// butterflies j = g + b T never reach the stage stride S (B T <= S for every plan) -- so every address derived from a
// result index is "one base + compile-time offset".
template <int N, int PTS = PTS16>
__host__ __device__ constexpr int final_off(int slot)
{
    using F = Final<N, PTS>;
    static_assert(F::B * (N / PTS) <= F::S, "j = g + b T must stay below the stage stride");
    return (slot % F::B) * (N / PTS) + bitrev(slot / F::B, ilog2(F::R)) * F::S;
}
template <int N, int SIGN, int PTS = PTS16, class Sync>
__device__ __forceinline__ void fft_line_inreg(cpk* v, float4* line, int g, const float4* tw2, const float2* tw3, Sync&& sync)
{
    using P = Plan<N, PTS>;
    auto to_smem = [&](int, int pidx, cpk val) { line[pidx] = make_float4(val.re.x, val.re.y, val.im.x, val.im.y); };
    stage1<N, SIGN, PTS>(v, g, to_smem);
    sync();
    load_line_regs<N, PTS>(v, line, g);
    sync();  // everyone has read before anyone overwrites (in-place exchange)
    if constexpr (P::R3 == 1) {
        stage2_compute<N, SIGN, PTS>(v, g, tw2);
    } else {
        stage2<N, SIGN, PTS>(v, g, tw2, to_smem);
        sync();
        load_line_regs<N, PTS>(v, line, g);
        sync();
        stage3_compute<N, SIGN, PTS>(v, g, tw3);
    }
}

// 16-byte asynchronous global -> shared copy (LDGSTS, L2 only: the data is streamed once)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

}  // namespace mwfft
