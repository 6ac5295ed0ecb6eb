// mw_cols_seam.cuh -- pass 2 without the halo line (sm_100a).
//
// k_cols_extract (mw_ocean_kernels.cuh) transforms W + 1 column lines per (A,B) slab: the extra one is the slab's east
// neighbour column, needed only for hds[index + 1] in the Jacobian of the slab's last column (FFTMesh.cs:264-267).  That
// column is transformed a second time by the CTA that owns it: 1/9 of pass 2's loads, exchanges and butterflies at W = 8,
// 1/5 at W = 4, and the reason narrow slabs (two CTAs per SM) never paid.
//
// Here the owner hands it over instead.  Every (A,B) CTA publishes the (dx, dz) / 2 values of its FIRST column to a small
// global array right after its transform (8 bytes per row) and raises a flag; its western neighbour picks them up between the
// part of the extraction that does not need them (hds, normals) and the part that does (Jacobian, whitecap).  A CTA can only
// ever wait for a CTA with a LOWER block index -- slabs are assigned in reverse order, east to west -- so the one it waits for
// has been dispatched before it and the wait cannot deadlock; the flags are cleared by pass 1 of the same tile (which runs
// between two pass 2s of a tile by stream order), so there is no epoch to carry and a replayed CUDA graph stays valid.
// The spin is bounded: a lost flag produces a wrong whitecap column that the parity tests catch, never a hung GPU.
#pragma once
#include "mw_ocean_kernels.cuh"

namespace mwk {

#ifndef MW_SEAM_SPIN_LIMIT
#define MW_SEAM_SPIN_LIMIT (1 << 20)
#endif

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float2 ld_cg2(const float2* p)  // L2 only: the data was written by another SM during this kernel
{
    float2 r;
    asm volatile("ld.global.cg.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p) : "memory");
    return r;
}

template <int N>
__host__ __device__ constexpr size_t seam_smem_bytes()
{
    return (size_t)Plan<N, fft_pts(N)>::TW_BYTES + (size_t)(slab_w(N) + 1) * mwfft::line_pitch(N, slab_w(N)) * sizeof(float4) +
           (size_t)((slab_w(N) * (N / fft_pts(N)) + 31) / 32) * 96 * nstage_slots(N) * sizeof(float);
}

// Without the halo group the CTA is W T threads, and 128 registers per thread fill the register file exactly at every
// resolution's CTAs per SM (1024: 512 threads x 1; 512: 256 x 2; 256: 128 x 4; 4-column slabs at 1024: 256 x 2).
__host__ __device__ constexpr int seam_maxreg(int N) { return fft_pts(N) == 32 ? 168 : 128; }

// W column lines per CTA, no halo group.  Thread <-> data as in k_cols_extract: thread tid owns line c = tid & (W - 1) and
// residue g = tid >> log2(W); a warp is 32 / W consecutive rows x W columns.
template <int N, int MINB, int OUTS>
__global__ void __launch_bounds__(slab_w(N) * (N / fft_pts(N)), MINB)
__maxnreg__(seam_maxreg(N)) k_cols_seam(const __grid_constant__ ColArgs a)
{
    constexpr int NS = nstage_slots(N);
    constexpr int PTS = fft_pts(N);
    using P = Plan<N, PTS>;
    constexpr int T = P::T;
    constexpr int W = slab_w(N);
    constexpr int NT = W * T;
    constexpr int LOGW = mwfft::ilog2(W);
    constexpr int LP = mwfft::line_pitch(N, W);
    constexpr bool LINEAR = (T % 16 == 0);
    extern __shared__ float4 smem4[];
    float4* tw2 = smem4;
    float2* tw3 = reinterpret_cast<float2*>(smem4 + P::TW2_F4);
    float4* lines = smem4 + P::TW_BYTES / 16;                              // [W][LP] + one more for the neighbour's (dx, dz) / 2
    float* nstage = reinterpret_cast<float*>(lines + (W + 1) * LP);         // [warps][NS][96]
    __shared__ uint64_t slab_bar;

    const int tile = a.tile0 + blockIdx.y;
    const int xt = blockIdx.y;
    const int tid = threadIdx.x;
    const int c = tid & (W - 1);
    const int g = tid >> LOGW;
    const size_t plane = (size_t)N * N;
    const size_t obase = (size_t)tile * plane;
    float4* line = lines + c * LP;
    auto cta_sync = [] { __syncthreads(); };

    const uint64_t pol = evict_first_policy();
    mwfft::cpk v[PTS];
    const int nab = a.ab_blocks;
    const bool is_ab = (int)blockIdx.x < nab;
    // (A,B) slabs east to west: block 0 owns the last slab, which has no east neighbour and waits for nobody
    const int slab = is_ab ? nab - 1 - (int)blockIdx.x : (int)blockIdx.x - nab;
    const int b0 = is_ab ? slab * W : slab * (4 * W);
    if (a.pdl == 1) pdl_trigger();
    pdl_wait();  // the intermediate is pass 1's (the predecessor in this stream)

    if (is_ab) {
        // the slab is one contiguous block of N * W * 16 bytes: one bulk copy into the line buffers (free until stage 1 writes them)
        float4* raw = lines;
        if (tid == 0) mbar_init(&slab_bar, 1);
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&slab_bar, (unsigned)(N * W * sizeof(float4)));
            bulk_g2s(raw, a.XAB + (size_t)xt * xab_tile_elems(N) + (size_t)slab * N * W, (unsigned)(N * W * sizeof(float4)), &slab_bar);
        }
        mwfft::load_twiddle_image<N, NT, PTS>(smem4, a.twimg);
        mbar_wait(&slab_bar, 0);
        const float4* src = raw + g * W + c;
#pragma unroll
        for (int k = 0; k < PTS; ++k) {
            const float4 e = src[(T * ((k + PTS / 2) & (PTS - 1))) * W];
            v[k].re = make_float2(e.x, e.y);
            v[k].im = make_float2(e.z, e.w);
        }
        __syncthreads();  // everyone has its inputs before the first stage overwrites the raw slab with the lines
    } else {
        // C slab: 4 W columns, line c = columns b0 + 4c .. 4c + 3 as two real-pair transforms (see k_cols_extract)
        const float4* src = reinterpret_cast<const float4*>(a.XC + (size_t)xt * plane + ((size_t)slab * N + g) * (4 * W)) + c;
#pragma unroll
        for (int k = 0; k < PTS; ++k) {
            const float4* p = src + (size_t)(T * ((k + PTS / 2) & (PTS - 1))) * (2 * W);
            const float4 e0 = ldg_fresh4(p), e1 = ldg_fresh4(p + W);
            v[k].re = make_float2(e0.x - e0.w, e1.x - e1.w);
            v[k].im = make_float2(e0.y + e0.z, e1.y + e1.z);
        }
        mwfft::load_twiddle_image<N, NT, PTS>(smem4, a.twimg);
    }
    mwfft::fft_line_inreg<N, +1, PTS>(v, line, g, tw2, tw3, cta_sync);
    if (a.pdl == 2) pdl_trigger();

    if (!is_ab) {
        // height (FFTMesh.cs:219) of four consecutive columns: (Re, Im) of lane x, (Re, Im) of lane y
        float* dst = a.height + obase + (size_t)g * N + b0 + 4 * c;
#pragma unroll
        for (int s = 0; s < PTS; ++s)
            st_once(reinterpret_cast<float4*>(dst + (size_t)mwfft::final_off<N, PTS>(s) * N),
                    make_float4(v[s].re.x, v[s].im.x, v[s].re.y, v[s].im.y), pol);
        return;
    }

    // ------------------------------------------------------------------ (A, B) slab
    const bool has_disp = OUTS < 0 ? a.disp != nullptr : (OUTS & 1) != 0;
    const bool has_normal = OUTS < 0 ? a.normal != nullptr : (OUTS & 2) != 0;
    const bool has_white = OUTS < 0 ? a.whitecap != nullptr : (OUTS & 4) != 0;
    const bool has_jac = OUTS < 0 ? a.jacobian != nullptr : (OUTS & 8) != 0;
    const bool need_d = has_white || has_jac;
    const bool last_slab = slab == nab - 1;  // CTA-uniform
    float2* D = reinterpret_cast<float2*>(line);
    float2* Dh = reinterpret_cast<float2*>(lines + W * LP);   // the east neighbour's first column
    const int pg = pad_idx(g);
    auto dpos = [&](int s) { return LINEAR ? pg + mwfft::pad_step(mwfft::final_off<N, PTS>(s)) : pad_idx(g + mwfft::final_off<N, PTS>(s)); };
    float2* seam_mine = a.seam + ((size_t)tile * nab + slab) * N;
    unsigned* flags = a.seam_flags + (size_t)tile * nab;
    if (need_d) {
        // (dx, dz) / 2 in place of the line for the neighbours' forward differences (FFTMesh.cs:260-267); "no neighbour =>
        // derivative 0" is data, not a branch: row N holds a copy of row N - 1, the last slab's east column a copy of column N - 1
#pragma unroll
        for (int s = 0; s < PTS; ++s) D[dpos(s)] = make_float2(0.5f * v[s].re.x, 0.5f * v[s].im.x);
        if (g == T - 1) D[pad_idx(N)] = make_float2(0.5f * v[PTS - 1].re.x, 0.5f * v[PTS - 1].im.x);  // the last slot is row g + N - T
        if (last_slab && c == W - 1) {
#pragma unroll
            for (int s = 0; s < PTS; ++s) Dh[dpos(s)] = make_float2(0.5f * v[s].re.x, 0.5f * v[s].im.x);
        }
        if (slab > 0 && c == 0) {  // my first column is my western neighbour's east column
#pragma unroll
            for (int s = 0; s < PTS; ++s)
                seam_mine[g + mwfft::final_off<N, PTS>(s)] = make_float2(0.5f * v[s].re.x, 0.5f * v[s].im.x);
        }
    }
    __syncthreads();
    if (need_d && slab > 0 && tid == 0) {
        __threadfence();                   // the column (written by other threads, ordered by the barrier) before the flag
        st_release_u32(flags + slab, 1u);
    }

    const int lane = tid & 31;
    const size_t o0 = obase + (size_t)g * N + b0 + c;
    float2* p_disp = has_disp ? a.disp + o0 : nullptr;
    float* p_white = has_white ? a.whitecap + o0 : nullptr;
    float* p_jac = has_jac ? a.jacobian + o0 : nullptr;
    // normals leave as 16-byte stores through a per-warp staging block (see k_cols_extract)
    constexpr int QR = 3 * W / 4;
    const int rr = lane / QR, qq = lane - QR * rr;
    float* wst = nstage + (tid >> 5) * (96 * NS);
    float4* p_nrm = has_normal ? reinterpret_cast<float4*>(a.normal) + (3 * (obase + (size_t)(g - (lane >> LOGW) + rr) * N + b0)) / 4 + qq
                               : nullptr;
    const bool nrm_lane = lane < 24 && (NT >= 32 || tid - lane + W * rr < NT);

    // ---- part A: hds and normals (FFTMesh.cs:247, :212-218) -- nothing here needs a neighbour ----
#pragma unroll
    for (int s0 = 0; s0 < PTS; s0 += NS) {
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            const int s = s0 + j;
            const int off = mwfft::final_off<N, PTS>(s);
            const float dx = v[s].re.x, sx = v[s].re.y, dz = v[s].im.x, sz = v[s].im.y;
            if (has_normal) {
                const float inv = rsqrt_ftz(fmaf(sx, sx, sz * sz) + 1.0f);  // nor = (sx, 1, sz) / |.|
                wst[96 * j + 3 * lane + 0] = sx * inv;
                wst[96 * j + 3 * lane + 1] = inv;
                wst[96 * j + 3 * lane + 2] = sz * inv;
            }
            if (has_disp) st_once(p_disp + (size_t)off * N, make_float2(dx, dz), pol);
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
    if (!need_d) return;

    // ---- the east neighbour's column: published by the CTA of slab + 1 (a lower block index) ----
    if (!last_slab) {
        if (tid == 0) {
            int spins = 0;
            while (ld_acquire_u32(flags + slab + 1) == 0u && ++spins < MW_SEAM_SPIN_LIMIT) { }
            if (spins >= MW_SEAM_SPIN_LIMIT) atomicAdd(a.seam_timeouts, 1u);   // reported by mw_ocean_sync as an error
        }
        __syncthreads();
        const float2* theirs = a.seam + ((size_t)tile * nab + slab + 1) * N;
#pragma unroll
        for (int r = tid; r < N; r += NT) Dh[pad_idx(r)] = ld_cg2(theirs + r);
        __syncthreads();
    }

    // ---- part B: Jacobian and whitecap (FFTMesh.cs:253-276) ----
    {
        const int dn = pad_idx(g + 1) - pg;  // padded distance to the next row (1 or 2)
        const float2* De = reinterpret_cast<const float2*>(line + LP);   // line c + 1, or the neighbour's column for c == W - 1
#pragma unroll
        for (int s = 0; s < PTS; ++s) {
            const int off = mwfft::final_off<N, PTS>(s);
            const float dx = v[s].re.x, sx = v[s].re.y, dz = v[s].im.x, sz = v[s].im.y;
            const int pa = dpos(s);
            const float2 nbs = D[LINEAR ? pa + dn : pad_idx(g + off + 1)];  // hds[index + resolution] / 2  (:260-263)
            const float2 nbe = De[pa];                                       // hds[index + 1] / 2           (:264-267)
            const float hx = 0.5f * dx, hz = 0.5f * dz;
            const float ddx_x = hx - nbs.x, ddx_y = hz - nbs.y, ddy_x = hx - nbe.x, ddy_y = hz - nbe.y;
            const float jac = fmaf(1.0f + ddx_x, 1.0f + ddy_y, -(ddx_y * ddy_x));  // :268
            if (has_jac) st_once(p_jac + (size_t)off * N, jac, pol);
            if (has_white) {
                const float r2 = fmaf(sx, sx, sz * sz);
                const float inv = rsqrt_ftz(r2 + 1.0f);
                const float noise = 0.3f * inv * sqrt_approx(r2);                  // :269-270
                float turb = fmaxf(1.0f - jac + noise, 0.0f);
                turb = fminf(turb, 1.0f);
                st_once(p_white + (size_t)off * N, turb * turb * fmaf(-2.0f, turb, 3.0f), pol);  // :273
            }
        }
    }
}

}  // namespace mwk
