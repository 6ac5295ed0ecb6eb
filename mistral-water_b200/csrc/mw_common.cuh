// mw_common.cuh -- shared helpers for libmistral_ocean.so (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/mistral_ocean.h"

// ---------------------------------------------------------------------------------------------
// error plumbing: thread-local message, negative status codes, no exceptions across the ABI
// ---------------------------------------------------------------------------------------------
void mw_set_error(const char* fmt, ...);
extern std::atomic<long long> g_mw_launches;

#define MW_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            mw_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return (_e == cudaErrorMemoryAllocation) ? MW_E_OOM : MW_E_CUDA;                       \
        }                                                                                          \
    } while (0)

#define MW_LAUNCH_CHECK()                                                                          \
    do {                                                                                           \
        g_mw_launches.fetch_add(1, std::memory_order_relaxed);                                     \
        cudaError_t _e = cudaGetLastError();                                                       \
        if (_e != cudaSuccess) {                                                                   \
            mw_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return MW_E_CUDA;                                                                      \
        }                                                                                          \
    } while (0)

// ---------------------------------------------------------------------------------------------
// constants of the reference (Scripts/FFTMesh.cs:50-54)
// ---------------------------------------------------------------------------------------------
#define MW_PI_F 3.1415926536f
#define MW_G_F 9.81f
#define MW_EPSILON_F 0.0001f

// ---------------------------------------------------------------------------------------------
// complex arithmetic on float2
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// Programmatic dependent launch (sm_90+): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may
// become resident while its predecessor in the stream is still running.  pdl_wait() blocks until the predecessor grid has
// completed and its writes are visible (a no-op for a normal launch); pdl_trigger() tells the scheduler that this CTA no
// longer minds the successor's CTAs being placed (they take free slots only and park at their own pdl_wait()).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Host side: launch `kern(arg)`; with pdl the launch carries the programmatic-serialisation attribute.
template <class Kern, class Arg>
static inline cudaError_t mw_launch(Kern kern, dim3 grid, unsigned threads, size_t smem, cudaStream_t st, bool pdl, const Arg& arg)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, arg);
}

// streaming (read-once / write-once) global accesses: keep them out of L1
__device__ __forceinline__ float4 ldg_stream4(const float4* p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float2 ldg_stream2(const float2* p)
{
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}

// Data written by the PREDECESSOR kernel of a programmatic dependent launch (the pass-1 -> pass-2 intermediate, the per-frame
// phase table, the renderer's images): read after pdl_wait() with plain (weak, coherent) loads.  ld.global.nc promises the
// data is read-only for the whole lifetime of the reading kernel, which does not hold while the producer grid is still
// resident next to it -- the non-.nc forms below are what the PTX memory model covers.
__device__ __forceinline__ float4 ldg_fresh4(const float4* p)
{
    float4 r;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ float2 ldg_fresh2(const float2* p)
{
    float2 r;
    asm volatile("ld.global.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ float4 ld_plain4(const float4* p)
{
    float4 r;
    asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ float ld_plain1(const float* p)
{
    float r;
    asm volatile("ld.global.f32 %0, [%1];" : "=f"(r) : "l"(p) : "memory");
    return r;
}

// Developer hooks (phase stamps, "switch one phase off" flags; tools/phase_timing.py): compiled in only with -DMW_DEVHOOKS=1,
// so that the production kernels carry no trace of them.
#ifndef MW_DEVHOOKS
#define MW_DEVHOOKS 0
#endif
#define MW_DBG(a, bits) (MW_DEVHOOKS && ((a).dbg_flags & (bits)))

// Streaming variants with an evict-first L2 priority: for data that is touched exactly once (the spectrum on its way in,
// the outputs on their way out), so that it does not push the pass-1 -> pass-2 intermediate out of the 126 MB L2.
#ifndef MW_EVICT_FIRST
#define MW_EVICT_FIRST 1
#endif
__device__ __forceinline__ uint64_t evict_first_policy()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float4 ldg_once4(const float4* p, uint64_t pol)
{
#if MW_EVICT_FIRST
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
    return r;
#else
    (void)pol;
    return ldg_stream4(p);
#endif
}
__device__ __forceinline__ void st_once(float* p, float v, uint64_t pol)
{
#if MW_EVICT_FIRST
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
#else
    (void)pol; *p = v;
#endif
}
__device__ __forceinline__ void st_once(float2* p, float2 v, uint64_t pol)
{
#if MW_EVICT_FIRST
    asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1,%2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
#else
    (void)pol; *p = v;
#endif
}
__device__ __forceinline__ void st_once(float4* p, float4 v, uint64_t pol)
{
#if MW_EVICT_FIRST
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
#else
    (void)pol; *p = v;
#endif
}

// ---------------------------------------------------------------------------------------------
// bulk asynchronous copy global -> shared (the TMA unit's 1-D form) completing on an mbarrier.  One thread issues it; the
// bytes in flight are tracked by the copy engine, not by the LSU's per-thread request queue.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* mbar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* mbar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* mbar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, unsigned parity)
{
    asm volatile("{\n.reg .pred P1;\nMW_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra MW_DONE;\nbra MW_WAIT;\nMW_DONE:\n}"
                 ::"r"(smem_u32(mbar)), "r"(parity) : "memory");
}

// 2-D tensor store shared -> global through a CUtensorMap (TMA), bulk-group completion
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, const void* smem_src, int x, int y)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(smem_u32(smem_src)) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11), the engine's stand-in for UnityEngine.Random.value.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// 24-bit uniform in (0, 1]
__host__ __device__ __forceinline__ float u32_to_unit_open0(uint32_t u)
{
    return (float)((u >> 8) + 1u) * (1.0f / 16777216.0f);
}
