// mw_renderer.cu -- handle management and C ABI of the OceanRenderer (GPU-shader convention) path
// (include/mistral_ocean.h, mw_renderer_*), plus mw_mesh_generate.
#include <math.h>
#include <stdlib.h>
#include <new>
#include <vector>

#include "mw_renderer_kernels.cuh"

struct mw_renderer {
    mw_renderer_params p;
    int R = 0;        // texture resolution = 8 * resolution (OceanRenderer.cs:136)
    int tiles = 1;
    size_t n2 = 0;
    bool device_ptrs = false;
    bool have_initial = false;
    cudaStream_t stream = nullptr;
    float4* initial = nullptr;  // [tiles][R][R]  initialTexture
    float* phase = nullptr;     // [tiles][R][R]  ping/pong phase textures (updated in place)
    float* rate = nullptr;      // [R][R]
    float* kw = nullptr;        // [R]
    float4* twimg = nullptr;
    float4* XAB = nullptr;      // intermediate: two slots of one tile group each
    float2* XC = nullptr;
    int group_tiles = 1, x_tiles = 1;
    bool pdl = true;   // programmatic dependent launch of the frame kernels (MW_PDL=0 switches it off)
    // tile-group pipelining as in the FFTMesh path: groups alternate between two streams and two slots of the
    // intermediate, so that pass 1 of one group overlaps pass 2 of the previous one and the intermediate stays in L2
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // scratch images (host-pointer mode, or inputs of pass 3 the caller did not ask for)
    float4* s_disp = nullptr; float4* s_height = nullptr; float4* s_normal = nullptr; float* s_white = nullptr;
    float4* s_white4 = nullptr; float* s_jac = nullptr;
};

template <class T>
static int r_ensure(T** p, size_t count)
{
    if (*p) return MW_OK;
    MW_CUDA(cudaMalloc((void**)p, count * sizeof(T)));
    return MW_OK;
}

#define MWR_CHECK(h)                                                               \
    do {                                                                           \
        if (!(h)) { mw_set_error("null handle"); return MW_E_INVALID_ARG; }        \
        MW_CUDA(cudaSetDevice((h)->p.device));                                     \
    } while (0)

extern "C" void mw_renderer_destroy(mw_renderer* r)
{
    if (!r) return;
    cudaSetDevice(r->p.device);
    if (r->stream) cudaStreamSynchronize(r->stream);
    void* ptrs[] = {r->initial, r->phase, r->rate, r->kw, r->twimg, r->XAB, r->s_disp, r->s_height, r->s_normal,
                    r->s_white, r->s_white4, r->s_jac};
    for (void* q : ptrs) if (q) cudaFree(q);
    if (r->aux_stream) { cudaStreamSynchronize(r->aux_stream); cudaStreamDestroy(r->aux_stream); }
    if (r->ev_fork) cudaEventDestroy(r->ev_fork);
    if (r->ev_join) cudaEventDestroy(r->ev_join);
    if (r->stream) cudaStreamDestroy(r->stream);
    delete r;
}

static int upload_tables(mw_renderer* r)
{
    const int R = r->R;

    mwr::k_r_rate<<<(unsigned)((r->n2 + 255) / 256), 256, 0, r->stream>>>(r->rate, R, r->p.length);
    MW_LAUNCH_CHECK();
    // GetWave component per texel index, same fp32 operation order as the device code
    std::vector<float> kw(R);
    for (int i = 0; i < R; ++i) {
        volatile float two_pi = 2.0f * 3.1415926536f;
        volatile float n = (float)(i < R / 2 ? i : i - R);
        volatile float num = two_pi * n;
        kw[i] = num / r->p.length;
    }
    MW_CUDA(cudaMemcpyAsync(r->kw, kw.data(), R * sizeof(float), cudaMemcpyHostToDevice, r->stream));
    MW_CUDA(cudaStreamSynchronize(r->stream));
    return MW_OK;
}

extern "C" int mw_renderer_create(const mw_renderer_params* params, mw_renderer** out)
{
    if (!params || !out) { mw_set_error("mw_renderer_create: null argument"); return MW_E_INVALID_ARG; }
    *out = nullptr;
    const mw_renderer_params& p = *params;
    const long long R = 8LL * p.resolution;
    if (p.resolution < 4 || R > 2048 || (R & (R - 1))) {
        mw_set_error("resolution must be a power of two in [4, 256] (texture side 8 * resolution in [32, 2048]), got %d", p.resolution);
        return MW_E_INVALID_ARG;
    }
    if (!(p.length > 0.f) || !isfinite(p.length)) { mw_set_error("length must be positive and finite"); return MW_E_INVALID_ARG; }
    if (p.tiles < 1 || p.tiles > 65535) { mw_set_error("tiles must be in [1, 65535], got %d", p.tiles); return MW_E_INVALID_ARG; }
    int ndev = 0;
    MW_CUDA(cudaGetDeviceCount(&ndev));
    if (p.device < 0 || p.device >= ndev) { mw_set_error("device %d out of range (%d devices)", p.device, ndev); return MW_E_INVALID_ARG; }
    MW_CUDA(cudaSetDevice(p.device));
    {
        cudaDeviceProp prop;
        MW_CUDA(cudaGetDeviceProperties(&prop, p.device));
        if (prop.major != 10) {
            mw_set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", p.device, prop.major, prop.minor);
            return MW_E_CUDA;
        }
    }
    mw_renderer* r = new (std::nothrow) mw_renderer();
    if (!r) { mw_set_error("out of host memory"); return MW_E_OOM; }
    r->p = p;
    r->R = (int)R;
    r->tiles = p.tiles;
    r->n2 = (size_t)R * R;
    r->device_ptrs = (p.flags & MW_DEVICE_PTRS) != 0;
    int rc = MW_OK;
    auto fail = [&](int code) { mw_renderer_destroy(r); return code; };
    if (cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking) != cudaSuccess) { mw_set_error("cudaStreamCreate failed"); return fail(MW_E_CUDA); }
    if ((rc = r_ensure(&r->initial, r->n2 * r->tiles))) return fail(rc);
    if ((rc = r_ensure(&r->phase, r->n2 * r->tiles))) return fail(rc);
    if ((rc = r_ensure(&r->rate, r->n2))) return fail(rc);
    if ((rc = r_ensure(&r->kw, (size_t)R))) return fail(rc);
    if ((rc = r_ensure(&r->twimg, (size_t)mwfft::twiddle_image_bytes((int)R) / 16))) return fail(rc);
    {
        long long gt = (32ll << 20) / (long long)(r->n2 * 24);
        if (gt < 1) gt = 1;
        if (gt > r->tiles) gt = r->tiles;
        r->group_tiles = (int)gt;
        if (const char* e = getenv("MW_PDL")) r->pdl = atoi(e) != 0;
        r->x_tiles = r->tiles <= r->group_tiles ? r->tiles : 2 * r->group_tiles;
        char* x = nullptr;
        const size_t xab_bytes = r->n2 * sizeof(float4) * r->x_tiles;
        if ((rc = r_ensure(&x, xab_bytes + r->n2 * r->x_tiles * 8))) return fail(rc);
        if (cudaStreamCreateWithFlags(&r->aux_stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&r->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&r->ev_join, cudaEventDisableTiming) != cudaSuccess) {
            mw_set_error("stream/event creation failed"); return fail(MW_E_CUDA);
        }
        r->XAB = reinterpret_cast<float4*>(x);
        r->XC = reinterpret_cast<float2*>(x + xab_bytes);
    }
    // RenderTextures start black: phase = 0 (OceanRenderer.cs:138-139)
    if (cudaMemsetAsync(r->phase, 0, r->n2 * r->tiles * sizeof(float), r->stream) != cudaSuccess) { mw_set_error("memset failed"); return fail(MW_E_CUDA); }
    {
        const double PI_D = 3.14159265358979323846;
        std::vector<float> img(mwfft::twiddle_image_bytes((int)R) / 4);
        mwfft::twiddle_image_host((int)R, -1, img.data(), [&](int x, float& c, float& s) {
            c = (float)cos(2.0 * PI_D * x / (double)R); s = (float)sin(2.0 * PI_D * x / (double)R); });
        if (cudaMemcpyAsync(r->twimg, img.data(), img.size() * sizeof(float), cudaMemcpyHostToDevice, r->stream) != cudaSuccess ||
            cudaStreamSynchronize(r->stream) != cudaSuccess) { mw_set_error("table upload failed"); return fail(MW_E_CUDA); }
    }
    if ((rc = upload_tables(r))) return fail(rc);
    *out = r;
    return MW_OK;
}

extern "C" int mw_renderer_sync(mw_renderer* r)
{
    MWR_CHECK(r);
    MW_CUDA(cudaStreamSynchronize(r->stream));
    return MW_OK;
}

// RenderInitial (OceanRenderer.cs:209-214) with _Amplitude = amplitude / 10000 (:149)
extern "C" int mw_renderer_render_initial(mw_renderer* r)
{
    MWR_CHECK(r);
    const size_t total = r->n2 * r->tiles;
    volatile float amp = r->p.amplitude / 10000.0f;
    mwr::k_r_initial<<<(unsigned)((total + 127) / 128), 128, 0, r->stream>>>(r->initial, r->R, r->tiles, r->p.length, amp,
                                                                             r->p.wind_x, r->p.wind_y, r->p.seed1, r->p.seed2);
    MW_LAUNCH_CHECK();
    if (!r->device_ptrs) MW_CUDA(cudaStreamSynchronize(r->stream));
    r->have_initial = true;
    return MW_OK;
}

static int copy_image(mw_renderer* r, void* dst, const void* src, size_t bytes, bool to_engine)
{
    const cudaMemcpyKind kind = r->device_ptrs ? cudaMemcpyDeviceToDevice : (to_engine ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost);
    MW_CUDA(cudaMemcpyAsync(dst, src, bytes, kind, r->stream));
    if (!r->device_ptrs) MW_CUDA(cudaStreamSynchronize(r->stream));
    return MW_OK;
}

extern "C" int mw_renderer_set_initial(mw_renderer* r, const float* rgba)
{
    MWR_CHECK(r);
    if (!rgba) { mw_set_error("null buffer"); return MW_E_INVALID_ARG; }
    int rc = copy_image(r, r->initial, rgba, r->n2 * r->tiles * sizeof(float4), true);
    if (rc == MW_OK) r->have_initial = true;
    return rc;
}
extern "C" int mw_renderer_get_initial(mw_renderer* r, float* rgba)
{
    MWR_CHECK(r);
    if (!rgba) { mw_set_error("null buffer"); return MW_E_INVALID_ARG; }
    if (!r->have_initial) { mw_set_error("initial spectrum not rendered: call mw_renderer_render_initial or mw_renderer_set_initial first"); return MW_E_STATE; }
    return copy_image(r, rgba, r->initial, r->n2 * r->tiles * sizeof(float4), false);
}
extern "C" int mw_renderer_set_phase(mw_renderer* r, const float* phase)
{
    MWR_CHECK(r);
    if (!phase) { mw_set_error("null buffer"); return MW_E_INVALID_ARG; }
    return copy_image(r, r->phase, phase, r->n2 * r->tiles * sizeof(float), true);
}
extern "C" int mw_renderer_get_phase(mw_renderer* r, float* phase)
{
    MWR_CHECK(r);
    if (!phase) { mw_set_error("null buffer"); return MW_E_INVALID_ARG; }
    return copy_image(r, phase, r->phase, r->n2 * r->tiles * sizeof(float), false);
}

// OceanRenderer.Update's parameter refresh (OceanRenderer.cs:94-109): choppiness and length take effect on the next
// frame; a change of length, wind or amplitude re-renders the initial spectrum.
extern "C" int mw_renderer_set_params(mw_renderer* r, float length, float choppiness, float amplitude, float wind_x, float wind_y)
{
    MWR_CHECK(r);
    if (!(length > 0.f) || !isfinite(length)) { mw_set_error("length must be positive and finite"); return MW_E_INVALID_ARG; }
    const bool rerender = length != r->p.length || amplitude != r->p.amplitude || wind_x != r->p.wind_x || wind_y != r->p.wind_y;
    const bool retable = length != r->p.length;
    r->p.length = length; r->p.choppiness = choppiness; r->p.amplitude = amplitude; r->p.wind_x = wind_x; r->p.wind_y = wind_y;
    if (retable) { int rc = upload_tables(r); if (rc) return rc; }
    if (rerender) return mw_renderer_render_initial(r);
    return MW_OK;
}

#ifndef MW_RMAPS_PER_GROUP
#define MW_RMAPS_PER_GROUP 1   // k_r_maps right behind each tile group's k_r_cols (its two images still in L2); 0 = one launch at the end
#endif
template <int N>
static int run_renderer_frame(mw_renderer* r, float dt, float4* d_disp, float4* d_height, const mwr::RMapArgs* maps)
{
    using P = mwfft::Plan<N>;
    constexpr int T = P::T;
    constexpr int W = mwk::slab_w(N);
    constexpr size_t smem_r = P::TW_BYTES + (size_t)3 * mwfft::line_pitch(N, 8) * sizeof(float4);
    constexpr size_t smem_c = P::TW_BYTES + (size_t)W * mwfft::line_pitch(N, W) * sizeof(float4);
    static bool attr_done[64] = {};
    if (!attr_done[r->p.device]) {
        MW_CUDA(cudaFuncSetAttribute(mwr::k_r_rows<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r));
        MW_CUDA(cudaFuncSetAttribute(mwr::k_r_cols<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
        MW_CUDA(cudaFuncSetAttribute(mwr::k_r_rows<N>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        MW_CUDA(cudaFuncSetAttribute(mwr::k_r_cols<N>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        attr_done[r->p.device] = true;
    }
    const int G = r->group_tiles;
    const int ngroups = (r->tiles + G - 1) / G;
    const bool dual = ngroups > 1;
    if (dual) {
        MW_CUDA(cudaEventRecord(r->ev_fork, r->stream));
        MW_CUDA(cudaStreamWaitEvent(r->aux_stream, r->ev_fork, 0));
    }
    for (int gi = 0; gi < ngroups; ++gi) {
        const int t0 = gi * G;
        const int nt = r->tiles - t0 < G ? r->tiles - t0 : G;
        const int slot = dual ? (gi & 1) : 0;
        cudaStream_t st = (dual && slot) ? r->aux_stream : r->stream;
        float4* xab = r->XAB + (size_t)slot * G * r->n2;
        float2* xc = r->XC + (size_t)slot * G * r->n2;
        mwr::RRowArgs ra{r->initial, r->phase, r->rate, r->kw, r->twimg, xab, xc, dt, r->p.choppiness, t0};
        MW_CUDA(mw_launch(mwr::k_r_rows<N>, dim3(N / 2, nt), 3 * T, smem_r, st, r->pdl, ra));
        MW_LAUNCH_CHECK();
        mwr::RColArgs ca{xab, xc, r->twimg, d_disp, d_height, t0, N / W};
        MW_CUDA(mw_launch(mwr::k_r_cols<N>, dim3(N / W + N / (2 * W), nt), W * T, smem_c, st, r->pdl, ca));
        MW_LAUNCH_CHECK();
        if (maps && MW_RMAPS_PER_GROUP) {
            // pass 3 of this group straight away: the group's displacement / height images (32 B per texel) are still in L2
            mwr::RMapArgs ma = *maps;
            ma.tile0 = t0;
            MW_CUDA(mw_launch(mwr::k_r_maps, dim3((N + 31) / 32, (N + 7) / 8, nt), 256, 0, st, r->pdl, ma));
            MW_LAUNCH_CHECK();
        }
    }
    if (dual) {
        MW_CUDA(cudaEventRecord(r->ev_join, r->aux_stream));
        MW_CUDA(cudaStreamWaitEvent(r->stream, r->ev_join, 0));
    }
    return MW_OK;
}

// GenerateTexture (OceanRenderer.cs:216-316)
extern "C" int mw_renderer_generate_texture(mw_renderer* r, float delta_time, const mw_renderer_out* out)
{
    MWR_CHECK(r);
    if (!out) { mw_set_error("mw_renderer_generate_texture: null output block"); return MW_E_INVALID_ARG; }
    if (!r->have_initial) { mw_set_error("initial spectrum not rendered: call mw_renderer_render_initial or mw_renderer_set_initial first"); return MW_E_STATE; }
    const size_t total = r->n2 * r->tiles;
    const bool dev = r->device_ptrs;
    int rc;
    auto pick4 = [&](float* user, float4** scratch, float4** d) -> int {
        if (dev && user) { *d = (float4*)user; return MW_OK; }
        int e = r_ensure(scratch, total);
        *d = *scratch;
        return e;
    };
    float4 *d_disp = nullptr, *d_height = nullptr, *d_normal = nullptr, *d_white4 = nullptr;
    float *d_white = nullptr, *d_jac = nullptr;
    if ((rc = pick4(out->displacement, &r->s_disp, &d_disp))) return rc;      // pass 3 always needs both images
    if ((rc = pick4(out->height, &r->s_height, &d_height))) return rc;
    const bool want_white = out->white || out->white_rgba || out->jacobian;
    if (out->normal && (rc = pick4(out->normal, &r->s_normal, &d_normal))) return rc;
    if (out->white_rgba && (rc = pick4(out->white_rgba, &r->s_white4, &d_white4))) return rc;
    if (out->white) { if (dev) d_white = out->white; else { if ((rc = r_ensure(&r->s_white, total))) return rc; d_white = r->s_white; } }
    if (out->jacobian) { if (dev) d_jac = out->jacobian; else { if ((rc = r_ensure(&r->s_jac, total))) return rc; d_jac = r->s_jac; } }

    volatile float dt = delta_time * r->p.mult;  // OceanRenderer.cs:223
    const mwr::RMapArgs ma{d_disp, d_height, d_normal, d_white, d_white4, d_jac, r->R, 0, r->R / r->p.resolution,
                           (r->p.flags & MW_WRAP_REPEAT) ? 1 : 0, r->p.length / (float)r->R};
    const mwr::RMapArgs* mp = (d_normal || want_white) ? &ma : nullptr;
    switch (r->R) {
        case 32: rc = run_renderer_frame<32>(r, dt, d_disp, d_height, mp); break;
        case 64: rc = run_renderer_frame<64>(r, dt, d_disp, d_height, mp); break;
        case 128: rc = run_renderer_frame<128>(r, dt, d_disp, d_height, mp); break;
        case 256: rc = run_renderer_frame<256>(r, dt, d_disp, d_height, mp); break;
        case 512: rc = run_renderer_frame<512>(r, dt, d_disp, d_height, mp); break;
        case 1024: rc = run_renderer_frame<1024>(r, dt, d_disp, d_height, mp); break;
        case 2048: rc = run_renderer_frame<2048>(r, dt, d_disp, d_height, mp); break;
        default: mw_set_error("unsupported texture resolution %d", r->R); return MW_E_INVALID_ARG;
    }
    if (rc) return rc;
    if (mp && !MW_RMAPS_PER_GROUP) {
        MW_CUDA(mw_launch(mwr::k_r_maps, dim3((r->R + 31) / 32, (r->R + 7) / 8, r->tiles), 256, 0, r->stream, r->pdl, ma));
        MW_LAUNCH_CHECK();
    }
    if (!dev) {
        auto d2h = [&](void* h, const void* d, size_t bytes) -> cudaError_t {
            return h ? cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, r->stream) : cudaSuccess;
        };
        MW_CUDA(d2h(out->displacement, d_disp, total * 16));
        MW_CUDA(d2h(out->height, d_height, total * 16));
        MW_CUDA(d2h(out->normal, d_normal, total * 16));
        MW_CUDA(d2h(out->white, d_white, total * 4));
        MW_CUDA(d2h(out->white_rgba, d_white4, total * 16));
        MW_CUDA(d2h(out->jacobian, d_jac, total * 4));
        MW_CUDA(cudaStreamSynchronize(r->stream));
    }
    return MW_OK;
}

// OceanRenderer.GenerateMesh (OceanRenderer.cs:172-207; identical topology in FFTMesh.cs:101-139). Host pointers.
extern "C" int mw_mesh_generate(int device, int32_t resolution, float unit_width, float* vertices, float* normals, float* uvs,
                                int32_t* indices)
{
    if (resolution < 2 || resolution > 4096) { mw_set_error("mw_mesh_generate: resolution must be in [2, 4096], got %d", resolution); return MW_E_INVALID_ARG; }
    MW_CUDA(cudaSetDevice(device));
    const size_t n2 = (size_t)resolution * resolution;
    const size_t ni = (size_t)(resolution - 1) * (resolution - 1) * 6;
    float *dv = nullptr, *dn = nullptr, *du = nullptr;
    int* di = nullptr;
    auto cleanup = [&]() { if (dv) cudaFree(dv); if (dn) cudaFree(dn); if (du) cudaFree(du); if (di) cudaFree(di); };
#define MWM_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { mw_set_error("%s failed: %s", #expr, cudaGetErrorString(_e)); cleanup(); return _e == cudaErrorMemoryAllocation ? MW_E_OOM : MW_E_CUDA; } } while (0)
    if (vertices) MWM_TRY(cudaMalloc((void**)&dv, n2 * 12));
    if (normals) MWM_TRY(cudaMalloc((void**)&dn, n2 * 12));
    if (uvs) MWM_TRY(cudaMalloc((void**)&du, n2 * 8));
    if (indices) MWM_TRY(cudaMalloc((void**)&di, ni * 4));
    mwr::k_mesh_generate<<<(unsigned)((n2 + 255) / 256), 256>>>(dv, dn, du, di, resolution, unit_width);
    g_mw_launches.fetch_add(1, std::memory_order_relaxed);
    MWM_TRY(cudaGetLastError());
    if (vertices) MWM_TRY(cudaMemcpy(vertices, dv, n2 * 12, cudaMemcpyDeviceToHost));
    if (normals) MWM_TRY(cudaMemcpy(normals, dn, n2 * 12, cudaMemcpyDeviceToHost));
    if (uvs) MWM_TRY(cudaMemcpy(uvs, du, n2 * 8, cudaMemcpyDeviceToHost));
    if (indices) MWM_TRY(cudaMemcpy(indices, di, ni * 4, cudaMemcpyDeviceToHost));
    cleanup();
    return MW_OK;
}
