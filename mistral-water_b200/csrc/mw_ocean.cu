// mw_ocean.cu -- handle management and the C ABI of the ocean path (include/mistral_ocean.h).
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include <vector>

#include "mw_ocean_kernels.cuh"
#include "mw_direct_kernels.cuh"
#include "mw_cols_seam.cuh"

// ---------------------------------------------------------------------------------------------
// error text (thread-local) and global launch counter
// ---------------------------------------------------------------------------------------------
static thread_local char t_err[512] = "";
std::atomic<long long> g_mw_launches{0};

void mw_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof t_err, fmt, ap);
    va_end(ap);
}
extern "C" const char* mw_last_error(void) { return t_err; }
extern "C" int mw_version(void) { return MW_VERSION; }
extern "C" int64_t mw_kernel_launch_count(void) { return (int64_t)g_mw_launches.load(); }

#ifndef MW_SEAM_DEFAULT
#define MW_SEAM_DEFAULT 1
#endif

// ---------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------
struct EvPair { cudaEvent_t a, b; int kid; };

struct mw_ocean {
    mw_ocean_params p;
    int N = 0;
    int tiles = 1;
    size_t n2 = 0;
    bool device_ptrs = false;
    bool profile = false;
    bool have_h0 = false;
    bool direct = false;                  // direct-sum frame (non-periodic / non-power-of-two / N < 32 grids): mw_direct_kernels.cuh
    float2* dH = nullptr;                 // [tiles][N*N] htilde(t) of the frame (direct mode)
    bool host_async = false;              // MW_HOST_ASYNC: host-pointer calls do not wait
    cudaStream_t copy_stream = nullptr;   // device -> host result copies (host-pointer mode)
    cudaEvent_t ev_computed = nullptr, ev_copied = nullptr;
    bool copies_pending = false;
    float2* s_stage = nullptr;            // [2][tiles][N*N] upload staging of set_h0 (host-pointer mode)
    float timer = 0.f;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    // device state
    float4* spec = nullptr;   // [tiles][N*N] (h0, h0conj) as given (get_h0 / evolve_spectrum)
    float4* spec_r = nullptr; // [tiles][N*N] -(h0, h0conj) * ramp[n + m]: what the frame kernels read
    float* omega = nullptr;   // [N*N]
    int* qidx = nullptr;      // [N*N] omega / w0 (integer)
    float2* ptab = nullptr;   // [q_max + 1] per-frame (cos, sin)(omega_q t)
    int q_entries = 0;
    float2* ramp = nullptr;   // [2N]
    float* kd = nullptr;      // [N]
    float2* tw = nullptr;     // [N]
    float4* twimg = nullptr;  // shared-memory twiddle tables of the frame kernels, ready to copy (sign +1)
    float4* XAB = nullptr;    // [tiles][N/8][N][8] intermediate, fields A and B (16 B / point)
    float2* XC = nullptr;     // [tiles][N/16][N][16] intermediate, field C (8 B / point); lives right behind XAB
    // scratch outputs (host-pointer mode, or inputs of k_mesh_outputs)
    float* s_height = nullptr; float2* s_disp = nullptr; float* s_normal = nullptr; float* s_white = nullptr;
    float* s_jac = nullptr; float* s_vert = nullptr; float4* s_col = nullptr; float2* s_h = nullptr;
    int dbg_flags = 0;
    // tile-group pipelining: the frame is issued as groups of `group_tiles` tiles, alternating between two
    // streams and two slots of the intermediate buffer, so that (a) the intermediate of a group stays in the
    // 126 MB L2 between pass 1 and pass 2 and (b) pass 1 of one group overlaps pass 2 of the previous one
    int group_tiles = 1, x_tiles = 1, slots = 2;
    // programmatic dependent launch of the frame kernels (MW_PDL=0 off, 1 = successors released at CTA start, 2 = late)
    int pdl = 1;
    cudaStream_t aux_stream[3] = {nullptr, nullptr, nullptr};   // streams 1..slots-1 (stream 0 is the caller's)
    cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
    long long* dbg_rows = nullptr; long long* dbg_cols = nullptr;  // developer phase timing (mw_debug_phase_buffers)
    // CUDA graph of a single-group frame (k_phase_table -> pass 1 -> pass 2 [-> mesh outputs]): what one FFTMesh.Update() costs
    // is launch latency, not bandwidth -- a 64^2 or 256^2 frame is three back-to-back launches.  Captured the second time the
    // same output pointers are seen; per frame only the phase table's `t` argument is patched (MW_GRAPH=0 disables).
    struct FrameGraph {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        cudaGraphNode_t phase_node = nullptr;
        cudaKernelNodeParams phase_params = {};
        float2* a_ptab = nullptr; int a_entries = 0; float a_length = 0.f, a_t = 0.f;   // argument storage of k_phase_table
        void* args[4] = {nullptr, nullptr, nullptr, nullptr};
        void* key[8] = {};
        void* seen[8] = {};
        bool have_seen = false;
        int kernels = 0;
    } fg;
    bool graph_enabled = true;
    // small single frames (N <= 256, at most 2^18 grid points per call): pass 1 evaluates e^{i omega t} itself and the
    // k_phase_table launch disappears -- such a frame is a chain of launch latencies (MW_INLINE_PHASE=0 disables)
    bool inline_phase = false;
    // pass 2 without the halo line (mw_cols_seam.cuh): the east neighbour's column is handed over by the CTA that owns it
    bool use_seam = false;
    float2* seam = nullptr;          // [tiles][N / W][N]
    unsigned* seam_flags = nullptr;  // [tiles][N / W]
    // profiling
    std::vector<EvPair> ev_pool; size_t ev_used = 0;
    double k_ms[MW_KERNEL_COUNT] = {0, 0, 0};
    int64_t k_n[MW_KERNEL_COUNT] = {0, 0, 0};
};

static int drain_events(mw_ocean* o)
{
    if (o->ev_used == 0) return MW_OK;
    MW_CUDA(cudaStreamSynchronize(o->stream));
    for (size_t i = 0; i < o->ev_used; ++i) {
        float ms = 0.f;
        MW_CUDA(cudaEventElapsedTime(&ms, o->ev_pool[i].a, o->ev_pool[i].b));
        o->k_ms[o->ev_pool[i].kid] += ms;
        o->k_n[o->ev_pool[i].kid] += 1;
    }
    o->ev_used = 0;
    return MW_OK;
}

struct ProfScope {
    mw_ocean* o; EvPair* e = nullptr;
    ProfScope(mw_ocean* o_, int kid) : o(o_)
    {
        if (!o->profile) return;
        if (o->ev_used == o->ev_pool.size()) {
            if (o->ev_pool.size() >= 4096) { drain_events(o); }
            else {
                EvPair p; p.kid = 0;
                if (cudaEventCreate(&p.a) != cudaSuccess || cudaEventCreate(&p.b) != cudaSuccess) return;
                o->ev_pool.push_back(p);
            }
        }
        e = &o->ev_pool[o->ev_used++];
        e->kid = kid;
        cudaEventRecord(e->a, o->stream);
    }
    ~ProfScope() { if (e) cudaEventRecord(e->b, o->stream); }
};

template <class T>
static int ensure(T** p, size_t count)
{
    if (*p) return MW_OK;
    MW_CUDA(cudaMalloc((void**)p, count * sizeof(T)));
    return MW_OK;
}

static bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

extern "C" int mw_ocean_create(const mw_ocean_params* params, mw_ocean** out)
{
    if (!params || !out) { mw_set_error("mw_ocean_create: null argument"); return MW_E_INVALID_ARG; }
    *out = nullptr;
    const mw_ocean_params& p = *params;
    if (!(p.unit_width > 0.f) || !(p.length > 0.f) || !isfinite(p.unit_width) || !isfinite(p.length)) {
        mw_set_error("unit_width and length must be positive and finite");
        return MW_E_INVALID_ARG;
    }
    // The transform path needs the periodic case (length == resolution * unit_width: only then is the reference's direct sum a
    // DFT) on a power-of-two grid of 32..2048.  Small grids that are not (the FFT Mesh scene's own 12 x 12,
    // length 12.39) run the same sum directly on the GPU (mw_direct_kernels.cuh); large ones are refused.
    const float want_len = (float)p.resolution * p.unit_width;
    const bool periodic = fabsf(p.length - want_len) <= 1e-6f * fabsf(want_len);
    const bool fft_ok = is_pow2(p.resolution) && p.resolution >= 32 && p.resolution <= 2048 && periodic;
    const bool direct = !fft_ok;
    if (direct && (p.resolution < 2 || p.resolution > MW_DIRECT_MAX_RESOLUTION)) {
        if (!is_pow2(p.resolution) || p.resolution > 2048 || p.resolution < 2)
            mw_set_error("resolution must be a power of two in [32, 2048] (transform path) or any value in [2, %d] (direct-sum path), got %d",
                         MW_DIRECT_MAX_RESOLUTION, p.resolution);
        else
            mw_set_error("length (%g) must equal resolution * unit_width (%g) above resolution %d: only the periodic case is an FFT",
                         p.length, want_len, MW_DIRECT_MAX_RESOLUTION);
        return MW_E_INVALID_ARG;
    }
    if (p.tiles < 1 || p.tiles > 65535) { mw_set_error("tiles must be in [1, 65535], got %d", p.tiles); return MW_E_INVALID_ARG; }
    if (!(p.t_division != 0.f)) { mw_set_error("t_division must be non-zero"); return MW_E_INVALID_ARG; }
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
    mw_ocean* o = new (std::nothrow) mw_ocean();
    if (!o) { mw_set_error("out of host memory"); return MW_E_OOM; }
    o->p = p;
    o->N = p.resolution;
    o->tiles = p.tiles;
    o->n2 = (size_t)o->N * o->N;
    o->device_ptrs = (p.flags & MW_DEVICE_PTRS) != 0;
    o->profile = (p.flags & MW_PROFILE) != 0;
    o->host_async = (p.flags & MW_HOST_ASYNC) != 0 && !o->device_ptrs;
    o->direct = direct;
    const int N = o->N;
    int rc = MW_OK;
    auto fail = [&](int code) { mw_ocean_destroy(o); return code; };
    if (cudaStreamCreateWithFlags(&o->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        mw_set_error("cudaStreamCreate failed"); return fail(MW_E_CUDA);
    }
    o->stream = o->own_stream;
    if (!o->device_ptrs) {
        if (cudaStreamCreateWithFlags(&o->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&o->ev_computed, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&o->ev_copied, cudaEventDisableTiming) != cudaSuccess) {
            mw_set_error("stream/event creation failed"); return fail(MW_E_CUDA);
        }
    }
    if ((rc = ensure(&o->spec, o->n2 * o->tiles))) return fail(rc);
    if ((rc = ensure(&o->omega, o->n2))) return fail(rc);
    if ((rc = ensure(&o->qidx, o->n2))) return fail(rc);
    if (direct) {
        // direct-sum path: the spectrum as given, the per-frame htilde(t) image, omega; none of the transform's tables
        if ((rc = ensure(&o->dH, o->n2 * o->tiles))) return fail(rc);
        mwk::k_dispersion<<<(unsigned)((o->n2 + 255) / 256), 256, 0, o->stream>>>(o->omega, o->qidx, N, p.length);
        g_mw_launches.fetch_add(1);
        if (cudaStreamSynchronize(o->stream) != cudaSuccess) { mw_set_error("k_dispersion failed: %s", cudaGetErrorString(cudaGetLastError())); return fail(MW_E_CUDA); }
        *out = o;
        return MW_OK;
    }
    if ((rc = ensure(&o->spec_r, o->n2 * o->tiles))) return fail(rc);
    if ((rc = ensure(&o->ramp, (size_t)2 * N))) return fail(rc);
    if ((rc = ensure(&o->kd, (size_t)N))) return fail(rc);
    if ((rc = ensure(&o->tw, (size_t)N))) return fail(rc);
    if ((rc = ensure(&o->twimg, (size_t)mwfft::twiddle_image_bytes(N, mwk::fft_pts(N)) / 16))) return fail(rc);
    {
        // group size: keep one group's intermediate (24 B per point) around 32 MB
        long long gt = (32ll << 20) / (long long)(o->n2 * 24);
        if (const char* e = getenv("MW_GROUP_TILES")) gt = atoll(e);
        if (gt < 1) gt = 1;
        if (gt > o->tiles) gt = o->tiles;
        o->group_tiles = (int)gt;
        if (const char* e = getenv("MW_PDL")) o->pdl = atoi(e);
        if (const char* e = getenv("MW_GRAPH")) o->graph_enabled = atoi(e) != 0;
        o->inline_phase = o->N <= 256 && o->n2 * (size_t)o->tiles <= ((size_t)1 << 18);
        if (const char* e = getenv("MW_INLINE_PHASE")) o->inline_phase = o->inline_phase && atoi(e) != 0;
        // pass 2 without the halo line: measured faster from N = 512 up (profiles/r02_seam_sweep.jsonl: 16 x 1024^2 402 -> 383 us,
        // 64 x 512^2 394 -> 342 us, 4 x 2048^2 536 -> 515 us) and neutral-to-slower below (256^2 x 256 equal, single 64^2 / 128^2
        // frames + 0.3 - 1 us for the hand-over)
        o->use_seam = MW_SEAM_DEFAULT != 0 && N >= 512;
        if (const char* e = getenv("MW_SEAM")) o->use_seam = atoi(e) != 0;
        if (o->use_seam) {
            const size_t nab = (size_t)(N / mwk::slab_w(N));
            if ((rc = ensure(&o->seam, (size_t)o->tiles * nab * N))) return fail(rc);
            if ((rc = ensure(&o->seam_flags, (size_t)o->tiles * nab + 1))) return fail(rc);   // + the time-out counter
            if (cudaMemset(o->seam_flags, 0, ((size_t)o->tiles * nab + 1) * sizeof(unsigned)) != cudaSuccess) { mw_set_error("cudaMemset failed"); return fail(MW_E_CUDA); }
        }
        if (const char* e = getenv("MW_SLOTS")) o->slots = atoi(e);
        if (o->slots < 2) o->slots = 2;
        if (o->slots > 4) o->slots = 4;
        o->x_tiles = o->tiles <= o->group_tiles ? o->tiles : o->slots * o->group_tiles;
        char* x = nullptr;  // one allocation: 16 + 8 B per grid point of x_tiles tiles
        const size_t xab_bytes = mwk::xab_tile_elems(o->N) * sizeof(float4) * o->x_tiles;
        if ((rc = ensure(&x, xab_bytes + o->n2 * o->x_tiles * 8))) return fail(rc);
        o->XAB = reinterpret_cast<float4*>(x);
        o->XC = reinterpret_cast<float2*>(x + xab_bytes);
        bool ok = cudaEventCreateWithFlags(&o->ev_fork, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; i < o->slots - 1 && ok; ++i)
            ok = cudaStreamCreateWithFlags(&o->aux_stream[i], cudaStreamNonBlocking) == cudaSuccess &&
                 cudaEventCreateWithFlags(&o->ev_join[i], cudaEventDisableTiming) == cudaSuccess;
        if (!ok) { mw_set_error("stream/event creation failed"); return fail(MW_E_CUDA); }
    }

    // small host-built tables
    std::vector<float2> tw(N), ramp(2 * N);
    std::vector<float> kd(N);
    const double PI_D = 3.14159265358979323846;
    for (int x = 0; x < N; ++x) tw[x] = make_float2((float)cos(2.0 * PI_D * x / N), (float)sin(2.0 * PI_D * x / N));
    for (int s = 0; s < 2 * N; ++s) {
        // exp(i pi s (1 - N) / N), angle reduced exactly: s (N-1) mod 2N
        const long long u = ((long long)s * (N - 1)) % (2LL * N);
        ramp[s] = make_float2((float)cos(-PI_D * (double)u / N), (float)sin(-PI_D * (double)u / N));
    }
    for (int i = 0; i < N; ++i) {
        // FFTMesh.cs:201  kx = 2 * PI * (i - resolution / 2.0f) / length   (strict fp32, source order)
        volatile float two_pi = 2.0f * MW_PI_F;
        volatile float d = (float)i - (float)N / 2.0f;
        volatile float num = two_pi * d;
        kd[i] = num / p.length;
    }
    std::vector<float> twimg(mwfft::twiddle_image_bytes(N, mwk::fft_pts(N)) / 4);
    mwfft::twiddle_image_host(N, +1, twimg.data(), [&](int x, float& c, float& s) { c = tw[x].x; s = tw[x].y; }, mwk::fft_pts(N));
    if (cudaMemcpy(o->tw, tw.data(), N * sizeof(float2), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(o->twimg, twimg.data(), twimg.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(o->ramp, ramp.data(), 2 * N * sizeof(float2), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(o->kd, kd.data(), N * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
        mw_set_error("table upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(MW_E_CUDA);
    }
    mwk::k_dispersion<<<(unsigned)((o->n2 + 255) / 256), 256, 0, o->stream>>>(o->omega, o->qidx, N, p.length);
    g_mw_launches.fetch_add(1);
    {
        // |k| is largest at grid point (0, 0), so qidx[0] is the largest multiple of w0 on the grid
        int qmax = -1;
        if (cudaMemcpyAsync(&qmax, o->qidx, sizeof(int), cudaMemcpyDeviceToHost, o->stream) != cudaSuccess ||
            cudaStreamSynchronize(o->stream) != cudaSuccess) {
            mw_set_error("k_dispersion failed: %s", cudaGetErrorString(cudaGetLastError()));
            return fail(MW_E_CUDA);
        }
        if (qmax < 0 || qmax > (1 << 24)) {
            mw_set_error("dispersion table would need %d entries (length %g, resolution %d): unsupported", qmax, p.length, N);
            return fail(MW_E_INVALID_ARG);
        }
        o->q_entries = qmax + 1;
        if ((rc = ensure(&o->ptab, (size_t)o->q_entries))) return fail(rc);
    }
    *out = o;
    return MW_OK;
}

extern "C" void mw_ocean_destroy(mw_ocean* o)
{
    if (!o) return;
    cudaSetDevice(o->p.device);
    if (o->stream) cudaStreamSynchronize(o->stream);
    if (o->copy_stream) { cudaStreamSynchronize(o->copy_stream); cudaStreamDestroy(o->copy_stream); }
    if (o->ev_computed) cudaEventDestroy(o->ev_computed);
    if (o->ev_copied) cudaEventDestroy(o->ev_copied);
    for (auto& e : o->ev_pool) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    if (o->fg.exec) cudaGraphExecDestroy(o->fg.exec);
    if (o->fg.graph) cudaGraphDestroy(o->fg.graph);
    void* ptrs[] = {o->seam, o->seam_flags, o->dH, o->s_stage, o->spec, o->spec_r, o->qidx, o->ptab, o->twimg, o->omega, o->ramp, o->kd, o->tw, o->XAB, o->s_height, o->s_disp, o->s_normal,
                    o->s_white, o->s_jac, o->s_vert, o->s_col, o->s_h};
    for (void* q : ptrs) if (q) cudaFree(q);
    for (int i = 0; i < 3; ++i) {
        if (o->aux_stream[i]) { cudaStreamSynchronize(o->aux_stream[i]); cudaStreamDestroy(o->aux_stream[i]); }
        if (o->ev_join[i]) cudaEventDestroy(o->ev_join[i]);
    }
    if (o->ev_fork) cudaEventDestroy(o->ev_fork);
    if (o->own_stream) cudaStreamDestroy(o->own_stream);
    delete o;
}

#define MW_CHECK_HANDLE(o)                                                         \
    do {                                                                           \
        if (!(o)) { mw_set_error("null handle"); return MW_E_INVALID_ARG; }        \
        MW_CUDA(cudaSetDevice((o)->p.device));                                     \
    } while (0)

extern "C" int mw_ocean_sync(mw_ocean* o)
{
    MW_CHECK_HANDLE(o);
    MW_CUDA(cudaStreamSynchronize(o->stream));
    if (o->copy_stream) MW_CUDA(cudaStreamSynchronize(o->copy_stream));
    o->copies_pending = false;
    if (o->use_seam && o->seam_flags) {
        // a pass-2 CTA that gave up waiting for its neighbour's column (never expected: mw_cols_seam.cuh) left a wrong whitecap column
        unsigned n = 0;
        MW_CUDA(cudaMemcpy(&n, o->seam_flags + (size_t)o->tiles * (o->N / mwk::slab_w(o->N)), sizeof n, cudaMemcpyDeviceToHost));
        if (n) {
            cudaMemset(o->seam_flags + (size_t)o->tiles * (o->N / mwk::slab_w(o->N)), 0, sizeof n);
            mw_set_error("pass 2: %u neighbour-column hand-overs timed out; the whitecap of those slabs' last column is invalid", n);
            return MW_E_CUDA;
        }
    }
    return MW_OK;
}

extern "C" int mw_ocean_set_stream(mw_ocean* o, void* cuda_stream)
{
    MW_CHECK_HANDLE(o);
    MW_CUDA(cudaStreamSynchronize(o->stream));
    int rc = drain_events(o);
    if (rc) return rc;
    o->stream = cuda_stream ? (cudaStream_t)cuda_stream : o->own_stream;
    o->fg.have_seen = false;
    return MW_OK;
}

extern "C" int mw_ocean_init_spectrum(mw_ocean* o)
{
    MW_CHECK_HANDLE(o);
    const size_t total = o->n2 * o->tiles;
    mwk::k_init_spectrum<<<(unsigned)((total + 127) / 128), 128, 0, o->stream>>>(
        o->spec, o->N, o->tiles, o->p.length, o->p.amplitude, o->p.wind_x, o->p.wind_y, o->p.seed);
    MW_LAUNCH_CHECK();
    if (!o->direct) {
        mwk::k_ramp_spectrum<<<(unsigned)((total + 255) / 256), 256, 0, o->stream>>>(o->spec, o->ramp, o->spec_r, o->N, (int64_t)total);
        MW_LAUNCH_CHECK();
    }
    if (!o->device_ptrs) MW_CUDA(cudaStreamSynchronize(o->stream));
    o->have_h0 = true;
    return MW_OK;
}

extern "C" int mw_ocean_set_h0(mw_ocean* o, const float* h0, const float* h0conj)
{
    MW_CHECK_HANDLE(o);
    if (!h0 || !h0conj) { mw_set_error("mw_ocean_set_h0: null buffer"); return MW_E_INVALID_ARG; }
    const size_t total = o->n2 * o->tiles;
    const float2 *d0 = (const float2*)h0, *d1 = (const float2*)h0conj;
    if (!o->device_ptrs) {
        // stage on the device, then interleave (h0, h0conj) into the packed spectrum
        int rc = ensure(&o->s_stage, 2 * total);
        if (rc) return rc;
        float2* stage = o->s_stage;
        MW_CUDA(cudaMemcpyAsync(stage, h0, total * sizeof(float2), cudaMemcpyHostToDevice, o->stream));
        MW_CUDA(cudaMemcpyAsync(stage + total, h0conj, total * sizeof(float2), cudaMemcpyHostToDevice, o->stream));
        d0 = stage; d1 = stage + total;
    }
    mwk::k_pack_h0<<<(unsigned)((total + 255) / 256), 256, 0, o->stream>>>(o->spec, d0, d1, (int64_t)total);
    MW_LAUNCH_CHECK();
    if (!o->direct) {
        mwk::k_ramp_spectrum<<<(unsigned)((total + 255) / 256), 256, 0, o->stream>>>(o->spec, o->ramp, o->spec_r, o->N, (int64_t)total);
        MW_LAUNCH_CHECK();
    }
    if (!o->device_ptrs && !o->host_async) MW_CUDA(cudaStreamSynchronize(o->stream));
    o->have_h0 = true;
    return MW_OK;
}

extern "C" int mw_ocean_get_h0(mw_ocean* o, float* h0, float* h0conj)
{
    MW_CHECK_HANDLE(o);
    if (!h0 || !h0conj) { mw_set_error("mw_ocean_get_h0: null buffer"); return MW_E_INVALID_ARG; }
    if (!o->have_h0) { mw_set_error("h0 not initialised: call mw_ocean_init_spectrum or mw_ocean_set_h0 first"); return MW_E_STATE; }
    const size_t total = o->n2 * o->tiles;
    float2 *d0 = (float2*)h0, *d1 = (float2*)h0conj;
    float2* stage = nullptr;
    if (!o->device_ptrs) {
        MW_CUDA(cudaMallocAsync((void**)&stage, 2 * total * sizeof(float2), o->stream));
        d0 = stage; d1 = stage + total;
    }
    mwk::k_unpack_h0<<<(unsigned)((total + 255) / 256), 256, 0, o->stream>>>(o->spec, d0, d1, (int64_t)total);
    MW_LAUNCH_CHECK();
    if (!o->device_ptrs) {
        MW_CUDA(cudaMemcpyAsync(h0, d0, total * sizeof(float2), cudaMemcpyDeviceToHost, o->stream));
        MW_CUDA(cudaMemcpyAsync(h0conj, d1, total * sizeof(float2), cudaMemcpyDeviceToHost, o->stream));
        MW_CUDA(cudaFreeAsync(stage, o->stream));
        MW_CUDA(cudaStreamSynchronize(o->stream));
    }
    return MW_OK;
}

extern "C" int mw_ocean_get_rest_vertices(mw_ocean* o, float* xyz)
{
    if (!o || !xyz) { mw_set_error("null argument"); return MW_E_INVALID_ARG; }
    const int N = o->N;
    const float uw = o->p.unit_width;
    // FFTMesh.cs:104-112: the half-cell offset applies to even resolutions only
    for (int i = 0; i < N; ++i) {
        volatile float hp = (float)(i - N / 2) * uw;
        for (int j = 0; j < N; ++j) {
            volatile float vp = (float)(j - N / 2) * uw;
            volatile float off = (N % 2 == 0) ? uw / 2.0f : 0.0f;
            float* v = xyz + 3 * ((size_t)i * N + j);
            v[0] = hp + off; v[1] = 0.f; v[2] = vp + off;
        }
    }
    return MW_OK;
}

extern "C" int mw_ocean_get_dispersion(mw_ocean* o, float* omega)
{
    MW_CHECK_HANDLE(o);
    if (!omega) { mw_set_error("null buffer"); return MW_E_INVALID_ARG; }
    MW_CUDA(cudaMemcpyAsync(omega, o->omega, o->n2 * sizeof(float), cudaMemcpyDeviceToHost, o->stream));
    MW_CUDA(cudaStreamSynchronize(o->stream));
    return MW_OK;
}

extern "C" int mw_ocean_evolve_spectrum(mw_ocean* o, float t, float* htilde)
{
    MW_CHECK_HANDLE(o);
    if (!htilde) { mw_set_error("null buffer"); return MW_E_INVALID_ARG; }
    if (!o->have_h0) { mw_set_error("h0 not initialised"); return MW_E_STATE; }
    const size_t total = o->n2 * o->tiles;
    float2* d = (float2*)htilde;
    if (!o->device_ptrs) { int rc = ensure(&o->s_h, total); if (rc) return rc; d = o->s_h; }
    mwk::k_evolve<<<(unsigned)((total + 255) / 256), 256, 0, o->stream>>>(o->spec, o->omega, d, (int64_t)o->n2, o->tiles, t);
    MW_LAUNCH_CHECK();
    if (!o->device_ptrs) {
        MW_CUDA(cudaMemcpyAsync(htilde, d, total * sizeof(float2), cudaMemcpyDeviceToHost, o->stream));
        MW_CUDA(cudaStreamSynchronize(o->stream));
    }
    return MW_OK;
}

// ---------------------------------------------------------------------------------------------
// per-frame launches
// ---------------------------------------------------------------------------------------------
template <int N, int RP, int MINB, bool INLINE_PHASE>
static int launch_rows_impl(mw_ocean* o, const mwk::RowArgs& a, int ntiles, cudaStream_t st)
{
    constexpr int PTS = mwk::fft_pts(N);
    constexpr int threads = RP * 3 * (N / PTS);
    constexpr size_t smem = mwfft::Plan<N, PTS>::TW_BYTES + (size_t)RP * 3 * mwfft::line_pitch(N, 8) * sizeof(float4);
    static bool attr_done[64] = {};
    if (!attr_done[o->p.device]) {
        MW_CUDA(cudaFuncSetAttribute(mwk::k_spectrum_rows<N, RP, MINB, INLINE_PHASE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MW_CUDA(cudaFuncSetAttribute(mwk::k_spectrum_rows<N, RP, MINB, INLINE_PHASE>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     cudaSharedmemCarveoutMaxShared));
        attr_done[o->p.device] = true;
    }
    dim3 grid(N / 2 / RP, ntiles);
    ProfScope ps(o, 0);
    MW_CUDA(mw_launch(mwk::k_spectrum_rows<N, RP, MINB, INLINE_PHASE>, grid, threads, smem, st, o->pdl != 0, a));
    MW_LAUNCH_CHECK();
    return MW_OK;
}
template <int N, int RP, int MINB>
static int launch_rows(mw_ocean* o, const mwk::RowArgs& a, int ntiles, cudaStream_t st)
{
    if constexpr (N <= 256) {
        if (o->inline_phase) return launch_rows_impl<N, RP, MINB, true>(o, a, ntiles, st);
    }
    return launch_rows_impl<N, RP, MINB, false>(o, a, ntiles, st);
}

template <int N, int MINB, int OUTS>
static int launch_cols_seam(mw_ocean* o, mwk::ColArgs a, int ntiles, cudaStream_t st)
{
    constexpr int W = mwk::slab_w(N);
    constexpr int threads = W * (N / mwk::fft_pts(N));
    constexpr size_t smem = mwk::seam_smem_bytes<N>();
    static bool attr_done[64] = {};
    if (!attr_done[o->p.device]) {
        MW_CUDA(cudaFuncSetAttribute(mwk::k_cols_seam<N, MINB, OUTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MW_CUDA(cudaFuncSetAttribute(mwk::k_cols_seam<N, MINB, OUTS>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     cudaSharedmemCarveoutMaxShared));
        attr_done[o->p.device] = true;
    }
    const bool want_ab = a.disp || a.normal || a.whitecap || a.jacobian;
    a.ab_blocks = want_ab ? N / W : 0;
    const int c_blocks = a.height ? N / (4 * W) : 0;
    if (a.ab_blocks + c_blocks == 0) return MW_OK;
    a.seam = o->seam;
    a.seam_flags = o->seam_flags;
    a.seam_timeouts = o->seam_flags + (size_t)o->tiles * (N / W);
    dim3 grid(a.ab_blocks + c_blocks, ntiles);
    ProfScope ps(o, 1);
    MW_CUDA(mw_launch(mwk::k_cols_seam<N, MINB, OUTS>, grid, threads, smem, st, o->pdl != 0, a));
    MW_LAUNCH_CHECK();
    return MW_OK;
}

template <int N, int MINB, int OUTS>
static int launch_cols_outs(mw_ocean* o, mwk::ColArgs a, int ntiles, cudaStream_t st)
{
    if (o->use_seam) return launch_cols_seam<N, MINB, OUTS>(o, a, ntiles, st);
    constexpr int W = mwk::slab_w(N);
    constexpr int PTS = mwk::fft_pts(N);
    constexpr int threads = (W + 1) * (N / PTS);
    constexpr size_t smem = mwfft::Plan<N, PTS>::TW_BYTES + (size_t)(W + 1) * mwfft::line_pitch(N, W) * sizeof(float4) +
                            mwk::cols_stage_bytes(N, OUTS, threads);
    static bool attr_done[64] = {};
    if (!attr_done[o->p.device]) {
        MW_CUDA(cudaFuncSetAttribute(mwk::k_cols_extract<N, MINB, OUTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MW_CUDA(cudaFuncSetAttribute(mwk::k_cols_extract<N, MINB, OUTS>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     cudaSharedmemCarveoutMaxShared));
        attr_done[o->p.device] = true;
    }
    const bool want_ab = a.disp || a.normal || a.whitecap || a.jacobian;
    a.ab_blocks = want_ab ? N / W : 0;
    const int c_blocks = a.height ? N / (4 * W) : 0;
    if (a.ab_blocks + c_blocks == 0) return MW_OK;
    dim3 grid(a.ab_blocks + c_blocks, ntiles);
    ProfScope ps(o, 1);
    MW_CUDA(mw_launch(mwk::k_cols_extract<N, MINB, OUTS>, grid, threads, smem, st, o->pdl != 0, a));
    MW_LAUNCH_CHECK();
    return MW_OK;
}

template <int N, int MINB>
static int launch_cols(mw_ocean* o, const mwk::ColArgs& a, int ntiles, cudaStream_t st)
{
    // the two output sets the reference's frame asks for get straight-line kernels; anything else decides at run time
    const int outs = (a.disp ? 1 : 0) | (a.normal ? 2 : 0) | (a.whitecap ? 4 : 0) | (a.jacobian ? 8 : 0);
    if (outs == 7) return launch_cols_outs<N, MINB, 7>(o, a, ntiles, st);   // hds + normal + whitecap (EvaluateWaves)
    if (outs == 3) return launch_cols_outs<N, MINB, 3>(o, a, ntiles, st);   // hds + normal
    return launch_cols_outs<N, MINB, -1>(o, a, ntiles, st);
}

template <int N, int RP, int RMINB, int CMINB>
static int run_frame_n(mw_ocean* o, mwk::RowArgs ra, mwk::ColArgs ca)
{
    int rc;
    const int G = o->group_tiles;
    const int ngroups = (o->tiles + G - 1) / G;
    // MW_PROFILE handles stay on one stream so that the per-kernel event times are not overlapped
    const bool dual = ngroups > 1 && !o->profile;
    const int S = o->slots;
    if (dual) {
        MW_CUDA(cudaEventRecord(o->ev_fork, o->stream));
        for (int i = 0; i < S - 1; ++i) MW_CUDA(cudaStreamWaitEvent(o->aux_stream[i], o->ev_fork, 0));
    }
    float4* xab0 = o->XAB;
    float2* xc0 = o->XC;
    for (int gi = 0; gi < ngroups; ++gi) {
        const int t0 = gi * G;
        const int nt = o->tiles - t0 < G ? o->tiles - t0 : G;
        const int slot = ngroups > 1 ? (gi % S) : 0;
        cudaStream_t st = (dual && slot) ? o->aux_stream[slot - 1] : o->stream;
        ra.tile0 = ca.tile0 = t0;
        ra.XAB = xab0 + (size_t)slot * G * mwk::xab_tile_elems(o->N);
        ra.XC = xc0 + (size_t)slot * G * o->n2;
        ca.XAB = ra.XAB;
        ca.XC = ra.XC;
        if ((rc = launch_rows<N, RP, RMINB>(o, ra, nt, st))) return rc;
        if ((rc = launch_cols<N, CMINB>(o, ca, nt, st))) return rc;
    }
    if (dual) {
        for (int i = 0; i < S - 1; ++i) {
            MW_CUDA(cudaEventRecord(o->ev_join[i], o->aux_stream[i]));
            MW_CUDA(cudaStreamWaitEvent(o->stream, o->ev_join[i], 0));
        }
    }
    return MW_OK;
}

// pass-1 CTAs per SM asked of the compiler (__launch_bounds__ minimum blocks = the register cap) per resolution
#ifndef MW_ROWS_RP_256
#define MW_ROWS_RP_256 4
#endif
#ifndef MW_ROWS_MINB_256
#define MW_ROWS_MINB_256 3
#endif
#ifndef MW_ROWS_MINB_512
#define MW_ROWS_MINB_512 3
#endif
#ifndef MW_ROWS_MINB_2048
#define MW_ROWS_MINB_2048 2
#endif
static int run_frame(mw_ocean* o, const mwk::RowArgs& ra, const mwk::ColArgs& ca)
{
    switch (o->N) {
        case 32: return run_frame_n<32, 16, 1, 1>(o, ra, ca);
        case 64: return run_frame_n<64, 8, 1, 1>(o, ra, ca);
        case 128: return run_frame_n<128, 8, 1, 1>(o, ra, ca);
        case 256: return run_frame_n<256, MW_ROWS_RP_256, MW_ROWS_MINB_256, 1>(o, ra, ca);
        case 512: return run_frame_n<512, 2, MW_ROWS_MINB_512, 1>(o, ra, ca);
        case 1024: {
            static const int minb = getenv("MW_ROWS_MINB") ? atoi(getenv("MW_ROWS_MINB")) : 3;
            constexpr int CM = MW_SLABW_1024 == 4 ? 2 : 1;
#if MW_PTS_1024 == 32 && defined(MW_ROWS_RP2)
            return run_frame_n<1024, 2, 2, CM>(o, ra, ca);   // radix-32 engine: 2 row pairs (6 lines, 192 threads) per CTA
#else
            return minb == 3 ? run_frame_n<1024, 1, 3, CM>(o, ra, ca) : run_frame_n<1024, 1, 4, CM>(o, ra, ca);
#endif
        }
        case 2048: return run_frame_n<2048, 1, MW_ROWS_MINB_2048, 1>(o, ra, ca);
    }
    mw_set_error("unsupported resolution %d", o->N);
    return MW_E_INVALID_ARG;
}

// CUtensorMap of one output plane seen as a 2-D float tensor [tiles * N rows][N * comp floats], box = 256 rows x (8 * comp) floats
static int encode_plane(CUtensorMap* tm, void* base, int N, int tiles, int comp)
{
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        MW_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        if (!p || qres != cudaDriverEntryPointSuccess) { mw_set_error("cuTensorMapEncodeTiled not available"); return MW_E_CUDA; }
        fn = (EncodeFn)p;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)N * comp, (cuuint64_t)tiles * N};
    const cuuint64_t gstride[1] = {(cuuint64_t)N * comp * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)(mwk::slab_w(N) * comp), 256u};
    const cuuint32_t estride[2] = {1u, 1u};
    const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { mw_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return MW_E_CUDA; }
    return MW_OK;
}

extern "C" int mw_ocean_generate(mw_ocean* o, float t, const mw_ocean_out* out)
{
    MW_CHECK_HANDLE(o);
    if (!out) { mw_set_error("mw_ocean_generate: null output block"); return MW_E_INVALID_ARG; }
    if (!o->have_h0) { mw_set_error("h0 not initialised: call mw_ocean_init_spectrum or mw_ocean_set_h0 first"); return MW_E_STATE; }
    const size_t total = o->n2 * o->tiles;
    const bool dev = o->device_ptrs;
    const bool mesh = out->vertices || out->colors;
    int rc;

    // where each field is produced on the device
    float* d_height = nullptr; float2* d_disp = nullptr; float* d_normal = nullptr; float* d_white = nullptr; float* d_jac = nullptr;
    if (out->height || out->vertices) {
        if (dev && out->height) d_height = out->height;
        else { if ((rc = ensure(&o->s_height, total))) return rc; d_height = o->s_height; }
    }
    if (out->disp || out->vertices) {
        if (dev && out->disp) d_disp = (float2*)out->disp;
        else { if ((rc = ensure(&o->s_disp, total))) return rc; d_disp = o->s_disp; }
    }
    if (out->normal) {
        if (dev) d_normal = out->normal;
        else { if ((rc = ensure(&o->s_normal, total * 3))) return rc; d_normal = o->s_normal; }
    }
    if (out->whitecap || out->colors) {
        if (dev && out->whitecap) d_white = out->whitecap;
        else { if ((rc = ensure(&o->s_white, total))) return rc; d_white = o->s_white; }
    }
    if (out->jacobian) {
        if (dev) d_jac = out->jacobian;
        else { if ((rc = ensure(&o->s_jac, total))) return rc; d_jac = o->s_jac; }
    }

    // host-pointer mode: the scratch outputs of the previous frame may still be on their way to the host
    if (!dev && o->copies_pending) MW_CUDA(cudaStreamWaitEvent(o->stream, o->ev_copied, 0));
    if (o->direct) {
        // the whitecap needs hds and the normal whatever the caller asked for
        if ((d_white || d_jac) && !d_disp) { if ((rc = ensure(&o->s_disp, total))) return rc; d_disp = o->s_disp; }
        if (d_white && !d_normal) { if ((rc = ensure(&o->s_normal, total * 3))) return rc; d_normal = o->s_normal; }
        mwk::k_direct_htilde<<<(unsigned)((total + 255) / 256), 256, 0, o->stream>>>(o->spec, o->omega, o->dH, (int64_t)o->n2, o->tiles, t);
        MW_LAUNCH_CHECK();
        {
            ProfScope ps(o, 0);
            mwk::k_direct_displace<<<dim3((unsigned)o->n2, (unsigned)o->tiles), mwk::DIRECT_THREADS, 0, o->stream>>>(
                o->dH, d_height, d_disp, d_normal, o->N, o->p.unit_width, o->p.length);
            MW_LAUNCH_CHECK();
        }
        if (d_white || d_jac) {
            ProfScope ps(o, 1);
            mwk::k_direct_whitecap<<<(unsigned)((total + 255) / 256), 256, 0, o->stream>>>(d_disp, d_normal, d_white, d_jac, o->N, o->tiles);
            MW_LAUNCH_CHECK();
        }
    }

    float* d_vert = nullptr; float4* d_col = nullptr;
    if (mesh) {
        if (out->vertices) {
            if (dev) d_vert = out->vertices;
            else { if ((rc = ensure(&o->s_vert, total * 3))) return rc; d_vert = o->s_vert; }
        }
        if (out->colors) {
            if (dev) d_col = (float4*)out->colors;
            else { if ((rc = ensure(&o->s_col, total))) return rc; d_col = o->s_col; }
        }
    }
    // the kernels of one frame (transform path): phase table, the two passes, the optional mesh-facing epilogue
    auto issue_frame = [&](float tt) -> int {
        if (!o->direct) {
            // e^{i omega t} for every distinct omega of the grid (one entry per multiple of w0), then the frame
            if (!o->inline_phase) {
                mwk::k_phase_table<<<(unsigned)((o->q_entries + 255) / 256), 256, 0, o->stream>>>(o->ptab, o->q_entries, o->p.length, tt);
                MW_LAUNCH_CHECK();
            }
            mwk::RowArgs ra{o->spec_r, o->qidx, o->ptab, o->kd, o->twimg, o->XAB, o->XC, 0, o->dbg_rows, o->dbg_flags, o->pdl,
                            o->p.length, tt, o->use_seam ? o->seam_flags : nullptr, o->N / mwk::slab_w(o->N)};
            mwk::ColArgs ca{};
            ca.XAB = o->XAB; ca.XC = o->XC; ca.twimg = o->twimg; ca.height = d_height; ca.disp = d_disp; ca.normal = d_normal;
            ca.whitecap = d_white; ca.jacobian = d_jac; ca.dbg = o->dbg_cols; ca.dbg_flags = o->dbg_flags; ca.pdl = o->pdl;
            if (mwk::cols_tma_store(o->N, (d_disp ? 1 : 0) | (d_normal ? 2 : 0) | (d_white ? 4 : 0) | (d_jac ? 8 : 0))) {
                int r2;
                if ((r2 = encode_plane(&ca.tm_white, d_white, o->N, o->tiles, 1)) || (r2 = encode_plane(&ca.tm_disp, d_disp, o->N, o->tiles, 2)) ||
                    (r2 = encode_plane(&ca.tm_normal, d_normal, o->N, o->tiles, 3))) return r2;
            }
            int r3 = run_frame(o, ra, ca);
            if (r3) return r3;
        }
        if (mesh) {
            ProfScope ps(o, 2);
            mwk::k_mesh_outputs<<<(unsigned)((total + 255) / 256), 256, 0, o->stream>>>(
                d_height, d_disp, d_white, d_vert, d_col, o->N, o->tiles, o->p.unit_width, o->p.choppiness);
            MW_LAUNCH_CHECK();
        }
        return MW_OK;
    };
    // Single-group frames (one launch per kernel on one stream) are replayed from a CUDA graph: first sighting of a set of
    // output pointers runs normally (it also sets the kernels' attributes), the second is captured, later ones patch `t`.
    const bool graphable = o->graph_enabled && !o->inline_phase && !o->direct && !o->profile && o->tiles <= o->group_tiles && !o->dbg_rows && !o->dbg_cols &&
                           o->dbg_flags == 0;
    void* key[8] = {d_height, d_disp, d_normal, d_white, d_jac, d_vert, d_col, (void*)o->stream};
    mw_ocean::FrameGraph& fg = o->fg;
    if (graphable && fg.exec && memcmp(key, fg.key, sizeof key) == 0) {
        fg.a_t = t;
        MW_CUDA(cudaGraphExecKernelNodeSetParams(fg.exec, fg.phase_node, &fg.phase_params));
        MW_CUDA(cudaGraphLaunch(fg.exec, o->stream));
        g_mw_launches.fetch_add(fg.kernels, std::memory_order_relaxed);
    } else if (graphable && fg.have_seen && memcmp(key, fg.seen, sizeof key) == 0) {
        if (fg.exec) { cudaGraphExecDestroy(fg.exec); fg.exec = nullptr; }
        if (fg.graph) { cudaGraphDestroy(fg.graph); fg.graph = nullptr; }
        const long long before = g_mw_launches.load();
        MW_CUDA(cudaStreamBeginCapture(o->stream, cudaStreamCaptureModeThreadLocal));
        rc = issue_frame(t);
        cudaGraph_t g = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(o->stream, &g);
        bool ok = rc == MW_OK && ce == cudaSuccess && g != nullptr;
        if (ok) {
            // the phase-table node is the only kernel node with k_phase_table as its function
            size_t nn = 0;
            ok = cudaGraphGetNodes(g, nullptr, &nn) == cudaSuccess && nn > 0;
            std::vector<cudaGraphNode_t> nodes(nn);
            ok = ok && cudaGraphGetNodes(g, nodes.data(), &nn) == cudaSuccess;
            fg.phase_node = nullptr;
            for (size_t i = 0; ok && i < nn; ++i) {
                cudaGraphNodeType ty;
                if (cudaGraphNodeGetType(nodes[i], &ty) != cudaSuccess || ty != cudaGraphNodeTypeKernel) continue;
                cudaKernelNodeParams kp = {};
                if (cudaGraphKernelNodeGetParams(nodes[i], &kp) != cudaSuccess) continue;
                if (kp.func == (void*)mwk::k_phase_table) { fg.phase_node = nodes[i]; fg.phase_params = kp; }
            }
            ok = ok && fg.phase_node != nullptr;
        }
        if (ok) {
            fg.a_ptab = o->ptab; fg.a_entries = o->q_entries; fg.a_length = o->p.length; fg.a_t = t;
            fg.args[0] = &fg.a_ptab; fg.args[1] = &fg.a_entries; fg.args[2] = &fg.a_length; fg.args[3] = &fg.a_t;
            fg.phase_params.kernelParams = fg.args;
            fg.phase_params.extra = nullptr;
            ok = cudaGraphInstantiate(&fg.exec, g, 0) == cudaSuccess;
        }
        if (ok) {
            fg.graph = g;
            fg.kernels = (int)(g_mw_launches.load() - before);
            memcpy(fg.key, key, sizeof key);
            MW_CUDA(cudaGraphLaunch(fg.exec, o->stream));
        } else {
            // capture not possible here (e.g. the caller's stream is already being captured): run the frame the plain way
            (void)cudaGetLastError();
            if (g) cudaGraphDestroy(g);
            if (fg.exec) { cudaGraphExecDestroy(fg.exec); fg.exec = nullptr; }
            o->graph_enabled = false;
            g_mw_launches.store(before);
            if (rc) return rc;
            if ((rc = issue_frame(t))) return rc;
        }
    } else {
        if ((rc = issue_frame(t))) return rc;
        memcpy(fg.seen, key, sizeof key);
        fg.have_seen = graphable;
    }

    if (!dev) {
        // results leave on the copy stream: the main stream is free for the next frame's upload meanwhile
        MW_CUDA(cudaEventRecord(o->ev_computed, o->stream));
        MW_CUDA(cudaStreamWaitEvent(o->copy_stream, o->ev_computed, 0));
        auto d2h = [&](void* h, const void* d, size_t bytes) -> cudaError_t {
            return h ? cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, o->copy_stream) : cudaSuccess;
        };
        MW_CUDA(d2h(out->height, d_height, total * 4));
        MW_CUDA(d2h(out->disp, d_disp, total * 8));
        MW_CUDA(d2h(out->normal, d_normal, total * 12));
        MW_CUDA(d2h(out->whitecap, d_white, total * 4));
        MW_CUDA(d2h(out->jacobian, d_jac, total * 4));
        MW_CUDA(d2h(out->vertices, d_vert, total * 12));
        MW_CUDA(d2h(out->colors, d_col, total * 16));
        MW_CUDA(cudaEventRecord(o->ev_copied, o->copy_stream));
        o->copies_pending = true;
        if (!o->host_async) {
            MW_CUDA(cudaStreamSynchronize(o->copy_stream));
            o->copies_pending = false;
        }
    }
    return MW_OK;
}

extern "C" int mw_ocean_update(mw_ocean* o, float delta_time, const mw_ocean_out* out)
{
    if (!o) { mw_set_error("null handle"); return MW_E_INVALID_ARG; }
    o->timer = o->timer + delta_time / o->p.t_division;  // FFTMesh.cs:70
    return mw_ocean_generate(o, o->timer, out);          // FFTMesh.cs:72
}
extern "C" int mw_ocean_reset_timer(mw_ocean* o)
{
    if (!o) { mw_set_error("null handle"); return MW_E_INVALID_ARG; }
    o->timer = 0.f;  // FFTMesh.cs:64
    return MW_OK;
}
extern "C" float mw_ocean_timer(const mw_ocean* o) { return o ? o->timer : 0.f; }

extern "C" int mw_ocean_kernel_times(mw_ocean* o, float ms[MW_KERNEL_COUNT], int64_t launches[MW_KERNEL_COUNT], int reset)
{
    MW_CHECK_HANDLE(o);
    if (!o->profile) { mw_set_error("handle was not created with MW_PROFILE"); return MW_E_STATE; }
    int rc = drain_events(o);
    if (rc) return rc;
    for (int k = 0; k < MW_KERNEL_COUNT; ++k) {
        if (ms) ms[k] = (float)o->k_ms[k];
        if (launches) launches[k] = o->k_n[k];
        if (reset) { o->k_ms[k] = 0; o->k_n[k] = 0; }
    }
    return MW_OK;
}

// Developer hooks (not in the public header; live only in -DMW_DEVHOOKS=1 builds): device buffers that receive 8 clock64 stamps
// per CTA, and the "switch one phase off" flags of tools/phase_timing.py.
extern "C" __attribute__((visibility("default"))) int mw_debug_phase_buffers(mw_ocean* o, long long* rows, long long* cols)
{
    if (!MW_DEVHOOKS) { mw_set_error("this library was built without developer hooks (-DMW_DEVHOOKS=1)"); return MW_E_STATE; }
    if (!o) return MW_E_INVALID_ARG;
    o->dbg_rows = rows;
    o->dbg_cols = cols;
    return MW_OK;
}
extern "C" __attribute__((visibility("default"))) int mw_debug_flags(mw_ocean* o, int flags)
{
    if (!MW_DEVHOOKS) { mw_set_error("this library was built without developer hooks (-DMW_DEVHOOKS=1)"); return MW_E_STATE; }
    if (!o) return MW_E_INVALID_ARG;
    o->dbg_flags = flags;
    return MW_OK;
}
