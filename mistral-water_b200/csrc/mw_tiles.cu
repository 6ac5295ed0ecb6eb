// mw_tiles.cu -- multi-GPU tile sets behind the C ABI (include/mistral_ocean.h, "Multi-GPU tile sets").
//
// SURVEY.md section 8e / BASELINE config 5: independent ocean tiles, one rank per GPU, ONE collective -- the in-place
// all-gather of the final float buffers.  The reference has no counterpart (Scripts/FFTMesh.cs runs one mesh on one device;
// nothing in EvaluateWaves :224-280 couples two meshes), so everything here is this engine's own design:
//
//   * a rank = one mw_ocean handle (device pointers) writing straight into its slot of a double-buffered gather buffer;
//   * the gather of frame k runs on communication streams under the generation of frame k + 1;
//   * MW_GATHER_NCCL: ncclAllGather in place (libnccl.so.2 resolved with dlopen: the library carries no link-time
//     dependency on NCCL and reports MW_E_NCCL when it is absent);
//   * MW_GATHER_PEER: every rank pushes its slot into the peers' buffers over NVLink (peer access inside a process, CUDA IPC
//     mappings between processes) -- k_push_slots_bulk (TMA bulk copies, default), k_push_slots (SM stores) or one copy-engine
//     transfer per peer.  Between processes the ranks are fenced by stream memory operations on flag words that live in the
//     exported allocation: "my buffer b is free" (cuStreamWriteValue32 into every peer) before a push may start
//     (cuStreamWaitValue32 on the local copy), "my slot has landed" after it.  No host round trip.
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>  // types and prototypes only: every NCCL entry point is resolved at run time
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <new>
#include <vector>

#include "mw_common.cuh"

// ---------------------------------------------------------------------------------------------
// NCCL, loaded on first use
// ---------------------------------------------------------------------------------------------
namespace {

struct NcclApi {
    void* lib = nullptr;
    bool ok = false;
    char why[256] = "";
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclCommGetAsyncError) CommGetAsyncError = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
};

NcclApi* nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return &api;
    tried = true;
    const char* names[] = {getenv("MW_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (!n || !*n) continue;
        api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) {
        snprintf(api.why, sizeof api.why, "libnccl.so.2 could not be loaded (%s)", dlerror());
        return &api;
    }
    bool all = true;
#define MW_NCCL_SYM(field, name)                                              \
    do {                                                                      \
        api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, name)); \
        if (!api.field) { all = false; snprintf(api.why, sizeof api.why, "symbol %s missing in libnccl", name); } \
    } while (0)
    MW_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    MW_NCCL_SYM(CommInitRank, "ncclCommInitRank");
    MW_NCCL_SYM(CommInitAll, "ncclCommInitAll");
    MW_NCCL_SYM(CommDestroy, "ncclCommDestroy");
    MW_NCCL_SYM(AllGather, "ncclAllGather");
    MW_NCCL_SYM(GroupStart, "ncclGroupStart");
    MW_NCCL_SYM(GroupEnd, "ncclGroupEnd");
    MW_NCCL_SYM(GetErrorString, "ncclGetErrorString");
    MW_NCCL_SYM(CommGetAsyncError, "ncclCommGetAsyncError");
    MW_NCCL_SYM(GetVersion, "ncclGetVersion");
#undef MW_NCCL_SYM
    api.ok = all;
    return &api;
}

#define MW_NCCL(expr)                                                                                   \
    do {                                                                                                \
        ncclResult_t _r = (expr);                                                                       \
        if (_r != ncclSuccess) {                                                                        \
            mw_set_error("%s failed: %s (%s:%d)", #expr, nccl_api()->GetErrorString(_r), __FILE__, __LINE__); \
            return MW_E_NCCL;                                                                           \
        }                                                                                               \
    } while (0)

// ---------------------------------------------------------------------------------------------
// stream memory operations (driver API, resolved through the runtime)
// ---------------------------------------------------------------------------------------------
typedef CUresult (*MemOp32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
struct MemOps {
    MemOp32Fn write32 = nullptr, wait32 = nullptr;
    bool ok = false;
};
MemOps* memops()
{
    static MemOps m;
    static bool tried = false;
    if (tried) return &m;
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &q) == cudaSuccess && p && q == cudaDriverEntryPointSuccess)
        m.write32 = (MemOp32Fn)p;
    p = nullptr;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess && p && q == cudaDriverEntryPointSuccess)
        m.wait32 = (MemOp32Fn)p;
    m.ok = m.write32 && m.wait32;
    return &m;
}
#define MW_CU(expr)                                                                      \
    do {                                                                                 \
        CUresult _r = (expr);                                                            \
        if (_r != CUDA_SUCCESS) {                                                        \
            mw_set_error("%s failed: CUresult %d (%s:%d)", #expr, (int)_r, __FILE__, __LINE__); \
            return MW_E_CUDA;                                                            \
        }                                                                                \
    } while (0)

// Fallback for drivers that refuse stream memory operations on peer (IPC-mapped) addresses: a one-thread kernel that stores
// the flag with system scope.  Chosen at connect time if the first cuStreamWriteValue32 on a peer address fails.
__global__ void k_flag_write(uint32_t* flag, uint32_t value)
{
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t*>(flag) = value;
    __threadfence_system();
}

// SM-side push (MW_TILES_PUSH=sm): one kernel reads this rank's slot once and stores it into every peer's gather buffer through
// the peer mappings -- (world - 1) 16-byte stores per 16-byte load, a few CTAs wide, running beside the frame kernels.  The
// alternative to one copy-engine transfer per peer, which reaches only 370-420 GB/s on this pool once several flows run at once.
struct PushArgs {
    const float4* src;
    float4* dst[MW_TILES_MAX_WORLD];
    int npeers;
    unsigned long long n16;   // 16-byte elements per slot
};
__global__ void __launch_bounds__(512) k_push_slots(const __grid_constant__ PushArgs a)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < a.n16; i += 4 * stride) {
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = ldg_stream4(a.src + i + k * stride);
        for (int p = 0; p < a.npeers; ++p) {
#pragma unroll
            for (int k = 0; k < 4; ++k) a.dst[p][i + k * stride] = v[k];
        }
    }
    for (; i < a.n16; i += stride) {
        const float4 v = ldg_stream4(a.src + i);
        for (int p = 0; p < a.npeers; ++p) a.dst[p][i] = v;
    }
}


// Bulk-copy push (MW_TILES_PUSH=tma): the same transfer driven by the TMA unit instead of by load/store instructions.  One
// thread per CTA runs a ring of PUSH_STAGES shared-memory stages: cp.async.bulk global -> shared of a chunk of this rank's
// slot (mbarrier completion), then (world - 1) cp.async.bulk shared -> peer global of the same chunk as one bulk group; a stage is
// reloaded as soon as the group that read it has drained (wait_group.read), so loads, the stores of the previous chunk and
// the fabric's own queues overlap.  Chunks are dealt to the CTAs round-robin: at any moment the grid works on one contiguous
// window of the slot.  No register, LSU or issue-slot traffic on the SMs that host it.
constexpr int PUSH_STAGES = 4;
struct PushBulkArgs {
    const char* src;
    char* dst[MW_TILES_MAX_WORLD];
    int npeers;
    unsigned chunk;             // bytes per chunk (multiple of 16)
    unsigned long long bytes;   // bytes per slot (multiple of 16)
};
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__global__ void __launch_bounds__(32) k_push_slots_bulk(const __grid_constant__ PushBulkArgs a)
{
    extern __shared__ __align__(128) unsigned char stage_mem[];
    __shared__ uint64_t full[PUSH_STAGES];
    if (threadIdx.x != 0) return;
#pragma unroll
    for (int s = 0; s < PUSH_STAGES; ++s) mbar_init(&full[s], 1);
    const unsigned long long nchunks = (a.bytes + a.chunk - 1) / a.chunk;
    const unsigned long long mine = nchunks > blockIdx.x ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto chunk_off = [&](unsigned long long k) { return (blockIdx.x + k * gridDim.x) * (unsigned long long)a.chunk; };
    auto chunk_len = [&](unsigned long long off) { return (unsigned)((a.bytes - off) < a.chunk ? (a.bytes - off) : a.chunk); };
    auto load = [&](unsigned long long k) {
        const int s = (int)(k % PUSH_STAGES);
        const unsigned long long off = chunk_off(k);
        const unsigned len = chunk_len(off);
        mbar_expect_tx(&full[s], len);
        bulk_g2s(stage_mem + (size_t)s * a.chunk, a.src + off, len, &full[s]);
    };
    for (unsigned long long k = 0; k < PUSH_STAGES && k < mine; ++k) load(k);
    for (unsigned long long k = 0; k < mine; ++k) {
        const int s = (int)(k % PUSH_STAGES);
        mbar_wait(&full[s], (unsigned)((k / PUSH_STAGES) & 1));
        const unsigned long long off = chunk_off(k);
        const unsigned len = chunk_len(off);
        for (int p = 0; p < a.npeers; ++p) bulk_s2g(a.dst[p] + off, stage_mem + (size_t)s * a.chunk, len);
        bulk_commit();
        if (k >= 1 && k - 1 + PUSH_STAGES < mine) {
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // chunk k-1's stores have read their stage
            load(k - 1 + PUSH_STAGES);
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // every store of this CTA is complete, not only read
    __threadfence_system();
}

constexpr int MAXW = MW_TILES_MAX_WORLD;
enum { PUSH_CE = 0, PUSH_SM = 1, PUSH_TMA = 2 };
constexpr int AUTO_PEER_MAX_WORLD = MW_TILES_MAX_WORLD;   // MW_GATHER_AUTO: the peer pushes up to this many ranks, ncclAllGather above
constexpr uint32_t BLOB_MAGIC = 0x4d575432u;  // "MWT2"
// flag words of one rank (uint32), written by its peers through their mappings
struct FlagWords {
    uint32_t free_[2][MAXW];   // free_[b][p]   >= s : rank p's buffer b may be overwritten by gather number s
    uint32_t landed[2][MAXW];  // landed[b][p]  >= s : rank p's slot of gather number s has landed in my buffer b
    uint32_t hello[MAXW];      // connect-time handshake
};

struct Blob {  // what mw_tiles_export writes; MW_TILES_BLOB_BYTES on the wire
    uint32_t magic, rank, world, device;
    uint64_t alloc_bytes, flags_off, slot_floats;
    uint32_t gather, tiles_per_rank, resolution, has_nccl_id;
    cudaIpcMemHandle_t mem;
    ncclUniqueId nccl_id;
};
static_assert(sizeof(Blob) <= MW_TILES_BLOB_BYTES, "blob too large");

struct TileRank {
    int rank = -1, device = -1;
    mw_ocean* ocean = nullptr;
    cudaStream_t s_user_own = nullptr, s_user = nullptr, s_gen = nullptr, s_comm = nullptr;
    cudaStream_t s_flag = nullptr;   // "my buffer is free" announcements: never queued behind a gather in progress
    cudaStream_t s_push[MAXW] = {};
    char* alloc = nullptr;           // [2][world][slot] floats + FlagWords
    float* gather[2] = {nullptr, nullptr};
    FlagWords* flags = nullptr;
    // peers as seen from this rank's device (peer mode)
    void* peer_base[MAXW] = {};      // IPC mappings to close (multi-process)
    float* peer_gather[2][MAXW] = {};
    FlagWords* peer_flags[MAXW] = {};
    cudaEvent_t ev_user = nullptr, ev_gen[2] = {}, ev_comm[2] = {}, ev_free[2] = {}, ev_push[MAXW] = {};
    bool gen_used[2] = {false, false}, comm_used[2] = {false, false};
    ncclComm_t comm = nullptr;
};

}  // namespace

struct mw_tiles {
    mw_tiles_params p;
    int N = 0, world = 1, tpr = 1, nlocal = 1, impl = MW_GATHER_NCCL;
    bool single = true, connected = false, async = false;
    bool kernel_flags = false;   // peer flag writes by k_flag_write instead of cuStreamWriteValue32
    int push_lanes = 0;          // MW_TILES_PUSH_LANES: 0 = one copy stream per peer; k > 0 = the pushes share k streams
    int push_mode = PUSH_TMA;    // how a rank's slot reaches its peers: mw_tiles_params.flags, or MW_TILES_PUSH=ce|sm|tma
    int push_ctas = 0;           // CTAs of the push kernel (MW_TILES_PUSH_CTAS; default by world, see mw_tiles_create)
    unsigned push_chunk = 16384; // MW_TILES_PUSH_CHUNK: bytes per bulk-copy chunk (tma)
    size_t n2 = 0, slot_floats = 0, alloc_bytes = 0, flags_off = 0;
    std::vector<TileRank> ranks;
    uint32_t frames = 0;     // frames generated so far
    uint32_t gathers = 0;    // gathers enqueued so far (sequence number of the flag protocol)
    int buf_of_frame[2] = {0, 0};  // [0] latest frame's buffer, [1] the one before
};

namespace {

void rotate_wind(float wx, float wy, float deg, float* ox, float* oy)
{
    const double a = (double)deg * 3.14159265358979323846 / 180.0;
    const double c = cos(a), s = sin(a);
    *ox = (float)(c * wx - s * wy);
    *oy = (float)(s * wx + c * wy);
}

int make_event(cudaEvent_t* e)
{
    MW_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    return MW_OK;
}

int create_rank(mw_tiles* t, TileRank& r, int rank, int device)
{
    r.rank = rank;
    r.device = device;
    MW_CUDA(cudaSetDevice(device));
    mw_ocean_params op = t->p.ocean;
    op.device = device;
    op.tiles = t->tpr;
    op.seed = t->p.ocean.seed + (uint64_t)rank * (uint64_t)t->tpr;
    op.flags = MW_DEVICE_PTRS | (t->p.ocean.flags & MW_PROFILE);
    rotate_wind(t->p.ocean.wind_x, t->p.ocean.wind_y, t->p.wind_step_deg * (float)(rank * t->tpr), &op.wind_x, &op.wind_y);
    int rc = mw_ocean_create(&op, &r.ocean);
    if (rc) return rc;
    MW_CUDA(cudaStreamCreateWithFlags(&r.s_user_own, cudaStreamNonBlocking));
    MW_CUDA(cudaStreamCreateWithFlags(&r.s_gen, cudaStreamNonBlocking));
    MW_CUDA(cudaStreamCreateWithFlags(&r.s_comm, cudaStreamNonBlocking));
    MW_CUDA(cudaStreamCreateWithFlags(&r.s_flag, cudaStreamNonBlocking));
    r.s_user = r.s_user_own;
    if ((rc = mw_ocean_set_stream(r.ocean, r.s_gen))) return rc;
    int prio_least = 0, prio_greatest = 0;
    MW_CUDA(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    const char* pe = getenv("MW_TILES_PUSH_PRIO");   // developer knob: 0 = default priority
    const int push_prio = (pe && atoi(pe) == 0) ? 0 : prio_greatest;
    for (int p = 0; p < t->world; ++p) {
        if (p == rank) continue;
        // the gather is the critical path of a multi-GPU step: its copies / push kernels go ahead of queued frame kernels
        MW_CUDA(cudaStreamCreateWithPriority(&r.s_push[p], cudaStreamNonBlocking, push_prio));
        if ((rc = make_event(&r.ev_push[p]))) return rc;
    }
    if ((rc = make_event(&r.ev_user))) return rc;
    for (int b = 0; b < 2; ++b)
        if ((rc = make_event(&r.ev_gen[b])) || (rc = make_event(&r.ev_comm[b])) || (rc = make_event(&r.ev_free[b]))) return rc;
    MW_CUDA(cudaFuncSetAttribute(k_push_slots_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(PUSH_STAGES * t->push_chunk)));
    MW_CUDA(cudaMalloc((void**)&r.alloc, t->alloc_bytes));
    MW_CUDA(cudaMemset(r.alloc + t->flags_off, 0, sizeof(FlagWords)));
    r.gather[0] = reinterpret_cast<float*>(r.alloc);
    r.gather[1] = r.gather[0] + (size_t)t->world * t->slot_floats;
    r.flags = reinterpret_cast<FlagWords*>(r.alloc + t->flags_off);
    return MW_OK;
}

void destroy_rank(mw_tiles* t, TileRank& r)
{
    if (r.device < 0) return;
    cudaSetDevice(r.device);
    cudaStream_t all[] = {r.s_gen, r.s_comm, r.s_flag, r.s_user_own};
    for (cudaStream_t s : all) if (s) cudaStreamSynchronize(s);
    for (int p = 0; p < MAXW; ++p) if (r.s_push[p]) cudaStreamSynchronize(r.s_push[p]);
    if (r.comm && nccl_api()->ok) nccl_api()->CommDestroy(r.comm);
    if (!t->single)
        for (int p = 0; p < MAXW; ++p) if (r.peer_base[p]) cudaIpcCloseMemHandle(r.peer_base[p]);
    if (r.ocean) mw_ocean_destroy(r.ocean);
    cudaSetDevice(r.device);
    for (int p = 0; p < MAXW; ++p) {
        if (r.s_push[p]) cudaStreamDestroy(r.s_push[p]);
        if (r.ev_push[p]) cudaEventDestroy(r.ev_push[p]);
    }
    for (cudaStream_t s : all) if (s) cudaStreamDestroy(s);
    if (r.ev_user) cudaEventDestroy(r.ev_user);
    for (int b = 0; b < 2; ++b) {
        if (r.ev_gen[b]) cudaEventDestroy(r.ev_gen[b]);
        if (r.ev_comm[b]) cudaEventDestroy(r.ev_comm[b]);
        if (r.ev_free[b]) cudaEventDestroy(r.ev_free[b]);
    }
    if (r.alloc) cudaFree(r.alloc);
    r.device = -1;
}

// cudaDeviceEnablePeerAccess once per ordered pair and process: a second tile set must not raise
// cudaErrorPeerAccessAlreadyEnabled again (harmless, but every API error shows up in memcheck reports); someone else -- NCCL,
// the host -- may still have enabled the pair before us.
int enable_peer_once(int from, int to)
{
    static bool enabled[64][64] = {};
    const bool tracked = from >= 0 && from < 64 && to >= 0 && to < 64;
    if (tracked && enabled[from][to]) return MW_OK;
    MW_CUDA(cudaSetDevice(from));
    const cudaError_t e = cudaDeviceEnablePeerAccess(to, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        mw_set_error("cudaDeviceEnablePeerAccess(%d -> %d): %s", from, to, cudaGetErrorString(e));
        return MW_E_CUDA;
    }
    (void)cudaGetLastError();
    if (tracked) enabled[from][to] = true;
    return MW_OK;
}

// all local ranks in one process: peer access between every pair of devices
int connect_single(mw_tiles* t)
{
    if (t->impl == MW_GATHER_PEER) {
        for (auto& a : t->ranks)
            for (auto& b : t->ranks) {
                if (a.rank == b.rank) continue;
                if (a.device != b.device) {
                    int can = 0;
                    MW_CUDA(cudaDeviceCanAccessPeer(&can, a.device, b.device));
                    if (!can) { mw_set_error("no peer access from device %d to device %d", a.device, b.device); return MW_E_CUDA; }
                    { int rc = enable_peer_once(a.device, b.device); if (rc) return rc; }
                }
                for (int k = 0; k < 2; ++k) a.peer_gather[k][b.rank] = b.gather[k];
                a.peer_flags[b.rank] = b.flags;
            }
    } else {
        NcclApi* n = nccl_api();
        if (!n->ok) { mw_set_error("MW_GATHER_NCCL: %s", n->why); return MW_E_NCCL; }
        std::vector<ncclComm_t> comms(t->world);
        std::vector<int> devs(t->world);
        for (int i = 0; i < t->world; ++i) devs[i] = t->ranks[i].device;
        MW_NCCL(n->CommInitAll(comms.data(), t->world, devs.data()));
        for (int i = 0; i < t->world; ++i) t->ranks[i].comm = comms[i];
    }
    t->connected = true;
    return MW_OK;
}

// the copy stream a push to peer p (the j-th peer visited) goes on
cudaStream_t push_stream(mw_tiles* t, TileRank& r, int p, int j)
{
    if (t->push_lanes <= 0) return r.s_push[p];
    const int lane = (j - 1) % t->push_lanes;          // j = 1 .. world - 1
    return r.s_push[(r.rank + 1 + lane) % t->world];   // reuse the streams of the first `lanes` peers
}

// "flag = value" in stream order, visible system-wide after everything the stream did before
int flag_write(mw_tiles* t, cudaStream_t s, uint32_t* flag, uint32_t value)
{
    if (!t->kernel_flags) {
        MW_CU(memops()->write32((CUstream)s, (CUdeviceptr)flag, value, 0));
        return MW_OK;
    }
    k_flag_write<<<1, 1, 0, s>>>(flag, value);
    MW_LAUNCH_CHECK();
    return MW_OK;
}

// wait for a stream with a deadline; false on timeout
bool stream_done_within(cudaStream_t s, double seconds)
{
    timespec t0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (;;) {
        const cudaError_t e = cudaStreamQuery(s);
        if (e == cudaSuccess) return true;
        if (e != cudaErrorNotReady) return false;
        timespec t1;
        clock_gettime(CLOCK_MONOTONIC, &t1);
        if ((t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec) > seconds) return false;
        timespec nap = {0, 200000};
        nanosleep(&nap, nullptr);
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" int mw_tiles_create(const mw_tiles_params* params, mw_tiles** out)
{
    if (!params || !out) { mw_set_error("mw_tiles_create: null argument"); return MW_E_INVALID_ARG; }
    *out = nullptr;
    const mw_tiles_params& p = *params;
    if (p.world < 1 || p.world > MAXW) { mw_set_error("world must be in [1, %d], got %d", MAXW, p.world); return MW_E_INVALID_ARG; }
    if (p.rank < -1 || p.rank >= p.world) { mw_set_error("rank %d out of range for world %d", p.rank, p.world); return MW_E_INVALID_ARG; }
    if (p.tiles_per_rank < 1) { mw_set_error("tiles_per_rank must be >= 1"); return MW_E_INVALID_ARG; }
    if (p.gather != MW_GATHER_NCCL && p.gather != MW_GATHER_PEER && p.gather != MW_GATHER_AUTO) {
        mw_set_error("gather must be MW_GATHER_NCCL, MW_GATHER_PEER or MW_GATHER_AUTO");
        return MW_E_INVALID_ARG;
    }
    // MW_GATHER_AUTO: see include/mistral_ocean.h -- every rank of a node resolves it the same way
    const bool peer_usable = p.rank < 0 || memops()->ok;   // between processes the flag protocol needs stream memory operations
    const int gather = p.gather != MW_GATHER_AUTO ? p.gather
                     : ((peer_usable && p.world <= AUTO_PEER_MAX_WORLD) || !nccl_api()->ok) ? MW_GATHER_PEER : MW_GATHER_NCCL;
    if (gather == MW_GATHER_NCCL && p.world > 1 && !nccl_api()->ok) { mw_set_error("MW_GATHER_NCCL: %s", nccl_api()->why); return MW_E_NCCL; }
    if (gather == MW_GATHER_PEER && p.rank >= 0 && p.world > 1 && !memops()->ok) {
        mw_set_error("MW_GATHER_PEER between processes needs cuStreamWriteValue32 / cuStreamWaitValue32");
        return MW_E_CUDA;
    }
    mw_tiles* t = new (std::nothrow) mw_tiles();
    if (!t) { mw_set_error("out of host memory"); return MW_E_OOM; }
    t->p = p;
    t->N = p.ocean.resolution;
    t->world = p.world;
    t->tpr = p.tiles_per_rank;
    t->single = p.rank < 0;
    t->nlocal = t->single ? p.world : 1;
    t->impl = gather;
    t->async = (p.flags & MW_TILES_ASYNC) != 0;
    if (const char* e = getenv("MW_TILES_PUSH_LANES")) t->push_lanes = atoi(e);   // (experiment knob, tools/gather_probe.py)
    if (t->push_lanes > p.world - 1) t->push_lanes = p.world - 1;
    t->push_mode = (p.flags & MW_TILES_PUSH_CE) ? PUSH_CE : (p.flags & MW_TILES_PUSH_SM) ? PUSH_SM : PUSH_TMA;
    // developer overrides (tools/gather_probe.py, tools/push_call.sh)
    if (const char* e = getenv("MW_TILES_PUSH")) t->push_mode = !strcmp(e, "sm") ? PUSH_SM : !strcmp(e, "tma") ? PUSH_TMA : PUSH_CE;
    // Enough CTAs to keep the links busy and no more: every SM that hosts a push CTA is lost to pass 2 (whose CTAs take a
    // whole register file).  Measured on B200 / NVSwitch (profiles/r02_push_probe_*.jsonl).
    t->push_ctas = t->push_mode == PUSH_SM ? 32 : (p.world <= 2 ? 32 : 64);
    if (const char* e = getenv("MW_TILES_PUSH_CTAS")) if (atoi(e) > 0) t->push_ctas = atoi(e);
    if (const char* e = getenv("MW_TILES_PUSH_CHUNK")) {
        const int c = atoi(e) & ~15;
        if (c >= 1024 && c * PUSH_STAGES <= 200 * 1024) t->push_chunk = (unsigned)c;
    }
    t->n2 = (size_t)t->N * t->N;
    t->slot_floats = (size_t)t->tpr * t->n2 * 7;
    t->flags_off = ((size_t)2 * t->world * t->slot_floats * sizeof(float) + 255) & ~(size_t)255;
    t->alloc_bytes = t->flags_off + ((sizeof(FlagWords) + 255) & ~(size_t)255);
    t->ranks.resize(t->nlocal);
    int rc = MW_OK;
    for (int i = 0; i < t->nlocal && !rc; ++i) {
        const int rank = t->single ? i : p.rank;
        rc = create_rank(t, t->ranks[i], rank, p.devices[rank]);
    }
    if (!rc && (t->single || t->world == 1)) rc = t->world > 1 ? connect_single(t) : (t->connected = true, MW_OK);
    if (rc) { mw_tiles_destroy(t); return rc; }
    *out = t;
    return MW_OK;
}

extern "C" void mw_tiles_destroy(mw_tiles* t)
{
    if (!t) return;
    for (auto& r : t->ranks) destroy_rank(t, r);
    delete t;
}

extern "C" int mw_tiles_disconnect(mw_tiles* t)
{
    if (!t) { mw_set_error("null handle"); return MW_E_INVALID_ARG; }
    int rc = mw_tiles_sync(t);
    for (auto& r : t->ranks) {
        if (r.device < 0) continue;
        cudaSetDevice(r.device);
        if (r.comm && nccl_api()->ok) { nccl_api()->CommDestroy(r.comm); r.comm = nullptr; }
        if (!t->single)
            for (int p = 0; p < MAXW; ++p)
                if (r.peer_base[p]) { cudaIpcCloseMemHandle(r.peer_base[p]); r.peer_base[p] = nullptr; }
    }
    if (t->world > 1) t->connected = false;
    return rc;
}

extern "C" int mw_tiles_get_layout(const mw_tiles* t, mw_tiles_layout* l)
{
    if (!t || !l) { mw_set_error("null argument"); return MW_E_INVALID_ARG; }
    const int64_t pts = (int64_t)t->tpr * (int64_t)t->n2;
    l->slot_floats = (int64_t)t->slot_floats;
    l->height_off = 0;
    l->disp_off = pts;
    l->normal_off = pts * 3;
    l->whitecap_off = pts * 6;
    l->world = t->world;
    l->tiles_per_rank = t->tpr;
    l->resolution = t->N;
    l->local_ranks = t->nlocal;
    return MW_OK;
}

extern "C" int mw_tiles_gather_impl(const mw_tiles* t) { return t ? t->impl : MW_E_INVALID_ARG; }

extern "C" mw_ocean* mw_tiles_ocean(mw_tiles* t, int local_rank)
{
    if (!t || local_rank < 0 || local_rank >= t->nlocal) return nullptr;
    return t->ranks[local_rank].ocean;
}

extern "C" int mw_tiles_export(mw_tiles* t, void* blob)
{
    if (!t || !blob) { mw_set_error("null argument"); return MW_E_INVALID_ARG; }
    if (t->single) { mw_set_error("mw_tiles_export: only for one-process-per-GPU handles (rank >= 0)"); return MW_E_STATE; }
    TileRank& r = t->ranks[0];
    MW_CUDA(cudaSetDevice(r.device));
    Blob b;
    memset(&b, 0, sizeof b);
    b.magic = BLOB_MAGIC; b.rank = (uint32_t)r.rank; b.world = (uint32_t)t->world; b.device = (uint32_t)r.device;
    b.alloc_bytes = t->alloc_bytes; b.flags_off = t->flags_off; b.slot_floats = t->slot_floats;
    b.gather = (uint32_t)t->impl; b.tiles_per_rank = (uint32_t)t->tpr; b.resolution = (uint32_t)t->N;
    if (t->impl == MW_GATHER_PEER && t->world > 1) MW_CUDA(cudaIpcGetMemHandle(&b.mem, r.alloc));
    if (t->impl == MW_GATHER_NCCL && r.rank == 0 && t->world > 1) {
        MW_NCCL(nccl_api()->GetUniqueId(&b.nccl_id));
        b.has_nccl_id = 1;
    }
    memset(blob, 0, MW_TILES_BLOB_BYTES);
    memcpy(blob, &b, sizeof b);
    return MW_OK;
}

extern "C" int mw_tiles_connect(mw_tiles* t, const void* blobs)
{
    if (!t || !blobs) { mw_set_error("null argument"); return MW_E_INVALID_ARG; }
    if (t->single) { mw_set_error("mw_tiles_connect: only for one-process-per-GPU handles (rank >= 0)"); return MW_E_STATE; }
    if (t->connected) return MW_OK;
    TileRank& r = t->ranks[0];
    MW_CUDA(cudaSetDevice(r.device));
    std::vector<Blob> all(t->world);
    for (int p = 0; p < t->world; ++p) {
        memcpy(&all[p], (const char*)blobs + (size_t)p * MW_TILES_BLOB_BYTES, sizeof(Blob));
        const Blob& b = all[p];
        if (b.magic != BLOB_MAGIC || (int)b.rank != p || (int)b.world != t->world || b.slot_floats != t->slot_floats ||
            (int)b.gather != t->impl || b.alloc_bytes != t->alloc_bytes) {
            mw_set_error("mw_tiles_connect: blob %d does not describe rank %d of this tile set (world %d, %d tiles of %d^2 per rank, gather %d)",
                         p, p, t->world, t->tpr, t->N, t->impl);
            return MW_E_INVALID_ARG;
        }
    }
    if (t->impl == MW_GATHER_NCCL) {
        if (!all[0].has_nccl_id) { mw_set_error("mw_tiles_connect: rank 0's blob carries no ncclUniqueId"); return MW_E_INVALID_ARG; }
        MW_NCCL(nccl_api()->CommInitRank(&r.comm, t->world, all[0].nccl_id, r.rank));
        t->connected = true;
        return MW_OK;
    }
    MemOps* m = memops();
    if (getenv("MW_TILES_FAIL_PEER_CONNECT")) {   // developer knob: lets a host rehearse its fallback to the NCCL arm
        mw_set_error("mw_tiles_connect: peer mappings refused (MW_TILES_FAIL_PEER_CONNECT is set)");
        return MW_E_CUDA;
    }
    for (int p = 0; p < t->world; ++p) {
        if (p == r.rank) continue;
        int can = 0;
        MW_CUDA(cudaDeviceCanAccessPeer(&can, r.device, (int)all[p].device));
        if (!can && (int)all[p].device != r.device) { mw_set_error("no peer access from device %d to device %u (rank %d)", r.device, all[p].device, p); return MW_E_CUDA; }
        MW_CUDA(cudaIpcOpenMemHandle(&r.peer_base[p], all[p].mem, cudaIpcMemLazyEnablePeerAccess));
        char* base = (char*)r.peer_base[p];
        r.peer_gather[0][p] = reinterpret_cast<float*>(base);
        r.peer_gather[1][p] = r.peer_gather[0][p] + (size_t)t->world * t->slot_floats;
        r.peer_flags[p] = reinterpret_cast<FlagWords*>(base + t->flags_off);
    }
    // handshake: say hello to every peer, wait for every peer's hello.  Proves that the memory operations reach peer
    // memory in both directions before a frame depends on them.
    for (int p = 0; p < t->world; ++p) {
        if (p == r.rank) continue;
        if (!t->kernel_flags && m->write32((CUstream)r.s_comm, (CUdeviceptr)&r.peer_flags[p]->hello[r.rank], 1u, 0) != CUDA_SUCCESS) {
            (void)cudaGetLastError();
            t->kernel_flags = true;   // this driver does not take memory operations on peer mappings
        }
        if (t->kernel_flags) {
            int rc = flag_write(t, r.s_comm, &r.peer_flags[p]->hello[r.rank], 1u);
            if (rc) return rc;
        }
    }
    for (int p = 0; p < t->world; ++p)
        if (p != r.rank) MW_CU(m->wait32((CUstream)r.s_comm, (CUdeviceptr)&r.flags->hello[p], 1u, CU_STREAM_WAIT_VALUE_GEQ));
    const double deadline = getenv("MW_TILES_CONNECT_TIMEOUT") ? atof(getenv("MW_TILES_CONNECT_TIMEOUT")) : 60.0;
    if (!stream_done_within(r.s_comm, deadline)) {
        // release the stream (satisfy the waits locally) so that the handle can still be destroyed
        std::vector<uint32_t> ones(MAXW, 1u);
        cudaMemcpy(r.flags->hello, ones.data(), sizeof(uint32_t) * MAXW, cudaMemcpyHostToDevice);
        cudaStreamSynchronize(r.s_comm);
        mw_set_error("mw_tiles_connect: peer flag handshake did not complete within %.0f s (rank %d)", deadline, r.rank);
        return MW_E_CUDA;
    }
    t->connected = true;
    return MW_OK;
}

#define MW_CHECK_TILES(t)                                                                       \
    do {                                                                                        \
        if (!(t)) { mw_set_error("null handle"); return MW_E_INVALID_ARG; }                     \
        if (!(t)->connected) { mw_set_error("tile set not connected: call mw_tiles_export / mw_tiles_connect on every rank first"); return MW_E_STATE; } \
    } while (0)

extern "C" int mw_tiles_init_spectrum(mw_tiles* t)
{
    if (!t) { mw_set_error("null handle"); return MW_E_INVALID_ARG; }
    for (auto& r : t->ranks) {
        int rc = mw_ocean_init_spectrum(r.ocean);
        if (rc) return rc;
    }
    return MW_OK;
}

extern "C" int mw_tiles_set_h0(mw_tiles* t, int local_rank, const float* h0, const float* h0conj)
{
    if (!t || local_rank < 0 || local_rank >= t->nlocal) { mw_set_error("mw_tiles_set_h0: bad handle or local rank"); return MW_E_INVALID_ARG; }
    TileRank& r = t->ranks[local_rank];
    MW_CUDA(cudaSetDevice(r.device));
    MW_CUDA(cudaEventRecord(r.ev_user, r.s_user));
    MW_CUDA(cudaStreamWaitEvent(r.s_gen, r.ev_user, 0));
    return mw_ocean_set_h0(r.ocean, h0, h0conj);
}

extern "C" int mw_tiles_set_stream(mw_tiles* t, void* const* streams)
{
    if (!t) { mw_set_error("null handle"); return MW_E_INVALID_ARG; }
    for (int i = 0; i < t->nlocal; ++i) {
        TileRank& r = t->ranks[i];
        MW_CUDA(cudaSetDevice(r.device));
        MW_CUDA(cudaStreamSynchronize(r.s_user));
        r.s_user = (streams && streams[i]) ? (cudaStream_t)streams[i] : r.s_user_own;
    }
    return MW_OK;
}

namespace {

// EvaluateWaves of every local rank into buffer b, ordered after the user stream and after the last gather of buffer b
int enqueue_generate(mw_tiles* t, int b, float time)
{
    for (auto& r : t->ranks) {
        MW_CUDA(cudaSetDevice(r.device));
        MW_CUDA(cudaEventRecord(r.ev_user, r.s_user));
        MW_CUDA(cudaStreamWaitEvent(r.s_gen, r.ev_user, 0));
        if (r.comm_used[b]) MW_CUDA(cudaStreamWaitEvent(r.s_gen, r.ev_comm[b], 0));
        float* slot = r.gather[b] + (size_t)r.rank * t->slot_floats;
        const size_t pts = (size_t)t->tpr * t->n2;
        mw_ocean_out o;
        memset(&o, 0, sizeof o);
        o.height = slot; o.disp = slot + pts; o.normal = slot + 3 * pts; o.whitecap = slot + 6 * pts;
        int rc = mw_ocean_generate(r.ocean, time, &o);
        if (rc) return rc;
        MW_CUDA(cudaSetDevice(r.device));
        MW_CUDA(cudaEventRecord(r.ev_gen[b], r.s_gen));
        r.gen_used[b] = true;
    }
    return MW_OK;
}

// launch the kernel push of `src` (slot_bytes) into dst[0 .. npeers) on stream s
int launch_push(mw_tiles* t, int push, cudaStream_t s, const float* src, char* const* dst, int npeers, size_t slot_bytes)
{
    if (push == PUSH_SM) {
        PushArgs pa;
        pa.src = reinterpret_cast<const float4*>(src);
        pa.npeers = npeers;
        pa.n16 = (unsigned long long)(slot_bytes / 16);
        for (int i = 0; i < npeers; ++i) pa.dst[i] = reinterpret_cast<float4*>(dst[i]);
        k_push_slots<<<t->push_ctas, 512, 0, s>>>(pa);
    } else {
        PushBulkArgs pa;
        pa.src = reinterpret_cast<const char*>(src);
        pa.npeers = npeers;
        pa.chunk = t->push_chunk;
        pa.bytes = (unsigned long long)slot_bytes;
        for (int i = 0; i < npeers; ++i) pa.dst[i] = dst[i];
        k_push_slots_bulk<<<t->push_ctas, 32, (size_t)PUSH_STAGES * t->push_chunk, s>>>(pa);   // (attribute set in create_rank)
    }
    MW_LAUNCH_CHECK();
    return MW_OK;
}

int enqueue_gather(mw_tiles* t, int b)
{
    if (t->world == 1) {
        for (auto& r : t->ranks) {
            MW_CUDA(cudaSetDevice(r.device));
            MW_CUDA(cudaStreamWaitEvent(r.s_comm, r.ev_gen[b], 0));
            MW_CUDA(cudaEventRecord(r.ev_comm[b], r.s_comm));
            r.comm_used[b] = true;
        }
        return MW_OK;
    }
    const uint32_t seq = ++t->gathers;
    const size_t slot_bytes = t->slot_floats * sizeof(float);
    // the kernel pushes move 16-byte units: slots of odd grids (direct-sum path) keep the copy engines
    const int push = (slot_bytes % 16 == 0) ? t->push_mode : PUSH_CE;
    // The user stream's position at this call bounds the readers of buffer b that the gather must not overtake.  The
    // "free" announcement travels on its own stream: it must not queue behind the previous gather's completion waits on
    // s_comm, or every step would pay a flag round trip between two gathers.  (Announcing before the previous gather into
    // this buffer has completed is safe: a peer's pushes into it are ordered on that peer's push stream.)
    for (auto& r : t->ranks) {
        MW_CUDA(cudaSetDevice(r.device));
        MW_CUDA(cudaEventRecord(r.ev_user, r.s_user));
        MW_CUDA(cudaStreamWaitEvent(r.s_comm, r.ev_user, 0));
        MW_CUDA(cudaStreamWaitEvent(r.s_flag, r.ev_user, 0));
    }
    if (t->impl == MW_GATHER_NCCL) {
        NcclApi* n = nccl_api();
        for (auto& r : t->ranks) {
            MW_CUDA(cudaSetDevice(r.device));
            if (r.gen_used[b]) MW_CUDA(cudaStreamWaitEvent(r.s_comm, r.ev_gen[b], 0));
        }
        if (t->nlocal > 1) MW_NCCL(n->GroupStart());
        for (auto& r : t->ranks) {
            MW_CUDA(cudaSetDevice(r.device));
            MW_NCCL(n->AllGather(r.gather[b] + (size_t)r.rank * t->slot_floats, r.gather[b], t->slot_floats, ncclFloat, r.comm, r.s_comm));
        }
        if (t->nlocal > 1) MW_NCCL(n->GroupEnd());
    } else if (t->single) {
        // phase 1: every rank's buffer b is free of readers from here on (its user stream has been waited for)
        for (auto& r : t->ranks) {
            MW_CUDA(cudaSetDevice(r.device));
            MW_CUDA(cudaEventRecord(r.ev_free[b], r.s_flag));
        }
        // phase 2: pushes, one copy-engine stream per (source, destination); destinations visited in a rank-dependent order
        for (auto& r : t->ranks) {
            MW_CUDA(cudaSetDevice(r.device));
            const float* src = r.gather[b] + (size_t)r.rank * t->slot_floats;
            if (push != PUSH_CE) {
                // one push kernel per source rank, storing through the peer-access mappings
                cudaStream_t s = r.s_push[(r.rank + 1) % t->world];
                char* dst[MAXW];
                int npeers = 0;
                if (r.gen_used[b]) MW_CUDA(cudaStreamWaitEvent(s, r.ev_gen[b], 0));
                for (int j = 1; j < t->world; ++j) {
                    const int p = (r.rank + j) % t->world;
                    MW_CUDA(cudaStreamWaitEvent(s, t->ranks[p].ev_free[b], 0));
                    dst[npeers++] = reinterpret_cast<char*>(t->ranks[p].gather[b] + (size_t)r.rank * t->slot_floats);
                }
                int rc = launch_push(t, push, s, src, dst, npeers, slot_bytes);
                if (rc) return rc;
                for (int p = 0; p < t->world; ++p) if (p != r.rank) MW_CUDA(cudaEventRecord(r.ev_push[p], s));
                continue;
            }
            for (int j = 1; j < t->world; ++j) {
                const int p = (r.rank + j) % t->world;
                TileRank& dst = t->ranks[p];
                cudaStream_t s = push_stream(t, r, p, j);
                if (r.gen_used[b]) MW_CUDA(cudaStreamWaitEvent(s, r.ev_gen[b], 0));
                MW_CUDA(cudaStreamWaitEvent(s, dst.ev_free[b], 0));
                MW_CUDA(cudaMemcpyPeerAsync(dst.gather[b] + (size_t)r.rank * t->slot_floats, dst.device, src, r.device, slot_bytes, s));
                MW_CUDA(cudaEventRecord(r.ev_push[p], s));
            }
        }
        // phase 3: a rank's gather is complete when its own slot is written, its pushes have been read out of it and every
        // peer's push has landed
        for (auto& r : t->ranks) {
            MW_CUDA(cudaSetDevice(r.device));
            if (r.gen_used[b]) MW_CUDA(cudaStreamWaitEvent(r.s_comm, r.ev_gen[b], 0));
            for (int p = 0; p < t->world; ++p) {
                if (p == r.rank) continue;
                MW_CUDA(cudaStreamWaitEvent(r.s_comm, r.ev_push[p], 0));
                MW_CUDA(cudaStreamWaitEvent(r.s_comm, t->ranks[p].ev_push[r.rank], 0));
            }
        }
    } else {
        MemOps* m = memops();
        TileRank& r = t->ranks[0];
        MW_CUDA(cudaSetDevice(r.device));
        // my buffer b is free of readers: tell every peer (ordered after ev_user on s_flag)
        for (int j = 1; j < t->world; ++j) {
            const int p = (r.rank + j) % t->world;
            int rc = flag_write(t, r.s_flag, &r.peer_flags[p]->free_[b][r.rank], seq);
            if (rc) return rc;
        }
        const float* src = r.gather[b] + (size_t)r.rank * t->slot_floats;
        if (push != PUSH_CE) {
            // one kernel on the first push stream: waits for every peer's "free", stores the slot into all of them, signals all
            cudaStream_t s = r.s_push[(r.rank + 1) % t->world];
            char* dst[MAXW];
            int npeers = 0;
            if (r.gen_used[b]) MW_CUDA(cudaStreamWaitEvent(s, r.ev_gen[b], 0));
            for (int j = 1; j < t->world; ++j) {
                const int p = (r.rank + j) % t->world;
                MW_CU(m->wait32((CUstream)s, (CUdeviceptr)&r.flags->free_[b][p], seq, CU_STREAM_WAIT_VALUE_GEQ));
                dst[npeers++] = reinterpret_cast<char*>(r.peer_gather[b][p] + (size_t)r.rank * t->slot_floats);
            }
            { int rc = launch_push(t, push, s, src, dst, npeers, slot_bytes); if (rc) return rc; }
            for (int j = 1; j < t->world; ++j) {
                const int p = (r.rank + j) % t->world;
                int rc = flag_write(t, s, &r.peer_flags[p]->landed[b][r.rank], seq);
                if (rc) return rc;
                MW_CUDA(cudaEventRecord(r.ev_push[p], s));
            }
        } else
        for (int j = 1; j < t->world; ++j) {
            const int p = (r.rank + j) % t->world;
            cudaStream_t s = push_stream(t, r, p, j);
            if (r.gen_used[b]) MW_CUDA(cudaStreamWaitEvent(s, r.ev_gen[b], 0));
            MW_CU(m->wait32((CUstream)s, (CUdeviceptr)&r.flags->free_[b][p], seq, CU_STREAM_WAIT_VALUE_GEQ));
            MW_CUDA(cudaMemcpyAsync(r.peer_gather[b][p] + (size_t)r.rank * t->slot_floats, src, slot_bytes, cudaMemcpyDeviceToDevice, s));
            int rc = flag_write(t, s, &r.peer_flags[p]->landed[b][r.rank], seq);
            if (rc) return rc;
            MW_CUDA(cudaEventRecord(r.ev_push[p], s));
        }
        if (r.gen_used[b]) MW_CUDA(cudaStreamWaitEvent(r.s_comm, r.ev_gen[b], 0));
        for (int p = 0; p < t->world; ++p) {
            if (p == r.rank) continue;
            MW_CUDA(cudaStreamWaitEvent(r.s_comm, r.ev_push[p], 0));
            MW_CU(m->wait32((CUstream)r.s_comm, (CUdeviceptr)&r.flags->landed[b][p], seq, CU_STREAM_WAIT_VALUE_GEQ));
        }
    }
    for (auto& r : t->ranks) {
        MW_CUDA(cudaSetDevice(r.device));
        MW_CUDA(cudaEventRecord(r.ev_comm[b], r.s_comm));
        r.comm_used[b] = true;
    }
    return MW_OK;
}

int host_wait(mw_tiles* t, bool comm)
{
    for (auto& r : t->ranks) {
        MW_CUDA(cudaSetDevice(r.device));
        MW_CUDA(cudaStreamSynchronize(comm ? r.s_comm : r.s_gen));
    }
    return MW_OK;
}

void report_buffers(mw_tiles* t, int b, void** gathered)
{
    if (!gathered) return;
    for (int i = 0; i < t->nlocal; ++i) gathered[i] = t->ranks[i].gather[b];
}

int next_buffer(mw_tiles* t)
{
    const int b = (int)(t->frames & 1u);
    t->frames += 1;
    t->buf_of_frame[1] = t->buf_of_frame[0];
    t->buf_of_frame[0] = b;
    return b;
}

}  // namespace

extern "C" int mw_tiles_generate_allgather(mw_tiles* t, float time, void** gathered)
{
    MW_CHECK_TILES(t);
    const int b = next_buffer(t);
    int rc;
    if ((rc = enqueue_generate(t, b, time)) || (rc = enqueue_gather(t, b))) return rc;
    report_buffers(t, b, gathered);
    return t->async ? MW_OK : host_wait(t, true);
}

extern "C" int mw_tiles_generate_local(mw_tiles* t, float time, void** gathered)
{
    MW_CHECK_TILES(t);
    const int b = next_buffer(t);
    int rc = enqueue_generate(t, b, time);
    if (rc) return rc;
    report_buffers(t, b, gathered);
    return t->async ? MW_OK : host_wait(t, false);
}

extern "C" int mw_tiles_allgather(mw_tiles* t)
{
    MW_CHECK_TILES(t);
    int rc = enqueue_gather(t, t->buf_of_frame[0]);
    if (rc) return rc;
    return t->async ? MW_OK : host_wait(t, true);
}

extern "C" int mw_tiles_wait(mw_tiles* t, int frames_back)
{
    MW_CHECK_TILES(t);
    if (frames_back < 0 || frames_back > 1) { mw_set_error("frames_back must be 0 or 1"); return MW_E_INVALID_ARG; }
    const int b = t->buf_of_frame[frames_back];
    for (auto& r : t->ranks) {
        MW_CUDA(cudaSetDevice(r.device));
        if (r.gen_used[b]) MW_CUDA(cudaStreamWaitEvent(r.s_user, r.ev_gen[b], 0));
        if (r.comm_used[b]) MW_CUDA(cudaStreamWaitEvent(r.s_user, r.ev_comm[b], 0));
    }
    return MW_OK;
}

extern "C" int mw_tiles_sync(mw_tiles* t)
{
    if (!t) { mw_set_error("null handle"); return MW_E_INVALID_ARG; }
    for (auto& r : t->ranks) {
        MW_CUDA(cudaSetDevice(r.device));
        MW_CUDA(cudaStreamSynchronize(r.s_gen));
        MW_CUDA(cudaStreamSynchronize(r.s_flag));
        for (int p = 0; p < MAXW; ++p) if (r.s_push[p]) MW_CUDA(cudaStreamSynchronize(r.s_push[p]));
        MW_CUDA(cudaStreamSynchronize(r.s_comm));
        if (r.comm) {
            ncclResult_t async = ncclSuccess;
            MW_NCCL(nccl_api()->CommGetAsyncError(r.comm, &async));
            if (async != ncclSuccess) {
                mw_set_error("NCCL asynchronous error on rank %d: %s", r.rank, nccl_api()->GetErrorString(async));
                return MW_E_NCCL;
            }
        }
    }
    return MW_OK;
}
