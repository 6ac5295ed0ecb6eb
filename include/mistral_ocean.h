/*
 * mistral_ocean.h -- C ABI of libmistral_ocean.so, the B200 (sm_100a) engine behind the
 * Mistral Water asset's ocean / pond hot path.
 *
 * The reference has no FFI layer: its hot path is the private
 * bodies of two C# MonoBehaviours and two Cg functions.  Each entry point below names the
 * reference code whose body it replaces; paths are relative to
 *   /root/reference/Assets/Mistral Water/
 * The C# side binds these with [DllImport("mistral_ocean")] (bindings/MistralOceanNative.cs,
 * INTEGRATION.md).  Plain C types only; no exceptions or C++ types cross the boundary.
 *
 * Conventions
 *   - every call returns MW_OK (0) or a negative MW_E_* code; mw_last_error() gives the text
 *     (thread-local);
 *   - grid layout is the reference's: idx = i * resolution + j, i <-> x, j <-> z
 *     (Scripts/FFTMesh.cs:110); Vector2 = 2 packed floats, Vector3 = 3, Color = 4;
 *   - a handle may hold `tiles` independent oceans (same parameters, seed + tile index);
 *     every per-grid buffer is then [tile][idx] contiguous;
 *   - buffer arguments are HOST pointers unless the handle was created with MW_DEVICE_PTRS,
 *     in which case they are device pointers on the handle's device and the call is
 *     asynchronous on the handle's stream (mw_ocean_sync to wait);
 *   - the transform path covers the periodic case: resolution a power of two in [32, 2048] and
 *     length == resolution * unit_width (only then is FFTMesh.Displacement's
 *     direct sum a DFT).  Small grids outside it -- any resolution in [2, MW_DIRECT_MAX_RESOLUTION],
 *     any length, e.g. the FFT Mesh demo scene's own 12 x 12 / 12.39 (Demo/FFT Mesh.unity:147,150) --
 *     run the same sum directly ON THE GPU, one thread block per vertex (O(N^4) work, sized for what
 *     the reference's CPU loop can run at all).  Anything else is MW_E_INVALID_ARG.  There is no CPU path.
 */
#ifndef MISTRAL_OCEAN_H
#define MISTRAL_OCEAN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define MW_VERSION 200 /* 0.2.0 */
#define MW_DIRECT_MAX_RESOLUTION 256 /* largest grid the direct-sum path accepts */

enum {
    MW_OK = 0,
    MW_E_INVALID_ARG = -1,
    MW_E_CUDA = -2,
    MW_E_OOM = -3,
    MW_E_STATE = -4, /* e.g. generate before h0 was initialised / set */
    MW_E_NCCL = -5
};

/* mw_ocean_params.flags */
enum {
    MW_DEVICE_PTRS = 1u << 0, /* buffer arguments are device pointers; calls are stream-async   */
    MW_PROFILE = 1u << 1,     /* record CUDA events around each kernel (mw_ocean_kernel_times) */
    MW_WRAP_REPEAT = 1u << 2, /* mw_renderer only: border taps wrap around instead of clamping       */
    MW_HOST_ASYNC = 1u << 3   /* host-pointer handles only: set_h0 / generate / update enqueue their copies
                                 and kernels and return at once; the host buffers must be page-locked and stay
                                 untouched until mw_ocean_sync.  Results leave on a separate copy stream, so the
                                 upload of the next frame's inputs overlaps the download of this frame's outputs */
};

/*
 * Parameter block: the serialized public fields of FFTMesh (Scripts/FFTMesh.cs:9-23), which
 * OceanRenderer repeats (Scripts/OceanRenderer.cs:10-19), one to one.
 */
typedef struct mw_ocean_params {
    int32_t resolution; /* FFTMesh.cs:13  grid points per side, N                              */
    float unit_width;   /* FFTMesh.cs:15  vertex spacing                                        */
    float length;       /* FFTMesh.cs:19  patch length L; must equal resolution * unit_width    */
    float choppiness;   /* FFTMesh.cs:9   scales the horizontal displacement of `vertices` only */
    float amplitude;    /* FFTMesh.cs:23  Phillips A                                            */
    float wind_x;       /* FFTMesh.cs:21  wind.x                                                */
    float wind_y;       /*                wind.y                                                */
    float t_division;   /* FFTMesh.cs:11  timer += deltaTime / tDivision (mw_ocean_update)      */
    uint64_t seed;      /* stand-in for UnityEngine.Random's hidden state (Philox4x32-10 key)   */
    int32_t device;     /* CUDA device ordinal                                                  */
    int32_t tiles;      /* independent oceans held by this handle (>= 1); tile k uses seed + k  */
    uint32_t flags;     /* MW_DEVICE_PTRS | MW_PROFILE                                          */
    uint32_t reserved;
} mw_ocean_params;

/*
 * Per-frame outputs of EvaluateWaves (Scripts/FFTMesh.cs:224-280).  NULL = not requested.
 * All are float32, [tiles][N*N] x components.
 */
typedef struct mw_ocean_out {
    float* height;   /* x1  hd.y                  (FFTMesh.cs:219, 243)                          */
    float* disp;     /* x2  hds = (hd.x, hd.z)    (FFTMesh.cs:247); NOT scaled by choppiness     */
    float* normal;   /* x3  normals               (FFTMesh.cs:218, 246)                          */
    float* whitecap; /* x1  colors[idx].r (all four channels are equal) (FFTMesh.cs:268-274)     */
    float* jacobian; /* x1  the Jacobian determinant before the smoothstep (FFTMesh.cs:268)      */
    float* vertices; /* x3  vertMeow, the displaced vertex (FFTMesh.cs:243-245)                  */
    float* colors;   /* x4  colors as Unity's Color[] (FFTMesh.cs:274), for mesh.colors          */
} mw_ocean_out;

typedef struct mw_ocean mw_ocean; /* opaque; owns device memory, stream, events */

int mw_version(void);
const char* mw_last_error(void);

/* Replaces FFTMesh.SetParams (FFTMesh.cs:90-99): validates, allocates all device state. */
int mw_ocean_create(const mw_ocean_params* params, mw_ocean** out);
void mw_ocean_destroy(mw_ocean* o);

/*
 * Replaces the h0 part of FFTMesh.GenerateMesh (FFTMesh.cs:114-116 -> htilde0 :168-176 ->
 * Phillips :149-166) on the device, drawing the four uniforms per grid point from
 * Philox4x32-10(key = seed + tile, counter = idx) in the reference's draw order.
 */
int mw_ocean_init_spectrum(mw_ocean* o);

/*
 * Alternative to init_spectrum: the host keeps UnityEngine.Random and hands over
 * verttilde / vertConj (FFTMesh.cs:35-36, filled at :114-116), [tiles][N*N] Vector2 each.
 * Together with `t` this is the engine's whole state, so get/set doubles as checkpoint/resume.
 */
int mw_ocean_set_h0(mw_ocean* o, const float* h0, const float* h0conj);
int mw_ocean_get_h0(mw_ocean* o, float* h0, float* h0conj);

/* Rest positions `vertices` of FFTMesh.GenerateMesh (FFTMesh.cs:107-112), [N*N] Vector3 (host). */
int mw_ocean_get_rest_vertices(mw_ocean* o, float* xyz);

/* Dispersion(n, m) (FFTMesh.cs:141-147) for the whole grid, [N*N] float (host). Bit-exact. */
int mw_ocean_get_dispersion(mw_ocean* o, float* omega);

/* htilde(t, n, m) (FFTMesh.cs:178-190) for the whole grid, [tiles][N*N] Vector2. */
int mw_ocean_evolve_spectrum(mw_ocean* o, float t, float* htilde);

/* Replaces FFTMesh.EvaluateWaves(t) (FFTMesh.cs:224-280). */
int mw_ocean_generate(mw_ocean* o, float t, const mw_ocean_out* out);

/*
 * Replaces the body of FFTMesh.Update (FFTMesh.cs:60-73): timer += delta_time / tDivision;
 * EvaluateWaves(timer).  mw_ocean_reset_timer is the `generate` toggle's timer = 0 (:64).
 */
int mw_ocean_update(mw_ocean* o, float delta_time, const mw_ocean_out* out);
int mw_ocean_reset_timer(mw_ocean* o);
float mw_ocean_timer(const mw_ocean* o);

/* Wait for everything queued on the handle's stream. */
int mw_ocean_sync(mw_ocean* o);

/*
 * Use a caller-owned cudaStream_t (e.g. the host framework's current stream) instead of the
 * handle's own; pass NULL to go back.  The previous stream is synchronised first.
 */
int mw_ocean_set_stream(mw_ocean* o, void* cuda_stream);

/*
 * With MW_PROFILE: accumulated device time (ms) and launch count per kernel since the last
 * reset.  Kernel ids: 0 = spectrum+row FFT, 1 = column FFT+extract(+whitecap),
 * 2 = mesh-output epilogue (vertices / colors, only when requested).
 */
#define MW_KERNEL_COUNT 3
int mw_ocean_kernel_times(mw_ocean* o, float ms[MW_KERNEL_COUNT], int64_t launches[MW_KERNEL_COUNT], int reset);
/* Kernel launches issued by this library in this process (all handles, all entry points). */
int64_t mw_kernel_launch_count(void);

/*
 * The engine's 2-D transform on caller data: `batch` complex N x N fields, [batch][N][N]
 * float2.  sign = -1: forward-sign un-normalised DFT, i.e. what the reference's radix-2
 * Stockham blit chain computes (Shaders/FFT/Stockham.shader:31-57 scheduled by
 * Scripts/OceanRenderer.cs:229-262); sign = +1: the conjugate transform used by the
 * FFTMesh synthesis.  Always host pointers.
 */
int mw_fft2d(int device, int32_t n, int32_t batch, int sign, const float* in, float* out);

/*
 * ---------------------------------------------------------------------------------------------
 * OceanRenderer path: the GPU-shader convention the Ocean Demo scene runs (SURVEY.md section 8,
 * rows a11-a13 + a10).  Replaces the blit chain of Scripts/OceanRenderer.cs GenerateTexture
 * (:216-316) over Shaders/FFT/{InitialSpectrum, Dispersion, Spectrum, SpectrumHeight, Stockham,
 * OceanNormal, WhiteCap}.shader.  It differs from the FFTMesh path in every convention (SURVEY.md
 * section 3.5): FFT-ordered k, capillary dispersion with an accumulated phase, damping 0.01,
 * amplitude / 10000, hash noise, forward-sign transform, choppiness inside the spectrum, stencil
 * normals, +-8-texel Jacobian -- each restated from the shader it comes from.
 *
 * Images are R x R RGBAFloat, R = 8 * resolution (OceanRenderer.cs:136), row-major [y][x] with x
 * (texcoord.x, the "horizontal" Stockham direction) contiguous: what Texture2D.LoadRawTextureData /
 * GetRawTextureData use.  [tiles] images back to back.
 */
typedef struct mw_renderer_params {
    int32_t resolution; /* OceanRenderer.cs:13  mesh resolution; textures are 8 x this (power of two, 4..256) */
    float unit_width;   /* :12  (mesh only)                                                        */
    float length;       /* :14  _Length                                                            */
    float choppiness;   /* :16  _Choppiness (inside the spectrum, Spectrum.shader:48-49)           */
    float amplitude;    /* :18  the shader receives amplitude / 10000 (:100, :149)                 */
    float wind_x;       /* :19                                                                     */
    float wind_y;
    float mult;         /* :11  _DeltaTime = deltaTime * mult (:223)                               */
    float seed1;        /* _RandomSeed1 = Random.value * 10 (:147); tile t uses seed + t           */
    float seed2;        /* _RandomSeed2 (:148)                                                     */
    int32_t device;
    int32_t tiles;      /* independent oceans held by this handle (>= 1)                           */
    uint32_t flags;     /* MW_DEVICE_PTRS | MW_WRAP_REPEAT                                         */
    uint32_t reserved;
} mw_renderer_params;

/* The four maps OceanRenderer binds to the ocean material (OceanRenderer.cs:310-313). NULL = not requested. */
typedef struct mw_renderer_out {
    float* displacement; /* x4  _Anim   = (Re hx, Im hx, Re hz, Im hz)   displacementTexture (:244)           */
    float* height;       /* x4  _Height = (Re h, Im h, Re h, Im h)       heightTexture (:280)                 */
    float* normal;       /* x4  _Bump   = (n, 1)                         OceanNormal.shader:55                */
    float* white;        /* x1  _White.r: the only channel ColorMask R lets through (WhiteCap.shader:14, :44) */
    float* white_rgba;   /* x4  (xx, xx, xx, 1): the fragment's return value, for a plain RGBA upload         */
    float* jacobian;     /* x1  the Jacobian before the smoothstep (WhiteCap.shader:38); test / debug output  */
} mw_renderer_out;

typedef struct mw_renderer mw_renderer; /* opaque */

/* OceanRenderer.SetParams (:116-170): validates, allocates; the phase images start black (zero). */
int mw_renderer_create(const mw_renderer_params* params, mw_renderer** out);
void mw_renderer_destroy(mw_renderer* r);
/* RenderInitial (:209-214): InitialSpectrum.shader on the device -> initialTexture. */
int mw_renderer_render_initial(mw_renderer* r);
/* initialTexture as data, [tiles][R*R] x4 = (h0, h0conj): upload the host's own (e.g. read back from Unity's
 * GPU, whose hash noise depends on its sin) / read ours back.  With the phase image this is the whole state. */
int mw_renderer_set_initial(mw_renderer* r, const float* rgba);
int mw_renderer_get_initial(mw_renderer* r, float* rgba);
/* the accumulated phase image (ping/pong RFloat textures, :60-61), [tiles][R*R] x1 */
int mw_renderer_set_phase(mw_renderer* r, const float* phase);
int mw_renderer_get_phase(mw_renderer* r, float* phase);
/* OceanRenderer.Update's parameter refresh (:94-109): re-renders the initial spectrum when length, wind or
 * amplitude changed. */
int mw_renderer_set_params(mw_renderer* r, float length, float choppiness, float amplitude, float wind_x, float wind_y);
/* GenerateTexture (:216-316): advances the phase by delta_time * mult and renders the requested maps. */
int mw_renderer_generate_texture(mw_renderer* r, float delta_time, const mw_renderer_out* out);
int mw_renderer_sync(mw_renderer* r);

/*
 * GenerateMesh (Scripts/OceanRenderer.cs:172-207; the same topology code is in Scripts/FFTMesh.cs:101-139):
 * rest vertices [N*N] x3, normals (0,1,0) [N*N] x3, uvs [N*N] x2, triangle indices [(N-1)^2 * 6] in the
 * reference's emission order.  Any pointer may be NULL.  Host pointers.
 */
int mw_mesh_generate(int device, int32_t resolution, float unit_width, float* vertices, float* normals, float* uvs,
                     int32_t* indices);

/*
 * Pond renderer: Gerstner sum-of-waves vertex displacement.
 * One wave = one term of Shaders/MistralWaterLib.cginc Gerstner (:71-99) or GerstnerLevelOne
 * (:101-125):   theta = freq * (dir . pos.xz) + rate * t
 *               offs.x += amp_xz * dir_x * cos(theta); offs.z += amp_xz * dir_y * cos(theta)
 *               offs.y += amp_y * sin(theta)
 * mw_gerstner_from_material / _level_one fill the table from the shader's own uniforms.
 */
#define MW_GERSTNER_MAX_WAVES 64
typedef struct mw_gerstner_wave {
    float dir_x, dir_y; /* D_w (not normalised by the reference either)                        */
    float freq;         /* Gerstner: _Frequency;      LevelOne: _Frequency * fs[i]             */
    float rate;         /* Gerstner: _WSpeed[w];      LevelOne: speeds[i] * _Frequency * fs[i] */
    float amp_xz;       /* Gerstner: steepness * amp; LevelOne: steepness*amp*steeps[i]*amps[i]*/
    float amp_y;        /* Gerstner: amp;             LevelOne: amp * amps[i]                  */
} mw_gerstner_wave;

/* mw_gerstner_params.flags, besides MW_DEVICE_PTRS: what out_nrm receives.  Default (neither bit): (0, 1, 0), what the
 * reference ships -- both shader variants overwrite their normal with it (MistralWaterLib.cginc:98, :121). */
enum {
    MW_GERSTNER_NORMAL_ANALYTIC = 1u << 4,  /* the exact normal of the displaced surface P(x, z) = (x + offs.x, offs.y, z + offs.z):
                                               normalize(dP/dz x dP/dx), from the same per-wave sin / cos -- what the commented
                                               attempt at :122-124 was after                              */
    MW_GERSTNER_NORMAL_DISCARDED = 1u << 5  /* the value Gerstner() computes at :92-97 and then throws away, literally:
                                               n = (0, 2, 0); n.x -= offs.x; n.y -= offs.z; n.xz *= smoothing; normalize(n)     */
};
typedef struct mw_gerstner_params {
    int32_t n_waves;
    int32_t device;
    uint32_t flags;  /* MW_DEVICE_PTRS | MW_GERSTNER_NORMAL_* */
    float smoothing; /* _Smoothing (MistralWaterLib.cginc:66), read by MW_GERSTNER_NORMAL_DISCARDED only */
    mw_gerstner_wave waves[MW_GERSTNER_MAX_WAVES];
} mw_gerstner_params;

/* Displacement()'s Gerstner branch (MistralWaterLib.cginc:168-177): amplitude is _Amplitude (the
 * 0.01 factor of :172 is applied here); fills 4 waves. */
int mw_gerstner_from_material(mw_gerstner_params* p, float amplitude, float frequency, float steepness,
                              const float w_speed[4], const float w_direction_ab[4], const float w_direction_cd[4]);
/* GerstnerLevelOne's constant tables (MistralWaterLib.cginc:105-109); appends 5 waves. */
int mw_gerstner_append_level_one(mw_gerstner_params* p, float amplitude, float frequency, float steepness);

/*
 * out_xyz[v] = pos_xyz[v] + offsets(pos_xyz[v].xz, t)   (MistralWaterLib.cginc:176)
 * out_nrm[v] = (0, 1, 0)  if non-NULL                   (:98 / :121 -- the reference discards
 *                                                        its computed normal), or the normal selected
 *                                                        by MW_GERSTNER_NORMAL_* in p->flags
 * n vertices of packed float3.
 */
int mw_gerstner_displace(const mw_gerstner_params* p, const float* pos_xyz, float* out_xyz, float* out_nrm,
                         int64_t n, float t, void* cuda_stream);

/*
 * Pond renderer, `Wave` displacement mode (Shaders/MistralWaterLib.cginc Wave :127-152 through Displacement
 * :160-164, keyword _DISPLACEMENTMODE_WAVE), for a mesh whose object and world frames coincide:
 *   out_xyz[v] = (x, y + (y + A sin(s t + f x) - A cos(s t + f z)), z),  A = amplitude * 0.01
 *   out_nrm[v] = normalize(cross(v2 - v0, v1 - v0)) of the two 0.05-offset neighbours after the _Smoothing blend.
 */
typedef struct mw_wave_params {
    float amplitude; /* _Amplitude */
    float frequency; /* _Frequency */
    float speed;     /* _Speed     */
    float smoothing; /* _Smoothing */
    int32_t device;
    uint32_t flags;  /* MW_DEVICE_PTRS */
} mw_wave_params;
int mw_wave_displace(const mw_wave_params* p, const float* pos_xyz, float* out_xyz, float* out_nrm, int64_t n, float t,
                     void* cuda_stream);

/*
 * ---------------------------------------------------------------------------------------------
 * Multi-GPU tile sets (BASELINE config 5): independent ocean tiles, one rank per GPU, and the ONE
 * collective of the path -- the in-place all-gather of the final float buffers [height | hds | normal | whitecap] (28 B per
 * grid point), after which every rank holds every tile.  No reference counterpart: Scripts/FFTMesh.cs runs one mesh on one
 * device, and nothing in EvaluateWaves (:224-280) couples two meshes, which is why tiles shard with no other exchange.
 *
 * Rank r generates tiles [r * tiles_per_rank, (r + 1) * tiles_per_rank): tile k uses seed `ocean.seed + k` and the wind
 * (ocean.wind_x, ocean.wind_y) rotated by `wind_step_deg * (r * tiles_per_rank)` degrees (all tiles of a rank share its wind).
 * The engine writes straight into the rank's slot of the gather buffer; the handle owns the buffers, the streams, the
 * peer mappings and the NCCL communicators.
 *
 * Two process models, one API:
 *   single process (rank == -1)   the reference's host model (one Unity process): the handle drives all `world` devices --
 *                                 ncclCommInitAll + grouped ncclAllGather, or peer pushes fenced by CUDA events;
 *   one process per GPU (rank>=0) each process creates its rank, then the ranks swap fixed-size blobs (CUDA IPC handles of
 *                                 the gather buffers and flag words; rank 0's blob carries the ncclUniqueId) by whatever
 *                                 means the host has -- mw_tiles_export / mw_tiles_connect.
 * Two ways to carry out the all-gather (`gather`):
 *   MW_GATHER_NCCL  one in-place ncclAllGather per frame (libnccl.so.2 is loaded at run time; absent => MW_E_NCCL);
 *   MW_GATHER_PEER  every rank pushes its slot into the peers' buffers over NVLink (peer access inside a process, CUDA IPC
 *                   mappings between processes): by default one small kernel per rank and frame whose CTAs each run a ring
 *                   of TMA bulk copies global -> shared -> every peer (MW_TILES_PUSH_TMA), or copy-engine transfers / plain
 *                   stores (MW_TILES_PUSH_CE / _SM).  Ranks in different processes are fenced by stream memory operations on
 *                   flag words in peer memory (cuStreamWriteValue32 / cuStreamWaitValue32): no host round trip.
 * Frames are double-buffered: the gather of frame k (communication streams) runs under the generation of frame k + 1.
 * Contract for the returned buffers: the buffer of frame k is rewritten by frame k + 2; all reads of it must have been
 * enqueued on the user stream (mw_tiles_set_stream; default: the handle's own) before the call that generates frame k + 2.
 */
#define MW_TILES_MAX_WORLD 16
#define MW_TILES_BLOB_BYTES 512
enum {
    MW_GATHER_NCCL = 0,
    MW_GATHER_PEER = 1,
    MW_GATHER_AUTO = 2 /* what measured fastest on 8 x B200 / NVSwitch (profiles/r02_push_probe_{2,4,8}gpu.jsonl, r02_bench_8gpu.json):
                          the peer pushes with the TMA kernel at every world size (2048^2 step at 2 / 4 / 8 GPUs: 0.21 / 0.54 /
                          1.28 ms against 0.37 / 0.64 / 1.45 ms with ncclAllGather); ncclAllGather if stream memory operations
                          are unavailable between processes, peer if NCCL cannot be loaded */
};
enum {
    MW_TILES_ASYNC = 1u << 0,   /* generate_allgather only enqueues; mw_tiles_wait / mw_tiles_sync order the results        */
    /* MW_GATHER_PEER: what moves a rank's slot into its peers' buffers (none of the three = MW_TILES_PUSH_TMA)             */
    MW_TILES_PUSH_CE = 1u << 1, /* one copy-engine transfer per peer (cudaMemcpyAsync / cudaMemcpyPeerAsync)                */
    MW_TILES_PUSH_SM = 1u << 2, /* one kernel: 16-byte loads, (world - 1) 16-byte stores through the peer mappings          */
    MW_TILES_PUSH_TMA = 1u << 3 /* one kernel: a ring of cp.async.bulk global -> shared -> (world - 1) peers per CTA, driven
                                   by one thread; the default                                                              */
};

typedef struct mw_tiles_params {
    mw_ocean_params ocean;   /* per-tile parameters; .tiles and .device are ignored (tiles_per_rank / devices[] below),
                                .flags may carry MW_PROFILE                                                              */
    int32_t world;           /* number of ranks = GPUs, 1..MW_TILES_MAX_WORLD                                           */
    int32_t rank;            /* -1: this process drives all ranks; r >= 0: this process is rank r                       */
    int32_t tiles_per_rank;
    int32_t gather;          /* MW_GATHER_NCCL | MW_GATHER_PEER | MW_GATHER_AUTO                                        */
    int32_t devices[MW_TILES_MAX_WORLD]; /* CUDA ordinal of rank r (rank >= 0: only devices[rank] is read)              */
    float wind_step_deg;     /* config 5: 45                                                                            */
    uint32_t flags;          /* MW_TILES_ASYNC | one of MW_TILES_PUSH_*                                                 */
} mw_tiles_params;

typedef struct mw_tiles_layout {
    int64_t slot_floats;     /* floats per rank slot = tiles_per_rank * N * N * 7                                       */
    int64_t height_off, disp_off, normal_off, whitecap_off; /* float offsets of the planar fields inside a slot:
                                field f of local tile l, grid index idx, component c sits at
                                f_off + (l * N * N + idx) * comps(f) + c                                                */
    int32_t world, tiles_per_rank, resolution, local_ranks; /* local_ranks: `world` (single process) or 1              */
} mw_tiles_layout;

typedef struct mw_tiles mw_tiles; /* opaque */

int mw_tiles_create(const mw_tiles_params* params, mw_tiles** out);
void mw_tiles_destroy(mw_tiles* t);
/* One process per GPU: waits for this rank's queued work, then drops its mappings of the peers' buffers and its NCCL
 * communicator.  Call it on every rank, synchronise the ranks (any barrier the host has), then mw_tiles_destroy: an exported
 * allocation must not be freed while a peer still maps it.  (mw_tiles_destroy alone does both steps without the barrier.) */
int mw_tiles_disconnect(mw_tiles* t);
int mw_tiles_get_layout(const mw_tiles* t, mw_tiles_layout* layout);
/* One process per GPU only: this rank's blob (MW_TILES_BLOB_BYTES), then all `world` blobs in rank order.  connect() is
 * collective (every rank must call it; it opens the peer mappings, runs a flag handshake, or ncclCommInitRank). */
int mw_tiles_export(mw_tiles* t, void* blob);
int mw_tiles_connect(mw_tiles* t, const void* blobs);
/* Device-side h0 init of every local tile (mw_ocean_init_spectrum). */
int mw_tiles_init_spectrum(mw_tiles* t);
/* mw_ocean_set_h0 for local rank i: DEVICE pointers on that rank's device, [tiles_per_rank][N*N] Vector2 each, read in
 * user-stream order (an upload enqueued on the user stream before this call is complete when the engine reads it). */
int mw_tiles_set_h0(mw_tiles* t, int local_rank, const float* h0, const float* h0conj);
/* The user stream(s) the calls below are ordered against: `streams[i]` for local rank i (cudaStream_t; NULL entries or a
 * NULL array select the handle's own). */
int mw_tiles_set_stream(mw_tiles* t, void* const* streams);
/*
 * One frame: EvaluateWaves(t) of every local tile into its rank's slot, then the all-gather.  gathered[i] receives local
 * rank i's gather buffer of this frame, [world][slot_floats] floats on that rank's device (may be NULL).  Without
 * MW_TILES_ASYNC the call returns when the buffers are complete; with it, see mw_tiles_wait.
 */
int mw_tiles_generate_allgather(mw_tiles* t, float time, void** gathered);
/* The two halves, for hosts (and benches) that want them apart: compute only / gather of the frame last generated. */
int mw_tiles_generate_local(mw_tiles* t, float time, void** gathered);
int mw_tiles_allgather(mw_tiles* t);
/* The user stream(s) wait (on the device, not the host) for the gather of the latest frame (frames_back = 0) or of the
 * one before (1).  mw_tiles_sync blocks the host until everything queued has completed and reports NCCL's asynchronous
 * errors (ncclCommGetAsyncError) as MW_E_NCCL. */
int mw_tiles_wait(mw_tiles* t, int frames_back);
int mw_tiles_sync(mw_tiles* t);
/* Which implementation runs the gather: MW_GATHER_NCCL or MW_GATHER_PEER (MW_GATHER_AUTO is resolved at create time). */
int mw_tiles_gather_impl(const mw_tiles* t);
/* The mw_ocean handle of local rank i (borrowed: owned by the tile set), e.g. for mw_ocean_set_h0 / mw_ocean_kernel_times. */
mw_ocean* mw_tiles_ocean(mw_tiles* t, int local_rank);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* MISTRAL_OCEAN_H */
