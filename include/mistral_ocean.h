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
 *   - the engine supports the periodic case only: resolution a power of two in [32, 2048] and
 *     length == resolution * unit_width.  Anything else is
 *     MW_E_INVALID_ARG -- there is no CPU or O(N^4) fallback.
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

#define MW_VERSION 100 /* 0.1.0 */

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

typedef struct mw_gerstner_params {
    int32_t n_waves;
    int32_t device;
    uint32_t flags; /* MW_DEVICE_PTRS */
    uint32_t reserved;
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
 *                                                        its computed normal)
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
 * Multi-GPU tile sets: peer-memory plumbing for the one collective of the path, the all-gather of the
 * final float buffers.  No reference counterpart (Scripts/FFTMesh.cs runs one mesh on one device; tiles never exchange data
 * while being generated).  One process per GPU: every rank exports the buffer its peers write its gathered slots into,
 * opens the peers' buffers from its own device, and pushes its slot with one asynchronous copy per peer -- copy engines over
 * NVLink, no SMs.  Fencing between ranks is the host's business (mistral-water_b200/tiles.py uses two 4-byte NCCL all-reduces).
 *   mw_peer_export : CUDA IPC handle of the allocation `dev_ptr` lives in + the pointer's offset inside it
 *   mw_peer_open   : map an exported allocation for direct access from `device` (once per allocation per process)
 *   mw_peer_copy   : asynchronous device-to-device copy on `cuda_stream` (either side may be a peer mapping)
 */
#define MW_PEER_HANDLE_BYTES 64
int mw_peer_export(const void* dev_ptr, void* handle64, uint64_t* offset);
int mw_peer_open(int device, const void* handle64, void** base);
int mw_peer_close(int device, void* base);
int mw_peer_copy(void* dst, const void* src, uint64_t bytes, void* cuda_stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* MISTRAL_OCEAN_H */
