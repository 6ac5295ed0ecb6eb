// MistralOceanNative.cs -- P/Invoke declarations for libmistral_ocean.so (include/mistral_ocean.h).
//
// Drop this file next to Assets/Mistral Water/Scripts/FFTMesh.cs and put the shared library where Unity's
// native-plugin loader finds it (Assets/Plugins/x86_64/libmistral_ocean.so).  It could not be compiled in the
// build image (no dotnet / mono / Unity there); the same exported symbols are exercised through ctypes by
// mistral-water_b200/native.py and tests/.  Struct layouts are asserted in tests/test_abi.py, and
// bindings/MistralOcean.Bindings.csproj compiles this file WITHOUT Unity (bindings/UnityShim.cs stands in for the
// four UnityEngine types used here) together with bindings/LayoutCheck.cs, which prints Marshal.SizeOf / OffsetOf of
// every struct next to the values the C header gives: `dotnet run --project bindings` on any machine with the SDK.
using System;
using System.Runtime.InteropServices;
using UnityEngine;

namespace MistralWater.Native
{
    [StructLayout(LayoutKind.Sequential)]
    public struct MwOceanParams            // mw_ocean_params: FFTMesh's public fields one to one (FFTMesh.cs:9-23)
    {
        public int resolution;             // FFTMesh.cs:13
        public float unitWidth;            // :15
        public float length;               // :19   must equal resolution * unitWidth
        public float choppiness;           // :9
        public float amplitude;            // :23
        public float windX, windY;         // :21
        public float tDivision;            // :11
        public ulong seed;                 // stand-in for UnityEngine.Random's state
        public int device;                 // CUDA device ordinal
        public int tiles;                  // independent oceans in this handle (1 for FFTMesh)
        public uint flags;                 // MW_DEVICE_PTRS = 1, MW_PROFILE = 2
        public uint reserved;
    }

    [StructLayout(LayoutKind.Sequential)]
    public struct MwOceanOut               // mw_ocean_out: IntPtr.Zero = not requested
    {
        public IntPtr height;              // float[N*N]
        public IntPtr disp;                // Vector2[N*N]   hds            (FFTMesh.cs:247)
        public IntPtr normal;              // Vector3[N*N]   normals        (:246)
        public IntPtr whitecap;            // float[N*N]
        public IntPtr jacobian;            // float[N*N]
        public IntPtr vertices;            // Vector3[N*N]   vertMeow       (:243-245)
        public IntPtr colors;              // Color[N*N]     colors         (:274)
    }

    [StructLayout(LayoutKind.Sequential)]
    public struct MwGerstnerWave { public float dirX, dirY, freq, rate, ampXZ, ampY; }

    [StructLayout(LayoutKind.Sequential)]
    public struct MwGerstnerParams
    {
        public int nWaves, device;
        public uint flags;                 // MW_DEVICE_PTRS = 1, MW_GERSTNER_NORMAL_ANALYTIC = 16, MW_GERSTNER_NORMAL_DISCARDED = 32
        public float smoothing;            // _Smoothing (MistralWaterLib.cginc:66), for MW_GERSTNER_NORMAL_DISCARDED
        [MarshalAs(UnmanagedType.ByValArray, SizeConst = 64)] public MwGerstnerWave[] waves;
    }

    [StructLayout(LayoutKind.Sequential)]
    public struct MwRendererParams         // mw_renderer_params: OceanRenderer's public fields (OceanRenderer.cs:10-19)
    {
        public int resolution;             // :13  mesh resolution; the maps are 8 x this
        public float unitWidth;            // :12
        public float length;               // :14
        public float choppiness;           // :16
        public float amplitude;            // :18  (the engine applies the / 10000 of :149)
        public float windX, windY;         // :19
        public float mult;                 // :11
        public float seed1, seed2;         // Random.value * 10 (:147-148)
        public int device, tiles;
        public uint flags, reserved;       // MW_DEVICE_PTRS = 1, MW_WRAP_REPEAT = 4
    }

    [StructLayout(LayoutKind.Sequential)]
    public struct MwRendererOut            // mw_renderer_out: IntPtr.Zero = not requested; R x R texels each
    {
        public IntPtr displacement;        // Color[]  -> _Anim    (OceanRenderer.cs:310)
        public IntPtr height;              // Color[]  -> _Height  (:313)
        public IntPtr normal;              // Color[]  -> _Bump    (:311)
        public IntPtr white;               // float[]  -> _White.r (:312)
        public IntPtr whiteRgba;           // Color[]  (xx, xx, xx, 1)
        public IntPtr jacobian;            // float[]
    }

    [StructLayout(LayoutKind.Sequential)]
    public struct MwTilesParams            // mw_tiles_params: multi-GPU tile sets (no reference counterpart; SURVEY.md section 8e)
    {
        public MwOceanParams ocean;        // per-tile parameters (.tiles / .device ignored)
        public int world;                  // ranks = GPUs
        public int rank;                   // -1: this process drives every GPU (the Unity host model)
        public int tilesPerRank;
        public int gather;                 // MW_GATHER_NCCL = 0, MW_GATHER_PEER = 1, MW_GATHER_AUTO = 2
        [MarshalAs(UnmanagedType.ByValArray, SizeConst = 16)] public int[] devices;
        public float windStepDeg;          // config 5: 45
        public uint flags;                 // MW_TILES_ASYNC = 1 | one of MW_TILES_PUSH_CE = 2, _SM = 4, _TMA = 8 (default)
    }

    [StructLayout(LayoutKind.Sequential)]
    public struct MwTilesLayout            // mw_tiles_layout
    {
        public long slotFloats, heightOff, dispOff, normalOff, whitecapOff;
        public int world, tilesPerRank, resolution, localRanks;
    }

    [StructLayout(LayoutKind.Sequential)]
    public struct MwWaveParams { public float amplitude, frequency, speed, smoothing; public int device; public uint flags; }

    public static class MistralOcean
    {
        const string Lib = "mistral_ocean";
        public const int MW_OK = 0, MW_E_INVALID_ARG = -1, MW_E_CUDA = -2, MW_E_OOM = -3, MW_E_STATE = -4, MW_E_NCCL = -5;

        [DllImport(Lib)] public static extern int mw_version();
        [DllImport(Lib)] static extern IntPtr mw_last_error();
        public static string LastError() { return Marshal.PtrToStringAnsi(mw_last_error()); }

        [DllImport(Lib)] public static extern int mw_ocean_create(ref MwOceanParams p, out IntPtr handle);
        [DllImport(Lib)] public static extern void mw_ocean_destroy(IntPtr handle);
        [DllImport(Lib)] public static extern int mw_ocean_init_spectrum(IntPtr handle);
        [DllImport(Lib)] public static extern int mw_ocean_set_h0(IntPtr handle, IntPtr h0, IntPtr h0conj);
        [DllImport(Lib)] public static extern int mw_ocean_get_h0(IntPtr handle, IntPtr h0, IntPtr h0conj);
        [DllImport(Lib)] public static extern int mw_ocean_get_rest_vertices(IntPtr handle, IntPtr xyz);
        [DllImport(Lib)] public static extern int mw_ocean_get_dispersion(IntPtr handle, IntPtr omega);
        [DllImport(Lib)] public static extern int mw_ocean_evolve_spectrum(IntPtr handle, float t, IntPtr htilde);
        [DllImport(Lib)] public static extern int mw_ocean_generate(IntPtr handle, float t, ref MwOceanOut o);
        [DllImport(Lib)] public static extern int mw_ocean_update(IntPtr handle, float deltaTime, ref MwOceanOut o);
        [DllImport(Lib)] public static extern int mw_ocean_reset_timer(IntPtr handle);
        [DllImport(Lib)] public static extern float mw_ocean_timer(IntPtr handle);
        [DllImport(Lib)] public static extern int mw_ocean_sync(IntPtr handle);
        [DllImport(Lib)] public static extern int mw_ocean_set_stream(IntPtr handle, IntPtr cudaStream);
        [DllImport(Lib)] public static extern int mw_ocean_kernel_times(IntPtr handle, float[] ms, long[] launches, int reset);
        [DllImport(Lib)] public static extern long mw_kernel_launch_count();
        [DllImport(Lib)] public static extern int mw_fft2d(int device, int n, int batch, int sign, IntPtr input, IntPtr output);
        [DllImport(Lib)] public static extern int mw_gerstner_from_material(ref MwGerstnerParams p, float amplitude, float frequency,
            float steepness, float[] wSpeed, float[] wDirectionAB, float[] wDirectionCD);
        [DllImport(Lib)] public static extern int mw_gerstner_append_level_one(ref MwGerstnerParams p, float amplitude, float frequency, float steepness);
        [DllImport(Lib)] public static extern int mw_gerstner_displace(ref MwGerstnerParams p, IntPtr posXyz, IntPtr outXyz, IntPtr outNrm,
            long n, float t, IntPtr cudaStream);

        [DllImport(Lib)] public static extern int mw_renderer_create(ref MwRendererParams p, out IntPtr handle);
        [DllImport(Lib)] public static extern void mw_renderer_destroy(IntPtr handle);
        [DllImport(Lib)] public static extern int mw_renderer_render_initial(IntPtr handle);
        [DllImport(Lib)] public static extern int mw_renderer_set_initial(IntPtr handle, IntPtr rgba);
        [DllImport(Lib)] public static extern int mw_renderer_get_initial(IntPtr handle, IntPtr rgba);
        [DllImport(Lib)] public static extern int mw_renderer_set_phase(IntPtr handle, IntPtr phase);
        [DllImport(Lib)] public static extern int mw_renderer_get_phase(IntPtr handle, IntPtr phase);
        [DllImport(Lib)] public static extern int mw_renderer_set_params(IntPtr handle, float length, float choppiness, float amplitude,
            float windX, float windY);
        [DllImport(Lib)] public static extern int mw_renderer_generate_texture(IntPtr handle, float deltaTime, ref MwRendererOut o);
        [DllImport(Lib)] public static extern int mw_renderer_sync(IntPtr handle);
        [DllImport(Lib)] public static extern int mw_mesh_generate(int device, int resolution, float unitWidth, IntPtr vertices,
            IntPtr normals, IntPtr uvs, IntPtr indices);
        [DllImport(Lib)] public static extern int mw_wave_displace(ref MwWaveParams p, IntPtr posXyz, IntPtr outXyz, IntPtr outNrm,
            long n, float t, IntPtr cudaStream);

        // multi-GPU tile sets (no reference counterpart): the handle owns gather buffers, streams, peer mappings and NCCL
        // communicators; rank = -1 drives all GPUs from this one process (ncclCommInitAll / peer copies fenced by events)
        public const int MW_GATHER_NCCL = 0, MW_GATHER_PEER = 1, MW_GATHER_AUTO = 2, MW_TILES_BLOB_BYTES = 512;
        public const uint MW_TILES_ASYNC = 1, MW_TILES_PUSH_CE = 2, MW_TILES_PUSH_SM = 4, MW_TILES_PUSH_TMA = 8;
        [DllImport(Lib)] public static extern int mw_tiles_create(ref MwTilesParams p, out IntPtr handle);
        [DllImport(Lib)] public static extern void mw_tiles_destroy(IntPtr handle);
        [DllImport(Lib)] public static extern int mw_tiles_disconnect(IntPtr handle);
        [DllImport(Lib)] public static extern int mw_tiles_get_layout(IntPtr handle, out MwTilesLayout layout);
        [DllImport(Lib)] public static extern int mw_tiles_export(IntPtr handle, byte[] blob512);
        [DllImport(Lib)] public static extern int mw_tiles_connect(IntPtr handle, byte[] blobsInRankOrder);
        [DllImport(Lib)] public static extern int mw_tiles_init_spectrum(IntPtr handle);
        [DllImport(Lib)] public static extern int mw_tiles_set_h0(IntPtr handle, int localRank, IntPtr h0Device, IntPtr h0conjDevice);
        [DllImport(Lib)] public static extern int mw_tiles_set_stream(IntPtr handle, IntPtr[] cudaStreams);
        [DllImport(Lib)] public static extern int mw_tiles_generate_allgather(IntPtr handle, float time, IntPtr[] gathered);
        [DllImport(Lib)] public static extern int mw_tiles_generate_local(IntPtr handle, float time, IntPtr[] gathered);
        [DllImport(Lib)] public static extern int mw_tiles_allgather(IntPtr handle);
        [DllImport(Lib)] public static extern int mw_tiles_wait(IntPtr handle, int framesBack);
        [DllImport(Lib)] public static extern int mw_tiles_sync(IntPtr handle);
        [DllImport(Lib)] public static extern int mw_tiles_gather_impl(IntPtr handle);
        [DllImport(Lib)] public static extern IntPtr mw_tiles_ocean(IntPtr handle, int localRank);

        public static void Check(int rc) { if (rc != MW_OK) throw new InvalidOperationException("mistral_ocean " + rc + ": " + LastError()); }
    }

    // The substitution inside OceanRenderer.cs (see INTEGRATION.md): the body of GenerateTexture.  The four maps go to
    // Texture2D(R, R, TextureFormat.RGBAFloat / RFloat) objects with LoadRawTextureData(pinned array) + Apply(false),
    // bound once to the ocean material exactly like the RenderTextures of :310-313.
    public sealed class OceanRendererEngine : IDisposable
    {
        IntPtr handle;
        GCHandle hd, hh, hn, hw;
        MwRendererOut outBlock;
        public readonly Color[] displacement, height, normal;
        public readonly float[] white;

        public OceanRendererEngine(int resolution, float unitWidth, float length, float choppiness, float amplitude, Vector2 wind, float mult)
        {
            var p = new MwRendererParams { resolution = resolution, unitWidth = unitWidth, length = length, choppiness = choppiness,
                amplitude = amplitude, windX = wind.x, windY = wind.y, mult = mult, seed1 = UnityEngine.Random.value * 10f,
                seed2 = UnityEngine.Random.value * 10f, device = 0, tiles = 1 };
            MistralOcean.Check(MistralOcean.mw_renderer_create(ref p, out handle));   // OceanRenderer.cs:116-170
            MistralOcean.Check(MistralOcean.mw_renderer_render_initial(handle));      // :209-214
            int texels = 64 * resolution * resolution;
            displacement = new Color[texels]; height = new Color[texels]; normal = new Color[texels]; white = new float[texels];
            hd = GCHandle.Alloc(displacement, GCHandleType.Pinned); hh = GCHandle.Alloc(height, GCHandleType.Pinned);
            hn = GCHandle.Alloc(normal, GCHandleType.Pinned); hw = GCHandle.Alloc(white, GCHandleType.Pinned);
            outBlock = new MwRendererOut { displacement = hd.AddrOfPinnedObject(), height = hh.AddrOfPinnedObject(),
                normal = hn.AddrOfPinnedObject(), white = hw.AddrOfPinnedObject() };
        }

        public void GenerateTexture(float deltaTime) { MistralOcean.Check(MistralOcean.mw_renderer_generate_texture(handle, deltaTime, ref outBlock)); }  // :216-316
        public void SetParams(float length, float choppiness, float amplitude, Vector2 wind)                                                          // :94-109
        { MistralOcean.Check(MistralOcean.mw_renderer_set_params(handle, length, choppiness, amplitude, wind.x, wind.y)); }

        public void Dispose()
        {
            if (handle != IntPtr.Zero) { MistralOcean.mw_renderer_destroy(handle); handle = IntPtr.Zero; }
            if (hd.IsAllocated) hd.Free(); if (hh.IsAllocated) hh.Free(); if (hn.IsAllocated) hn.Free(); if (hw.IsAllocated) hw.Free();
        }
    }

    // BASELINE config 5 from ONE host process: `world` GPUs, one independent tile each, every GPU ends with every tile
    // (device memory: the returned pointers are for CUDA-side consumers, e.g. graphics interop).
    public sealed class TileSetEngine : IDisposable
    {
        IntPtr handle;
        public readonly MwTilesLayout layout;
        public readonly IntPtr[] gathered;               // per GPU: [world][slotFloats] floats of the latest frame

        // flags: 0 = blocking calls + the default push engine (TMA bulk copies); MW_TILES_ASYNC and one of MW_TILES_PUSH_* may be or-ed in
        public TileSetEngine(int world, int resolution, float unitWidth, float choppiness, float amplitude, Vector2 wind, ulong seed,
                             int gather = MistralOcean.MW_GATHER_AUTO, uint flags = 0)
        {
            var p = new MwTilesParams { ocean = new MwOceanParams { resolution = resolution, unitWidth = unitWidth,
                length = resolution * unitWidth, choppiness = choppiness, amplitude = amplitude, windX = wind.x, windY = wind.y,
                tDivision = 1f, seed = seed, tiles = 1 }, world = world, rank = -1, tilesPerRank = 1,
                gather = gather, devices = new int[16], windStepDeg = 45f, flags = flags };
            for (int i = 0; i < world; ++i) p.devices[i] = i;
            MistralOcean.Check(MistralOcean.mw_tiles_create(ref p, out handle));
            MistralOcean.Check(MistralOcean.mw_tiles_get_layout(handle, out layout));
            MistralOcean.Check(MistralOcean.mw_tiles_init_spectrum(handle));
            gathered = new IntPtr[world];
        }

        public void GenerateAllGather(float t) { MistralOcean.Check(MistralOcean.mw_tiles_generate_allgather(handle, t, gathered)); }
        public void Dispose() { if (handle != IntPtr.Zero) { MistralOcean.mw_tiles_destroy(handle); handle = IntPtr.Zero; } }
    }

    // The substitution inside FFTMesh.cs (see INTEGRATION.md): bodies of SetParams / GenerateMesh / EvaluateWaves.
    public sealed class FFTMeshEngine : IDisposable
    {
        IntPtr handle;
        GCHandle hv, hn, hc;               // pinned managed arrays: Vector3[], Vector3[], Color[] are blittable
        MwOceanOut outBlock;

        public FFTMeshEngine(int resolution, float unitWidth, float length, float choppiness, float amplitude, Vector2 wind,
                             float tDivision, ulong seed, Vector3[] vertMeow, Vector3[] normals, Color[] colors)
        {
            var p = new MwOceanParams { resolution = resolution, unitWidth = unitWidth, length = length, choppiness = choppiness,
                amplitude = amplitude, windX = wind.x, windY = wind.y, tDivision = tDivision, seed = seed, device = 0, tiles = 1 };
            MistralOcean.Check(MistralOcean.mw_ocean_create(ref p, out handle));      // FFTMesh.cs:90-99
            MistralOcean.Check(MistralOcean.mw_ocean_init_spectrum(handle));          // FFTMesh.cs:114-116
            hv = GCHandle.Alloc(vertMeow, GCHandleType.Pinned);
            hn = GCHandle.Alloc(normals, GCHandleType.Pinned);
            hc = GCHandle.Alloc(colors, GCHandleType.Pinned);
            outBlock = new MwOceanOut { vertices = hv.AddrOfPinnedObject(), normal = hn.AddrOfPinnedObject(), colors = hc.AddrOfPinnedObject() };
        }

        public void EvaluateWaves(float t) { MistralOcean.Check(MistralOcean.mw_ocean_generate(handle, t, ref outBlock)); }  // FFTMesh.cs:224-280

        public void Dispose()
        {
            if (handle != IntPtr.Zero) { MistralOcean.mw_ocean_destroy(handle); handle = IntPtr.Zero; }
            if (hv.IsAllocated) hv.Free(); if (hn.IsAllocated) hn.Free(); if (hc.IsAllocated) hc.Free();
        }
    }
}
