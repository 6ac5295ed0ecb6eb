// LayoutCheck.cs -- `dotnet run --project bindings`: prints the marshalled size / key offsets of every P/Invoke struct
// next to the value include/mistral_ocean.h gives (the same numbers tests/test_abi.py asserts for the ctypes mirror),
// and, if libmistral_ocean.so is on the loader path, calls mw_version() through the binding.  Exit code 0 = all match.
#if !UNITY_5_3_OR_NEWER
using System;
using System.Runtime.InteropServices;
using MistralWater.Native;

public static class LayoutCheck
{
    static int bad = 0;
    static void Expect(string what, long got, long want)
    {
        Console.WriteLine("{0,-44} {1,6} (header: {2}){3}", what, got, want, got == want ? "" : "   <-- MISMATCH");
        if (got != want) bad++;
    }

    public static int Main()
    {
        Expect("sizeof(mw_ocean_params)", Marshal.SizeOf(typeof(MwOceanParams)), 56);
        Expect("offsetof(mw_ocean_params, seed)", (long)Marshal.OffsetOf(typeof(MwOceanParams), "seed"), 32);
        Expect("sizeof(mw_ocean_out)", Marshal.SizeOf(typeof(MwOceanOut)), 7 * 8);
        Expect("sizeof(mw_gerstner_wave)", Marshal.SizeOf(typeof(MwGerstnerWave)), 24);
        Expect("sizeof(mw_gerstner_params)", Marshal.SizeOf(typeof(MwGerstnerParams)), 16 + 64 * 24);
        Expect("sizeof(mw_renderer_params)", Marshal.SizeOf(typeof(MwRendererParams)), 56);
        Expect("offsetof(mw_renderer_params, flags)", (long)Marshal.OffsetOf(typeof(MwRendererParams), "flags"), 48);
        Expect("sizeof(mw_renderer_out)", Marshal.SizeOf(typeof(MwRendererOut)), 6 * 8);
        Expect("sizeof(mw_wave_params)", Marshal.SizeOf(typeof(MwWaveParams)), 24);
        Expect("sizeof(mw_tiles_params)", Marshal.SizeOf(typeof(MwTilesParams)), 56 + 4 * 4 + 16 * 4 + 8);
        Expect("offsetof(mw_tiles_params, devices)", (long)Marshal.OffsetOf(typeof(MwTilesParams), "devices"), 72);
        Expect("sizeof(mw_tiles_layout)", Marshal.SizeOf(typeof(MwTilesLayout)), 5 * 8 + 4 * 4);
        try { Console.WriteLine("mw_version() through the binding: {0}", MistralOcean.mw_version()); }
        catch (DllNotFoundException) { Console.WriteLine("libmistral_ocean.so not on the loader path: layout check only"); }
        return bad == 0 ? 0 : 1;
    }
}
#endif
