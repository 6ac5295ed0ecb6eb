// UnityShim.cs -- the four UnityEngine types MistralOceanNative.cs touches, so that the binding compiles with plain
// `dotnet build` (no Unity).  NOT for use inside Unity: there the real UnityEngine.dll provides these (exclude this file,
// or keep it out of Assets/).  Layouts are Unity's: sequential floats, blittable.
#if !UNITY_5_3_OR_NEWER
using System.Runtime.InteropServices;

namespace UnityEngine
{
    [StructLayout(LayoutKind.Sequential)] public struct Vector2 { public float x, y; public Vector2(float x, float y) { this.x = x; this.y = y; } }
    [StructLayout(LayoutKind.Sequential)] public struct Vector3 { public float x, y, z; public Vector3(float x, float y, float z) { this.x = x; this.y = y; this.z = z; } }
    [StructLayout(LayoutKind.Sequential)] public struct Color { public float r, g, b, a; }
    public static class Random
    {
        static readonly System.Random rng = new System.Random(1234);
        public static float value { get { return (float)rng.NextDouble(); } }
    }
}
#endif
