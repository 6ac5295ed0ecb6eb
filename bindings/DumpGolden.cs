// DumpGolden.cs -- the one route by which parity with the REAL reference can be pinned.
//
// The reference (Assets/Mistral Water/Scripts/FFTMesh.cs) holds no tests or golden vectors and cannot run outside Unity,
// so every fixture in tests/golden/ was produced by this repository's own restatement (oracle/ref_fftmesh.c).  This
// MonoBehaviour closes that gap on a machine that has Unity: attach it to the GameObject that carries the UNMODIFIED
// FFTMesh component (e.g. in Demo/FFT Mesh.unity), press Play, and it writes what the reference itself computed:
//
//     <out>/fftmesh_unity_N{N}_meta.json                 parameters, times, Unity version
//     <out>/fftmesh_unity_N{N}_h0.npy, _h0conj.npy       verttilde / vertConj   (FFTMesh.cs:35-36, filled at :114-116)
//     <out>/fftmesh_unity_N{N}_rest.npy                  vertices               (:107-112)
//     <out>/fftmesh_unity_N{N}_t{k}_vertMeow.npy         the displaced vertices (:243-245)   for each time t_k
//     <out>/fftmesh_unity_N{N}_t{k}_normals.npy          normals                (:246)
//     <out>/fftmesh_unity_N{N}_t{k}_colors.npy           colors                 (:274)
//
// Copy the files into tests/golden/unity/ of this repository: tests/test_reference_vectors.py then checks BOTH the CPU
// oracle and the CUDA engine (through mw_ocean_set_h0, so UnityEngine.Random's draws are taken as given) against them and
// stops skipping.  Private members are reached by reflection: FFTMesh.cs itself stays byte-for-byte the reference's.
//
// The grids are periodic (length = resolution * unitWidth, power-of-two resolution): the case the engine supports and the
// one in which the reference's direct sum is a DFT (SURVEY.md section 3.4).
using System;
using System.Globalization;
using System.IO;
using System.Reflection;
using System.Text;
using UnityEngine;

public class DumpGolden : MonoBehaviour
{
    public int[] resolutions = { 16, 32 };
    public float[] times = { 0f, 1.7f, 60f };
    public string outputDirectory = "";      // empty: Application.persistentDataPath/mistral_golden
    public bool quitWhenDone = true;

    const BindingFlags Priv = BindingFlags.NonPublic | BindingFlags.Instance;

    void Start()
    {
        var fm = GetComponent<FFTMesh>();
        if (fm == null) { Debug.LogError("DumpGolden: no FFTMesh on this GameObject"); return; }
        string dir = string.IsNullOrEmpty(outputDirectory) ? Path.Combine(Application.persistentDataPath, "mistral_golden") : outputDirectory;
        Directory.CreateDirectory(dir);
        fm.enabled = false;                                   // no Update() in between: we drive EvaluateWaves ourselves
        foreach (int n in resolutions) Dump(fm, n, dir);
        Debug.Log("DumpGolden: wrote " + dir);
#if UNITY_EDITOR
        if (quitWhenDone) UnityEditor.EditorApplication.isPlaying = false;
#else
        if (quitWhenDone) Application.Quit();
#endif
    }

    void Dump(FFTMesh fm, int n, string dir)
    {
        Type T = typeof(FFTMesh);
        fm.resolution = n;                                     // FFTMesh.cs:13
        fm.unitWidth = 1f;                                     // :15
        fm.length = n;                                         // :19  periodic: length == resolution * unitWidth
        fm.choppiness = 1f; fm.tDivision = 1f;                 // :9, :11
        fm.wind = new Vector2(5f, 3f);                         // :21  (FFT Mesh.unity:151)
        fm.amplitude = 0.01f;                                  // :23  (FFT Mesh.unity:152)
        UnityEngine.Random.InitState(1234 + n);                // makes the dump repeatable on the same Unity build
        T.GetMethod("SetParams", Priv).Invoke(fm, null);       // :90-99
        T.GetMethod("GenerateMesh", Priv).Invoke(fm, null);    // :101-139
        string stem = Path.Combine(dir, "fftmesh_unity_N" + n);
        WriteNpy(stem + "_h0.npy", Flatten((Vector2[])T.GetField("verttilde", Priv).GetValue(fm)), n * n, 2);
        WriteNpy(stem + "_h0conj.npy", Flatten((Vector2[])T.GetField("vertConj", Priv).GetValue(fm)), n * n, 2);
        WriteNpy(stem + "_rest.npy", Flatten((Vector3[])T.GetField("vertices", Priv).GetValue(fm)), n * n, 3);
        MethodInfo eval = T.GetMethod("EvaluateWaves", Priv);  // :224-280
        Mesh mesh = GetComponent<MeshFilter>().mesh;
        for (int k = 0; k < times.Length; ++k)
        {
            eval.Invoke(fm, new object[] { times[k] });
            WriteNpy(stem + "_t" + k + "_vertMeow.npy", Flatten((Vector3[])T.GetField("vertMeow", Priv).GetValue(fm)), n * n, 3);
            WriteNpy(stem + "_t" + k + "_normals.npy", Flatten(mesh.normals), n * n, 3);   // normals are set on the mesh at :278
            WriteNpy(stem + "_t" + k + "_colors.npy", Flatten(mesh.colors), n * n, 4);     // colors at :279
        }
        var sb = new StringBuilder();
        sb.Append("{\"resolution\": ").Append(n).Append(", \"unit_width\": 1.0, \"length\": ").Append(n)
          .Append(", \"choppiness\": 1.0, \"amplitude\": 0.01, \"wind\": [5.0, 3.0], \"times\": [");
        for (int k = 0; k < times.Length; ++k) sb.Append(k > 0 ? ", " : "").Append(times[k].ToString("R", CultureInfo.InvariantCulture));
        sb.Append("], \"unity_version\": \"").Append(Application.unityVersion).Append("\", \"source\": \"unmodified FFTMesh.cs via DumpGolden.cs\"}");
        File.WriteAllText(stem + "_meta.json", sb.ToString());
    }

    static float[] Flatten(Vector2[] a) { var f = new float[a.Length * 2]; for (int i = 0; i < a.Length; ++i) { f[2 * i] = a[i].x; f[2 * i + 1] = a[i].y; } return f; }
    static float[] Flatten(Vector3[] a) { var f = new float[a.Length * 3]; for (int i = 0; i < a.Length; ++i) { f[3 * i] = a[i].x; f[3 * i + 1] = a[i].y; f[3 * i + 2] = a[i].z; } return f; }
    static float[] Flatten(Color[] a) { var f = new float[a.Length * 4]; for (int i = 0; i < a.Length; ++i) { f[4 * i] = a[i].r; f[4 * i + 1] = a[i].g; f[4 * i + 2] = a[i].b; f[4 * i + 3] = a[i].a; } return f; }

    // NumPy .npy format 1.0, little-endian float32, C order, shape (rows, cols)
    static void WriteNpy(string path, float[] data, int rows, int cols)
    {
        string dict = "{'descr': '<f4', 'fortran_order': False, 'shape': (" + rows + ", " + cols + "), }";
        int unpadded = 10 + dict.Length + 1;
        int pad = (64 - unpadded % 64) % 64;
        string header = dict + new string(' ', pad) + "\n";
        using (var w = new BinaryWriter(File.Create(path)))
        {
            w.Write(new byte[] { 0x93, (byte)'N', (byte)'U', (byte)'M', (byte)'P', (byte)'Y', 1, 0 });
            w.Write((ushort)header.Length);
            w.Write(Encoding.ASCII.GetBytes(header));
            var bytes = new byte[data.Length * 4];
            Buffer.BlockCopy(data, 0, bytes, 0, bytes.Length);
            if (!BitConverter.IsLittleEndian) for (int i = 0; i < bytes.Length; i += 4) Array.Reverse(bytes, i, 4);
            w.Write(bytes);
        }
    }
}
