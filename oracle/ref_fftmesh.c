/*
 * oracle/ref_fftmesh.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * A literal, line-by-line C restatement of the reference's CPU ocean path,
 *   /root/reference/Assets/Mistral Water/Scripts/FFTMesh.cs
 * written from a reading of that file (no source copied; it is C#, this is C).
 * Each function names the FFTMesh.cs lines it follows.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / the timed CPU baseline.  The
 * product (libmistral_ocean.so) never links, loads or calls anything in oracle/.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or seeds
 * and cannot be executed here (C# against closed-source UnityEngine; no mono/dotnet in the
 * image).  This restatement therefore *defines* "reference results" under these stated
 * UnityEngine semantics:
 *   - all storage and +,-,*,/ are IEEE fp32, evaluated in source order, no FMA contraction
 *     (build with -ffp-contract=off, no -ffast-math);
 *   - Mathf.F(x) == (float)F((double)x) for Sqrt/Exp/Log/Cos/Sin/Floor;
 *   - Vector2.magnitude == Mathf.Sqrt(x*x + y*y); normalized/Normalize return v/|v| when
 *     |v| > 1e-5, else zero;
 *   - Mathf.SmoothStep(a,b,t): t=clamp01(t); t=-2ttt+3tt; b*t + a*(1-t);
 *   - UnityEngine.Random.value is replaced by an explicit array of uniforms (4 per grid
 *     point, draw order of FFTMesh.cs:114-115) or by the Philox4x32-10 stream defined in
 *     mw_philox.h (shared verbatim with nothing in the product: the CUDA side has its own
 *     implementation of the same published algorithm).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "ref_philox.h"

/* FFTMesh.cs:50-54 */
static const float PI = 3.1415926536f;
static const float G = 9.81f;
static const float EPSILON = 0.0001f;

typedef struct {
    int32_t resolution; /* FFTMesh.cs:13 */
    float unit_width;   /* :15 */
    float length;       /* :19 */
    float choppiness;   /* :9  */
    float amplitude;    /* :23 */
    float wind_x;       /* :21 */
    float wind_y;
} ref_params;

/* ---- UnityEngine.Mathf / Vector semantics ---- */
static inline float mSqrt(float x) { return (float)sqrt((double)x); }
static inline float mExp(float x) { return (float)exp((double)x); }
static inline float mLog(float x) { return (float)log((double)x); }
static inline float mCos(float x) { return (float)cos((double)x); }
static inline float mSin(float x) { return (float)sin((double)x); }
static inline float mFloor(float x) { return (float)floor((double)x); }
static inline float mAbs(float x) { return fabsf(x); }
static inline float mMax(float a, float b) { return a > b ? a : b; }
static inline float mClamp01(float t) { return t < 0.f ? 0.f : (t > 1.f ? 1.f : t); }
static inline float mSmoothStep(float from, float to, float t)
{
    t = mClamp01(t);
    t = -2.0f * t * t * t + 3.0f * t * t;
    return to * t + from * (1.0f - t);
}
typedef struct { float x, y; } v2;
typedef struct { float x, y, z; } v3;
static inline float v2mag(v2 a) { return mSqrt(a.x * a.x + a.y * a.y); }
static inline v2 v2normalized(v2 a)
{
    float m = v2mag(a);
    v2 r = {0.f, 0.f};
    if (m > 1e-5f) { r.x = a.x / m; r.y = a.y / m; }
    return r;
}
static inline float v2dot(v2 a, v2 b) { return a.x * b.x + a.y * b.y; }
static inline v3 v3normalize(v3 a)
{
    float m = mSqrt(a.x * a.x + a.y * a.y + a.z * a.z);
    v3 r = {0.f, 0.f, 0.f};
    if (m > 1e-5f) { r.x = a.x / m; r.y = a.y / m; r.z = a.z / m; }
    return r;
}

/* FFTMesh.cs:141-147  Dispersion(n, m): omega quantised to multiples of w0 = 2 pi / L */
float ref_dispersion(const ref_params* p, int n, int m)
{
    float w = 2 * PI / (p->length);
    float kx = PI * (2 * n - p->resolution) / p->length;
    float kz = PI * (2 * m - p->resolution) / p->length;
    return mFloor(mSqrt(G * mSqrt(kx * kx + kz * kz)) / w) * w;
}

/* FFTMesh.cs:149-166  Phillips(n, m) */
float ref_phillips(const ref_params* p, int n, int m)
{
    v2 k = {(2 * n - p->resolution) / p->length * PI, (2 * m - p->resolution) / p->length * PI};
    float k_length = v2mag(k);
    if (k_length < EPSILON) return 0.0f;
    float k_length2 = k_length * k_length;
    float k_length4 = k_length2 * k_length2;

    v2 wind = {p->wind_x, p->wind_y};
    float kDotW = v2dot(v2normalized(k), v2normalized(wind));
    float kDotW2 = kDotW * kDotW;
    float w_length = v2mag(wind);
    float l = w_length * w_length / G;
    float l2 = l * l;
    float damping = 0.001f;
    float L2 = l2 * damping * damping;
    return p->amplitude * mExp(-1.f / (k_length2 * l2)) / k_length4 * kDotW2 * mExp(-k_length2 * L2);
}

/* FFTMesh.cs:168-176  htilde0(n, m) with the two uniforms passed in (Random.value x2) */
static v2 htilde0(const ref_params* p, int n, int m, float z1, float z2)
{
    v2 r;
    r.x = mSqrt(-2.f * mLog(z1)) * mCos(2 * PI * z2);
    r.y = mSqrt(-2.f * mLog(z1)) * mSin(2 * PI * z2);
    float s = mSqrt(ref_phillips(p, n, m) / 2.f);
    r.x = r.x * s;
    r.y = r.y * s;
    return r;
}

/* The stand-in for UnityEngine.Random.value: uniform k (0..3) of grid point idx. */
void ref_uniforms(uint64_t seed, int64_t idx, float u[4])
{
    uint32_t r[4];
    ref_philox4x32_10((uint32_t)(uint64_t)idx, (uint32_t)((uint64_t)idx >> 32), 0u, 0u,
                      (uint32_t)seed, (uint32_t)(seed >> 32), r);
    for (int k = 0; k < 4; ++k) u[k] = ref_u32_to_unit_open0(r[k]);
}

/*
 * FFTMesh.cs:101-116  GenerateMesh(): rest positions + h0 + h0conj.
 * uniforms: 4 floats per grid point in draw order (z1,z2 for h0; z1,z2 for h0conj), or NULL
 * to draw them from Philox(seed).  Outputs: vertices N*N*3, h0 N*N*2, h0conj N*N*2.
 */
void ref_generate_mesh(const ref_params* p, const float* uniforms, uint64_t seed,
                       float* vertices, float* h0, float* h0conj)
{
    int resolution = p->resolution;
    int halfResolution = resolution / 2;
    for (int i = 0; i < resolution; i++) {
        float horizontalPosition = (i - halfResolution) * p->unit_width;
        for (int j = 0; j < resolution; j++) {
            int currentIdx = i * resolution + j;
            float verticalPosition = (j - halfResolution) * p->unit_width;
            float off = (resolution % 2 == 0) ? p->unit_width / 2.f : 0.f;
            if (vertices) {
                vertices[3 * currentIdx + 0] = horizontalPosition + off;
                vertices[3 * currentIdx + 1] = 0.f;
                vertices[3 * currentIdx + 2] = verticalPosition + off;
            }
            float u[4];
            if (uniforms) memcpy(u, uniforms + 4 * (size_t)currentIdx, sizeof u);
            else ref_uniforms(seed, currentIdx, u);
            v2 a = htilde0(p, i, j, u[0], u[1]);
            v2 temp = htilde0(p, resolution - i, resolution - j, u[2], u[3]);
            h0[2 * currentIdx + 0] = a.x;
            h0[2 * currentIdx + 1] = a.y;
            h0conj[2 * currentIdx + 0] = temp.x;
            h0conj[2 * currentIdx + 1] = -temp.y;
        }
    }
}

/* FFTMesh.cs:178-190  htilde(t, n, m) */
static inline v2 htilde(const ref_params* p, const float* h0, const float* h0conj, float t, int n, int m)
{
    int index = n * p->resolution + m;
    v2 a = {h0[2 * index], h0[2 * index + 1]};
    v2 b = {h0conj[2 * index], h0conj[2 * index + 1]};
    float omegat = ref_dispersion(p, n, m) * t;
    float _cos = mCos(omegat);
    float _sin = mSin(omegat);
    v2 c0 = {_cos, _sin};
    v2 c1 = {_cos, -_sin};
    v2 res = {a.x * c0.x - a.y * c0.y + b.x * c1.x - b.y * c1.y,
              a.x * c0.y + a.y * c0.x + b.x * c1.y + b.y * c1.x};
    return res;
}

void ref_htilde(const ref_params* p, const float* h0, const float* h0conj, float t, float* out)
{
    int N = p->resolution;
    for (int n = 0; n < N; ++n)
        for (int m = 0; m < N; ++m) {
            v2 r = htilde(p, h0, h0conj, t, n, m);
            out[2 * (n * N + m)] = r.x;
            out[2 * (n * N + m) + 1] = r.y;
        }
}

/* FFTMesh.cs:192-220  Displacement(x, t, out nor): the O(N^2)-per-vertex direct sum */
static v3 displacement(const ref_params* p, const float* h0, const float* h0conj, v2 x, float t, v3* nor)
{
    int resolution = p->resolution;
    float length = p->length;
    v2 h = {0.f, 0.f};
    v2 d = {0.f, 0.f};
    v3 n = {0.f, 0.f, 0.f};
    for (int i = 0; i < resolution; i++) {
        float kx = 2 * PI * (i - resolution / 2.0f) / length;
        for (int j = 0; j < resolution; j++) {
            float kz = 2 * PI * (j - resolution / 2.0f) / length;
            v2 k = {kx, kz};
            float k_length = v2mag(k);
            float kDotX = v2dot(k, x);
            v2 c = {mCos(kDotX), mSin(kDotX)};
            v2 temp = htilde(p, h0, h0conj, t, i, j);
            v2 htilde_c = {temp.x * c.x - temp.y * c.y, temp.x * c.y + temp.y * c.x};
            h.x = h.x + htilde_c.x;
            h.y = h.y + htilde_c.y;
            n.x = n.x + (-kx * htilde_c.y);
            n.y = n.y + 0.f;
            n.z = n.z + (-kz * htilde_c.y);
            if (k_length < EPSILON) continue;
            d.x = d.x + (kx / k_length * htilde_c.y);
            d.y = d.y + (-kz / k_length * htilde_c.y);
        }
    }
    v3 up_minus_n = {0.f - n.x, 1.f - n.y, 0.f - n.z};
    *nor = v3normalize(up_minus_n);
    v3 r = {d.x, h.x, d.y};
    return r;
}

/*
 * FFTMesh.cs:224-249 (first half of EvaluateWaves) for the vertex range [v_begin, v_end):
 * vertMeow (displaced vertex), normals, hds.  threads<=1 is the reference's behaviour (Unity
 * runs Update() on the main thread); threads>1 splits the vertex loop with OpenMP and is
 * labelled "not reference behaviour" wherever it is reported.
 */
void ref_evaluate_vertices(const ref_params* p, const float* vertices, const float* h0, const float* h0conj,
                           float t, int64_t v_begin, int64_t v_end, int threads,
                           float* vertMeow, float* normals, float* hds)
{
#ifdef _OPENMP
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads)
#endif
    for (int64_t index = v_begin; index < v_end; ++index) {
        v3 nor = {0.f, 0.f, 0.f};
        v2 x = {vertices[3 * index + 0], vertices[3 * index + 2]};
        v3 hd = displacement(p, h0, h0conj, x, t, &nor);
        vertMeow[3 * index + 1] = hd.y;
        vertMeow[3 * index + 2] = vertices[3 * index + 2] - hd.z * p->choppiness;
        vertMeow[3 * index + 0] = vertices[3 * index + 0] - hd.x * p->choppiness;
        normals[3 * index + 0] = nor.x;
        normals[3 * index + 1] = nor.y;
        normals[3 * index + 2] = nor.z;
        hds[2 * index + 0] = hd.x;
        hds[2 * index + 1] = hd.z;
    }
    (void)threads;
}

/* FFTMesh.cs:253-276  whitecap colours from hds + normals (needs the full grid). */
void ref_whitecaps(const ref_params* p, const float* hds, const float* normals, float* jacobian_out, float* colors)
{
    int resolution = p->resolution;
    for (int i = 0; i < resolution; i++) {
        for (int j = 0; j < resolution; j++) {
            int index = i * resolution + j;
            v2 dDdx = {0.f, 0.f};
            v2 dDdy = {0.f, 0.f};
            if (i != resolution - 1) {
                dDdx.x = 0.5f * (hds[2 * index] - hds[2 * (index + resolution)]);
                dDdx.y = 0.5f * (hds[2 * index + 1] - hds[2 * (index + resolution) + 1]);
            }
            if (j != resolution - 1) {
                dDdy.x = 0.5f * (hds[2 * index] - hds[2 * (index + 1)]);
                dDdy.y = 0.5f * (hds[2 * index + 1] - hds[2 * (index + 1) + 1]);
            }
            float jacobian = (1 + dDdx.x) * (1 + dDdy.y) - dDdx.y * dDdy.x;
            v2 noise = {mAbs(normals[3 * index + 0]) * 0.3f, mAbs(normals[3 * index + 2]) * 0.3f};
            float turb = mMax(1.f - jacobian + v2mag(noise), 0.f);
            float xx = mSmoothStep(0.f, 1.f, turb); /* :271-272 are dead stores */
            if (jacobian_out) jacobian_out[index] = jacobian;
            colors[4 * index + 0] = xx;
            colors[4 * index + 1] = xx;
            colors[4 * index + 2] = xx;
            colors[4 * index + 3] = xx;
        }
    }
}

/* FFTMesh.cs:224-280  EvaluateWaves(t), whole grid. */
void ref_evaluate_waves(const ref_params* p, const float* vertices, const float* h0, const float* h0conj,
                        float t, int threads, float* vertMeow, float* normals, float* hds,
                        float* jacobian, float* colors)
{
    int64_t n2 = (int64_t)p->resolution * p->resolution;
    ref_evaluate_vertices(p, vertices, h0, h0conj, t, 0, n2, threads, vertMeow, normals, hds);
    ref_whitecaps(p, hds, normals, jacobian, colors);
}

int ref_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
