/*
 * oracle/ref_gerstner.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * Literal C restatement of the pond renderer's Gerstner displacement, read from
 *   /root/reference/Assets/Mistral Water/Shaders/MistralWaterLib.cginc
 *     Gerstner          :71-99   (4 waves, uniforms _WSpeed/_WDirectionAB/_WDirectionCD)
 *     GerstnerLevelOne  :101-125 (5 waves, constant tables :105-109)
 *     Displacement      :154-180 (sVertex = worldPos.xzz, amplitude * 0.01, vertex += offsets)
 * plus the W-wave generalisation (ref_gerstner_table) both of them are special cases of.
 * `half` is fp32 on desktop targets; arithmetic is fp32 in source order, cos/sin evaluated in
 * double and rounded (the shader compiler's own transcendental accuracy is not specified, so
 * parity for this path is a tolerance, stated in tests/test_gerstner.py).
 *
 * PARITY UNPINNED: the reference has no tests or golden vectors for this path.
 */
#include <math.h>
#include <stdint.h>

static inline float fcos(float x) { return (float)cos((double)x); }
static inline float fsin(float x) { return (float)sin((double)x); }

/* MistralWaterLib.cginc:71-99 ; pos_xyz is the world position, sVertex.xz = (pos.x, pos.z) */
void ref_gerstner4(const float* pos_xyz, int64_t n, float time_y,
                   float amplitude, float frequency, float steepness,
                   const float speed[4], const float dirAB[4], const float dirCD[4],
                   float* offsets_xyz)
{
    for (int64_t v = 0; v < n; ++v) {
        float sx = pos_xyz[3 * v + 0], sz = pos_xyz[3 * v + 2];
        float AB[4], CD[4];
        for (int k = 0; k < 4; ++k) { AB[k] = steepness * amplitude * dirAB[k]; CD[k] = steepness * amplitude * dirCD[k]; }
        float dotABCD[4] = {
            frequency * (dirAB[0] * sx + dirAB[1] * sz),
            frequency * (dirAB[2] * sx + dirAB[3] * sz),
            frequency * (dirCD[0] * sx + dirCD[1] * sz),
            frequency * (dirCD[2] * sx + dirCD[3] * sz)};
        float COS[4], SIN[4];
        for (int k = 0; k < 4; ++k) {
            float t = time_y * speed[k];
            COS[k] = fcos(dotABCD[k] + t);
            SIN[k] = fsin(dotABCD[k] + t);
        }
        /* offs.x = dot(COS, (AB.x, AB.z, CD.x, CD.z)); offs.z = dot(COS, (AB.y, AB.w, CD.y, CD.w)) */
        float ox = COS[0] * AB[0] + COS[1] * AB[2] + COS[2] * CD[0] + COS[3] * CD[2];
        float oz = COS[0] * AB[1] + COS[1] * AB[3] + COS[2] * CD[1] + COS[3] * CD[3];
        float oy = SIN[0] * amplitude + SIN[1] * amplitude + SIN[2] * amplitude + SIN[3] * amplitude;
        offsets_xyz[3 * v + 0] = ox;
        offsets_xyz[3 * v + 1] = oy;
        offsets_xyz[3 * v + 2] = oz;
    }
}

/* MistralWaterLib.cginc:101-125 */
void ref_gerstner_level_one(const float* pos_xyz, int64_t n, float time_y,
                            float amplitude, float frequency, float steepness, float* offsets_xyz)
{
    static const float amps[5] = {0.7f, 0.6f, 0.6f, 0.7f, 0.9f};
    static const float steeps[5] = {0.95f, 0.615f, 0.821f, 0.462f, 0.611f};
    static const float speeds[5] = {-2.112f, 0.6124f, -0.878f, -3.6234f, 1.f};
    static const float dir[5][2] = {{1.f, -0.2f}, {-0.9f, 1.f}, {0.2f, 0.2f}, {-1.0f, 0.77f}, {0.99f, -1.145f}};
    static const float fs[5] = {0.954f, 1.52f, 0.44f, 0.21f, 0.8f};
    for (int64_t v = 0; v < n; ++v) {
        float sx = pos_xyz[3 * v + 0], sz = pos_xyz[3 * v + 2];
        float ox = 0.f, oy = 0.f, oz = 0.f;
        for (int i = 0; i < 5; ++i) {
            float d = sx * dir[i][0] + sz * dir[i][1];
            float th = frequency * fs[i] * d + speeds[i] * frequency * fs[i] * time_y;
            ox += steepness * amplitude * steeps[i] * amps[i] * dir[i][0] * fcos(th);
            oz += steepness * amplitude * steeps[i] * amps[i] * dir[i][1] * fcos(th);
            oy += amplitude * amps[i] * fsin(th);
        }
        offsets_xyz[3 * v + 0] = ox;
        offsets_xyz[3 * v + 1] = oy;
        offsets_xyz[3 * v + 2] = oz;
    }
}

/*
 * W-wave table form.  Per wave w: theta = freq[w] * (dir.x*x + dir.y*z) + rate[w] * t;
 * offs.x += amp_xz[w]*dir.x*cos; offs.z += amp_xz[w]*dir.y*cos; offs.y += amp_y[w]*sin.
 *   Gerstner        : freq=frequency, rate=speed_w, amp_xz=steepness*amplitude, amp_y=amplitude
 *   GerstnerLevelOne: freq=frequency*fs_i, rate=speeds_i*frequency*fs_i,
 *                     amp_xz=steepness*amplitude*steeps_i*amps_i, amp_y=amplitude*amps_i
 * wave layout: 6 floats {dir_x, dir_y, freq, rate, amp_xz, amp_y}.
 * out_xyz = pos + offsets (MistralWaterLib.cginc:176); out_nrm = (0,1,0) (:98, :121) if non-NULL.
 */
void ref_gerstner_table(const float* waves, int n_waves, const float* pos_xyz, int64_t n, float t,
                        float* out_xyz, float* out_nrm)
{
    for (int64_t v = 0; v < n; ++v) {
        float sx = pos_xyz[3 * v + 0], sz = pos_xyz[3 * v + 2];
        float ox = 0.f, oy = 0.f, oz = 0.f;
        for (int w = 0; w < n_waves; ++w) {
            const float* W = waves + 6 * w;
            float th = W[2] * (W[0] * sx + W[1] * sz) + W[3] * t;
            float c = fcos(th), s = fsin(th);
            ox += W[4] * W[0] * c;
            oz += W[4] * W[1] * c;
            oy += W[5] * s;
        }
        out_xyz[3 * v + 0] = pos_xyz[3 * v + 0] + ox;
        out_xyz[3 * v + 1] = pos_xyz[3 * v + 1] + oy;
        out_xyz[3 * v + 2] = pos_xyz[3 * v + 2] + oz;
        if (out_nrm) { out_nrm[3 * v + 0] = 0.f; out_nrm[3 * v + 1] = 1.f; out_nrm[3 * v + 2] = 0.f; }
    }
}

/*
 * The two normals the reference has code for but does not ship (both variants end with normal = (0, 1, 0), :98 / :121):
 *   mode 2 "analytic":  the exact normal of the displaced surface P(x, z) = (x + offs.x, offs.y, z + offs.z) the table form
 *                       defines, normalize(dP/dz x dP/dx) -- what the commented attempt at :122-124 was after; evaluated in
 *                       double here (it is a derived quantity, parity is a tolerance);
 *   mode 3 "discarded": Gerstner() :92-97 literally, fp32 in source order:
 *                       normal = (0,2,0); normal.x -= offs.x; normal.y -= offs.z; normal.xz *= _Smoothing; normalize.
 */
void ref_gerstner_table_normals(const float* waves, int n_waves, const float* pos_xyz, int64_t n, float t, int mode,
                                float smoothing, float* out_nrm)
{
    for (int64_t v = 0; v < n; ++v) {
        float sx = pos_xyz[3 * v + 0], sz = pos_xyz[3 * v + 2];
        if (mode == 3) {
            float ox = 0.f, oz = 0.f;
            for (int w = 0; w < n_waves; ++w) {
                const float* W = waves + 6 * w;
                float th = W[2] * (W[0] * sx + W[1] * sz) + W[3] * t;
                float c = fcos(th);
                ox += W[4] * W[0] * c;
                oz += W[4] * W[1] * c;
            }
            float nx = 0.f, ny = 2.f, nz = 0.f;
            nx -= ox;
            ny -= oz;
            nx *= smoothing; nz *= smoothing;
            float len = (float)sqrt((double)(nx * nx + ny * ny + nz * nz));
            out_nrm[3 * v + 0] = nx / len; out_nrm[3 * v + 1] = ny / len; out_nrm[3 * v + 2] = nz / len;
        } else {
            double jxx = 0, jxz = 0, jzz = 0, hx = 0, hz = 0;
            for (int w = 0; w < n_waves; ++w) {
                const float* W = waves + 6 * w;
                float th = W[2] * (W[0] * sx + W[1] * sz) + W[3] * t;   /* the same fp32 phase the displacement uses */
                double s = sin((double)th), c = cos((double)th), f = W[2];
                jxx += (double)W[4] * W[0] * W[0] * f * s;
                jxz += (double)W[4] * W[0] * W[1] * f * s;
                jzz += (double)W[4] * W[1] * W[1] * f * s;
                hx += (double)W[5] * W[0] * f * c;
                hz += (double)W[5] * W[1] * f * c;
            }
            double a[3] = {-jxz, hz, 1.0 - jzz}, b[3] = {1.0 - jxx, hx, -jxz};
            double c3[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
            double len = sqrt(c3[0] * c3[0] + c3[1] * c3[1] + c3[2] * c3[2]);
            out_nrm[3 * v + 0] = (float)(c3[0] / len); out_nrm[3 * v + 1] = (float)(c3[1] / len); out_nrm[3 * v + 2] = (float)(c3[2] / len);
        }
    }
}

/* MistralWaterLib.cginc:127-152 Wave + its call site Displacement :160-164 (keyword _DISPLACEMENTMODE_WAVE), with
 * unity_ObjectToWorld = unity_WorldToObject = identity (a pond mesh placed at the origin, unrotated, unscaled):
 *   sVertex = worldPos = vertex;  out.y = vertex.y + offsets.y  where offsets = displaced v0 (so out.y = 2 y + wave);
 *   normal = normalize(cross(v2 - v0, v1 - v0)) of the 0.05-offset neighbours after the _Smoothing blend. */
void ref_wave(const float* pos_xyz, int64_t n, float time_y, float amplitude, float frequency, float s, float smoothing,
              float* out_xyz, float* out_nrm)
{
    for (int64_t v = 0; v < n; ++v) {
        float v0[3] = {pos_xyz[3 * v], pos_xyz[3 * v + 1], pos_xyz[3 * v + 2]};
        float v1[3] = {v0[0] + 0.05f, v0[1], v0[2]};
        float v2[3] = {v0[0], v0[1], v0[2] + 0.05f};
        float speed = s * time_y;
        float amp = amplitude * 0.01f;
        v0[1] += fsin(speed + (v0[0] * frequency)) * amp;
        v1[1] += fsin(speed + (v1[0] * frequency)) * amp;
        v2[1] += fsin(speed + (v2[0] * frequency)) * amp;
        v0[1] -= fcos(speed + (v0[2] * frequency)) * amp;
        v1[1] -= fcos(speed + (v1[2] * frequency)) * amp;
        v2[1] -= fcos(speed + (v2[2] * frequency)) * amp;
        v1[1] -= (v1[1] - v0[1]) * (1.0f - smoothing);
        v2[1] -= (v2[1] - v0[1]) * (1.0f - smoothing);
        float a[3] = {v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2]};
        float b[3] = {v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2]};
        float c[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
        float len = (float)sqrt((double)(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]));
        out_xyz[3 * v + 0] = pos_xyz[3 * v];
        out_xyz[3 * v + 1] = pos_xyz[3 * v + 1] + v0[1];   /* v.vertex.y += offsets.y (:162) */
        out_xyz[3 * v + 2] = pos_xyz[3 * v + 2];
        if (out_nrm) { out_nrm[3 * v] = c[0] / len; out_nrm[3 * v + 1] = c[1] / len; out_nrm[3 * v + 2] = c[2] / len; }
    }
}
