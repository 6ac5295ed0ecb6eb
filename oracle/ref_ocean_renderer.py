"""oracle/ref_ocean_renderer.py -- TEST INFRASTRUCTURE ONLY (CPU oracle, numpy).

Restatement of the reference's GPU-shader ocean path (SURVEY.md section 8 rows a11-a13 + a10), the
one the Ocean Demo scene runs: Scripts/OceanRenderer.cs GenerateTexture (:216-316) blitting
Shaders/FFT/{InitialSpectrum,Dispersion,Spectrum,SpectrumHeight,Stockham,OceanNormal,WhiteCap}.shader
with the helpers of Shaders/FFT/FFTCommon.cginc.  Every function cites the lines it follows.

Images are [y][x][channel] arrays: a fragment at texel (x, y) has texcoord = ((x + .5) / R, (y + .5) / R),
so `texcoord.x * _Resolution` is x + .5 exactly (R is a power of two) and GetWave's `n -= 0.5`
(FFTCommon.cginc:61-62) recovers the integer texel index.  "Horizontal" Stockham passes run along x
(the last spatial axis, contiguous in memory).

Two evaluations of the same chain:
  * dtype=np.float32 -- literal: every arithmetic step in IEEE fp32 in source order, the radix-2 Stockham
    blits stage by stage; transcendentals (sin, cos, exp, log, sqrt) are "fp64 libm of the fp32 argument,
    rounded to fp32" (the same convention as oracle/ref_fftmesh.c uses for Mathf);
  * dtype=np.float64 -- the same formulas in fp64 with numpy.fft.fft2 for the transform; this is what the
    CUDA path is held to (tolerances in tests/test_renderer_gpu.py), and the literal fp32 chain must pass
    the same check against it (tests/test_oracle_renderer.py).

Stated semantics the repo cannot verify (Unity 2017.2 / HLSL, closed source):
  * RenderTextures start black (phase = 0 on the first frame, OceanRenderer.cs:138-139) and sample with
    wrapMode = Clamp (Unity's RenderTexture default; nothing in OceanRenderer.cs changes it), so the +-1 /
    +-8 texel taps of OceanNormal / WhiteCap clamp at the image border; wrap="repeat" gives the periodic
    alternative (what a host that sets TextureWrapMode.Repeat would get);
  * taps land exactly on texel centres, so bilinear filtering returns the texel itself;
  * HLSL normalize(v) = v / sqrt(dot(v, v)); smoothstep(0, 1, t) = s*s*(3 - 2*s), s = saturate(t);
    fmod(a, b) = a - b * trunc(a / b);
  * GPU sin/cos/exp/log are implementation-defined approximations: UVRandom's frac(sin(x) * 43758.5453)
    (FFTCommon.cginc:37-41) amplifies their error by 4e4, so the reference's own h0 differs from GPU to GPU.
    The hash here (and in the CUDA init kernel) is defined with a correctly rounded sin; hosts that want their
    own spectrum upload it (mw_renderer_set_initial).

PARITY UNPINNED: the reference has no golden vectors for this path.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
F64 = np.float64


def _c(x, dtype):
    """A source-code literal in the working precision."""
    return dtype(x)


def _fn(name, x, dtype):
    """Transcendental of the working-precision argument: fp64 libm, rounded to the working precision."""
    return getattr(np, name)(np.asarray(x, dtype=F64)).astype(dtype)


def texcoords(R: int, dtype=F32):
    """texcoord.x, texcoord.y of every fragment ([R][R] each, [y][x])."""
    c = ((np.arange(R, dtype=F64) + 0.5) / R).astype(dtype)
    return np.broadcast_to(c[None, :], (R, R)), np.broadcast_to(c[:, None], (R, R))


# --------------------------------------------------------------------------------------------
# FFTCommon.cginc
# --------------------------------------------------------------------------------------------
def get_wave(n, m, length, res, dtype=F32):
    """FFTCommon.cginc:58-67 GetWave: FFT-ordered wave vector 2 PI (n, m) / len."""
    PI = _c(3.1415926536, dtype)  # :7
    n = (n - _c(0.5, dtype)).astype(dtype)
    m = (m - _c(0.5, dtype)).astype(dtype)
    half = _c(res, dtype) * _c(0.5, dtype)
    n = np.where(n < half, n, n - _c(res, dtype)).astype(dtype)
    m = np.where(m < half, m, m - _c(res, dtype)).astype(dtype)
    two_pi = _c(2, dtype) * PI
    return (two_pi * n / _c(length, dtype)).astype(dtype), (two_pi * m / _c(length, dtype)).astype(dtype)


def _length2(x, y, dtype):
    return _fn("sqrt", (x * x + y * y).astype(dtype), dtype)


def phillips(n, m, amp, wind, res, length, dtype=F32):
    """FFTCommon.cginc:69-85 Phillips (damping 0.01)."""
    G = _c(9.81, dtype)  # :9
    EPS = _c(0.0001, dtype)  # :8
    kx, kz = get_wave(n, m, length, res, dtype)
    klen = _length2(kx, kz, dtype)
    klen2 = (klen * klen).astype(dtype)
    klen4 = (klen2 * klen2).astype(dtype)
    live = klen >= EPS  # :75-76
    safe = np.where(live, klen, _c(1, dtype))
    wx, wy = _c(wind[0], dtype), _c(wind[1], dtype)
    wlen = _length2(wx, wy, dtype)
    nkx, nkz = (kx / safe).astype(dtype), (kz / safe).astype(dtype)  # normalize(k)
    nwx, nwy = dtype(wx / wlen), dtype(wy / wlen)  # normalize(wind)
    kdw = (nkx * nwx + nkz * nwy).astype(dtype)
    kdw2 = (kdw * kdw).astype(dtype)
    l = dtype(dtype(wlen * wlen) / G)
    l2 = dtype(l * l)
    damping = _c(0.01, dtype)
    L2 = dtype(dtype(l2 * damping) * damping)
    safe2 = np.where(live, klen2, _c(1, dtype))
    safe4 = np.where(live, klen4, _c(1, dtype))
    e1 = _fn("exp", (_c(-1, dtype) / (safe2 * l2).astype(dtype)).astype(dtype), dtype)
    e2 = _fn("exp", ((-safe2) * L2).astype(dtype), dtype)
    val = ((((_c(amp, dtype) * e1).astype(dtype) / safe4).astype(dtype) * kdw2).astype(dtype) * e2).astype(dtype)
    return np.where(live, val, _c(0, dtype)).astype(dtype)


def uv_random(u, v, salt, rnd, dtype=F32):
    """FFTCommon.cginc:37-41 UVRandom: frac(sin(dot(uv + (salt, random), (12.9898, 78.233))) * 43758.5453)."""
    uu = (u + _c(salt, dtype)).astype(dtype)
    vv = (v + _c(rnd, dtype)).astype(dtype)
    d = ((uu * _c(12.9898, dtype)).astype(dtype) + (vv * _c(78.233, dtype)).astype(dtype)).astype(dtype)
    s = (_fn("sin", d, dtype) * _c(43758.5453, dtype)).astype(dtype)
    return (s - np.floor(s)).astype(dtype)


def htilde0(u, v, r1, r2, phi, dtype=F32):
    """FFTCommon.cginc:87-99 hTilde0: Box-Muller on two hashes clamped to [0.01, 1], times sqrt(phi / 2)."""
    PI = _c(3.1415926536, dtype)
    rand1 = np.clip(uv_random(u, v, 10.612, r1, dtype), _c(0.01, dtype), _c(1, dtype))
    rand2 = np.clip(uv_random(u, v, 11.899, r2, dtype), _c(0.01, dtype), _c(1, dtype))
    x = _fn("sqrt", (_c(-2, dtype) * _fn("log", rand1, dtype)).astype(dtype), dtype)
    y = ((_c(2, dtype) * PI) * rand2).astype(dtype)
    s = _fn("sqrt", (phi / _c(2, dtype)).astype(dtype), dtype)
    return ((x * _fn("cos", y, dtype)).astype(dtype) * s).astype(dtype), ((x * _fn("sin", y, dtype)).astype(dtype) * s).astype(dtype)


def initial_spectrum(R, length, amplitude, wind, seed1, seed2, dtype=F32):
    """InitialSpectrum.shader:42-54 -> [R][R][4] = (h0.x, h0.y, h0conj.x, h0conj.y).

    `amplitude` is the material's _Amplitude, i.e. OceanRenderer.amplitude / 10000 (OceanRenderer.cs:149).
    Note :47: Phillips(_Resolution - n, _Resolution - m) with n = x + .5 reaches texel index R - 1 - x.
    """
    u, v = texcoords(R, dtype)
    n = (u * _c(R, dtype)).astype(dtype)
    m = (v * _c(R, dtype)).astype(dtype)
    phi1 = phillips(n, m, amplitude, wind, R, length, dtype)
    phi2 = phillips((_c(R, dtype) - n).astype(dtype), (_c(R, dtype) - m).astype(dtype), amplitude, wind, R, length, dtype)
    s1, s2 = _c(seed1, dtype), _c(seed2, dtype)
    ax, ay = htilde0(u, v, dtype(s1 / _c(2, dtype)), dtype(s2 * _c(2, dtype)), phi1, dtype)  # :49
    bx, by = htilde0(u, v, s1, s2, phi2, dtype)  # :50
    return np.stack([ax, ay, bx, -by], -1).astype(dtype)


def dispersion_rate(R, length, dtype=F32):
    """FFTCommon.cginc:106-114 CalcDispersion without the `* dt`: sqrt(G |k| (1 + |k|^2 / 370 / 370))."""
    G = _c(9.81, dtype)
    u, v = texcoords(R, dtype)
    kx, kz = get_wave((u * _c(R, dtype)).astype(dtype), (v * _c(R, dtype)).astype(dtype), length, R, dtype)
    wlen = _length2(kx, kz, dtype)
    cap = (_c(1, dtype) + ((wlen * wlen).astype(dtype) / _c(370, dtype)).astype(dtype) / _c(370, dtype)).astype(dtype)
    return _fn("sqrt", ((G * wlen).astype(dtype) * cap).astype(dtype), dtype)


def dispersion_step(phase, R, length, dt, dtype=F32):
    """Dispersion.shader:32-41 + GetDispersion (FFTCommon.cginc:101-104): fmod(phase + rate * dt, 2 PI)."""
    PI = _c(3.1415926536, dtype)
    two_pi = _c(2, dtype) * PI
    s = (phase.astype(dtype) + (dispersion_rate(R, length, dtype) * _c(dt, dtype)).astype(dtype)).astype(dtype)
    return (s - two_pi * np.trunc(s / two_pi)).astype(dtype)


def _h_of_t(initial, phase, dtype):
    """Spectrum.shader:40-45 / SpectrumHeight.shader:40-45: h = h0 * pv + h0conj * conj(pv), pv = e^{i phase}."""
    c, s = _fn("cos", phase, dtype), _fn("sin", phase, dtype)
    a, b, p, q = (initial[..., i].astype(dtype) for i in range(4))
    hx = ((a * c - b * s).astype(dtype) + (p * c + q * s).astype(dtype)).astype(dtype)  # MultComplex, FFTCommon.cginc:43-46
    hy = ((a * s + b * c).astype(dtype) + (q * c - p * s).astype(dtype)).astype(dtype)
    return hx, hy


def spectrum(initial, phase, R, length, choppiness, dtype=F32):
    """Spectrum.shader:34-51 -> [R][R][4] = (hx, hz), h{x,z} = -MultByI(h * wave.{x,y} / w) * _Choppiness."""
    u, v = texcoords(R, dtype)
    kx, kz = get_wave((u * _c(R, dtype)).astype(dtype), (v * _c(R, dtype)).astype(dtype), length, R, dtype)
    w = np.maximum(_c(0.0001, dtype), _length2(kx, kz, dtype))  # :47
    hr, hi = _h_of_t(initial, phase, dtype)
    ch = _c(choppiness, dtype)
    out = np.empty((R, R, 4), dtype)
    for o, k in ((0, kx), (2, kz)):
        tr = ((hr * k).astype(dtype) / w).astype(dtype)
        ti = ((hi * k).astype(dtype) / w).astype(dtype)
        out[..., o] = (ti * ch).astype(dtype)       # -MultByI(t) = (t.y, -t.x)
        out[..., o + 1] = ((-tr) * ch).astype(dtype)
    return out


def spectrum_height(initial, phase, dtype=F32):
    """SpectrumHeight.shader:34-47 -> [R][R][4] = (h, h)."""
    hr, hi = _h_of_t(initial, phase, dtype)
    return np.stack([hr, hi, hr, hi], -1).astype(dtype)


# --------------------------------------------------------------------------------------------
# Stockham.shader + its schedule
# --------------------------------------------------------------------------------------------
def stockham_blit(tex, sub, horizontal, dtype=F32):
    """Stockham.shader:31-57, one blit on an [R][R][4] image (two complex fields xy / zw)."""
    R = tex.shape[0]
    index = np.arange(R)
    even = (index // sub) * (sub // 2) + index % (sub // 2)  # :41
    ang = (_c(-2, dtype) * _c(3.1415926536, dtype) * (index.astype(dtype) / _c(sub, dtype)).astype(dtype)).astype(dtype)  # :51, GetTwiddle
    tc, ts = _fn("cos", ang, dtype), _fn("sin", ang, dtype)
    if horizontal:
        ev, od = tex[:, even, :], tex[:, even + R // 2, :]
        tc, ts = tc[None, :], ts[None, :]
    else:
        ev, od = tex[even, :, :], tex[even + R // 2, :, :]
        tc, ts = tc[:, None], ts[:, None]
    out = np.empty_like(tex, dtype=dtype)
    for o in (0, 2):  # outputA = even.xy + MultComplex(twiddle, odd.xy); outputB likewise on .zw
        out[..., o] = (ev[..., o] + (tc * od[..., o] - ts * od[..., o + 1]).astype(dtype)).astype(dtype)
        out[..., o + 1] = (ev[..., o + 1] + (tc * od[..., o + 1] + ts * od[..., o]).astype(dtype)).astype(dtype)
    return out


def stockham_chain(tex, dtype=F32):
    """OceanRenderer.cs:229-262: log2 R horizontal blits, then log2 R vertical ones (literal, any dtype)."""
    R = tex.shape[0]
    stages = int(np.log2(R))
    y = tex.astype(dtype)
    for horizontal in (True, False):
        for s in range(stages):
            y = stockham_blit(y, 2 ** (s + 1), horizontal, dtype)
    return y


def transform(tex, dtype=F32):
    """The Stockham chain: literal blits in fp32, numpy.fft.fft2 in fp64 (the same forward un-normalised DFT)."""
    if dtype == F32:
        return stockham_chain(tex, F32)
    z = np.stack([tex[..., 0] + 1j * tex[..., 1], tex[..., 2] + 1j * tex[..., 3]])
    Z = np.fft.fft2(z.astype(np.complex128))
    return np.stack([Z[0].real, Z[0].imag, Z[1].real, Z[1].imag], -1)


# --------------------------------------------------------------------------------------------
# OceanNormal.shader, WhiteCap.shader
# --------------------------------------------------------------------------------------------
def _tap(img, dx, dy, wrap):
    """tex2D(img, texcoord + (dx, dy) texels) for every fragment."""
    R = img.shape[0]
    ix = np.arange(R) + dx
    iy = np.arange(R) + dy
    if wrap == "clamp":
        ix, iy = np.clip(ix, 0, R - 1), np.clip(iy, 0, R - 1)
    elif wrap == "repeat":
        ix, iy = ix % R, iy % R
    else:
        raise ValueError(wrap)
    return img[iy][:, ix]


def ocean_normal(disp, height, R, length, wrap="clamp", dtype=F32):
    """OceanNormal.shader:32-56 -> [R][R][4] = (normalize(sum of four cross products), 1)."""
    ts = dtype(_c(length, dtype) / _c(R, dtype))  # :42 texelSize
    center = disp[..., 0:3].astype(dtype)  # :44 -- .rgb: (Re hx, Im hx, Re hz), as written

    def vec(dx, dy):  # GetVec :32-37
        d, h = _tap(disp, dx, dy, wrap), _tap(height, dx, dy, wrap)
        return np.stack([d[..., 0], h[..., 0], d[..., 2]], -1).astype(dtype)

    def off(x, z):
        return np.array([x, 0, z], dtype)

    right = (off(ts, 0) + vec(1, 0)).astype(dtype) - center
    left = (off(-ts, 0) + vec(-1, 0)).astype(dtype) - center
    top = (off(0, -ts) + vec(0, -1)).astype(dtype) - center
    bottom = (off(0, ts) + vec(0, 1)).astype(dtype) - center

    def cross(a, b):
        return np.stack([(a[..., 1] * b[..., 2]).astype(dtype) - (a[..., 2] * b[..., 1]).astype(dtype),
                         (a[..., 2] * b[..., 0]).astype(dtype) - (a[..., 0] * b[..., 2]).astype(dtype),
                         (a[..., 0] * b[..., 1]).astype(dtype) - (a[..., 1] * b[..., 0]).astype(dtype)], -1).astype(dtype)

    s = (((cross(right, top) + cross(top, left)).astype(dtype) + cross(left, bottom)).astype(dtype) + cross(bottom, right)).astype(dtype)
    mag = _fn("sqrt", ((s[..., 0] * s[..., 0] + s[..., 1] * s[..., 1]).astype(dtype) + s[..., 2] * s[..., 2]).astype(dtype), dtype)
    n = (s / mag[..., None]).astype(dtype)
    return np.concatenate([n, np.ones((R, R, 1), dtype)], -1)


def white_cap(disp, bump, R, mesh_resolution, wrap="clamp", dtype=F32):
    """WhiteCap.shader:33-45 -> [R][R] (the R channel; ColorMask R).

    texelSize = 1 / _Length with _Length = the *mesh* resolution (OceanRenderer.cs:306) while the image is
    R = 8 * mesh_resolution wide (:136): the taps sit R / mesh_resolution texels away.
    """
    step = R // int(mesh_resolution)
    half, eight = _c(-0.5, dtype), _c(8, dtype)
    rb = (0, 2)
    dy = ((half * (_tap(disp, 0, -step, wrap)[..., rb] - _tap(disp, 0, step, wrap)[..., rb]).astype(dtype)).astype(dtype) / eight).astype(dtype)  # :35
    dx = ((half * (_tap(disp, -step, 0, wrap)[..., rb] - _tap(disp, step, 0, wrap)[..., rb]).astype(dtype)).astype(dtype) / eight).astype(dtype)  # :36
    nx = (_c(0.3, dtype) * bump[..., 0]).astype(dtype)  # :37 noise = 0.3 * bump.xz
    nz = (_c(0.3, dtype) * bump[..., 2]).astype(dtype)
    one = _c(1, dtype)
    jac = (((one + dx[..., 0]) * (one + dy[..., 1])).astype(dtype) - (dx[..., 1] * dy[..., 0]).astype(dtype)).astype(dtype)  # :38
    turb = np.maximum(_c(0, dtype), ((one - jac).astype(dtype) + _length2(nx, nz, dtype)).astype(dtype))  # :39
    s = np.clip(turb, _c(0, dtype), one)  # :42 smoothstep(0, 1, turb) -- :40-41 are overwritten
    return ((s * s).astype(dtype) * (_c(3, dtype) - (_c(2, dtype) * s).astype(dtype)).astype(dtype)).astype(dtype), jac


# --------------------------------------------------------------------------------------------
# OceanRenderer.GenerateTexture
# --------------------------------------------------------------------------------------------
class RendererState:
    """What OceanRenderer keeps between frames: the initial-spectrum image and the phase image."""

    def __init__(self, mesh_resolution, length, choppiness, amplitude, wind, seed1, seed2, mult=2.0, dtype=F32,
                 wrap="clamp", initial=None):
        self.mesh_resolution = int(mesh_resolution)
        self.R = 8 * self.mesh_resolution  # OceanRenderer.cs:136
        self.length, self.choppiness, self.mult = length, choppiness, mult
        self.dtype, self.wrap = dtype, wrap
        # RenderInitial (:209-214) with _Amplitude = amplitude / 10000 (:149)
        amp = F32(amplitude) / F32(10000.0)
        self.initial = (initial_spectrum(self.R, length, amp, wind, seed1, seed2, F32) if initial is None
                        else np.asarray(initial, F32).reshape(self.R, self.R, 4))
        self.phase = np.zeros((self.R, self.R), dtype)  # ping/pong phase textures start black

    def generate_texture(self, delta_time):
        """OceanRenderer.cs:216-316 -> dict(displacement, height, normal, white[, jacobian])."""
        dt = F32(F32(delta_time) * F32(self.mult))  # :223
        d, R = self.dtype, self.R
        self.phase = dispersion_step(self.phase, R, self.length, dt, d)  # :220-224
        disp = transform(spectrum(self.initial, self.phase, R, self.length, self.choppiness, d), d)  # :226-262
        height = transform(spectrum_height(self.initial, self.phase, d), d)  # :264-298
        normal = ocean_normal(disp, height, R, self.length, self.wrap, d)  # :300-302
        white, jac = white_cap(disp, normal, R, self.mesh_resolution, self.wrap, d)  # :303-307
        return {"displacement": disp, "height": height, "normal": normal, "white": white, "jacobian": jac,
                "phase": self.phase.copy()}


def generate_mesh(resolution, unit_width):
    """OceanRenderer.cs:172-207 (identical to FFTMesh.cs:101-139 minus the spectrum): vertices, normals, uvs, indices."""
    N = int(resolution)
    uw = F32(unit_width)
    half = N // 2
    off = uw / F32(2) if N % 2 == 0 else F32(0)
    pos = ((np.arange(N) - half).astype(F32) * uw + off).astype(F32)
    vertices = np.zeros((N, N, 3), F32)
    vertices[..., 0] = pos[:, None]
    vertices[..., 2] = pos[None, :]
    normals = np.zeros((N, N, 3), F32)
    normals[..., 1] = 1
    frac = (np.arange(N).astype(F32) * F32(1.0) / F32(N - 1)).astype(F32)
    uvs = np.stack(np.broadcast_arrays(frac[:, None], frac[None, :]), -1).astype(F32)
    indices = []
    for i in range(N):
        for j in range(N - 1):
            cur = i * N + j
            if i != N - 1:
                indices += [cur, cur + 1, cur + N]
            if i != 0:
                indices += [cur, cur - N + 1, cur + 1]
    return vertices.reshape(-1, 3), normals.reshape(-1, 3), uvs.reshape(-1, 2), np.asarray(indices, np.int32)
