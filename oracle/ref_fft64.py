"""oracle/ref_fft64.py -- TEST INFRASTRUCTURE ONLY (CPU oracle, numpy).

fp64 restatement of the reference's CPU ocean path (FFTMesh.cs, cited per function) in its
*transform* form, so that sizes where the literal O(N^4) loop of oracle/ref_fftmesh.c costs
minutes-to-hours can still be checked.  Everything FFTMesh.cs defines in fp32 *before* the
direct sum (omega, omega*t, the k multipliers) is reproduced in strict fp32 here; the sum itself
(and everything downstream) is done in fp64.

The identity used (checked against the literal loop in
tests/test_oracle.py): for even N (N % 4 == 0) and length == resolution * unitWidth,

    S[a,b] = sum_{n,m} G'[n,m] * exp(i (k_n x_a + k_m z_b))
           = sigma[a,b] * N^2 * ifft2(G'[n,m] * r[n] * r[m])[a,b]
    r[n] = exp(i pi n (1 - N) / N),  sigma[a,b] = -(-1)^(a+b)

with k_n = 2 pi (n - N/2) / L (FFTMesh.cs:201) and x_a = (a - N/2 + 1/2) * unitWidth (:107,112).

PARITY UNPINNED: the reference has no golden vectors; see oracle/ref_fftmesh.c header.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
PI = F32(3.1415926536)  # FFTMesh.cs:50
G = F32(9.81)  # :52
EPSILON = F32(0.0001)  # :54


def omega_f32(N: int, L: float) -> np.ndarray:
    """FFTMesh.cs:141-147 Dispersion(n, m) for the whole grid, strict fp32 -> [N, N] float32."""
    L = F32(L)
    w = F32(2) * PI / L
    idx = (2 * np.arange(N, dtype=np.int64) - N).astype(F32)
    k1 = PI * idx / L  # PI * (2n - N) / length
    kx = k1[:, None]
    kz = k1[None, :]
    mag = np.sqrt(kx * kx + kz * kz, dtype=F32)
    return (np.floor(np.sqrt(G * mag, dtype=F32) / w) * w).astype(F32)


def k_displacement_f32(N: int, L: float) -> np.ndarray:
    """FFTMesh.cs:201/204: kx = 2 * PI * (i - N / 2.0f) / length, strict fp32 -> [N] float32."""
    i = np.arange(N, dtype=np.int64).astype(F32)
    return (F32(2) * PI * (i - F32(N) / F32(2.0)) / F32(L)).astype(F32)


def htilde(h0: np.ndarray, h0conj: np.ndarray, N: int, L: float, t: float) -> np.ndarray:
    """FFTMesh.cs:178-190: h(k,t) = h0 e^{i w t} + h0conj e^{-i w t} -> [N, N] complex128.

    omegat is the fp32 product Dispersion(n,m) * t exactly as in :183; cos/sin are then taken
    in fp64 (the reference rounds them to fp32, a <= 6e-8 relative perturbation).
    """
    om_t = (omega_f32(N, L) * F32(t)).astype(F32).astype(np.float64)
    e = np.cos(om_t) + 1j * np.sin(om_t)
    a = h0.reshape(N, N, 2).astype(np.float64)
    b = h0conj.reshape(N, N, 2).astype(np.float64)
    return (a[..., 0] + 1j * a[..., 1]) * e + (b[..., 0] + 1j * b[..., 1]) * np.conj(e)


def _ramp(N: int) -> np.ndarray:
    n = np.arange(N, dtype=np.float64)
    return np.exp(1j * np.pi * n * (1.0 - N) / N)


def direct_transform(Gs: np.ndarray) -> np.ndarray:
    """S[a,b] = sum_{n,m} G[n,m] e^{i(k_n x_a + k_m z_b)} (any leading batch dims) via ifft2."""
    N = Gs.shape[-1]
    assert Gs.shape[-2] == N and N % 4 == 0
    r = _ramp(N)
    sgn = 1.0 - 2.0 * (np.arange(N) % 2)
    sigma = -(sgn[:, None] * sgn[None, :])
    return sigma * (N * N) * np.fft.ifft2(Gs * (r[:, None] * r[None, :]), axes=(-2, -1))


def fields(h0: np.ndarray, h0conj: np.ndarray, N: int, L: float, t: float) -> dict:
    """FFTMesh.cs:192-220 Displacement() for every vertex, in transform form (fp64).

    Returns height, dx, dz (the `d` accumulators, i.e. hds), sx, sz (= -n.x, -n.z so that the
    un-normalised normal is (sx, 1, sz)).
    """
    H = htilde(h0, h0conj, N, L, t)
    k1 = k_displacement_f32(N, L)
    kx = np.broadcast_to(k1[:, None], (N, N))
    kz = np.broadcast_to(k1[None, :], (N, N))
    klen = np.sqrt(kx * kx + kz * kz, dtype=F32)  # Vector2.magnitude in fp32 (:206)
    live = klen >= EPSILON  # :213-214
    safe = np.where(live, klen, F32(1))
    ux = np.where(live, (kx / safe).astype(F32), F32(0)).astype(np.float64)  # kx / k_length (:215)
    uz = np.where(live, (kz / safe).astype(F32), F32(0)).astype(np.float64)
    stack = np.stack([H, ux * H, uz * H, kx.astype(np.float64) * H, kz.astype(np.float64) * H])
    S = direct_transform(stack)
    return {
        "height": S[0].real,  # :211, :219
        "dx": S[1].imag,  # :215  d.x += kx/|k| * Im
        "dz": -S[2].imag,  # :215  d.y += -kz/|k| * Im
        "sx": S[3].imag,  # :212  n.x += -kx * Im  => up - n = (+sum kx Im, 1, +sum kz Im)
        "sz": S[4].imag,
    }


def evaluate_waves(h0, h0conj, N: int, L: float, unit_width: float, choppiness: float, t: float) -> dict:
    """FFTMesh.cs:224-276 EvaluateWaves(t) in fp64: vertMeow, normals, hds, jacobian, colors."""
    f = fields(h0, h0conj, N, L, t)
    uw = F32(unit_width)
    half = N // 2
    pos = ((np.arange(N) - half).astype(F32) * uw + uw / F32(2)).astype(np.float64)  # :107,111-112
    vx = np.broadcast_to(pos[:, None], (N, N))
    vz = np.broadcast_to(pos[None, :], (N, N))
    vert = np.stack([vx - f["dx"] * float(F32(choppiness)), f["height"], vz - f["dz"] * float(F32(choppiness))], -1)
    mag = np.sqrt(f["sx"] ** 2 + 1.0 + f["sz"] ** 2)
    normals = np.stack([f["sx"] / mag, 1.0 / mag, f["sz"] / mag], -1)  # :218
    hds = np.stack([f["dx"], f["dz"]], -1)  # :247
    dDdx = np.zeros((N, N, 2))
    dDdy = np.zeros((N, N, 2))
    dDdx[:-1] = 0.5 * (hds[:-1] - hds[1:])  # :260-263 (index + resolution = next i)
    dDdy[:, :-1] = 0.5 * (hds[:, :-1] - hds[:, 1:])  # :264-267
    jac = (1 + dDdx[..., 0]) * (1 + dDdy[..., 1]) - dDdx[..., 1] * dDdy[..., 0]  # :268
    noise = 0.3 * np.sqrt(normals[..., 0] ** 2 + normals[..., 2] ** 2)  # :269
    turb = np.maximum(1.0 - jac + noise, 0.0)  # :270
    tt = np.clip(turb, 0.0, 1.0)
    white = -2.0 * tt ** 3 + 3.0 * tt ** 2  # :273 Mathf.SmoothStep(0, 1, turb)
    return {
        "height": f["height"], "hds": hds, "normals": normals, "vertMeow": vert,
        "jacobian": jac, "whitecap": white, "colors": np.repeat(white[..., None], 4, -1),
    }


# --------------------------------------------------------------------------------------------
# Stockham.shader restatement (a10): one radix-2 autosort stage as the fragment program does it.
# --------------------------------------------------------------------------------------------
def stockham_stage(x: np.ndarray, sub: int, axis: int = -1) -> np.ndarray:
    """Stockham.shader:31-57 for _SubTransformSize = sub along `axis` (fp64 complex).

    index = texel index; evenIndex = floor(index/S)*(S/2) + fmod(index, S/2) (:41);
    out = in[even] + twiddle * in[even + N/2] with twiddle angle -2 pi index / S (:51-54).
    """
    x = np.moveaxis(np.asarray(x, dtype=np.complex128), axis, -1)
    N = x.shape[-1]
    index = np.arange(N)
    even = (index // sub) * (sub // 2) + index % (sub // 2)
    ang = -2.0 * np.pi * index / sub
    out = x[..., even] + (np.cos(ang) + 1j * np.sin(ang)) * x[..., even + N // 2]
    return np.moveaxis(out, -1, axis)


def stockham_fft2d(x: np.ndarray) -> np.ndarray:
    """OceanRenderer.cs:229-262 schedule: log2(N) horizontal stages, then log2(N) vertical.

    'Horizontal' indexes texcoord.x; with the engine's row-major [i][j] layout we take x <-> j
    (last axis).  The result equals numpy.fft.fft2 (forward sign, un-normalised).
    """
    N = x.shape[-1]
    stages = int(np.log2(N))
    y = np.asarray(x, dtype=np.complex128)
    for axis in (-1, -2):
        for s in range(stages):
            y = stockham_stage(y, 2 ** (s + 1), axis)
    return y


def direct_sum_fp64(Gs: np.ndarray, N: int, L: float, unit_width: float) -> np.ndarray:
    """Literal fp64 double loop of FFTMesh.cs:199-217 for one spectrum (small N only)."""
    k = 2.0 * np.pi * (np.arange(N) - N / 2.0) / L
    x = (np.arange(N) - N // 2) * unit_width + unit_width / 2.0
    E = np.exp(1j * np.outer(x, k))  # [a, n]
    return E @ Gs @ E.T
