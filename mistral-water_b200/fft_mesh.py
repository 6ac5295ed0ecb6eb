"""Host-side mirror of the reference's FFTMesh MonoBehaviour (Scripts/FFTMesh.cs).

Same public fields (FFTMesh.cs:9-23), same lifecycle (Awake / Update), same private method names
(SetParams / GenerateMesh / EvaluateWaves) with the same meaning -- but every numeric body is one
call into libmistral_ocean.so.  A C# maintainer makes exactly these substitutions in FFTMesh.cs
(INTEGRATION.md); this Python class exists so that the parity tests read like tests of the
reference component, since no C# toolchain is available in this image.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .ocean import Ocean


@dataclass
class Mesh:
    """The three arrays FFTMesh hands to UnityEngine.Mesh (FFTMesh.cs:134-137, 277-279)."""
    vertices: np.ndarray | None = None  # Vector3[N*N]
    normals: np.ndarray | None = None   # Vector3[N*N]
    colors: np.ndarray | None = None    # Color[N*N]
    uv: np.ndarray | None = None
    indices: np.ndarray | None = None


@dataclass
class FFTMesh:
    # ---- Parameter Variables, FFTMesh.cs:9-23 (same names, same defaults) ----
    choppiness: float = 1.0
    tDivision: float = 1.0
    resolution: int = 50
    unitWidth: float = 1.0
    generate: bool = False
    length: float = 1.0
    wind: tuple = (1.0, 1.0)
    amplitude: float = 1.0
    # ---- not in the reference: which GPU, and the stand-in for UnityEngine.Random's state ----
    device: int = 0
    seed: int = 0
    mesh: Mesh = field(default_factory=Mesh)

    def __post_init__(self):
        self.timer = 0.0  # FFTMesh.cs:44
        self._ocean: Ocean | None = None
        self.vertices = None
        self.verttilde = None
        self.vertConj = None
        self.hds = None
        self.jacobian = None

    # FFTMesh.cs:75-84
    def Awake(self, verttilde=None, vertConj=None) -> None:
        self.SetParams()
        self.GenerateMesh(verttilde, vertConj)

    # FFTMesh.cs:60-73
    def Update(self, deltaTime: float) -> None:
        if self.generate:
            self.timer = 0.0
            self.SetParams()
            self.GenerateMesh()
            self.generate = False
        self.timer = float(np.float32(self.timer) + np.float32(deltaTime) / np.float32(self.tDivision))
        self.EvaluateWaves(self.timer)

    # FFTMesh.cs:90-99: (re)allocate everything for the current public fields
    def SetParams(self) -> None:
        if self._ocean is not None:
            self._ocean.close()
        self._ocean = Ocean(self.resolution, self.unitWidth, self.length, self.choppiness, self.amplitude,
                            self.wind, self.tDivision, seed=self.seed, device=self.device)
        n2 = self.resolution * self.resolution
        self._bufs = {
            "vertices": np.empty((1, n2, 3), np.float32), "normal": np.empty((1, n2, 3), np.float32),
            "colors": np.empty((1, n2, 4), np.float32), "disp": np.empty((1, n2, 2), np.float32),
            "jacobian": np.empty((1, n2, 1), np.float32),
        }

    # FFTMesh.cs:101-139.  verttilde / vertConj may be supplied by the host (it then keeps its own
    # RNG, e.g. UnityEngine.Random); otherwise they are drawn on the device (Philox, `seed`).
    def GenerateMesh(self, verttilde=None, vertConj=None) -> None:
        o = self._ocean
        self.vertices = o.rest_vertices()                      # :107-112
        if verttilde is not None:
            o.set_h0(np.ascontiguousarray(verttilde, np.float32), np.ascontiguousarray(vertConj, np.float32))
        else:
            o.init_spectrum()                                  # :114-116
        h0, hc = o.get_h0()
        self.verttilde, self.vertConj = h0[0], hc[0]
        N = self.resolution
        self.mesh.vertices = self.vertices
        self.mesh.normals = np.tile(np.array([0.0, 1.0, 0.0], np.float32), (N * N, 1))  # :113
        i, j = np.divmod(np.arange(N * N), N)
        self.mesh.uv = np.stack([i / np.float32(N - 1), j / np.float32(N - 1)], -1).astype(np.float32)  # :117

    # FFTMesh.cs:224-280
    def EvaluateWaves(self, t: float) -> None:
        self._ocean.generate(t, self._bufs)
        self.hds = self._bufs["disp"][0]
        self.jacobian = self._bufs["jacobian"][0, :, 0]
        self.mesh.vertices = self._bufs["vertices"][0]         # :277
        self.mesh.normals = self._bufs["normal"][0]            # :278
        self.mesh.colors = self._bufs["colors"][0]             # :279

    def close(self) -> None:
        if self._ocean is not None:
            self._ocean.close()
            self._ocean = None
