"""Regenerates tests/golden/*.npz from the literal C oracle (oracle/ref_fftmesh.c).

The reference ships no golden vectors and cannot run here (C# + closed-source UnityEngine), so
these fixtures pin the *restatement*: they freeze what oracle/ref_fftmesh.c produced when it was
written, so that later edits to the oracle (or to the CUDA path) cannot drift silently.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def ocean(N, seed, ts, name, wind=(5.0, 3.0), amplitude=0.01, choppiness=1.0):
    p = cref.params(N, unit_width=1.0, length=float(N), choppiness=choppiness, amplitude=amplitude, wind=wind)
    v, h0, hc = cref.generate_mesh(p, seed=seed)
    out = {"N": N, "seed": seed, "wind": np.array(wind, np.float32), "amplitude": np.float32(amplitude),
           "choppiness": np.float32(choppiness), "ts": np.array(ts, np.float32),
           "vertices": v, "h0": h0, "h0conj": hc, "omega": cref.dispersion(p)}
    for k, t in enumerate(ts):
        r = cref.evaluate_waves(p, v, h0, hc, float(np.float32(t)), threads=cref.max_threads())
        out[f"htilde_{k}"] = cref.htilde(p, h0, hc, float(np.float32(t)))
        for key in ("vertMeow", "normals", "hds", "jacobian", "whitecap"):
            out[f"{key}_{k}"] = r[key]
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, {k: getattr(v, "shape", v) for k, v in out.items() if k.endswith("_0")})


def gerstner():
    rng = np.random.default_rng(3)
    pos = rng.uniform(-50, 50, (1001, 3)).astype(np.float32)
    m = dict(amplitude=10.0, frequency=2.58, steepness=0.99, speed=(1.2, 0.71, 1.1, 0.73),
             dirAB=(0.3, 0.73, 0.85, 0.25), dirCD=(-0.25, 1.11, 0.5, 0.5))  # Pond Water Mat.mat
    g4 = cref.gerstner4(pos, 1.7, m["amplitude"] * 0.01, m["frequency"], m["steepness"], m["speed"], m["dirAB"], m["dirCD"])
    g5 = cref.gerstner_level_one(pos, 1.7, 0.1, 2.58, 0.99)
    np.savez_compressed(os.path.join(HERE, "gerstner_pond.npz"), pos=pos, t=np.float32(1.7), offsets4=g4, offsets5=g5)
    print("gerstner_pond.npz", g4.shape)


def renderer():
    """OceanRenderer path, Ocean Demo scene values (Demo/Ocean Demo.unity:296-302) at mesh resolution 8 (R = 64):
    the literal fp32 blit chain of oracle/ref_ocean_renderer.py after 3 frames of 16 ms."""
    from oracle import ref_ocean_renderer as R
    cfg = dict(resolution=8, length=434.48, choppiness=0.46, amplitude=0.41, wind=(14.45, 12.0), seed1=3.7, seed2=8.1, mult=1.5)
    s = R.RendererState(cfg["resolution"], cfg["length"], cfg["choppiness"], cfg["amplitude"], cfg["wind"], cfg["seed1"],
                        cfg["seed2"], cfg["mult"], np.float32)
    out = {k: np.float32(v) if not isinstance(v, tuple) else np.array(v, np.float32) for k, v in cfg.items()}
    out["resolution"] = np.int32(cfg["resolution"])
    out.update(initial=s.initial.copy(), frames=np.int32(3), dt=np.float32(0.016))
    for _ in range(3):
        m = s.generate_texture(0.016)
    for k in ("displacement", "height", "normal", "white", "jacobian", "phase"):
        out[k] = m[k].astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "renderer_r64.npz"), **out)
    print("renderer_r64.npz", {k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    if "--renderer-only" in sys.argv:
        renderer()
        sys.exit(0)
    ocean(32, 1234, [0.0, 1.7, 60.0], "fftmesh_n32.npz")
    ocean(64, 1234, [1.7], "fftmesh_n64.npz")
    ocean(32, 99, [3.25], "fftmesh_n32_wind.npz", wind=(-2.0, 7.5), amplitude=0.0004, choppiness=0.46)
    gerstner()
    renderer()
