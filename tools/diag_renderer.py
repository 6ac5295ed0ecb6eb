import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
import mistral_water_b200 as mw
from oracle import ref_ocean_renderer as ror
F64=np.float64
res=int(sys.argv[1]) if len(sys.argv)>1 else 128
r = mw.Renderer(res, 434.48, 0.46, 0.41, (14.45, 12.0), 1.5, seed1=3.7, seed2=8.1)
r.render_initial(); ini = r.get_initial()[0]
s = ror.RendererState(res, 434.48, 0.46, 0.41, (14.45, 12.0), 3.7, 8.1, 1.5, F64, initial=ini)
for f in range(2):
    got = r.generate_texture(0.02, names=("displacement","height","normal","white","jacobian"))
    b = s.generate_texture(0.02)
def rel(a,b): return np.linalg.norm(a.astype(F64).ravel()-b.ravel())/np.linalg.norm(b.ravel())
for k in ("displacement","height","normal"):
    print(k, "rel", rel(got[k][0], b[k]), "maxabs", np.abs(got[k][0]-b[k]).max(), "max", np.abs(b[k]).max())
for k in ("white","jacobian"):
    d = np.abs(got[k][0,...,0]-b[k]); i = np.unravel_index(d.argmax(), d.shape)
    print(k, "maxabs", d.max(), "at", i, "val", b[k][i], "jac there", b["jacobian"][i], "mean err", d.mean())
print("phase err", np.abs(r.get_phase()[0]-b["phase"]).max())
