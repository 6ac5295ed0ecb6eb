"""Developer script: aggregate `ncu --page source --csv --print-source cuda,sass` by CUDA source line.
Usage: ncu -i rep --page source --csv --print-source cuda,sass > src.csv; python tools/ncu_source_hot.py src.csv [kernel-substr]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
filt = sys.argv[2] if len(sys.argv) > 2 else ""
fn, fpath, hdr = None, None, None
agg = collections.defaultdict(lambda: collections.Counter())
text = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1]; continue
    if r[0] == "Function Name": fn = r[1]; hdr = None; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or filt not in (fn or ""): continue
    d = dict(zip(hdr, r))
    # rows with an Address are SASS rows attributed to the preceding source line
    key = (fn.split("(")[0][-40:], fpath.split("/")[-1], d["Line No"])
    try: samples = int(d.get("# Samples") or 0)
    except ValueError: samples = 0
    if d.get("Address"):
        continue
    agg[key]["samples"] += samples
    try: agg[key]["inst"] += int(d.get("Instructions Executed") or 0)
    except ValueError: pass
    for k in d:
        if k.startswith("stall_") and "Not Issued" not in k:
            try: agg[key][k] += int(d[k] or 0)
            except ValueError: pass
    text[key] = d["Source"][:110]
tot = sum(v["samples"] for v in agg.values()) or 1
print("total samples", tot)
for key, v in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:45]:
    st = sorted(((k, c) for k, c in v.items() if k.startswith("stall_")), key=lambda kc: -kc[1])[:3]
    print(f"{100*v['samples']/tot:5.1f}% inst={v['inst']:>9} {key[1]}:{key[2]:>4} {' '.join(f'{k[6:]}={c}' for k,c in st):40s} | {text[key].strip()}")
