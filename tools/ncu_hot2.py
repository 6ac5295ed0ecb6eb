"""Developer script: hot source lines of one kernel from `ncu -i rep --page source --csv --print-source cuda,sass`.
Usage: python tools/ncu_hot2.py src.csv <kernel-substr> [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
filt = sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
fpath = fn = hdr = None
agg = collections.defaultdict(collections.Counter); text = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1]; continue
    if r[0] == "Function Name": fn = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or filt not in (fn or "") or not r[0]: continue   # source rows only (SASS rows have an empty Line No)
    d = {}
    for k, v in zip(hdr, r):
        d.setdefault(k, v)
    key = (fpath.split("/")[-1], int(r[0]))
    try: agg[key]["samples"] += int(d["# Samples"] or 0); agg[key]["inst"] += int(d["Instructions Executed"] or 0)
    except ValueError: pass
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k:
            try: agg[key][k] += int(v or 0)
            except ValueError: pass
    text[key] = r[1][:100]
tot = sum(v["samples"] for v in agg.values()) or 1
ti = sum(v["inst"] for v in agg.values()) or 1
st_tot = collections.Counter()
for v in agg.values():
    for k, c in v.items():
        if k.startswith("stall_"): st_tot[k] += c
print("total samples", tot, "warp instructions", ti, {k[6:]: round(100 * c / tot, 1) for k, c in st_tot.most_common(9)})
for key, v in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(((k, c) for k, c in v.items() if k.startswith("stall_")), key=lambda kc: -kc[1])[:3]
    print(f"{100*v['samples']/tot:5.1f}% inst={100*v['inst']/ti:4.1f}% {key[0]}:{key[1]:>4} {' '.join(f'{k[6:]}={c}' for k,c in st):44s} | {text[key].strip()}")
