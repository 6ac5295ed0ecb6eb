"""Developer script: where a kernel's warp-stall samples fall, from `ncu --page source --csv --print-source sass`.
Prints the samples per region of `chunk` SASS instructions (with the dominant opcodes / stall reasons) and the top instructions.
Usage: python tools/ncu_sass_hot.py sass.csv kernel-substr [chunk]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
filt = sys.argv[2]; chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 100
fn, hdr, inst = None, None, []
for r in rows:
    if not r: continue
    if r[0] in ("Function Name", "Kernel Name"):
        if fn == r[1] and inst: break  # second copy of the same kernel (another launch)
        fn = r[1]; hdr = None; continue
    if r[0] in ("Address", "Line No") or (hdr is None and "Source" in r): hdr = r; continue
    if hdr is None or filt not in (fn or "") or len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    try: s = int(d["# Samples"])
    except (ValueError, KeyError): s = 0
    st = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v)}
    inst.append((d["Source"].strip(), s, st, d.get("Instructions Executed", "0")))
tot = sum(s for _, s, _, _ in inst) or 1
print(f"{filt}: {len(inst)} instructions, {tot} samples")
for i in range(0, len(inst), chunk):
    seg = inst[i:i + chunk]
    s = sum(x[1] for x in seg)
    stc = collections.Counter()
    for x in seg: stc.update(x[2])
    ops = collections.Counter(x[0].split()[0] if not x[0].startswith("@") else x[0].split()[1] for x in seg)
    print(f"[{i:5d}] {100*s/tot:5.1f}%  stalls: {' '.join(f'{k}={v}' for k, v in stc.most_common(4)):60s} ops: {' '.join(f'{k}:{v}' for k, v in ops.most_common(5))}")
print("top instructions:")
for j, (src, s, st, n) in sorted(enumerate(inst), key=lambda t: -t[1][1])[:25]:
    print(f"  #{j:5d} {100*s/tot:5.2f}% exec={n:>8} {src[:70]:70s} {' '.join(f'{k}={v}' for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])}")
