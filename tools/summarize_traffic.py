"""Sums dram__bytes_{read,write}.sum over the launches of an ncu --csv log (tools/traffic_frame.py) -> JSON on stdout."""
import csv, json, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}
out = {}
for r in rows[1:]:
    k = r[ik].split("<")[0].split("(")[0].split("::")[-1].strip()
    d = out.setdefault(k, {"launches": 0, "dram_read": 0.0, "dram_write": 0.0, "time_us": 0.0})
    v = float(r[iv].replace(",", "")) * mult.get(r[iu], 1)
    if r[im] == "dram__bytes_read.sum": d["dram_read"] += v; d["launches"] += 1
    elif r[im] == "dram__bytes_write.sum": d["dram_write"] += v
    elif r[im] == "gpu__time_duration.sum": d["time_us"] += v
tot = {"dram_read": sum(d["dram_read"] for d in out.values()), "dram_write": sum(d["dram_write"] for d in out.values())}
print(json.dumps({"kernels": out, "total_bytes": tot["dram_read"] + tot["dram_write"], **tot}, indent=1))
