"""Developer script: opcode histogram of one kernel's SASS, whole and per region of `--chunk` instructions.
Usage: cuobjdump -sass -fun <mangled> file.o | python tools/sass_hist.py [--chunk 300]"""
import re, sys, collections, argparse
ap = argparse.ArgumentParser(); ap.add_argument("--chunk", type=int, default=300); a = ap.parse_args()
pat = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)")
ops = []
for line in sys.stdin:
    m = pat.match(line)
    if m: ops.append(m.group(2))
print("total", len(ops), dict(collections.Counter(ops).most_common(25)))
for i in range(0, len(ops), a.chunk):
    c = collections.Counter(ops[i:i + a.chunk])
    print(f"[{i:5d}] " + " ".join(f"{k}:{v}" for k, v in c.most_common(12)))
