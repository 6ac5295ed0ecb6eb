"""Runs here (no GPU): turns the ncu reports brought back in gpurun_out/ into the tracked summaries under profiles/.
Usage: python tools/summarize_profiles.py r01"""
import csv, json, os, subprocess, sys, collections

R = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sector_hit_rate.pct",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__occupancy_limit_barriers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.avg.per_cycle_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def stalls(d):
    s = {}
    for k, v in d.items():
        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k:
            try:
                s[k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(v)
            except ValueError:
                pass
    return sorted(s.items(), key=lambda kv: -kv[1])[:6]


md = [f"# profiles/{R}: ncu summaries (B200, sm_100a)\n",
      "Source reports: `gpurun_out/" + R + "_*.ncu-rep` (scratch, not tracked). Commands: `tools/profile_round.sh`.",
      "Durations under ncu are cold-cache and serialised; use them for shares and traffic, not for throughput.\n"]
traffic = {}
for tag, title in (("frame_grouped", "frame kernels, default scheduling (1 tile of 1024^2 per launch, --cache-control none: intermediate L2-resident as in the timed region)"),
                   ("frame_batched", "frame kernels, one launch for all 16 tiles (MW_GROUP_TILES=16: steady state over many waves, intermediate through HBM)"),
                   ("frame_2048", "frame kernels, one 2048^2 tile (BASELINE configs[4] per-rank work)"),
                   ("frame_256", "frame kernels, 256 x 256^2 in one launch (MW_GROUP_TILES=256)"),
                   ("gerstner", "k_gerstner, 32 waves x 1048576 vertices"),
                   ("renderer", "OceanRenderer path, 16 x (1024^2 maps) per call: k_r_rows, k_r_cols, k_r_maps")):
    rep = os.path.join(G, f"{R}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        continue
    recs, units = raw(rep)
    md.append(f"## {title}\n")
    for d in recs:
        name = d["Kernel Name"].split("(")[0].strip()
        md.append(f"### `{name}`\n")
        md.append("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in d:
                md.append(f"| {k} | {d[k]} | {units.get(k, '')} |")
        md.append("")
        md.append("top warp-stall reasons (warps per issue-active cycle): " + ", ".join(f"{k} {v:.2f}" for k, v in stalls(d)) + "\n")
        if False and tag == "frame_batched":
            def tobytes(x, u):
                return float(x) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            short = "k_cols" if "k_cols" in name else "k_spectrum_rows"
            tot = tobytes(d["dram__bytes_read.sum"], units["dram__bytes_read.sum"]) + tobytes(d["dram__bytes_write.sum"], units["dram__bytes_write.sum"])
            traffic[short] = {"bytes_per_tile": tot / 16.0, "source": f"profiles/{R}_summary.md: dram__bytes_read.sum + dram__bytes_write.sum of one 16-tile launch "
                              "(ncu --set full, MW_GROUP_TILES=16) / 16; a one-tile launch under ncu leaves its writes dirty in L2, so it undercounts"}
# launch list shares
lc = os.path.join(G, f"{R}_launches.csv")
if os.path.exists(lc):
    rows = [r for r in csv.reader(open(lc)) if r and r[0].isdigit()]
    # columns: ID, Process ID, Process Name, Host Name, Kernel Name, Context, Stream, Block Size, Grid Size, Device, CC, Section, Metric Name, Unit, Value
    agg = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "").strip()
        val = float(r[-1].replace(",", ""))
        unit = r[-2]
        val_us = val / 1e3 if unit in ("ns", "nsecond") else val * (1e3 if unit in ("ms", "msecond") else 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += val_us
    tot = sum(v[1] for v in agg.values())
    md.append(f"## launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs --e2e-steps 1` ({len(rows)} launches)\n")
    md.append("| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        md.append(f"| `{k[:90]}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% |")
    md.append("")
    with open(os.path.join(P, f"{R}_launches.csv"), "w") as f:
        f.write(open(lc).read())
ph = os.path.join(G, f"{R}_phases.txt")
if os.path.exists(ph):
    md.append("## phases switched off one at a time (`MW_GROUP_TILES=16 python tools/phase_timing.py`, CUDA events per launch, 16 tiles of 1024^2)\n")
    md.append("```\n" + "".join(l for l in open(ph) if l.startswith("flags=")) + "```\n")
for fn in (f"{R}_bench.json", f"{R}_extra.json", f"{R}_smi.csv"):
    src = os.path.join(G, fn)
    if os.path.exists(src):
        open(os.path.join(P, fn), "w").write(open(src).read())
# DRAM traffic of one whole frame in the timed scheduling (tools/traffic_frame.py under ncu --replay-mode application)
for tag in ("traffic_grouped", "traffic_batched"):
    src = os.path.join(G, f"{R}_{tag}.json")
    if os.path.exists(src):
        t = json.load(open(src))
        open(os.path.join(P, f"{R}_{tag}.json"), "w").write(open(src).read())
        md.append(f"## DRAM traffic of one 16 x 1024^2 frame, `{tag}` (ncu --replay-mode application --cache-control none, summed over the frame's launches)\n")
        md.append("| kernel | launches | dram read MB | dram write MB | serialised us |\n|---|---|---|---|---|")
        for k, v in t["kernels"].items():
            md.append(f"| `{k}` | {v['launches']} | {v['dram_read'] / 1e6:.1f} | {v['dram_write'] / 1e6:.1f} | {v['time_us']:.1f} |")
        md.append(f"\ntotal {t['total_bytes'] / 1e6:.1f} MB per frame against 738.2 MB algorithmic (44 B x 16 777 216 points): x{t['total_bytes'] / 738197504:.2f}\n")
open(os.path.join(P, f"{R}_summary.md"), "w").write("\n".join(md) + "\n")
print("\n".join(md)[:3000])
