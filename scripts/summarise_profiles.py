"""Turn gpurun_out ncu captures into the committed summaries under profiles/.

usage: python scripts/summarise_profiles.py <round tag> <launches.csv> <full.ncu-rep> [<more.ncu-rep> ...]
(the first launch of every kernel name over all reports is summarised)
"""
import collections, csv, json, os, subprocess, sys
tag, launches_csv, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

def short(name):
    return name.split("(")[0].replace("void ", "").replace("<unnamed>::", "")

# ---- launch list -------------------------------------------------------------------------
rows = list(csv.reader(open(launches_csv)))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[idx["Metric Value"]].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[idx["Metric Unit"]], 1.0)
    a = agg.setdefault((short(r[idx["Kernel Name"]]), r[idx["Grid Size"]], r[idx["Block Size"]]), [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(v[1] for v in agg.values())
lines = [f"# ncu launch list, round {tag}", "",
         "`ncu --metrics gpu__time_duration.sum --clock-control none -k regex:\"encode_|decode_|nms_\" -c 400` on",
         "`python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline` (B = 4096 images per step; the first",
         "16 encode launches with grid 256 are the benchmark's own input preparation).  Times are cold-cache and",
         "serialised: compare shares, not absolutes.", "",
         "| kernel | grid | block | launches | total us | share |", "|---|---|---|---|---|---|"]
for (k, g, b), v in agg.items():
    lines.append(f"| `{k}` | {g} | {b} | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% |")
open(os.path.join(out_dir, f"{tag}_launches.md"), "w").write("\n".join(lines) + "\n")

# ---- full capture --------------------------------------------------------------------------
want = ["gpu__time_duration.sum", "launch__waves_per_multiprocessor", "sm__cycles_active.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed" ,
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
def num(v):
    try: return float(v.replace(",", ""))
    except Exception: return None
seen = {}
md = [f"# ncu --set full summary, round {tag}", "",
      "`ncu --set full --clock-control none --import-source on` on the same bench command; one launch per kernel shown",
      "(B = 4096 images: encode kernels run per 1024-image chunk, decode / NMS once per step).", ""]
traffic = {}
all_rows = []
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]; units = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
    all_rows += [(r, units, idx, os.path.basename(rep)) for r in rows[2:]]
for r, units, idx, rep_name in all_rows:
    k = short(r[idx["Kernel Name"]])
    if rep_name.endswith("dense.ncu-rep"): k += " [dense random head, 256 images]"
    if k in seen: continue
    seen[k] = 1
    md.append(f"## `{k}`\n")
    md.append("| metric | value | unit |\n|---|---|---|")
    for w in want:
        if w in idx: md.append(f"| {w} | {r[idx[w]]} | {units[idx[w]]} |")
    rd, wr = num(r[idx["dram__bytes_read.sum"]]), num(r[idx["dram__bytes_write.sum"]])
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    rd *= scale[units[idx["dram__bytes_read.sum"]]]; wr *= scale[units[idx["dram__bytes_write.sum"]]]
    grid = num(r[idx["launch__grid_size"]])
    md.append(f"\nDRAM traffic of this launch: {(rd + wr) / 1e6:.1f} MB (read {rd / 1e6:.1f}, write {wr / 1e6:.1f}).\n")
    traffic[k] = {"dram_bytes_per_launch": rd + wr, "grid": grid}
open(os.path.join(out_dir, f"{tag}_ncu_full.md"), "w").write("\n".join(md) + "\n")
json.dump(traffic, open(os.path.join(out_dir, f"{tag}_traffic_raw.json"), "w"), indent=1)
print(open(os.path.join(out_dir, f"{tag}_launches.md")).read())
print(json.dumps(traffic, indent=1))
