"""Aggregate `ncu --page source --print-source cuda,sass --csv` per CUDA source line.
usage: ncu_lines.py file.csv [top_n] [sort: inst|samples]"""
import csv, sys, collections
path = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
key = sys.argv[3] if len(sys.argv) > 3 else "inst"
rows = list(csv.reader(open(path)))
cur = None; hdr = None; line = None
agg = collections.defaultdict(lambda: [0.0, 0.0, ""])
def f(v):
    try: return float(v.replace(',', ''))
    except Exception: return 0.0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed"); continue
    if hdr is None or len(r) <= si: continue
    if r[0].strip():
        line = (cur, int(r[0])); agg[line][2] = r[1].strip()
    elif len(r) > 2 and r[2].strip() and line is not None:     # a SASS row under the current line
        agg[line][0] += f(r[si]); agg[line][1] += f(r[ii])
tot_i = sum(v[1] for v in agg.values()) or 1; tot_s = sum(v[0] for v in agg.values()) or 1
print(f"total warp-instructions {tot_i:.0f}  samples {tot_s:.0f}")
k = 1 if key == "inst" else 0
for kk, v in sorted(agg.items(), key=lambda kv: -kv[1][k])[:top_n]:
    print(f"{100*v[1]/tot_i:5.1f}% inst {100*v[0]/tot_s:5.1f}% smp  {kk[0]}:{kk[1]}  {v[2][:100]}")
