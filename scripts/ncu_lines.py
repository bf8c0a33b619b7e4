"""Aggregate `ncu --page source --print-source cuda,sass --csv` samples per CUDA source line."""
import csv, sys, collections
path = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
files = {}
cur = None
hdr = None
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
    if r[0].strip():            # a source line row: remember it
        line = (cur, r[0]); src = r[1]
        agg[line][2] = src.strip()
    if len(r) > 2 and r[2].strip():   # sass row under current line
        agg[line][0] += f(r[si]); agg[line][1] += f(r[ii])
tot = sum(v[0] for v in agg.values())
print("total samples", tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top_n]:
    print(f"{v[0]:8.0f} {100*v[0]/tot:5.1f}%  inst {v[1]:12.0f}  {k[0]}:{k[1]}  {v[2][:100]}")
