"""Stall-reason totals and per-source-line stall attribution of one kernel from
`ncu -i rep --page source --print-source cuda,sass --csv`.
usage: ncu_stalls.py file.csv [top_n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 10
hdr = None; cur = None; line = None
tot = collections.Counter(); per = collections.defaultdict(collections.Counter); src = {}
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0].strip():
        line = (cur, int(r[0])); src[line] = r[1].strip(); continue
    if r[2].strip().startswith('0x') and line:
        for i, h in enumerate(hdr):
            if h.startswith('stall_') and '(Not' not in h:
                try: v = float(r[i])
                except ValueError: continue
                tot[h] += v; per[h][line] += v
s = sum(tot.values()) or 1
print("stall reason totals (all samples)")
for k, v in tot.most_common(8): print(f"  {k:26s} {100*v/s:5.1f}%")
for h, _ in tot.most_common(4):
    t = sum(per[h].values()) or 1
    print("==", h)
    for k, v in per[h].most_common(top): print(f"  {100*v/t:5.1f}% {k[0]}:{k[1]} {src[k][:96]}")
