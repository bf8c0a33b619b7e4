"""Aggregate `ncu --page source --print-source sass --csv` by SASS opcode.
usage: ncu_opcodes.py file.csv [top_n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(r for r in rows if r and r[0] == "Address")
ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
agg = collections.defaultdict(lambda: [0.0, 0.0])
for r in rows:
    if not r or not r[0].startswith("0x") or len(r) <= ii: continue
    parts = r[1].split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    op = ".".join(op.split(".")[:2]) if op.split(".")[0] in ("MUFU", "F2F", "F2I", "I2F", "LDS", "STS", "LDG", "STG", "SHFL", "VOTE") else op.split(".")[0]
    try: agg[op][0] += float(r[ii].replace(",", "")); agg[op][1] += float(r[si].replace(",", ""))
    except ValueError: pass
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values()) or 1
print(f"total {ti/1e6:.1f}M warp-instructions, {ts:.0f} samples")
for op, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{op:14s} {n/1e6:8.2f}M {100*n/ti:5.1f}% inst {100*s/ts:5.1f}% smp")
