#!/usr/bin/env python
"""Hot source lines / SASS of one launch in an .ncu-rep: python scripts/ncu_hot_lines.py rep launch_idx [n]"""
import csv, io, subprocess, sys
rep, li = sys.argv[1], int(sys.argv[2]); n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(li), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if '# Samples' in r)
hdr = rows[hi]; ci = {}
for i, h in enumerate(hdr): ci.setdefault(h, i)
S = ci['# Samples']; E = ci['Instructions Executed']
def f(r, c):
    try: return float(r[c])
    except Exception: return 0.0
src = [r for r in rows[hi + 1:] if len(r) > S and r[0].strip().isdigit()]
sass = [r for r in rows[hi + 1:] if len(r) > S and not r[0].strip().isdigit()]
tot = sum(f(r, S) for r in src) or 1
print(f"total samples {tot:.0f}")
print("--- source lines")
for r in sorted(src, key=lambda r: -f(r, S))[:n]:
    print(f"{f(r,S):8.0f} {f(r,S)/tot*100:5.1f}% exec={f(r,E):9.0f} L{r[0]:>4} {r[1].strip()[:120]}")
print("--- sass")
seen = set()
for r in sorted(sass, key=lambda r: -f(r, S)):
    key = (r[ci['Address']], r[3])
    if key in seen: continue
    seen.add(key)
    if len(seen) > n: break
    print(f"{f(r,S):8.0f} exec={f(r,E):9.0f} {r[ci['Address']][-6:]} {r[3][:100]}")
