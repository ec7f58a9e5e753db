#!/usr/bin/env python
"""Per-kernel SHARES of one eager step from an ncu launch list (cold-cache, serialised launches: shares only, never absolute times).

    python scripts/ncu_launch_shares.py ncu.csv out.txt "<header text>" [anchor-regex]

Takes the launches between the last two launches matching `anchor-regex` (default: the 16384 -> 4096 D-FPS kernel, which opens a
KITTI step), groups them by kernel name and prints time, share, launches and DRAM bytes."""
import csv, re, sys
from collections import OrderedDict


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
    ix = {h: i for i, h in enumerate(rows[0])}
    launches = OrderedDict()
    for r in rows[1:]:
        d = launches.setdefault(int(r[ix["ID"]]), {"name": r[ix["Kernel Name"]].split("(")[0]})
        try:
            d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
        except ValueError:
            pass
    seq = [launches[k] for k in sorted(launches)]
    anchor = re.compile(sys.argv[4] if len(sys.argv) > 4 else r"fps_pruned_kernel<32")
    hits = [i for i, k in enumerate(seq) if anchor.search(k["name"])]
    step = seq[hits[-2]:hits[-1]] if len(hits) >= 2 else seq
    tot = sum(k.get("gpu__time_duration.sum", 0.0) for k in step) / 1e3
    agg = OrderedDict()
    for k in step:
        a = agg.setdefault(k["name"], [0.0, 0, 0.0])
        a[0] += k.get("gpu__time_duration.sum", 0.0) / 1e3
        a[1] += 1
        a[2] += k.get("dram__bytes_read.sum", 0.0) + k.get("dram__bytes_write.sum", 0.0)
    with open(sys.argv[2], "w") as f:
        f.write(sys.argv[3].rstrip() + f"\n# {len(step)} launches, {tot:.0f} us serialised\n")
        for name, (us, n, by) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            f.write(f"{us:10.1f} us  {100 * us / tot:5.1f}%  x{n:3d}  dram {by / 1e6:8.2f} MB  {name[-90:]}\n")


if __name__ == "__main__":
    main()
