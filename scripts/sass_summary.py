#!/usr/bin/env python
"""Blackwell-specific SASS mnemonics per kernel of libspsk.so (no GPU needed):  python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, re, subprocess, sys
from pathlib import Path

LIB = Path(__file__).resolve().parents[1] / "spsnet_b200" / "_C" / "libspsk.so"
MN = ["UTCHMMA.2CTA", "UTCHMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "STAS", "UCGABAR", "SYNCS", "LDGSTS", "CREDUX", "REDUX"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    per, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for mn in MN:
            if re.search(r"\b" + re.escape(mn) + r"\b", line) and not (mn == "UTCHMMA" and "UTCHMMA.2CTA" in line) and not (mn == "REDUX" and "CREDUX" in line):
                per[cur][mn] += 1
                break
    tot = collections.Counter()
    for c in per.values():
        tot.update(c)
    print("# cuobjdump -sass spsnet_b200/_C/libspsk.so (sm_100a, nvcc 12.9): Blackwell-specific SASS mnemonics per kernel (round 2, final kernels).")
    print("# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM / STTM = tcgen05.ld / tcgen05.st (TMEM), UTCBAR = tcgen05.commit,")
    print("# UBLKCP = cp.async.bulk (1-D bulk copy on mbarrier), STAS = st.async (DSMEM), UCGABAR = cluster barrier, SYNCS = mbarrier ops,")
    print("# LDGSTS = cp.async, REDUX/CREDUX = warp reduce.  No UTMALDG: the A operands are per-row gathers, weights are pre-tiled 1-D bulk copies.")
    print("# totals: " + "  ".join(f"{k}={v}" for k, v in sorted(tot.items())))
    for name, c in per.items():
        if c:
            print(f"{name[:78]:78s} " + "  ".join(f"{k}={v}" for k, v in sorted(c.items())))


if __name__ == "__main__":
    main()
