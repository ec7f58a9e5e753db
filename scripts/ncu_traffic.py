#!/usr/bin/env python
"""DRAM traffic per kernel INSTANCE for bench.py's `roofline.traffic`.

    python scripts/ncu_traffic.py calls.json ncu.csv out.json [summary.txt]

calls.json : the library calls of one eager step in launch order (bench.py with SPSK_DUMP_CALLS=calls.json)
ncu.csv    : `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum ...` of the same command
             (every launch of the run; the LAST complete step is used)
out.json   : {bench kernel key (with its shape): dram bytes read + written per launch}
A library call maps to one or more consecutive spsk:: kernels (e.g. the grid ball query = build + query)."""
import csv, json, re, sys

PATTERNS = {  # library call -> kernel name regexes, in launch order (each matches one launch; optional ones end with '?')
    "spsk_farthest_point_sampling": [r"fps_"],
    "spsk_sa_mma_forward": [r"sa_mma_kernel"],
    "spsk_pw_mma_forward": [r"pw_mma_kernel"],
    "spsk_ball_query_msg_grid": [r"bq_grid_build_kernel", r"bq_grid_query_kernel"],
    "spsk_ball_query_msg": [r"ball_query_msg_kernel|ball_query_kernel"],
    "spsk_ball_query": [r"ball_query_kernel"],
    "spsk_score_topk": [r"score_topk_kernel"],
    "spsk_gather_rows": [r"gather_rows_kernel"],
    "spsk_gather_points": [r"gather_cols_kernel|gather_points"],
    "spsk_ball_query_grid_workspace_bytes": [], "spsk_nms_workspace_bytes": [], "spsk_detect_workspace_bytes": [],
    "spsk_scatter_grad_workspace_bytes": [], "spsk_sa_mma_config": [], "spsk_fp16_overflow_poll": [],
    "spsk_make_twin": [r"make_twin_kernel"],
    "spsk_grouped_linear": [r"linear_ffma|grouped"],
    "spsk_pointwise_linear": [r"linear_ffma|pointwise"],
    "spsk_detect_postprocess": [r"detect_sort_kernel", r"nms_mask", r"nms_reduce"],
}


def main():
    calls = json.load(open(sys.argv[1]))
    rows = [r for r in csv.reader(open(sys.argv[2])) if len(r) > 10]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    launches = {}
    for r in rows[1:]:
        lid = int(r[ix["ID"]])
        d = launches.setdefault(lid, {"name": r[ix["Kernel Name"]]})
        try:
            d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
        except ValueError:
            pass
    seq = [launches[k] for k in sorted(launches) if "spsk" in launches[k]["name"] or re.search(r"fps_|sa_mma|pw_mma|bq_grid|ball_query|score_topk|gather_|make_twin|linear_ffma|detect_|nms_", launches[k]["name"])]
    need = sum(len(PATTERNS.get(c.split("[")[0], [r".*"])) for c in calls)
    seq = seq[-need:]   # the last complete step
    out, lines, pos = {}, [], 0
    for c in calls:
        pats = PATTERNS.get(c.split("[")[0], [r".*"])
        tot, us, names = 0.0, 0.0, []
        for pat in pats:
            if pos >= len(seq) or not re.search(pat, seq[pos]["name"]):
                raise SystemExit(f"launch list does not line up with the call order at {c}: expected /{pat}/, got {seq[pos]['name'] if pos < len(seq) else 'end'}")
            k = seq[pos]
            tot += k.get("dram__bytes_read.sum", 0.0) + k.get("dram__bytes_write.sum", 0.0)
            us += k.get("gpu__time_duration.sum", 0.0) / 1e3
            names.append(k["name"].split("(")[0][-50:])
            pos += 1
        prev = out.get(c)
        out[c] = tot if prev is None else max(prev, tot)
        lines.append(f"{c:60s} {us:9.1f} us  dram {tot/1e6:9.2f} MB   {' + '.join(names)}")
    json.dump(out, open(sys.argv[3], "w"), indent=1)
    if len(sys.argv) > 4:
        open(sys.argv[4], "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
