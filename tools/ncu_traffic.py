#!/usr/bin/env python
"""profiles/r2_ncu_graph_step.csv (ncu --graph-profiling graph: one replay of bench.py's whole step graph as ONE
workload) -> profiles/r2_roofline_traffic.json, which bench.py reads into roofline.traffic.

  python tools/ncu_traffic.py profiles/r2_ncu_graph_step.csv --pairs 262144 --cells <cells of one step of that run>

bench.py cannot run ncu inside itself; the capture is of the CURRENT round's build and of the same command at a smaller
shard (ncu replays the graph once per metric pass), so the figure is carried per DP cell and scaled by the step's exact
cell count."""
import argparse
import csv
import json
import os


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--pairs", type=int, required=True)
    ap.add_argument("--cells", type=float, required=True, help="exact DP cells of one step of the profiled run")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles",
                                                   "r2_roofline_traffic.json"))
    args = ap.parse_args()
    rows = [r for r in csv.reader(open(args.csv)) if len(r) > 10 and r[0].isdigit()]
    per_id = {}
    for r in rows:
        per_id.setdefault(r[0], {})[r[-3]] = (r[-2], float(r[-1].replace(",", "")))
    graphs = [m for m in per_id.values() if "dram__bytes_read.sum" in m]
    if not graphs:
        raise SystemExit("no workload with dram metrics in %s" % args.csv)
    g = max(graphs, key=lambda m: m["gpu__time_duration.sum"][1])        # the step graph (not a stray fill kernel)

    def to_bytes(m):
        unit, v = m
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]

    rd, wr = to_bytes(g["dram__bytes_read.sum"]), to_bytes(g["dram__bytes_write.sum"])
    out = {"source": "ncu --graph-profiling graph over one replay of the step graph (%s, --pairs %d)" % (os.path.basename(args.csv), args.pairs),
           "dram_bytes_read": rd, "dram_bytes_write": wr, "cells_per_step": args.cells,
           "dram_bytes_per_cell": (rd + wr) / args.cells,
           "counters": {k: {"unit": v[0], "value": v[1]} for k, v in g.items()}}
    json.dump(out, open(args.out, "w"), indent=1)
    print(json.dumps(out)[:600])


if __name__ == "__main__":
    main()
