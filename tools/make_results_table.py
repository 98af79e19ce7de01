#!/usr/bin/env python
"""Regenerates BASELINE.md section 5 (the results table) from the JSON lines committed under profiles/ -- one table,
one source.  python tools/make_results_table.py"""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def load(name):
    p = os.path.join(P, name)
    if not os.path.exists(p):
        return None
    txt = open(p).read().strip()
    if name.endswith(".jsonl"):
        return [json.loads(l) for l in txt.splitlines() if l.strip().startswith("{")]
    return json.loads(txt)


def f0(x):
    return "–" if x is None else "%.0f" % x


def ext_row(label, d, ref=None):
    if d is None:
        return None
    e = d["e2e"]
    cpu = d.get("cpu_baseline") or (ref or {}).get("cpu_baseline")
    return "| %s | %d | **%s** | %.2f / %.2f | %.1f M | %s (pinned), %s (pageable), %d callers/GPU | %s | bit-exact, cells equal |" % (
        label, d["n_gpus"], f0(d["value"]), d["roofline"]["frac_alu_pipe"], d["roofline"]["frac_dual_issue"],
        d["read_pairs_per_s"] / 1e6, f0(e["value"]), f0(e["pageable"]["value"]), e["caller_threads_per_gpu"],
        ("%.2f GCUPS (%s, %d cores)" % (cpu["value"], cpu["kind"], cpu["cores"])) if cpu else "–")


def main():
    rows = []
    full = load("r2_bench_full.json")
    refarm = load("r2_bench_reference_arm.json")
    rows.append(ext_row("**C2** extension (1M x 2, 151 bp, 4096 reads/call)", full, refarm))
    for n in (2, 4, 8):
        rows.append(ext_row("C2 extension, %d GPUs (1M x 2 per GPU, weak)" % n, load("r2_bench_%dgpu.json" % n)))
    rows.append(ext_row("C1 extension (101 bp)", load("r2_bench_C1.json")))
    rows.append(ext_row("C5 extension (250 bp, 5 % error)", load("r2_bench_C5.json")))
    out = ["| config | GPUs | kernel GCUPS (resident) | fraction of int roofline (ALU pipe / dual issue) | pairs/s resident | end to end, host buffers (GCUPS) | CPU arm | parity |",
           "|---|---|---|---|---|---|---|---|"]
    out += [r for r in rows if r]
    ms = load("r2_matesw_bench.jsonl") or []
    big = load("r2_matesw_C3_1M.json")
    two = load("r2_matesw_C3_2gpu.json")
    for m in ms + ([big] if big else []) + ([two] if two else []):
        rc = m.get("reference_cpu") or {}
        out.append("| %s | %d | **%s** | %.2f / – | %.2f M | %s (4096-pair calls, 4 callers), %s (10-pair calls, 64 callers) | %s | bit-exact (7 fields) |" % (
            m["workload"].split(":")[0].replace(" mate-SW", "") + " mate-SW, " + ("%d pairs%s" % (m["jobs"] // 2, " per GPU" if m.get("n_gpus", 1) > 1 else "")), m.get("n_gpus", 1), f0(m["kernel_gcups"]),
            m["roofline_frac_alu"], m.get("read_pairs_per_s", 0) / 1e6,
            f0((m.get("host_abi_large") or {}).get("gcups")), f0((m.get("host_abi_sbatch10") or {}).get("gcups")),
            ("%.1f GCUPS (SSE2 ksw_align2, %d cores)" % (rc["gcups"], rc["cores"])) if "gcups" in rc else "–"))
    if full and "swglobal" in full and "kernel_gcups" in full["swglobal"]:
        g = full["swglobal"]
        out.append("| next row: SWGlobal (CIGAR), 151 bp | 1 | %s | %.2f / – | %.1f M reads/s | – | %.1f GCUPS (oracle) | bit-exact (score + CIGAR) |" % (
            f0(g["kernel_gcups"]), g["roofline_frac_alu"], g["reads_per_s"] / 1e6, g["cpu_oracle_gcups"]))
    c = load("r2_coords_bench.json")
    if c:
        out.append("| next row: coordinate seam vs wire seam (%s) | 1 | – | – | – | wire %s, wire incl. C packer %s, coords %s GCUPS; H2D %.0f vs %.0f B/task | – | replies identical |" % (
            c["workload"], f0(c["wire"]["gcups"]), f0(c.get("wire_incl_packing", {}).get("gcups")), f0(c["coords"]["gcups"]),
            c["coords"]["h2d_bytes_per_task"], c["wire"]["h2d_bytes_per_task"]))
    c2 = load("r2_chain2aln_bench.json")
    if c2:
        out.append("| next row: round-flattened chain -> alignment driver | 1 | – | – | %.1f M reads/s | – | %.0f k reads/s (oracle, 1 core) | regions identical |" % (
            c2["reads_per_s"] / 1e6, c2["oracle_single_core_reads_per_s"] / 1e3))
    # A/B runs of the last session (pair steps with fewer ALU-pipe instructions), one box each, reduced shards
    ab = (load("r2_p2_variants_ab.jsonl") or [])
    by = {(r["run"].split("_v")[0], r["variant"]): r["gcups"] for r in ab}
    for wl in ("C2", "C1"):
        if (wl, 0) in by and (wl, 3) in by:
            out.append("| A/B, extension %s, 262144 pairs resident: pair step before / after the ALU-pipe relief (the rows above predate it) | 1 | %s -> **%s** | – | – | – | – | bit-exact (all kernel paths) |" % (
                wl, f0(by[(wl, 0)]), f0(by[(wl, 3)])))
    ab2 = (load("r2_alu_relief_ab.jsonl") or [])
    ms2 = {(r["workload"][:2], "exp" in r["build"]): r["gcups"] for r in ab2 if r["kernel"] == "mate-SW"}
    for wl in ("C3", "C1"):
        if (wl, False) in ms2 and (wl, True) in ms2:
            out.append("| A/B, mate-SW %s windows, 8192 pairs resident: systolic step before / after the ALU-pipe relief (the rows above predate it) | 1 | %s -> **%s** | – / %.2f | – | – | – | bit-exact (7 fields) |" % (
                wl, f0(ms2[(wl, False)]), f0(ms2[(wl, True)]), [r["roofline_frac_alu"] for r in ab2 if r["kernel"] == "mate-SW" and r["workload"][:2] == wl and "exp" in r["build"]][0]))
    table = "\n".join(out)
    p = os.path.join(ROOT, "BASELINE.md")
    s = open(p).read()
    beg, end = "<!-- results:begin (tools/make_results_table.py) -->", "<!-- results:end -->"
    if beg in s:
        s = s[:s.index(beg) + len(beg)] + "\n" + table + "\n" + s[s.index(end):]
    else:
        raise SystemExit("markers missing in BASELINE.md")
    open(p, "w").write(s)
    print(table)


if __name__ == "__main__":
    main()
