#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 SW hot path (driver contract in the task brief).

Workload (BASELINE.json configs[1]): pair-end 1M x 2 reads of 151 bp simulated from
a 100 Mbp synthetic reference (eps = 1 %, insert N(400, 50)); seed-extension tasks for the longest
seed of every read, one seam call per 4096 reads (-bSWExtSize 4096).  One "step" = one pass of the
batched-extension hot path over ALL seam calls of the rank's shard.

  value : GCUPS, whole job, inputs resident in HBM, CUDA-event timed (exact DP cells counted by
          the kernels, checked against the oracle on the cpu_baseline sample)
  e2e   : same metric through the reference-facing C ABI (csbwa_extend_batch) with HOST buffers,
          H2D / D2H inside the timed region, several caller threads like Spark task threads
  N > 1 : weak scaling -- every rank owns a full shard (its own 1M pairs), no collective on the
          data path (reads shard naturally); time = max over ranks.

`--impl reference` times the reference's own CPU implementation of the path (its bwa-0.7.8 C
ksw_extend2, compiled from /root/reference into oracle/_ref, driven with the extension() control
flow; the oracle port when _ref is absent) on all host cores, on a bounded sample of the workload.
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

# more hardware queues than the default 8: the seam's submission streams and their per-class aux
# streams are independent (must be set before the CUDA context exists)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# rank 0 prints ONE JSON line on stdout: everything else that writes to file descriptor 1 (NCCL's
# version banner, library chatter) is sent to stderr, and the result line goes to the saved descriptor
_RESULT_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _RESULT_OUT.write(json.dumps(line) + "\n")
    _RESULT_OUT.flush()


ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

OPS_PER_CELL = 13          # BASELINE.md section 2 / SURVEY.md 8(d)
READS_PER_CALL = 4096
CFGS = {   # BASELINE.md section 3; the headline (and the default) is C2
    "C1": dict(name="C1", L=101, ref_bp=5_000_000, eps=0.01, mu=300, sigma=30, seed=20260102),
    "C2": dict(name="C2", L=151, ref_bp=100_000_000, eps=0.01, mu=400, sigma=50, seed=20260103),
    "C5": dict(name="C5", L=250, ref_bp=100_000_000, eps=0.05, mu=600, sigma=60, seed=20260106),
}
CFG = CFGS["C2"]


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def task_count(buf):
    return int(np.frombuffer(buf[8:12].tobytes(), dtype="<i4")[0])


def gen_workload(pkg, n_pairs, rank, ref=None, numpy_packer=False):
    """numpy_packer: build the wire buffers with the numpy restatement of the packer (byte-identical, tested) --
    the reference arm uses it so that its process never maps libcsbwa_sw.so."""
    return pkg.workload.ext_workload(n_pairs, CFG["L"], CFG["ref_bp"], CFG["eps"], CFG["mu"], CFG["sigma"],
                                     pkg.shard.shard_seed(CFG["seed"], rank), reads_per_call=READS_PER_CALL, ref=ref,
                                     chunk_pairs=max(65536, READS_PER_CALL // 2), numpy_packer=numpy_packer)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = []
        for line in open(self.f.name):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 9:
                rows.append(parts)
        os.unlink(self.f.name)
        if not rows:
            return out
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in rows:
            for k, nm in enumerate(names):
                if r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        out.update(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=float(rows[0][2]) if rows[0][2].isdigit() else None,
                   reasons=sorted(reasons), samples=len(rows),
                   power_w_max=max(float(r[3]) for r in rows if r[3].replace(".", "").isdigit()) if rows else None)
        return out


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline (the ONLY places that execute oracle/)
# ---------------------------------------------------------------------------------------------
def cpu_extension_run(bufs, n_threads, use_ref):
    """Time the CPU implementation over the given seam calls.  Returns (seconds, kind):
    kind "reference" = the reference's compiled ksw_extend2 (oracle/_ref) under the extension()
    control flow, "port" = the oracle's restatement of the Scala."""
    from oracle import oracle as O
    kind = "reference" if (use_ref and O.ref_available()) else "port"
    t0 = time.perf_counter()
    for b in bufs:
        if kind == "reference":
            O.extend_wire_ref(b, n_threads=n_threads)
        else:
            O.extend_wire(b, n_threads=n_threads, want_stats=False)
    return time.perf_counter() - t0, kind


def count_cells(bufs, n_threads):
    """Exact DP cells of the given calls, counted by the oracle (untimed)."""
    from oracle import oracle as O
    return sum(int(O.extend_wire(b, n_threads=n_threads)[1].sum()) for b in bufs)


def size_cpu_sample(bufs, n_threads, use_ref, target_s):
    """Pick how many seam calls make ~target_s seconds of CPU work (calibrated on the first call)."""
    dt, _ = cpu_extension_run(bufs[:1], n_threads, use_ref)
    n = int(max(1, min(len(bufs), round(target_s / max(dt, 1e-4)))))
    return n


def run_reference_arm(args, pkg):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    cores = os.cpu_count() or 1
    n_pairs = min(args.pairs, 262144)
    w = gen_workload(pkg, n_pairs, 0, numpy_packer=True)
    n_calls = size_cpu_sample(w["bufs"], cores, True, args.cpu_step_seconds)
    sample = w["bufs"][:n_calls]
    for _ in range(args.warmup):
        cpu_extension_run(sample[:max(1, n_calls // 4)], cores, True)
    cells_sample = count_cells(sample, cores)
    t_total, cells_total, kind = 0.0, 0, "port"
    for _ in range(args.steps):
        dt, kind = cpu_extension_run(sample, cores, True)
        t_total += dt
        cells_total += cells_sample
    gcups = cells_total / t_total / 1e9
    n_tasks = sum(task_count(b) for b in sample)
    line = {
        "impl": "reference", "metric": "seed-extension SW throughput (whole job)", "value": gcups, "unit": "GCUPS",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": config_block(args, args.pairs),          # the b200 arm's config; the sample is stated below
        "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": cores, "kind": kind,
                         "sample": "%d seam calls (%d tasks) per step, drawn from %d pairs generated with the "
                                   "workload's own generator and parameters" % (n_calls, n_tasks, n_pairs)},
        "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "read_pairs_per_s": (n_calls * READS_PER_CALL / 2) / (t_total / args.steps),
        "gpu_launches": 0,
    }
    emit(line)


def config_block(args, n_pairs):
    return {"workload": "%s: pair-end %dx2 %d bp reads vs %d Mbp synthetic reference, eps=%.2f, seed extension, "
                        "%d reads per seam call" % (CFG["name"], n_pairs, CFG["L"], CFG["ref_bp"] // 1000000, CFG["eps"], READS_PER_CALL),
            "pairs_per_gpu": n_pairs, "reads_per_call": READS_PER_CALL,
            "l2_policy": "inputs larger than L2 (all seam-call buffers of the shard stay resident, > 126 MB)",
            "parallelism": "reads sharded per GPU, no collective"}


# ---------------------------------------------------------------------------------------------
# the other BASELINE read lengths, as sub-objects of the default line
# ---------------------------------------------------------------------------------------------
def resident_leg(pkg, L, torch, dev, cfg_name, n_pairs, steps, nstreams_max, gsz, alu_peak):
    """Resident-input throughput of another BASELINE config (C1: 101 bp, C5: 250 bp at 5 % error) on the launch
    path of the headline leg -- all seam calls of a shard resident, `gsz` calls per launch sequence, launch sequences
    round-robin over streams, one CUDA graph per step -- at a reduced shard, so that the default run shows the
    kernel's fraction of the roofline at every read length.  The first call's replies and exact cell counts are
    checked against the oracle."""
    global CFG
    saved = CFG
    CFG = CFGS[cfg_name]
    try:
        w = gen_workload(pkg, n_pairs, 0)
        cfg_txt = config_block(None, n_pairs)["workload"]
    finally:
        CFG = saved
    bufs = w["bufs"]
    ntasks = [task_count(b) for b in bufs]
    CALL = pkg._lib.CALL_DTYPE
    offs, pos = [], 0
    for b in bufs:
        offs.append(pos)
        pos += (b.size + 255) & ~255
    d_in = torch.empty(pos, dtype=torch.uint8, device=dev)
    for b, o in zip(bufs, offs):
        d_in[o:o + b.size].copy_(torch.from_numpy(b))
    ooffs = np.concatenate([[0], np.cumsum([10 * n for n in ntasks])]).astype(np.int64)
    d_out = torch.zeros(int(ooffs[-1]), dtype=torch.int16, device=dev)
    d_cells = torch.zeros(1, dtype=torch.int64, device=dev)
    groups = []
    for g0 in range(0, len(bufs), gsz):
        idx = range(g0, min(len(bufs), g0 + gsz))
        tab = np.zeros(len(idx), dtype=CALL)
        tb = 0
        for j, i in enumerate(idx):
            tab[j] = (offs[i] - offs[g0], bufs[i].size, ntasks[i], int(ooffs[i] - ooffs[g0]), tb, 0)
            tb += ntasks[i]
        groups.append((g0, tab, torch.from_numpy(tab.view(np.uint8).copy()).to(dev), tb))
    nstreams = max(1, min(nstreams_max, len(groups)))
    scr_bytes = max(L.csbwa_extend_scratch_bytes(g[3], 0) for g in groups) + (64 << 20)
    scratch = [torch.empty(scr_bytes, dtype=torch.uint8, device=dev) for _ in range(nstreams)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(nstreams)]

    def enqueue_step(main_stream):
        fork = torch.cuda.Event()
        fork.record(main_stream)
        for s in streams:
            s.wait_event(fork)
        for gi, (g0, tab, d_tab, _nt) in enumerate(groups):
            rc = L.csbwa_extend_multi_device(d_in.data_ptr() + offs[g0], tab.ctypes.data, d_tab.data_ptr(), len(tab),
                                             d_out.data_ptr() + 2 * int(ooffs[g0]), d_cells.data_ptr(),
                                             scratch[gi % nstreams].data_ptr(), scr_bytes,
                                             C.c_void_p(streams[gi % nstreams].cuda_stream))
            if rc != 0:
                raise RuntimeError("csbwa_extend_multi_device: %d %s" % (rc, L.csbwa_last_error().decode()))
        for s in streams:
            j = torch.cuda.Event()
            j.record(s)
            main_stream.wait_event(j)

    main_stream = torch.cuda.current_stream()
    enqueue_step(main_stream)
    torch.cuda.synchronize()
    cells_per_step = int(d_cells.item())
    graph = None
    try:
        g = torch.cuda.CUDAGraph()
        cap_stream = torch.cuda.Stream(device=dev)
        with torch.cuda.graph(g, stream=cap_stream):
            enqueue_step(torch.cuda.current_stream())
        graph = g
    except Exception as e:
        sys.stderr.write("bench: CUDA graph capture failed in the %s leg (%s); using direct launches\n" % (cfg_name, e))
        torch.cuda.synchronize()

    def run_step():
        if graph is not None:
            graph.replay()
        else:
            enqueue_step(main_stream)

    for _ in range(3):
        run_step()
    d_out.zero_()
    d_cells.zero_()
    torch.cuda.synchronize()
    # Same timing as the headline leg when a step touches more than L2 holds (wire in + replies out): `steps` replays
    # back to back under one event pair.  A smaller shard is timed step by step with an L2 flush in between.
    step_bytes = pos + 2 * int(ooffs[-1])
    big = step_bytes > (132 << 20)                        # L2 is 126 MB
    if big:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            run_step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    else:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        evs = []
        for _ in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run_step()
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in evs)
    cells = int(d_cells.item())
    if cells != cells_per_step * steps:
        raise RuntimeError("cell counter differs between steps: %d vs %d x %d" % (cells, cells_per_step, steps))
    from oracle import oracle as O
    O.build()
    oref, ocells, _ = O.extend_wire(bufs[0], n_threads=os.cpu_count() or 1)
    got = d_out[:10 * ntasks[0]].cpu().numpy()
    d_cells.zero_()
    ms3 = (C.c_float * 3)()
    g0, tab, d_tab, _nt = groups[0]
    rc = L.csbwa_extend_profile_device(d_in.data_ptr(), tab.ctypes.data, d_tab.data_ptr(), 1, d_out.data_ptr(),
                                       d_cells.data_ptr(), scratch[0].data_ptr(), scr_bytes, C.c_void_p(0), ms3)
    torch.cuda.synchronize()
    cells_ok = rc == 0 and int(d_cells.item()) == int(ocells.sum())
    gcups = cells / (ms * 1e-3) / 1e9
    return {"workload": cfg_txt, "pairs": n_pairs, "steps": steps, "ms_per_step": ms / steps,
            "tasks_per_step": int(sum(ntasks)), "cells_per_step": cells_per_step, "value": gcups, "unit": "GCUPS",
            "read_pairs_per_s": (w["n_reads"] / 2) * steps / (ms * 1e-3),
            "roofline_frac_alu": (gcups * OPS_PER_CELL / alu_peak) if alu_peak else None,
            "cuda_graph": graph is not None, "streams": nstreams, "calls_per_launch_sequence": gsz,
            "inputs": "resident in HBM (%d MB of wire buffers + %d MB of replies per step)" % (pos >> 20, (2 * int(ooffs[-1])) >> 20),
            "l2_policy": ("inputs larger than L2: the steps run back to back under one CUDA-event pair, like the headline leg" if big else
                          "L2 flushed between the timed steps (256 MB write); each step timed by its own CUDA-event pair"),
            "parity_first_call_ok": bool(np.array_equal(got, oref)), "cells_match_oracle": bool(cells_ok)}


# ---------------------------------------------------------------------------------------------
# own arm
# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=1_000_000, help="read pairs per GPU")
    ap.add_argument("--streams", type=int, default=6)
    ap.add_argument("--group-calls", type=int, default=32, help="seam calls coalesced per launch sequence (resident leg)")
    ap.add_argument("--threads", type=int, default=0, help="caller threads per GPU of the e2e leg (0 = 64)")
    ap.add_argument("--e2e-seconds", type=float, default=3.0, help="minimum length of one timed e2e repetition")
    ap.add_argument("--cpu-step-seconds", type=float, default=6.0)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--reads-per-call", type=int, default=4096, help="experiments only; the headline is 4096")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="C2", choices=sorted(CFGS), help="BASELINE config (the headline is C2)")
    ap.add_argument("--no-matesw", action="store_true", help="skip the short mate-SW (C3 shape) leg reported under 'matesw'")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs under ncu only)")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the short C1 / C5 resident legs reported under 'other_configs'")
    ap.add_argument("--other-pairs", type=int, default=524288, help="read pairs of the C1 / C5 legs under 'other_configs'")
    ap.add_argument("--ext-mode", type=int, default=-1, help="extension core: 1 column pairs (s16x2), 0 one column per step (u8); -1 library default")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    globals()["READS_PER_CALL"] = args.reads_per_call
    globals()["CFG"] = CFGS[args.workload]
    pkg = importlib.import_module("cloud-scale-bwamem_b200")
    if args.impl == "reference":
        run_reference_arm(args, pkg)
        return

    import torch
    import torch.distributed as dist
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = pkg.lib()
    if L.csbwa_init(0) < 1:
        raise SystemExit("csbwa_init failed: " + L.csbwa_last_error().decode())
    if args.ext_mode in (0, 1):
        L.csbwa_set_ext_mode(args.ext_mode)
    ext_mode = L.csbwa_set_ext_mode(-1)

    # ---- workload (untimed) ----
    w = gen_workload(pkg, args.pairs, rank)
    bufs = w["bufs"]
    ntasks = [task_count(b) for b in bufs]
    total_tasks = sum(ntasks)
    in_bytes = sum(b.size for b in bufs)
    out_bytes = 20 * total_tasks

    # ---- device-resident state ----
    # All seam-call buffers of the shard stay resident; the launch path is the one the host seam
    # uses: calls are COALESCED, `group_calls` seam calls per multi-call launch sequence.
    CALL = pkg._lib.CALL_DTYPE
    offs, pos = [], 0
    for b in bufs:
        offs.append(pos)
        pos += (b.size + 255) & ~255
    d_in = torch.empty(pos, dtype=torch.uint8, device=dev)
    for b, o in zip(bufs, offs):
        d_in[o:o + b.size].copy_(torch.from_numpy(b))
    ooffs = np.concatenate([[0], np.cumsum([10 * n for n in ntasks])]).astype(np.int64)
    d_out = torch.zeros(int(ooffs[-1]), dtype=torch.int16, device=dev)
    d_cells = torch.zeros(1, dtype=torch.int64, device=dev)
    gsz = max(1, args.group_calls)
    groups = []          # (first call, host table, device table, n tasks)
    for g0 in range(0, len(bufs), gsz):
        idx = range(g0, min(len(bufs), g0 + gsz))
        tab = np.zeros(len(idx), dtype=CALL)
        tb = 0
        for j, i in enumerate(idx):
            tab[j] = (offs[i] - offs[g0], bufs[i].size, ntasks[i], int(ooffs[i] - ooffs[g0]), tb, 0)
            tb += ntasks[i]
        groups.append((g0, tab, torch.from_numpy(tab.view(np.uint8).copy()).to(dev), tb))
    nstreams = max(1, min(args.streams, len(groups)))
    scr_bytes = max(L.csbwa_extend_scratch_bytes(g[3], 0) for g in groups) + (64 << 20)
    scratch = [torch.empty(scr_bytes, dtype=torch.uint8, device=dev) for _ in range(nstreams)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(nstreams)]
    main_stream = torch.cuda.current_stream()

    def launch_group(gi, s_idx, cu_stream):
        g0, tab, d_tab, _nt = groups[gi]
        rc = L.csbwa_extend_multi_device(d_in.data_ptr() + offs[g0], tab.ctypes.data, d_tab.data_ptr(), len(tab),
                                         d_out.data_ptr() + 2 * int(ooffs[g0]), d_cells.data_ptr(),
                                         scratch[s_idx].data_ptr(), scr_bytes, C.c_void_p(cu_stream))
        if rc != 0:
            raise RuntimeError("csbwa_extend_multi_device: %d %s" % (rc, L.csbwa_last_error().decode()))

    def enqueue_step(main_stream):
        """All seam calls of the shard, group by group, round-robin over the streams (fork/join)."""
        fork = torch.cuda.Event()
        fork.record(main_stream)
        for s in streams:
            s.wait_event(fork)
        for gi in range(len(groups)):
            launch_group(gi, gi % nstreams, streams[gi % nstreams].cuda_stream)
        for s in streams:
            j = torch.cuda.Event()
            j.record(s)
            main_stream.wait_event(j)

    # warm-up (also sets kernel attributes before any capture)
    enqueue_step(main_stream)
    torch.cuda.synchronize()
    cells_per_step = int(d_cells.item())
    graph = None
    if not args.no_graph:
        try:
            g = torch.cuda.CUDAGraph()
            cap_stream = torch.cuda.Stream(device=dev)
            with torch.cuda.graph(g, stream=cap_stream):
                enqueue_step(torch.cuda.current_stream())
            graph = g
        except Exception as e:   # capture unsupported -> plain launches
            graph = None
            sys.stderr.write("bench: CUDA graph capture failed (%s); using direct launches\n" % e)
            torch.cuda.synchronize()

    def run_step():
        if graph is not None:
            graph.replay()
        else:
            enqueue_step(main_stream)

    for _ in range(args.warmup - 1):
        run_step()
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    launches0 = pkg.stats()["kernel_launches"]
    d_cells.zero_()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # profiling runs only (ncu --profile-from-start off): the timed replays are the profiled range
    prof = os.environ.get("CSBWA_PROFILE_STEP") == "1"
    if prof:
        torch.cuda.profiler.start()
    e0.record()
    for _ in range(args.steps):
        run_step()
    e1.record()
    torch.cuda.synchronize()
    if prof:
        torch.cuda.profiler.stop()
    if world > 1:
        dist.barrier()
    ms_total = e0.elapsed_time(e1)
    cells_total = int(d_cells.item())
    assert cells_total == cells_per_step * args.steps, (cells_total, cells_per_step)
    kernels_per_step = len(groups) * L.csbwa_extend_launches_per_call()
    result_dev = d_out.cpu().numpy()

    # ---- e2e through the C ABI with host buffers ----
    # Concurrent seam callers: Spark task threads block inside the JNI call, so executors oversubscribe the cores.
    # The number of caller threads PER GPU is the same at every N (default 64), reported as e2e.caller_threads_per_gpu.
    # Headline leg: the callers' wire and reply buffers live in pinned host memory (csbwa_host_alloc), which the seam
    # serves without a staging copy; every step still moves every input byte host -> device and every reply byte
    # device -> host inside the timed region.  Second leg (e2e.pageable): ordinary numpy buffers, one staging copy each
    # way made by the caller thread -- what a JVM caller gets through GetByteArrayRegion / SetShortArrayRegion.
    nthreads = args.threads or 64
    in_sizes1 = [b.size for b in bufs]
    out_sizes1 = [10 * n for n in ntasks]
    expect = result_dev

    def e2e_leg(ins, outs, min_seconds, reps):
        """reps timed repetitions of >= min_seconds each; a repetition = `passes` passes over the shard's call list as
        ONE stream of blocking calls from nthreads native caller threads (executor task threads persist from batch to
        batch).  Returns per-repetition seconds, passes per repetition, and the stats delta of the last repetition."""
        def mk(k):
            in_ptrs = (C.c_void_p * (len(ins) * k))(*([b.ctypes.data for b in ins] * k))
            out_ptrs = (C.c_void_p * (len(ins) * k))(*([o.ctypes.data for o in outs] * k))
            isz = np.array(in_sizes1 * k, dtype=np.int32)
            osz = np.array(out_sizes1 * k, dtype=np.int32)
            keep = (in_ptrs, out_ptrs, isz, osz)
            return lambda: (keep, L.csbwa_extend_calls(in_ptrs, isz.ctypes.data, out_ptrs, osz.ctypes.data, len(ins) * k, nthreads, local))[1]

        def run(fn):
            rc = fn()
            if rc != 0:
                raise RuntimeError("e2e call failed: %d %s" % (rc, L.csbwa_last_error().decode()))

        run(mk(2))                                           # warm-up (graphs, staging, thread start-up)
        t0 = time.perf_counter()
        run(mk(2))
        est = (time.perf_counter() - t0) / 2
        passes = int(max(args.steps, np.ceil(min_seconds / max(est, 1e-4))))
        if world > 1:                                        # same pass count on every rank
            passes = int(pkg.shard.reduce_job([passes], [0], device=dev)[0][0])
        timed = mk(passes)
        secs, st0, st1 = [], None, None
        for _ in range(reps):
            for o in outs:
                o[:] = 0                                     # a stale reply from an earlier pass cannot satisfy the check below
            torch.cuda.synchronize()
            st0 = pkg.stats()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            run(timed)
            torch.cuda.synchronize()
            secs.append(time.perf_counter() - t0)
            st1 = pkg.stats()
            if world > 1:
                dist.barrier()
            if not np.array_equal(np.concatenate(outs), expect):
                raise RuntimeError("device-resident and host-buffer paths disagree")
        return secs, passes, st0, st1

    e2e_secs, e2e_passes, st_e2e0, st_e2e1 = [float("nan")], 1, pkg.stats(), pkg.stats()
    pg_secs, pg_passes = [float("nan")], 1
    if not args.no_e2e:
        arena = pkg._lib.PinnedArena(sum(x + 512 for x in in_sizes1) + sum(2 * x + 512 for x in out_sizes1) + 4096)
        pin_ins, pin_outs = [], []
        for b, n in zip(bufs, out_sizes1):
            pi = arena.take(b.size)
            pi[:] = b
            pin_ins.append(pi)
            pin_outs.append(arena.take(2 * n, np.int16))
        e2e_secs, e2e_passes, st_e2e0, st_e2e1 = e2e_leg(pin_ins, pin_outs, args.e2e_seconds, 3)
        pg_outs = [np.zeros(n, dtype=np.int16) for n in out_sizes1]
        pg_secs, pg_passes, _, _ = e2e_leg(bufs, pg_outs, args.e2e_seconds / 2, 1)
    clocks = sampler.stop()

    # ---- phase split of the dominant kernel (roofline) ----
    ms3 = (C.c_float * 3)()
    prep = left = right = 0.0
    sample_groups = list(range(0, len(groups), max(1, len(groups) // 8)))
    sample_cells0 = int(d_cells.item())
    for gi in sample_groups:
        g0, tab, d_tab, _nt = groups[gi]
        rc = L.csbwa_extend_profile_device(d_in.data_ptr() + offs[g0], tab.ctypes.data, d_tab.data_ptr(), len(tab),
                                           d_out.data_ptr() + 2 * int(ooffs[g0]), d_cells.data_ptr(),
                                           scratch[0].data_ptr(), scr_bytes, C.c_void_p(0), ms3)
        assert rc == 0, rc
        prep += ms3[0]; left += ms3[1]; right += ms3[2]
    torch.cuda.synchronize()
    sample_cells = int(d_cells.item()) - sample_cells0
    st_after = pkg.stats()

    # ---- reduce over ranks ----
    red_t, (cells_all, tasks_all, reads_all, inb_all, outb_all) = pkg.shard.reduce_job(
        [ms_total] + [x * 1e3 for x in e2e_secs] + [x * 1e3 for x in pg_secs],
        [cells_total, total_tasks, w["n_reads"], in_bytes, out_bytes], device=dev)
    ms_total_max = red_t[0]
    e2e_rep_ms = red_t[1:1 + len(e2e_secs)]                  # per repetition: max over ranks
    pg_rep_ms = red_t[1 + len(e2e_secs):]

    if rank == 0:
        peaks = {}
        try:
            peaks = pkg._lib.int_peak(local)
        except Exception as e:
            sys.stderr.write("int_peak failed: %s\n" % e)
        mp = {}
        mp_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(mp_path):
            mp = json.load(open(mp_path))
        gcups = cells_all / (ms_total_max * 1e-3) / 1e9
        cells_step_all = cells_all / args.steps               # all ranks, one pass over every shard
        e2e_rep_gcups = [cells_step_all * e2e_passes / (t * 1e-3) / 1e9 for t in e2e_rep_ms]
        e2e_gcups = float(np.median(e2e_rep_gcups))
        e2e_ms_step = float(np.median(e2e_rep_ms)) / e2e_passes
        pg_gcups = cells_step_all * pg_passes / (pg_rep_ms[0] * 1e-3) / 1e9
        n_sub = max(1, st_e2e1["ext_groups"] - st_e2e0["ext_groups"])
        alu_peak = peaks.get("VIADDMNMX")                      # 1e9 ALU-pipe thread-instr/s (max-class ops)
        dual_peak = peaks.get("IADD3")                         # adds dual-issue onto the FMA pipe as IMAD.IADD
        side_ms = left + right
        side_gops = sample_cells * OPS_PER_CELL / (side_ms * 1e-3) / 1e9 if side_ms > 0 else None
        hbm_peak = mp.get("hbm_gbs", 6650.0)
        # DRAM bytes of the dominant kernels come from an ncu capture of the CURRENT round's build (profiles/,
        # written by tools/ncu_traffic.py from the committed CSV); bench.py cannot run ncu inside itself
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "r2_roofline_traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            traffic = tj["dram_bytes_per_cell"] * cells_all / args.steps / world      # per GPU, one step
            traffic_src = tj.get("source", "") + "; carried per DP cell (%.4f B/cell) and scaled by this step's cells" % tj["dram_bytes_per_cell"]
        line = {
            "metric": "seed-extension SW throughput (whole job)", "value": gcups, "unit": "GCUPS",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16 (DPX s16x2 lanes, int32 insertion chain)" if ext_mode == 1 else "int32 (u8-range scores, DPX)",
            "data": "synthetic", "config": config_block(args, args.pairs),
            "read_pairs_per_s": (reads_all / 2) * args.steps / (ms_total_max * 1e-3),
            "tasks_per_step": tasks_all, "cells_per_step": cells_all / args.steps,
            "e2e": {"value": e2e_gcups, "unit": "GCUPS", "h2d_bytes_per_step": inb_all, "d2h_bytes_per_step": outb_all,
                    "read_pairs_per_s": (reads_all / 2) / (e2e_ms_step * 1e-3),
                    "ms_per_step": e2e_ms_step, "caller_threads_per_gpu": nthreads,
                    "repetitions_gcups": e2e_rep_gcups, "min": min(e2e_rep_gcups), "max": max(e2e_rep_gcups),
                    "passes_per_repetition": e2e_passes, "seconds_per_repetition": [t * 1e-3 for t in e2e_rep_ms],
                    "host_buffers": "pinned (csbwa_host_alloc): zero host staging copies; every pass moves all input bytes H2D and "
                                    "all reply bytes D2H inside the timed region",
                    "zero_copy_calls_frac": (st_e2e1["ext_zero_copy_calls"] - st_e2e0["ext_zero_copy_calls"]) /
                                            max(1, st_e2e1["ext_calls"] - st_e2e0["ext_calls"]),
                    "pageable": {"value": pg_gcups, "unit": "GCUPS", "passes": pg_passes, "seconds": pg_rep_ms[0] * 1e-3,
                                 "host_buffers": "ordinary numpy arrays: one staging copy each way by the caller thread"},
                    "api": "csbwa_extend_batch (blocking, host buffers; concurrent calls coalesced into one device submission: "
                           "gather kernel H2D + launch sequence + scatter kernel D2H per group); value = median of 3 repetitions, "
                           "each one stream of blocking calls from caller threads that persist across passes"},
            "gpu_launches": int(kernels_per_step * args.steps * world),
            "ext_core": {0: "u8, one column per step",
                         1: "p2 s16x2 (two adjacent query columns per DPX instruction)"}.get(ext_mode),
            "cuda_graph": graph is not None, "streams": nstreams, "calls_per_launch_sequence": gsz,
            "e2e_calls_per_device_submission": (st_e2e1["ext_calls"] - st_e2e0["ext_calls"]) / n_sub,
            "e2e_ms_launch_to_done_per_submission": (st_e2e1["host_ms"] - st_e2e0["host_ms"]) / n_sub,
            "host_cpus": os.cpu_count(),
            "clocks": clocks,
            "roofline": {
                "bound": "int_alu", "kernel": "k_ext_side (left + right, all size classes)",
                "achieved": gcups / world * OPS_PER_CELL, "peak": (alu_peak or 0), "unit": "Gop/s per GPU",
                "frac": (gcups / world * OPS_PER_CELL / alu_peak) if alu_peak else None,
                "frac_alu_pipe": (gcups / world * OPS_PER_CELL / alu_peak) if alu_peak else None,
                "frac_dual_issue": (gcups / world * OPS_PER_CELL / dual_peak) if dual_peak else None,
                "ops_per_cell": OPS_PER_CELL,
                "how": "PER GPU: algorithmic int32 ops of the timed steps (13 x exact DP cells, all ranks) / number of GPUs / "
                       "CUDA-event time of the steps (max over ranks); the side kernels overlap across streams, their share of "
                       "a step is side_kernel_share.  frac = frac_alu_pipe: against the measured max-class (VIADDMNMX) issue rate, "
                       "one ALU pipe; frac_dual_issue: against the measured IADD3 rate, where adds dual-issue on the FMA pipe. "
                       "Packed s16x2 instructions update two cells each, so neither fraction is a pipe-utilisation counter.",
                "isolated_launch_gops": side_gops,
                "isolated_launch_frac": (side_gops / alu_peak) if (side_gops and alu_peak) else None,
                "peak_source": "measured here: dependent-free VIADDMNMX stream on all SMs (csbwa_int_peak), 1e9 thread-instr/s",
                "phase_ms_sample": {"prepare": prep, "left": left, "right": right, "groups": len(sample_groups),
                                    "side_kernel_share": (left + right) / max(prep + left + right, 1e-9)},
                "traffic": traffic, "traffic_source": traffic_src,
                "hbm": {"achieved": (inb_all + outb_all) / world * args.steps / (ms_total_max * 1e-3) / 1e9, "peak": hbm_peak,
                        "unit": "GB/s", "peak_source": "MEASURED_PEAKS.json" if mp else "fallback"},
            },
            "int_peaks_ginstr_per_s": peaks,
        }
        if world == 1 and not args.no_matesw:
            # the other half of the north-star path (pair-end mate rescue, SWAlign2), C3 window shape, short run
            try:
                from tools import bench_matesw
                m = bench_matesw.run_config(pkg, "C3", 8192, pairs_per_call=4096, steps=3, cpu_sample_jobs=256, peaks=peaks, device=local)
                line["matesw"] = {k: m[k] for k in ("workload", "jobs", "kernel_gcups", "roofline_frac_alu", "host_abi_large",
                                                    "host_abi_sbatch10", "cpu_oracle_gcups", "parity_sample_ok")}
                line["matesw"]["reference_cpu"] = bench_matesw.run_reference(pkg, "C3", 1024, 3.0)
            except Exception as e:       # never lose the headline line over the extra leg
                line["matesw"] = {"error": repr(e)}
            # and the next row of the path (SWGlobal / CIGAR generation), short run
            try:
                from tools import bench_global
                g = bench_global.run(pkg, pairs=65536, steps=3, cpu_sample_jobs=1024, device=local, peaks=peaks)
                line["swglobal"] = {k: g[k] for k in ("workload", "jobs", "kernel_gcups", "reads_per_s", "roofline_frac_alu",
                                                      "cpu_oracle_gcups", "parity_sample_ok")}
            except Exception as e:
                line["swglobal"] = {"error": repr(e)}
        if world == 1 and args.workload == "C2" and not args.no_other_configs and not args.no_matesw:
            # the other read lengths BASELINE names (C1: 101 bp, C5: 250 bp at 5 % error): resident-input legs on the
            # headline's launch path at a reduced shard -- the kernel's fraction of the roofline is length-dependent
            line["other_configs"] = {}
            for name, pairs_x, steps_x in (("C1", 2 * args.other_pairs, 5), ("C5", args.other_pairs, 3)):
                try:                     # C1 at twice the pairs: its tasks are small, 1M pairs make a step of ~ 145 MB (> L2)
                    line["other_configs"][name] = resident_leg(pkg, L, torch, dev, name, pairs_x, steps_x, args.streams,
                                                               gsz, alu_peak)
                except Exception as e:   # never lose the headline line over the extra legs
                    line["other_configs"][name] = {"error": repr(e)}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle as O
            O.build()
            cores = os.cpu_count() or 1
            n_calls = size_cpu_sample(bufs, cores, True, 12.0)
            ccells1 = count_cells(bufs[:n_calls], cores)
            dt, ccells, passes, kind = 0.0, 0, 0, "port"
            while dt < 10.0 and passes < 64:           # ~10-30 s of CPU work on the bounded sample
                d1, kind = cpu_extension_run(bufs[:n_calls], cores, True)
                dt += d1; ccells += ccells1; passes += 1
            # the kernels' cell counter must agree with the oracle's on the same calls
            ocells = count_cells(bufs[:min(4, n_calls, gsz)], cores)
            d_cells.zero_()
            g0, tab, d_tab, _nt = groups[0]
            k = min(4, n_calls, len(tab))
            rc = L.csbwa_extend_profile_device(d_in.data_ptr() + offs[g0], tab.ctypes.data, d_tab.data_ptr(), k,
                                               d_out.data_ptr() + 2 * int(ooffs[g0]), d_cells.data_ptr(),
                                               scratch[0].data_ptr(), scr_bytes, C.c_void_p(0), ms3)
            assert rc == 0, rc
            torch.cuda.synchronize()
            line["cells_match_oracle"] = bool(int(d_cells.item()) == ocells)
            line["cpu_baseline"] = {"value": ccells / dt / 1e9, "unit": "GCUPS", "cores": cores, "kind": kind,
                                    "sample": "first %d seam calls (%d tasks) of the same workload x %d passes, %.1f s" %
                                              (n_calls, sum(ntasks[:n_calls]), passes, dt)}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
