"""Seam-call coalescing (csrc/coalesce.hpp) exercised on the CPU with a host executor: many caller
threads, small staging limits so groups fill up and slots recycle, every reply routed back to its
own caller bit-exactly."""
import threading

import pytest

import numpy as np

from tests import emu_lib, util


def test_coalescer_many_threads(pkg, oracle, emu):
    rng = np.random.default_rng(61)
    wires = []
    for i in range(24):
        tuples = [util.rand_ext_task(rng, L=int(rng.choice([76, 101, 151]))) for _ in range(int(rng.integers(1, 60)))]
        wires.append(pkg.jni.packTasks(util.make_ext_params(pkg, tuples)))
    refs = [oracle.extend_wire(w)[0] for w in wires]
    co = emu_lib.EmuCoalescer(emu, n_slots=3, max_inflight=2, max_bytes=64 * 1024, max_tasks=400, max_calls=6, delay_us=1500)
    try:
        assert all(co.fits(w) for w in wires)
        errs = []

        def work(tid):
            for rep in range(6):
                for i in range(tid, len(wires), 8):
                    rc, out = co.submit(wires[i], zero_copy=(i % 3 == 0))
                    if rc != 0 or not np.array_equal(out, refs[i]):
                        errs.append((tid, i, rc))

        th = [threading.Thread(target=work, args=(t,)) for t in range(8)]
        [t.start() for t in th]
        [t.join() for t in th]
        assert not errs, errs[:5]
        groups, calls = co.stats()
        assert calls == 6 * len(wires)
        assert groups < calls            # calls really travelled together
        # what the executor is told about the load (it picks the small-group kernel by it): never fewer groups than it
        # really has in flight, and with two permits at most one other
        low, seen = co.others()
        assert low == 0 and seen <= 1
    finally:
        co.close()


def test_coalescer_single_caller_no_added_latency(pkg, oracle, emu):
    """An isolated call is committed immediately as a group of one."""
    rng = np.random.default_rng(62)
    wire = pkg.jni.packTasks(util.make_ext_params(pkg, [util.rand_ext_task(rng) for _ in range(10)]))
    co = emu_lib.EmuCoalescer(emu, delay_us=0)
    try:
        for _ in range(5):
            rc, out = co.submit(wire)
            assert rc == 0 and np.array_equal(out, oracle.extend_wire(wire)[0])
        assert co.stats() == (5, 5)
        assert co.others() == (0, 0)     # an isolated caller always finds the device idle
    finally:
        co.close()


def test_coalescer_rejects_mixed_headers_into_separate_groups(pkg, oracle, emu):
    """Calls with different option bytes must never share a group (the device uses one option set per group)."""
    rng = np.random.default_rng(63)
    w1 = pkg.jni.packTasks(util.make_ext_params(pkg, [util.rand_ext_task(rng) for _ in range(20)]))
    w2 = w1.copy()
    w2[7] = 1; w2[12] = 0; w2[13] = 0        # optional header: zdrop = 0
    r1, r2 = oracle.extend_wire(w1)[0], oracle.extend_wire(w2)[0]
    co = emu_lib.EmuCoalescer(emu, n_slots=4, max_inflight=1, delay_us=3000)
    try:
        res = {}

        def work(k, w):
            res[k] = co.submit(w)

        th = [threading.Thread(target=work, args=(k, w1 if k % 2 == 0 else w2)) for k in range(8)]
        [t.start() for t in th]
        [t.join() for t in th]
        for k in range(8):
            rc, out = res[k]
            assert rc == 0 and np.array_equal(out, r1 if k % 2 == 0 else r2)
    finally:
        co.close()


def test_coalescer_bad_record_fails_only_its_call(pkg, oracle, emu):
    """Fault isolation: a call whose record points outside its buffer gets BADWIRE; the calls it was
    coalesced with get their correct replies."""
    rng = np.random.default_rng(64)
    good = pkg.jni.packTasks(util.make_ext_params(pkg, [util.rand_ext_task(rng) for _ in range(12)]))
    bad = good.copy()
    bad[32 + 8:32 + 12] = np.frombuffer(np.int32(10 ** 8).tobytes(), dtype=np.uint8)     # taskPos of record 0 far outside
    ref = oracle.extend_wire(good)[0]
    co = emu_lib.EmuCoalescer(emu, n_slots=3, max_inflight=1, delay_us=4000)
    try:
        res = {}

        def work(k):
            res[k] = co.submit(bad if k == 3 else good, zero_copy=(k % 2 == 0))

        th = [threading.Thread(target=work, args=(k,)) for k in range(8)]
        [t.start() for t in th]
        [t.join() for t in th]
        groups, calls = co.stats()
        assert calls == 8 and groups < 8            # the bad call really shared a group with good ones
        for k in range(8):
            rc, out = res[k]
            if k == 3:
                assert rc == -3
            else:
                assert rc == 0 and np.array_equal(out, ref)
    finally:
        co.close()


def test_coalescer_thread_sanitizer(tmp_path):
    """The queueing code (pump thread, futex wake-ups, slot recycling, zero-copy and staged calls mixed) under
    ThreadSanitizer with a host executor: 12 caller threads, 3 slots, 4800 calls -- no data race, every reply routed."""
    import os, shutil, subprocess
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stress", "co_stress.cpp")
    exe = str(tmp_path / "co_stress")
    b = subprocess.run([gxx, "-O1", "-g", "-std=c++17", "-fsanitize=thread", "-pthread", "-o", exe, src], capture_output=True, text=True)
    if b.returncode != 0:
        pytest.skip("toolchain has no ThreadSanitizer runtime: " + b.stderr[-200:])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ThreadSanitizer" not in r.stderr, (r.stdout[-500:], r.stderr[-2000:])
    assert "bad 0" in r.stdout
