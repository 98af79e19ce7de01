#!/usr/bin/env python
"""Device time of ONE coalesced group (k seam calls of 4096 reads) alone on the GPU: the latency floor
of the host seam.  Prints total ms (graph-free launch sequence with aux streams) and the
prepare / left / right split of the serialised profile path."""
import ctypes as C
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    pkg = importlib.import_module("cloud-scale-bwamem_b200")
    L = pkg.lib()
    assert L.csbwa_init(1) >= 1
    dev = torch.device("cuda:0")
    w = pkg.workload.ext_workload(65536, 151, 20_000_000, 0.01, 400, 50, 20260103, reads_per_call=4096)
    bufs = w["bufs"]
    out = []
    ks = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else (1, 2, 4, 8, 16, 32)
    for k in ks:
        sel = bufs[:k]
        nt = [int(np.frombuffer(b[8:12].tobytes(), dtype="<i4")[0]) for b in sel]
        tab = np.zeros(k, dtype=pkg._lib.CALL_DTYPE)
        pos = opos = tb = 0
        for i, b in enumerate(sel):
            tab[i] = (pos, b.size, nt[i], opos, tb, 0)
            pos += (b.size + 255) & ~255; opos += 10 * nt[i]; tb += nt[i]
        h_in = np.zeros(pos, dtype=np.uint8)
        for i, b in enumerate(sel):
            h_in[tab[i]["in_off"]:tab[i]["in_off"] + b.size] = b
        d_in = torch.from_numpy(h_in).to(dev)
        d_tab = torch.from_numpy(tab.view(np.uint8).copy()).to(dev)
        d_out = torch.zeros(opos, dtype=torch.int16, device=dev)
        d_cells = torch.zeros(1, dtype=torch.int64, device=dev)
        scr = torch.empty(L.csbwa_extend_scratch_bytes(tb, pos), dtype=torch.uint8, device=dev)
        st = torch.cuda.current_stream().cuda_stream

        def run():
            rc = L.csbwa_extend_multi_device(d_in.data_ptr(), tab.ctypes.data, d_tab.data_ptr(), k, d_out.data_ptr(),
                                             d_cells.data_ptr(), scr.data_ptr(), scr.numel(), C.c_void_p(st))
            assert rc == 0
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        tot = 0.0
        for _ in range(reps):
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        ms3 = (C.c_float * 3)()
        rc = L.csbwa_extend_profile_device(d_in.data_ptr(), tab.ctypes.data, d_tab.data_ptr(), k, d_out.data_ptr(),
                                           d_cells.data_ptr(), scr.data_ptr(), scr.numel(), C.c_void_p(st), ms3)
        assert rc == 0
        out.append({"calls": k, "tasks": tb, "group_ms": tot / reps, "serial_prepare_left_right_ms": [round(x, 3) for x in ms3]})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
