#!/usr/bin/env python
"""Host -> device staging bandwidth of the box for the copy sizes of the host seam (1-8 MB per coalesced
group): copy engine vs a pull kernel over mapped pinned memory (csbwa_h2d_probe)."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("cloud-scale-bwamem_b200")
L = pkg.lib()
assert L.csbwa_init(1) >= 1
for nbytes in (1 << 20, 2600000 // 16 * 16, 8 << 20):
    for ns in (1, 4, 8):
        a = L.csbwa_h2d_probe(nbytes, 50, 0, ns, 0)
        res = [round(L.csbwa_h2d_probe(nbytes, 50, 1, ns, g), 1) for g in (8, 32, 148)]
        print(nbytes, "bytes, streams", ns, "copy engine %.1f GB/s" % a, "pull kernel (grid 8/32/148)", res)
