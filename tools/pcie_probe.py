"""Host<->device copy bandwidth of the box (pinned memory), to put the e2e leg in context."""
import json
import torch
out = {}
for mb in (1, 2, 16, 256):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for name, (src, dst) in (("h2d", (h, d)), ("d2h", (d, h))):
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        out["%s_%dMB_GBps" % (name, mb)] = round(n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9, 2)
print(json.dumps(out))
