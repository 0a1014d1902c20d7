"""Pinned host <-> device copy bandwidth of the box (one GPU): the ceiling of the end-to-end step's text upload (1.26 GB per step)
and fragment download (0.76 GB). usage: python tools/pcie_peak.py"""
import json, torch
n = 1264 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
out = {}
for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
    best = 0.0
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = max(best, n / (a.elapsed_time(b) * 1e-3) / 1e9)
    out[name + "_GBs"] = round(best, 1)
# four slices on four streams, as the pipeline's workers issue them
streams = [torch.cuda.Stream() for _ in range(4)]
q = n // 4
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for i, s in enumerate(streams):
    s.wait_event(a)
    with torch.cuda.stream(s):
        d[i * q:(i + 1) * q].copy_(h[i * q:(i + 1) * q], non_blocking=True)
for s in streams:
    torch.cuda.current_stream().wait_stream(s)
b.record(); torch.cuda.synchronize()
out["h2d_4streams_GBs"] = round(n / (a.elapsed_time(b) * 1e-3) / 1e9, 1)
print(json.dumps(out))
