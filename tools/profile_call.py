"""One warm-up call and one profiled call of dofs3d_process_dev on synthetic 1080p video (for ncu launch lists):
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/profile_call.py [pairs]
Also prints the per-stage CUDA-event times of the second call when run without ncu."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import denseopticalflowsegmentation3d_b200 as d
from denseopticalflowsegmentation3d_b200.capi import BOX_DTYPE, STATS_DTYPE

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
W, H, MAXB = 1920, 1080, 256
dev = torch.device("cuda", 0)
with d.Context(W, H, max_pairs=n) as c:
    fr = torch.empty((n + 1, H, W, 3), dtype=torch.uint8, device=dev)
    c.synth_frames_dev(1234, 8, 0, n + 1, fr.data_ptr())
    labels = torch.empty((n, H, W), dtype=torch.int32, device=dev)
    boxes = torch.empty((n, MAXB * BOX_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    nbox = torch.empty((n,), dtype=torch.int32, device=dev)
    stats = torch.empty((n, STATS_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    for it in range(2):
        c.set_timing(it == 1)
        c.process_dev(fr.data_ptr(), n + 1, labels.data_ptr(), boxes.data_ptr(), nbox.data_ptr(), MAXB, stats.data_ptr())
        c.sync()
    t = c.timing()
    tot = sum(v[0] for v in t.values())
    print(f"{n} pairs: {tot:.2f} ms, {1e3 * n / tot:.1f} pairs/s alone; boxes {nbox.sum().item()}")
    for k, (ms, cnt) in sorted(t.items(), key=lambda kv: -kv[1][0]):
        print(f"  {k:22s} {ms:8.3f} ms  ({cnt} marks)")
