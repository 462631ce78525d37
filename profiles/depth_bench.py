"""Depth pre-processing kernels alone (vn/env_lab.py:420-434,78-90,69-76,295-299): 64 frames of 512x640, noise given / drawn.
usage: python profiles/depth_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dgvit_b200 as dg
n = 64
raws = [torch.rand(n, 512, 640, device="cuda") * 10 for _ in range(3)]
nzs = [torch.randn(n, 512, 640, device="cuda") * 50 for _ in range(3)]
rng = torch.tensor([3407, 0], dtype=torch.int64, device="cuda")
for name, fn in (("noise given", lambda i: dg.depth_augment(raws[i % 3], noise=nzs[i % 3])),
                 ("noise drawn", lambda i: dg.depth_augment(raws[i % 3], rng_state=rng))):
    fn(0); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(10):
        fn(i + 1)
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) * 100
    by = n * (1392640 + (512 * 640 * 4 if name == "noise given" else 0))
    print(f"{name}: {us:.1f} us per {n} frames, {by / us / 1e3:.0f} GB/s")
