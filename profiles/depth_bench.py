"""Depth pre-processing kernels alone (vn/env_lab.py:420-434,78-90,69-76,295-299): 64 frames of 512x640, noise given / drawn.
usage: python profiles/depth_bench.py [strip ...]   (strip = output rows per streaming strip, 0 = every row through the tiles)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dgvit_b200 as dg
from dgvit_b200 import _lib as L
n = 64
raws = [torch.rand(n, 512, 640, device="cuda") * 10 for _ in range(3)]
nzs = [torch.randn(n, 512, 640, device="cuda") * 50 for _ in range(3)]
rng = torch.tensor([3407, 0], dtype=torch.int64, device="cuda")
for arg in sys.argv[1:] or ["8"]:
    strip, _, skip = arg.partition(":")       # "8:3" = strip 8, skip mask 3 (1 band, 2 streaming, 4 min/max kernel off)
    strip = int(strip)
    L.check(L.lib().dgvit_set_option(b"depth_strip", strip), "set_option")
    L.check(L.lib().dgvit_set_option(b"depth_skip", int(skip or 0)), "set_option")
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
        # the same ten calls replayed from a CUDA graph: GPU time without the host side of the calls
        outs = [torch.empty(n, 128, 160, device="cuda") for _ in range(3)]
        fg = (lambda i: dg.depth_augment(raws[i % 3], noise=nzs[i % 3], out=outs[i % 3])) if name == "noise given" else \
             (lambda i: dg.depth_augment(raws[i % 3], rng_state=rng, out=outs[i % 3]))
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(10):
                fg(i)
        for _ in range(20):
            g.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(15):
            a.record(); g.replay(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 100)
        usg = sorted(ts)[len(ts) // 2]
        print(f"strip {arg:>4s} {name}: eager {us:.1f} us, graph {usg:.1f} us per {n} frames, {by / usg / 1e3:.0f} GB/s")
