"""Fused MLP forward / backward micro-benchmark through the C ABI (CUDA events, L2 flushed between iterations).
usage: python profiles/mlp_bench.py [rows ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgvit_b200 import _lib as L

dev = "cuda"
lib = L.lib()
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def run(rows, hid=2048):
    x = torch.randn(rows, 64, device=dev).bfloat16()
    W1 = (torch.randn(hid, 64, device=dev) * 0.125).bfloat16()
    W2 = (torch.randn(64, hid, device=dev) * hid ** -0.5).bfloat16()
    b1, b2 = torch.randn(hid, device=dev) * 0.3, torch.randn(64, device=dev)
    resid = torch.randn(rows, 64, device=dev)
    out = torch.empty(rows, 64, device=dev)
    dy = (torch.randn(rows, 64, device=dev) * 1e-3).bfloat16()
    dx = torch.empty(rows, 64, device=dev)
    dw = torch.empty(hid * 129, device=dev)
    db2 = torch.empty(64, device=dev)
    part = torch.empty(lib.dgvit_mlp_partial_floats(rows, hid), device=dev)
    def f():
        L.check(lib.dgvit_mlp_bf16(x.data_ptr(), W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr(), resid.data_ptr(),
                                   out.data_ptr(), None, None, None, None, None, rows, hid, st))
    def b():
        L.check(lib.dgvit_mlp_bf16(x.data_ptr(), W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), None, None, None, dy.data_ptr(),
                                   dx.data_ptr(), dw.data_ptr(), db2.data_ptr(), part.data_ptr(), rows, hid, st))
    W2h = W2.float().half()
    def f16():
        L.check(lib.dgvit_mlp_fwd_f16w2(x.data_ptr(), W1.data_ptr(), b1.data_ptr(), W2h.data_ptr(), b2.data_ptr(), resid.data_ptr(),
                                        out.data_ptr(), rows, hid, st))
    tf, tb, tf16 = timeit(f), timeit(b), timeit(f16)
    fl = 4.0 * rows * 64 * hid
    print(f"rows={rows:6d} hid={hid}: fwd (f16 hidden tile) {tf16:7.1f} us ({fl / tf16 / 1e6:6.1f} TFLOP/s)")
    print(f"rows={rows:6d} hid={hid}: fwd {tf:7.1f} us ({fl / tf / 1e6:6.1f} TFLOP/s)   bwd (dX + dW + 2 reduces) {tb:7.1f} us "
          f"({2 * fl / tb / 1e6:6.1f} TFLOP/s algorithmic)")


for r in ([int(a) for a in sys.argv[1:]] or [256, 16640, 65 * 2048]):
    run(r)
