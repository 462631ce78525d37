"""Kernel micro-benchmarks through the C ABI (CUDA events, L2 flushed between iterations).
usage: python profiles/microbench.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgvit_b200 import _lib as L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = B * 65
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


def linear(rows, N, K, epi, kn=0):
    x = torch.randn(rows, K, device=dev).bfloat16()
    W = (torch.randn(K, N, device=dev) if kn else torch.randn(N, K, device=dev)).bfloat16()
    y = torch.empty(rows, N, device=dev, dtype=torch.bfloat16)
    y2 = torch.empty_like(y)
    bias = torch.randn(N, device=dev)
    aux = torch.randn(rows, N, device=dev).bfloat16()
    def fn():
        L.check(lib.dgvit_linear_bf16(x.data_ptr(), W.data_ptr(), y.data_ptr(), rows, N, K, epi, bias.data_ptr(),
                                      aux.data_ptr(), y2.data_ptr(), kn, st))
    return timeit(fn)


def attention(bwd):
    H = 4
    qkv = torch.randn(T, 3 * H * 64, device=dev).bfloat16()
    o = torch.empty(T, H * 64, device=dev, dtype=torch.bfloat16)
    do = torch.randn(T, H * 64, device=dev).bfloat16()
    dq = torch.empty_like(qkv)
    def f():
        L.check(lib.dgvit_attention_bf16(qkv.data_ptr(), o.data_ptr(), None, None, B, 65, H, 64, 1, None, st))
    def b():
        L.check(lib.dgvit_attention_bf16(qkv.data_ptr(), o.data_ptr(), do.data_ptr(), dq.data_ptr(), B, 65, H, 64, 1, None, st))
    f()
    return timeit(b if bwd else f)


import os
if os.environ.get("DBG"): L.check(lib.dgvit_set_option(b"debug_epilogue", int(os.environ["DBG"])))
print(f"B={B} T={T} DBG={os.environ.get('DBG')}")
for name, args in [("QKV      x[T,64]  W[768,64]           none", (T, 768, 64, 0)),
                   ("fc1      x[T,64]  W[2048,64]   bias+gelu2", (T, 2048, 64, 1)),
                   ("dH       dy[T,64] W2[64,2048]    gelu_bwd", (T, 2048, 64, 2, 1)),
                   ("dH+Hact  dy[T,64] W2[64,2048]   gelu_bwd2", (T, 2048, 64, 3, 1)),
                   ("dO       dy[T,64] Wo[64,256]         none", (T, 256, 64, 0, 1))]:
    us = linear(*args)
    rows, N, K = args[:3]
    print(f"{name}: {us:8.1f} us   {2 * rows * N * K / us / 1e6:7.1f} TFLOP/s   out {rows * N * 2 / us / 1e3:7.1f} GB/s")
print(f"attention fwd: {attention(False):8.1f} us")
print(f"attention bwd: {attention(True):8.1f} us")

# ---- HBM-bound kernels: achieved bandwidth on ALGORITHMIC bytes
import ctypes as C
import dgvit_b200 as dg
store = dg.ReplayStore(30000, (128, 160), 2, 2, dev, seed=1)
store.fill_synthetic(30000, seed=2)
for Bg in (256, 4096):
    idx = torch.randint(0, 30000, (Bg,), device=dev)
    out = {k: torch.empty(Bg, w, device=dev) for k, w in dict(obs=20480, next_obs=20480, pobs=2, next_pobs=2, act=2, rew=1, done=1).items()}
    us = timeit(lambda: store.gather(idx, out))
    print(f"replay gather B={Bg}: {us:8.1f} us   {Bg * 327752 / us / 1e3:7.1f} GB/s (algorithmic 327,752 B/sample)")
for n in (1, 64):
    raw = torch.rand(n, 512, 640, device=dev) * 8
    noise = torch.randn(n, 512, 640, device=dev) * 50
    us = timeit(lambda: dg.depth_augment(raw, noise))
    print(f"depth augment n={n}: {us:8.1f} us   {n * (1392640 + 1310720) / us / 1e3:7.1f} GB/s (algorithmic 2,703,360 B/frame with injected noise)")
    rng = torch.tensor([1, 2], dtype=torch.int64, device=dev)
    us = timeit(lambda: dg.depth_augment(raw, None, rng))
    print(f"depth augment n={n} (in-kernel noise): {us:8.1f} us   {n * 1392640 / us / 1e3:7.1f} GB/s (algorithmic 1,392,640 B/frame)")
