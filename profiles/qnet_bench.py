"""CNN twin-Q critic (QNetwork) micro-benchmark: forward, forward+backward and one SAC.learn with critic_type='CNN'
(CUDA events, L2 flushed between iterations).  usage: python profiles/qnet_bench.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dgvit_b200 as dg

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10):
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


img, goal = torch.rand(B, 128, 160, device=dev), torch.rand(B, 2, device=dev)
FLOP = 2 * (62 * 78 * 16 * 25 + 29 * 37 * 64 * 400 + 13 * 17 * 256 * 1600) + 2 * 2 * (290 * 128 + 128 * 32 + 32 * 2)
for prec in ("bf16", "fp32"):
    m = dg.QNetwork(2, 2).to(dev)
    m.precision = prec
    act = torch.rand(B, 2, device=dev, requires_grad=True)

    def fwd():
        with torch.no_grad():
            m([img, goal, act])

    def fwdbwd():
        q1, q2 = m([img, goal, act])
        (q1.sum() + q2.sum()).backward()

    tf, tb = timeit(fwd), timeit(fwdbwd)
    print(f"QNetwork B={B} {prec}: forward {tf:8.1f} us ({B * FLOP / tf / 1e6:6.1f} TFLOP/s algorithmic), "
          f"forward+backward {tb:8.1f} us ({3 * B * FLOP / tb / 1e6:6.1f} TFLOP/s)")
for graph in (False, True):
    ag = dg.SAC(2, 2, "GaussianTransformer", "CNN", False, False, False, 3407, BUFFER_SIZE=4096, block=4, head=4, l_f_size=64,
                precision="bf16", use_cuda_graph=graph)
    ag.replay_buffer.fill_synthetic(4096)
    for _ in range(4):
        ag.learn_async(B)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ag.learn_async(B)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e3 / 20
    print(f"SAC.learn_async critic_type='CNN' (dgvit_sac_update, graph={graph}) B={B}: {t / 1e3:.2f} ms/update = {B / t * 1e6:.0f} samples/s")
    ag.close()
