"""Where the batch-1 act latency goes: wall per call, GPU time of the graph replay, host-side floor (empty graph + sync)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import dgvit_b200 as dg

ag = dg.SAC(2, 2, "GaussianTransformer", "Transformer", False, False, False, 3407, BUFFER_SIZE=64, block=4, head=4, l_f_size=64,
            precision="bf16")
frame = np.random.rand(128, 160, 1).astype(np.float32)
goal = np.array([0.3, -0.2], dtype=np.float32)
for _ in range(20):
    ag.choose_action(frame, goal, evaluate=True)
torch.cuda.synchronize()
N = 500
ts = []
for _ in range(N):
    t0 = time.perf_counter(); ag.choose_action(frame, goal, evaluate=True); ts.append(time.perf_counter() - t0)
ts.sort()
print(f"wall per call: p50 {ts[N // 2] * 1e6:.1f} us  p10 {ts[N // 10] * 1e6:.1f} us")
st = ag.policy._act_state
g = st["graph"]
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
gs = []
for _ in range(N):
    ev[0].record(); g.replay(); ev[1].record(); ev[1].synchronize(); gs.append(ev[0].elapsed_time(ev[1]) * 1e3)
gs.sort()
print(f"graph replay, GPU time between events: p50 {gs[N // 2]:.1f} us")
# host floor: replay + record + synchronize of a graph with one tiny kernel
x = torch.zeros(8, device="cuda")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    x += 1
torch.cuda.synchronize()
g0 = torch.cuda.CUDAGraph()
with torch.cuda.graph(g0):
    x += 1
done = torch.cuda.Event()
hs = []
for _ in range(N):
    t0 = time.perf_counter(); g0.replay(); done.record(); done.synchronize(); hs.append(time.perf_counter() - t0)
hs.sort()
print(f"one-kernel graph replay + event sync, wall: p50 {hs[N // 2] * 1e6:.1f} us")
hs = []
for _ in range(N):
    t0 = time.perf_counter()
    st["img_host"].numpy()[...] = frame.reshape(1, 128, 160)
    st["ps_host"].numpy()[...] = goal.reshape(1, 2)
    hs.append(time.perf_counter() - t0)
hs.sort()
print(f"host staging writes: p50 {hs[N // 2] * 1e6:.1f} us")
