"""Tiny driver for ncu captures of the batch-1 act path (SAC.choose_action): N calls, no timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import dgvit_b200 as dg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ag = dg.SAC(2, 2, "GaussianTransformer", "Transformer", False, False, False, 3407, BUFFER_SIZE=64, block=4, head=4, l_f_size=64,
            precision="bf16")
frame = np.random.rand(128, 160, 1).astype(np.float32)
goal = np.array([0.3, -0.2], dtype=np.float32)
for _ in range(n):
    a = ag.choose_action(frame, goal, evaluate=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    a = ag.choose_action(frame, goal, evaluate=True)
t1 = time.perf_counter()
print("done", a, f"{(t1 - t0) / 200 * 1e6:.1f} us per call")
