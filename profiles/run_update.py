"""Tiny driver for ncu captures: N fused updates at B=256 (bf16) on one GPU, no timing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dgvit_b200 as dg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
ag = dg.SAC(2, 2, "GaussianTransformer", "Transformer", False, False, False, 3407, LR_C=1e-3, LR_A=1e-3, LR_ALPHA=1e-4,
            BUFFER_SIZE=2048, TAU=5e-4, POLICY_FREQ=1, GAMMA=0.999, ALPHA=1.0, block=4, head=4, l_f_size=64,
            precision="bf16")
ag.replay_buffer.fill_synthetic(2048)
for _ in range(n):
    ag.learn_async(B)
torch.cuda.synchronize()
print("done", ag._losses.tolist())
