import sys, os
sys.path.insert(0, "/root/repo")
import torch
import dgvit_b200 as dg
B=256
m = dg.QNetwork(2, 2).to("cuda"); m.precision = "bf16"
img, goal = torch.rand(B, 128, 160, device="cuda"), torch.rand(B, 2, device="cuda")
act = torch.rand(B, 2, device="cuda", requires_grad=True)
for _ in range(2):
    q1, q2 = m([img, goal, act]); (q1.sum() + q2.sum()).backward()
torch.cuda.synchronize()
