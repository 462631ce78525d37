"""graph-replayed vs eager trajectories (the comparison of tests/test_gpu_round2.py::test_graph_buffers_follow_the_batch_size),
fresh process state per call: python profiles/race_check2.py mode [opts]   mode: gg (graph vs graph), ge (graph vs eager), ges (with syncs)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dgvit_b200 as dg
from dgvit_b200 import _lib as L
mode = sys.argv[1]
for kv in filter(None, (sys.argv[2] if len(sys.argv) > 2 else "").split(",")):
    k, v = kv.split("=")
    L.check(L.lib().dgvit_set_option(k.encode(), int(v)), "opt")
def run(graph, sync=False):
    ag = dg.SAC(2, 2, "GaussianTransformer", "Transformer", False, False, False, 11, BUFFER_SIZE=300, TAU=5e-4,
                POLICY_FREQ=1, GAMMA=0.999, ALPHA=1.0, block=2, head=2, l_f_size=32, precision="bf16", use_cuda_graph=graph)
    ag.replay_buffer.fill_synthetic(300, seed=3)
    outs = []
    for B in (64, 64, 128, 64, 128, 128, 64, 64, 128):
        ag.learn_async(B)
        if sync:
            torch.cuda.synchronize()
        outs.append(ag._loss_buffer().clone())
    torch.cuda.synchronize()
    return torch.cat([ag.policy._arena.flatten(), ag.critic._arena.flatten()] + outs).clone()
a = run(True, mode == "ges")
b = run(mode == "gg", mode == "ges")
print("EQUAL" if torch.equal(a, b) else f"DIFF first at {int((a != b).nonzero()[0])} of {a.numel()}, n={int((a != b).sum())}")
