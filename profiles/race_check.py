"""Run-to-run determinism of the eager (graph-free) update under library options: python profiles/race_check.py "opt=v,opt=v" ..."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dgvit_b200 as dg
from dgvit_b200 import _lib as L
def run(graph, D=32):
    ag = dg.SAC(2, 2, "GaussianTransformer", "Transformer", False, False, False, 11, BUFFER_SIZE=300, TAU=5e-4,
                POLICY_FREQ=1, GAMMA=0.999, ALPHA=1.0, block=2, head=2, l_f_size=D, precision="bf16", use_cuda_graph=graph)
    ag.replay_buffer.fill_synthetic(300, seed=3)
    for B in (64, 64, 128, 64, 128, 128, 64, 64, 128):
        ag.learn_async(B)
    torch.cuda.synchronize()
    return torch.cat([ag.policy._arena.flatten(), ag.critic._arena.flatten(), ag._loss_buffer().flatten()]).clone()
base = dict(pdl=1, target_fork=1, bwd_side=1, fork_streams=1, actor_s_when=1)
for arg in sys.argv[1:] or [""]:
    opts = dict(base)
    D = 32
    for kv in filter(None, arg.split(",")):
        k, v = kv.split("=")
        if k == "D":
            D = int(v)
        else:
            opts[k] = int(v)
    for k, v in opts.items():
        L.check(L.lib().dgvit_set_option(k.encode(), v), "opt")
    ref = run(False, D)
    bad = sum(0 if torch.equal(run(False, D), ref) else 1 for _ in range(12))
    print(f"[{arg}] eager: {bad}/12 runs differ from the first", flush=True)
