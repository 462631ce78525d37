"""Where two graph-replayed trajectories of the same agent first differ (step, tensor)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dgvit_b200 as dg
from dgvit_b200 import _lib as L
for kv in filter(None, (sys.argv[1] if len(sys.argv) > 1 else "").split(",")):
    k, v = kv.split("=")
    L.check(L.lib().dgvit_set_option(k.encode(), int(v)), "opt")
STEPS = (64, 64, 128, 64, 128, 128, 64, 64, 128)
def run(graph=True):
    ag = dg.SAC(2, 2, "GaussianTransformer", "Transformer", False, False, False, 11, BUFFER_SIZE=300, TAU=5e-4,
                POLICY_FREQ=1, GAMMA=0.999, ALPHA=1.0, block=2, head=2, l_f_size=32, precision="bf16", use_cuda_graph=graph)
    ag.replay_buffer.fill_synthetic(300, seed=3)
    outs = []
    for B in STEPS:
        ag.learn_async(B)
        outs.append((ag.policy._garena.clone(), ag.critic._garena.clone(), ag._loss_buffer().clone(), ag.critic_target._arena.clone()))
    torch.cuda.synchronize()
    return ag, outs
ag0, ref = run()
nbad = 0
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 16):
    ag, outs = run()
    for si, (a, b) in enumerate(zip(outs, ref)):
        msgs = []
        for which, (x, y) in enumerate(zip(a, b)):
            if not torch.equal(x, y):
                d = (x != y).nonzero().flatten()
                mod = (ag.policy, ag.critic, None, ag.critic_target)[which]
                where = ""
                if mod is not None:
                    offs = sorted((off, n) for n, off in mod._named_offsets())
                    where = sorted({[n for off, n in offs if off <= i][-1] for i in d[:3000].tolist()})[:10]
                msgs.append(f"{('actor grads', 'critic grads', 'losses', 'target params')[which]}: {d.numel()} differ, max {float((x - y).abs().max()):.2e} {where}")
        if msgs:
            nbad += 1
            print(f"rep {rep}: first difference at step {si} (B={STEPS[si]}): " + " | ".join(msgs), flush=True)
            break
print(f"{nbad} differing trajectories")
