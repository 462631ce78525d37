import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import dgvit_b200 as dg
for graph in (False, True):
    a2 = dg.SAC(2, 2, "GaussianTransformer", "Transformer", False, False, True, 3407, BUFFER_SIZE=4096, buffer_size_expert=2048,
                precision="bf16", use_cuda_graph=graph, block=4, head=4, l_f_size=64, TAU=5e-4, POLICY_FREQ=1, GAMMA=0.999, ALPHA=1.0)
    a2.replay_buffer.fill_synthetic(4096, seed=1)
    a2.replay_buffer.engage_host[:4096:7] = 1.0
    a2.replay_buffer_expert.fill_synthetic(2048, seed=2)
    for _ in range(10):
        a2.learn_guidence(False, 256)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(30):
        a2.learn_guidence(False, 256)
    torch.cuda.synchronize(); t = (time.perf_counter() - t0) / 30
    print(f"graph={graph}: {t*1e3:.2f} ms per learn_guidence(256) = {256/t:.0f} samples/s, graphs kept {len(a2._graphs)}")
    a2.close()
