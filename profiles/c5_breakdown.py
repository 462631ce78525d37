"""Where the time of one update of the wide / deep variant (BASELINE config 5: D=128, 6 blocks, 6 heads, 257 tokens) goes:
CUDA events around the launches of each kernel family in an eager single-stream pass (dgvit_prof_begin / _end).
usage: python profiles/c5_breakdown.py [batch]"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dgvit_b200 as dg
from dgvit_b200 import _lib as L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
lib = L.lib()
ag = dg.SAC(2, 2, "GaussianTransformer", "Transformer", False, False, False, 1, BUFFER_SIZE=1024, precision="bf16", image_size=(256, 320),
            block=6, head=6, l_f_size=128, TAU=5e-4, POLICY_FREQ=1, GAMMA=0.999, ALPHA=1.0)
ag.replay_buffer.fill_synthetic(1024, seed=2)
for _ in range(2):
    ag.learn_async(B)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ag.learn_async(B); e1.record(); torch.cuda.synchronize()
print(f"B={B}: eager step {e0.elapsed_time(e1):.2f} ms")
L.check(lib.dgvit_set_option(b"fork_streams", 0))
for name, tag in (("attention", L.PROF_ATTENTION), ("gemm_all", L.PROF_GEMM_ALL), ("mlp_fused", L.PROF_MLP_FUSED), ("gemm_mlp", L.PROF_GEMM_MLP),
                  ("ln_bwd", L.PROF_LN_BWD), ("embed", L.PROF_EMBED), ("adam", L.PROF_ADAM)):
    L.check(lib.dgvit_prof_begin(tag, 4000))
    ag.learn_async(B)
    torch.cuda.synchronize()
    ms, n, fl, by = C.c_double(), C.c_longlong(), C.c_double(), C.c_double()
    L.check(lib.dgvit_prof_end(C.byref(ms), C.byref(n), C.byref(fl), C.byref(by)))
    print(f"  {name:10s} {ms.value:9.2f} ms  {n.value:5d} launches  {fl.value / 1e9 / max(ms.value, 1e-9):8.1f} TFLOP/s")
