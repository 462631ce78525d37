"""Per-CTA wall-clock stamps (globaltimer) of consecutive graph-replayed launches of the fused MLP forward kernel
(DGVIT_MLP_TRACE build): when does each CTA enter, leave griddepcontrol.wait, and exit; how long is the gap between launches."""
import ctypes as C, os, sys
import torch
here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(here, "dgvit-depth-goal-guided-vision-transformer-_b200", "libdgvit_trace.so"))
lib.dgvit_mlp_bf16.argtypes = [C.c_void_p] * 12 + [C.c_int64, C.c_int, C.c_void_p]
lib.dgvit_debug_set_trace.argtypes = [C.c_void_p]
rows, hid, dev = 16640, 2048, "cuda"
x = torch.randn(rows, 64, device=dev).bfloat16()
W1 = (torch.randn(hid, 64, device=dev) * 0.125).bfloat16(); W2 = (torch.randn(64, hid, device=dev) * hid ** -0.5).bfloat16()
b1, b2 = torch.randn(hid, device=dev) * 0.3, torch.randn(64, device=dev)
resid = torch.randn(rows, 64, device=dev); out = torch.empty(rows, 64, device=dev)
NL = 4
traces = [torch.zeros(2048 + 1024, dtype=torch.int64, device=dev) for _ in range(NL)]
def launch(i):
    lib.dgvit_debug_set_trace(traces[i].data_ptr())
    st = torch.cuda.current_stream().cuda_stream
    assert lib.dgvit_mlp_bf16(x.data_ptr(), W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr(), resid.data_ptr(),
                              out.data_ptr(), None, None, None, None, None, rows, hid, st) == 0
for i in range(NL): launch(i)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i in range(NL): launch(i)
g.replay(); torch.cuda.synchronize()
g.replay(); torch.cuda.synchronize()
n = (rows + 127) // 128
T = [t.cpu()[2048:2048 + n * 4].view(n, 4).double() for t in traces]
t0 = min(float(t[:, 0].min()) for t in T)
for i, t in enumerate(T):
    e, w, x_ = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3, (t[:, 2] - t0) / 1e3
    print(f"launch {i}: entry {e.min():7.2f}..{e.max():7.2f} us | wait done {w.min():7.2f}..{w.max():7.2f} | exit {x_.min():7.2f}..{x_.max():7.2f} "
          f"| body (wait->exit) per CTA {float((x_ - w).min()):5.2f}..{float((x_ - w).max()):5.2f} mean {float((x_ - w).mean()):5.2f}")
