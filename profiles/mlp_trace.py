"""Pipeline timeline of the fused MLP forward kernel (CTA 0), from a DGVIT_MLP_TRACE build (`make trace`).
Prints, per hidden chunk, SM-clock offsets of: epilogue group wait start / accumulator ready / compute done /
H buffer free / H written, and for the MMA thread: wait h_ready start / h_ready / w2_full / GEMM2 issued."""
import ctypes as C, os, sys
import torch
here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(here, "dgvit-depth-goal-guided-vision-transformer-_b200", "libdgvit_trace.so"))
lib.dgvit_mlp_bf16.argtypes = [C.c_void_p] * 12 + [C.c_int64, C.c_int, C.c_void_p]
lib.dgvit_debug_set_trace.argtypes = [C.c_void_p]
rows, hid = int(sys.argv[1]) if len(sys.argv) > 1 else 16640, 2048
dev = "cuda"
x = torch.randn(rows, 64, device=dev).bfloat16()
W1 = (torch.randn(hid, 64, device=dev) * 0.125).bfloat16()
W2 = (torch.randn(64, hid, device=dev) * hid ** -0.5).bfloat16()
b1, b2 = torch.randn(hid, device=dev) * 0.3, torch.randn(64, device=dev)
resid = torch.randn(rows, 64, device=dev); out = torch.empty(rows, 64, device=dev)
trace = torch.zeros(16 * 64, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream().cuda_stream
def f():
    rc = lib.dgvit_mlp_bf16(x.data_ptr(), W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr(), resid.data_ptr(),
                            out.data_ptr(), None, None, None, None, None, rows, hid, st)
    assert rc == 0, rc
for _ in range(3): f()
lib.dgvit_debug_set_trace(trace.data_ptr())
f(); torch.cuda.synchronize()
t = trace.cpu().view(16, 64)
t0 = int(t[t > 0].min())
names = {0: "epi wait", 1: "acc_full", 2: "computed", 3: "h_free", 4: "h_written", 8: "mma: wait h_ready", 9: "mma: h_ready", 10: "mma: w2_full", 11: "mma: gemm2 issued"}
print("chunk " + " ".join(f"{names[k]:>18s}" for k in names))
for c in range(hid // 128):
    print(f"{c:5d} " + " ".join(f"{int(t[k, c]) - t0 if t[k, c] > 0 else -1:18d}" for k in names))
ev = ["entry", "after tmem alloc+sync", "after pdl_wait", "bias staged", "loop done", "y_full", "stored", "final sync", "mma: x_full", "mma: gemm1(0) issued", "before cluster sync", "after cluster sync", "dsmem reduced"]
print("one-off events (clocks since first loop event): " + ", ".join(f"{n}={int(t[12, i]) - t0}" for i, n in enumerate(ev)))
