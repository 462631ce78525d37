#!/usr/bin/env python
"""bench.py — DGViT actor-critic update throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA kernels via the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU PyTorch path
                                                               # (oracle port, all host threads)

A "step" is one full off-policy SAC update (vn/DRL.py:373-437): replay-index gather ->
TD target -> critic fwd/bwd/Adam -> actor fwd + critic(s,pi) -> policy/alpha losses -> actor
bwd/Adam -> alpha Adam -> Polyak, on one minibatch of 256 samples per GPU (BASELINE config[1];
weak scaling: every rank runs 256 samples/step and gradients are all-reduced over NCCL).
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "actor-critic update samples/sec"
UNIT = "samples/s"
PRESET = dict(block=4, head=4, l_f_size=64)      # vn/config.yaml:5,62-63
HP = dict(LR_C=1e-3, LR_A=1e-3, LR_ALPHA=1e-4, TAU=5e-4, POLICY_FREQ=1, GAMMA=0.999, ALPHA=1.0)   # config.yaml:12-17,39-41
SEED = 3407
# algorithmic FLOPs of one update per sample: 4 F_a + 5 F_c
F_A, F_C = 190_371_072, 190_371_328
FLOP_PER_SAMPLE = 4 * F_A + 5 * F_C
# dram__bytes_read.sum + dram__bytes_write.sum of one mlp_fwd_tc_kernel launch (16640 rows, no pre-activation
# store) from the `ncu --set full` capture summarised in profiles/ (None until captured)
NCU_TRAFFIC_BYTES = 34_890_000  # profiles/r2_02_mlp_fwd_ncu.md (mlp_fwd_tc_kernel<0,1,1>, 16640 rows: 13.39 MB read + 21.5 MB written per launch)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=mx, reasons=sorted(reasons),
                    samples=len(sm))


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port of SAC.learn on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_update_rate(budget_s: float, batch: int, max_steps: int, warmup: int = 1):
    """Times oracle.SACOracle.learn (restatement of the reference's CPU PyTorch path, pinned
    against the imported reference) with all host threads.  Returns (samples/s, cores, sample str, ms/step)."""
    from oracle import dgvit_oracle as O
    from oracle.init_params import reference_sac_init, synthetic_batch, synthetic_noise
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.Cfg(dim=PRESET["l_f_size"], depth=PRESET["block"], heads=PRESET["head"])
    actor, critic = reference_sac_init(cfg, SEED)
    orc = O.SACOracle(actor, critic, cfg, lr_a=HP["LR_A"], lr_c=HP["LR_C"], lr_alpha=HP["LR_ALPHA"], gamma=HP["GAMMA"],
                      tau=HP["TAU"], alpha=HP["ALPHA"], policy_freq=HP["POLICY_FREQ"])
    b = synthetic_batch(cfg, batch, SEED)
    nz = synthetic_noise(cfg, batch, SEED + 1)
    for _ in range(warmup):
        orc.learn(b, nz)
    t0 = time.perf_counter()
    n = 0
    while n < max_steps and (n == 0 or time.perf_counter() - t0 < budget_s):
        orc.learn(b, nz)
        n += 1
    dt = time.perf_counter() - t0
    return batch * n / dt, cores, f"{n} update step(s) of batch {batch} (+{warmup} warm-up), fp32, torch CPU {cores} threads", dt / n * 1e3


def cpu_act_latency(calls: int = 60):
    """p50 of the oracle's batch-1 actor sample on the host cores (reference GoTPolicy.choose_action path)."""
    from oracle import dgvit_oracle as O
    from oracle.init_params import reference_init
    cfg = O.Cfg(dim=PRESET["l_f_size"], depth=PRESET["block"], heads=PRESET["head"])
    p = reference_init("actor", cfg, SEED)
    img, goal, eps = torch.rand(1, 128, 160), torch.rand(1, 2), torch.randn(1, 2)
    mask = torch.ones(1, cfg.n_tokens, cfg.dim)
    ts = []
    with torch.no_grad():
        for i in range(calls + 10):
            t0 = time.perf_counter()
            O.actor_sample(p, img, goal, eps, cfg, mask)
            if i >= 10:
                ts.append(time.perf_counter() - t0)
    ts.sort()
    return ts[len(ts) // 2] * 1e3


def gpu_eager_rate(batch: int, dev, autocast: bool = False):
    """Informational: the same oracle (= the reference's ATen/cuBLAS library path, eager fp32, or eager under
    torch.autocast(bfloat16): "the existing Blackwell path through libraries") on the GPU."""
    from oracle import dgvit_oracle as O
    from oracle.init_params import reference_sac_init, synthetic_batch, synthetic_noise
    cfg = O.Cfg(dim=PRESET["l_f_size"], depth=PRESET["block"], heads=PRESET["head"])
    actor, critic = reference_sac_init(cfg, SEED)
    b = {k: v.to(dev) for k, v in synthetic_batch(cfg, batch, SEED).items()}
    nz = {k: v.to(dev) for k, v in synthetic_noise(cfg, batch, SEED + 1).items()}
    with torch.device(dev):
        orc = O.SACOracle({k: v.to(dev) for k, v in actor.items()}, {k: v.to(dev) for k, v in critic.items()}, cfg,
                          gamma=HP["GAMMA"], tau=HP["TAU"], alpha=HP["ALPHA"], policy_freq=HP["POLICY_FREQ"])
        orc.log_alpha = orc.log_alpha.to(dev)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            for _ in range(3):
                orc.learn(b, nz)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            n = 10
            for _ in range(n):
                orc.learn(b, nz)
            torch.cuda.synchronize(dev)
    return dict(value=batch * n / (time.perf_counter() - t0), unit=UNIT,
                what="oracle restatement in eager PyTorch %s on the same GPU" % ("under torch.autocast(bfloat16)" if autocast else "fp32"))


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    # bounded sample: shrink the per-step batch so (steps+warmup) steps end within a few minutes
    t0 = time.perf_counter()
    rate_probe, cores, _, ms = cpu_update_rate(0.0, 32, 1, warmup=1)
    per_sample_s = 1.0 / rate_probe
    batch = args.batch
    while batch > 16 and per_sample_s * batch * (args.steps + args.warmup) > 200.0:
        batch //= 2
    rate, cores, sample, ms = cpu_update_rate(1e9, batch, args.steps, warmup=max(args.warmup, 1))
    line = dict(impl="reference", metric=METRIC, value=rate, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic",
                config=dict(workload=f"full SAC update (DGViT actor + Transformer critic, D=64 L=4 H=4), "
                                     f"CPU oracle port of vn/DRL.py:373-437, batch {batch} per step"),
                cpu_baseline=dict(value=rate, unit=UNIT, cores=cores, kind="port", sample=sample),
                e2e=dict(value=rate, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# this repo
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import dgvit_b200 as dg
    from dgvit_b200 import _lib as L
    rank, local_rank, world = dist_env()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL writes its banner ("NCCL version ...") to stdout at NCCL_DEBUG=VERSION: stdout carries the JSON line only
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        # ... and whatever NCCL still prints while the communicator comes up goes to stderr (fd-level: it is C code)
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    B = args.batch
    pk = peaks()

    ag = dg.SAC(2, 2, "GaussianTransformer", "Transformer", False, False, False, SEED, BUFFER_SIZE=args.replay,
                precision=args.precision, device=dev, distributed=world > 1, use_cuda_graph=not args.no_graph,
                **HP, **PRESET)
    ag.replay_buffer.fill_synthetic(args.replay, seed=SEED + rank)
    lib = L.lib()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- device-resident throughput (`value`)
    for _ in range(args.warmup):
        ag.learn_async(B)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        ag.learn_async(B)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    barrier()
    # the same step looped for >= 2 s (blocks of `steps` steps, median block): what the number looks like once the
    # power governor has settled; the clocks are sampled through both regions
    sustained = None
    if not args.no_sustained:
        blocks = []
        # every rank runs the SAME number of blocks (the steps hold collectives): derived from the max-over-ranks time above
        tm = torch.tensor([ms], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        nblocks = max(3, int(args.sustained_seconds * 1e3 / max(float(tm.item()), 1e-3)) + 1)
        for _ in range(nblocks):
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(args.steps):
                ag.learn_async(B)
            s1.record()
            torch.cuda.synchronize(dev)
            blocks.append(s0.elapsed_time(s1))
        tb = torch.tensor([statistics.median(blocks)], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        sustained = dict(value=B * world * args.steps / (float(tb.item()) / 1e3), unit=UNIT, blocks=len(blocks),
                         steps_per_block=args.steps, seconds=sum(blocks) / 1e3,
                         what="median block of the same step looped back to back for >= %.0f s" % args.sustained_seconds)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    # per-kernel timing of the dominant kernel family + launch count: the same steps launched eagerly
    # (CUDA events around individual launches cannot be recorded inside a replayed graph)
    graph_flag, ag.use_cuda_graph = ag.use_cuda_graph, False
    L.check(lib.dgvit_set_option(b"fork_streams", 0), "set_option")     # one stream: clean per-kernel durations
    psteps = min(args.steps, 3)
    prof = {}
    n0 = lib.dgvit_launch_count()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    tags = (("mlp_fused", L.PROF_MLP_FUSED), ("gemm_all", L.PROF_GEMM_ALL), ("attention", L.PROF_ATTENTION),
            ("adam_polyak", L.PROF_ADAM), ("layernorm_bwd", L.PROF_LN_BWD), ("patch_embed", L.PROF_EMBED),
            ("patchify", L.PROF_PATCH), ("replay_gather", L.PROF_GATHER))
    for tag_name, tag in tags:
        L.check(lib.dgvit_prof_begin(tag, psteps * 400), "prof_begin")
        for _ in range(psteps):
            ag.learn_async(B)
        torch.cuda.synchronize(dev)
        pms, pl, pfl, pby = C.c_double(), C.c_longlong(), C.c_double(), C.c_double()
        L.check(lib.dgvit_prof_end(C.byref(pms), C.byref(pl), C.byref(pfl), C.byref(pby)), "prof_end")
        prof[tag_name] = dict(ms_per_step=pms.value / psteps, launches_per_step=pl.value / psteps,
                              tflops=(pfl.value / 1e12) / (pms.value / 1e3) if pms.value > 0 else 0.0,
                              flop_per_step=pfl.value / psteps)
        if pl.value == 0 and tag_name in ("patchify",):
            prof.pop(tag_name, None)      # (the bf16 update has no patch matrix: nothing to report)
        elif pby.value > 0:       # HBM-bound family: algorithmic bytes / event time around every launch
            gbps = pby.value / 1e9 / (pms.value / 1e3) if pms.value > 0 else 0.0
            prof[tag_name] = dict(us_per_launch=pms.value * 1e3 / max(pl.value, 1), launches_per_step=pl.value / psteps,
                                  bytes_per_launch=pby.value / max(pl.value, 1), GBps=gbps, frac=gbps / pk["hbm"])
    p1.record()
    torch.cuda.synchronize(dev)
    launches = (lib.dgvit_launch_count() - n0) * args.steps // (len(tags) * psteps)
    eager_ms_per_step = p0.elapsed_time(p1) / (len(tags) * psteps)
    L.check(lib.dgvit_set_option(b"fork_streams", 1), "set_option")
    ag.use_cuda_graph = graph_flag
    barrier()
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = B * world * args.steps / (ms / 1e3)
    losses = ag._loss_buffer().tolist()

    # ---------------- end to end: pinned HOST minibatches -> H2D -> update -> D2H losses
    f = ag.replay_buffer.obs.shape[1]
    nbuf = 2
    shapes = dict(obs=(B, f), next_obs=(B, f), pobs=(B, 2), next_pobs=(B, 2), act=(B, 2), rew=(B, 1), done=(B, 1))
    g = torch.Generator().manual_seed(SEED + 7 + rank)
    host = [{k: torch.rand(*s, generator=g).pin_memory() for k, s in shapes.items()} for _ in range(4)]
    devb = [{k: torch.empty(*s, device=dev) for k, s in shapes.items()} for _ in range(nbuf)]
    h2d_bytes = sum(4 * s[0] * s[1] for s in shapes.values())
    loss_host = torch.zeros(4).pin_memory()
    copy_stream = torch.cuda.Stream(dev)
    ready = [torch.cuda.Event() for _ in range(nbuf)]
    freed = [torch.cuda.Event() for _ in range(nbuf)]
    main = torch.cuda.current_stream(dev)

    def e2e_steps(n):
        for i in range(n):
            j = i % nbuf
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[j])
                for k in shapes:
                    devb[j][k].copy_(host[i % len(host)][k], non_blocking=True)
                ready[j].record(copy_stream)
            main.wait_event(ready[j])
            ls = ag.update_from_batch_graphed(devb[j], j)
            freed[j].record(main)
            loss_host.copy_(ls, non_blocking=True)         # D2H read of the step's result
        torch.cuda.synchronize(dev)

    for j in range(nbuf):
        freed[j].record(main)
    e2e_steps(max(args.warmup, 3) * nbuf)      # each device buffer: eager warm-up, graph capture, replay
    barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    t0.record()
    e2e_steps(args.steps)
    t1.record()
    torch.cuda.synchronize(dev)
    w1 = time.perf_counter()
    e2e_ms = max(t0.elapsed_time(t1), (w1 - w0) * 1e3)     # host-inclusive: take the longer clock
    t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = B * world * args.steps / (float(t.item()) / 1e3)
    barrier()

    # ---------------- batch-1 act latency (BASELINE metric, second half): numpy in -> numpy out
    act = None
    if rank == 0:
        import numpy as np
        rs = np.random.RandomState(SEED)
        frame = rs.rand(128, 160, 1).astype(np.float32)
        goal = np.array([0.4, -0.3], dtype=np.float32)
        act = {}
        for name, ev in (("evaluate", True), ("sample", False)):
            for _ in range(100):
                ag.choose_action(frame, goal, ev)
            ts = []
            for _ in range(1000):
                t0 = time.perf_counter()
                ag.choose_action(frame, goal, ev)
                ts.append(time.perf_counter() - t0)
            ts.sort()
            act[name] = dict(p50_ms=ts[500] * 1e3, p99_ms=ts[990] * 1e3)
        act["calls"] = 1000
        act["path"] = "SAC.choose_action: pinned host frame -> CUDA graph (kernels read the pinned frame and write the action to pinned host memory: zero-copy) -> sync"
    barrier()

    dp_parity, c4, c5, next_rows = None, None, None, None
    if not args.no_extras:
        dp_parity, c4, c5 = run_extras(args, ag, dev, rank, world, dist, barrier)
        if world == 1:
            next_rows = run_next_rows(dev)

    if rank != 0:
        teardown(ag, dist)
        return

    # the dominant kernel again, 20 launches back to back between ONE pair of events (no per-launch event gap; weights
    # L2-warm as inside the step, activations rotating through 8 buffer sets = 85 MB)
    rows = B * 65
    gsrc = torch.Generator(device=dev).manual_seed(SEED)
    nset = 8
    xs = [torch.randn(rows, 64, device=dev, generator=gsrc).bfloat16() for _ in range(nset)]
    rs_ = [torch.randn(rows, 64, device=dev, generator=gsrc) for _ in range(nset)]
    outs = [torch.empty(rows, 64, device=dev) for _ in range(nset)]
    W1 = (torch.randn(2048, 64, device=dev, generator=gsrc) * 0.125).bfloat16()
    W2 = (torch.randn(64, 2048, device=dev, generator=gsrc) * 2048 ** -0.5).half()       # f16 copy, as the update reads it
    b1, b2 = torch.randn(2048, device=dev, generator=gsrc) * 0.3, torch.randn(64, device=dev, generator=gsrc)
    def mlp_burst(n):
        st_ = torch.cuda.current_stream(dev).cuda_stream
        for i in range(n):
            j = i % nset
            L.check(lib.dgvit_mlp_fwd_f16w2(xs[j].data_ptr(), W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr(),
                                            rs_[j].data_ptr(), outs[j].data_ptr(), rows, 2048, st_), "mlp_fwd_f16w2")
    mlp_burst(8)
    torch.cuda.synchronize(dev)
    gb = torch.cuda.CUDAGraph()           # replayed from a graph like the real step: no host launch cost between kernels
    with torch.cuda.graph(gb):
        mlp_burst(20)
    gb.replay()
    torch.cuda.synchronize(dev)
    b0_, b1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0_.record(); gb.replay(); b1_.record()
    torch.cuda.synchronize(dev)
    burst_us = b0_.elapsed_time(b1_) * 1e3 / 20
    del xs, rs_, outs

    # ---------------- the HBM-bound kernels of the path against the measured copy bandwidth (north_star: gather and
    # elementwise kernels "at a stated HBM-bandwidth fraction"): replay gather from the 2.46 GB store (random rows, > L2) and
    # the depth pre-processing kernel on fresh 512x640 frames (vn/env_lab.py:420-434)
    def hbm_time(fn, reps=10):
        fn(0)
        torch.cuda.synchronize(dev)
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        for i in range(reps):
            fn(i + 1)
        h1.record()
        torch.cuda.synchronize(dev)
        return h0.elapsed_time(h1) * 1e-3 / reps
    hbm = {}
    try:
        frame_floats = ag.replay_buffer.obs.shape[1]
        for gB in (B, 4096):
            gidx = [torch.randint(0, args.replay - 1, (gB,), device=dev, dtype=torch.int64) for _ in range(11)]
            z = lambda *sh: torch.zeros(*sh, dtype=torch.float32, device=dev)
            gout = dict(obs=z(gB, frame_floats), next_obs=z(gB, frame_floats), pobs=z(gB, 2), next_pobs=z(gB, 2), act=z(gB, 2),
                        rew=z(gB, 1), done=z(gB, 1))
            t = hbm_time(lambda i: ag.replay_buffer.gather(gidx[i], gout))
            by = gB * (4.0 * frame_floats * 4 + 9 * 4 * 2)          # two frames read + two written, plus the small fields
            hbm["replay_gather_B%d" % gB] = dict(us=t * 1e6, GBps=by / t / 1e9, frac=by / t / 1e9 / pk["hbm"],
                                                bytes_per_sample=by / gB)
            del gout, gidx
        nfr = 64
        raws = [torch.rand(nfr, 512, 640, device=dev) * 10.0 for _ in range(3)]      # 3 x 84 MB: rotates past L2
        douts = [torch.empty(nfr, 128, 160, device=dev) for _ in range(3)]
        drng = torch.tensor([SEED, 0], dtype=torch.int64, device=dev)

        def graph_time(fn, reps=10):
            """The same `reps` calls replayed from one CUDA graph (three launches per call: min/max pass, band rows, streamed
            rows): device time of the kernels without the host side of the eager calls; median of 9 replays."""
            fn(0)
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(reps):
                    fn(i + 1)
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize(dev)
            h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ts = []
            for _ in range(9):
                h0.record(); g.replay(); h1.record()
                torch.cuda.synchronize(dev)
                ts.append(h0.elapsed_time(h1) * 1e-3 / reps)
            return sorted(ts)[len(ts) // 2]

        f_rng = lambda i: dg.depth_augment(raws[i % 3], rng_state=drng, out=douts[i % 3])
        t, te = graph_time(f_rng), hbm_time(f_rng)
        by = nfr * 1392640.0                                         # bytes per frame (raw read once + states written)
        hbm["depth_augment_64x512x640"] = dict(us=t * 1e6, us_eager_calls=te * 1e6, GBps=by / t / 1e9, frac=by / t / 1e9 / pk["hbm"],
                                               bytes_per_frame=1392640,
                                               note="noise drawn in the kernel (Philox4x32-10 + Box-Muller per pixel): ALU-bound")
        nzs = [torch.randn(nfr, 512, 640, device=dev) * 50.0 for _ in range(3)]
        f_nz = lambda i: dg.depth_augment(raws[i % 3], noise=nzs[i % 3], out=douts[i % 3])
        t, te = graph_time(f_nz), hbm_time(f_nz)
        by = nfr * (1392640.0 + 512 * 640 * 4)
        hbm["depth_augment_64x512x640_noise_given"] = dict(us=t * 1e6, us_eager_calls=te * 1e6, GBps=by / t / 1e9,
                                                           frac=by / t / 1e9 / pk["hbm"], bytes_per_frame=1392640 + 512 * 640 * 4,
                                                           note="N(0,50) draws read from HBM (the parity-test mode)")
        del douts
        del raws, nzs
    except Exception as e:      # reporting only: never lose the bench line over it
        hbm["error"] = repr(e)[:200]

    # ---------------- roofline of the dominant kernel (the MLP GEMM family), timed live above
    dom = prof["mlp_fused"]
    ach = dom["tflops"]
    roof = dict(bound="tensor", achieved=ach, peak=pk["tf_sust"], unit="TFLOP/s", frac=ach / pk["tf_sust"],
                traffic=NCU_TRAFFIC_BYTES,
                kernel="mlp::mlp_fwd_tc_kernel (out-projection + LayerNorm-2 prologue, fused fc1 + GELU + fc2 (f16 hidden tile) + residual, next LayerNorm-1; tcgen05/TMEM)",
                how="CUDA events around every launch of the kernel in an eager single-stream replica of the timed steps "
                    "(includes the per-launch event gap; ncu reports 34-35 us per cold launch incl. the prologue)",
                back_to_back=dict(us_per_launch=burst_us, tflops=4.0 * rows * 64 * 2048 / burst_us / 1e6,
                                  frac=4.0 * rows * 64 * 2048 / burst_us / 1e6 / pk["tf_burst"], peak=pk["tf_burst"],
                                  what="20 graph-replayed launches between one event pair, %d token rows, burst bf16 peak" % rows),
                limiter=dict(what="stall-bound (ncu, profiles/r2_02_mlp_fwd_ncu.md: XU 30 %, FMA 21 %, tensor 16 % of peak; long-scoreboard "
                                  "and barrier waits with 4.5 warps per scheduler); the MUFU.TANH floor below (one per hidden element, "
                                  "16 per clock per SM) is the next hard limit: floor = rows*2048/(148*16) clocks",
                             floor_us=rows * 2048 / (148 * 16) / ((clocks or {}).get("sm_mhz") or 1965.0),
                             at_sm_mhz=(clocks or {}).get("sm_mhz") or 1965.0),
                algorithmic_flop_per_launch="4*rows*64*2048 + 2*rows*64*256 (8.72 + 0.55 GFLOP at 16640 token rows)",
                launches_per_step=dom["launches_per_step"], kernel_ms_per_step=dom["ms_per_step"],
                # share of the SERIALISED step (same single-stream eager pass the kernel was timed in: comparable with the
                # ncu launch list in profiles/, which is serialised too); in the graph-replayed step kernels of different
                # streams overlap, so kernel time / step time is not a share
                share_of_step=dom["ms_per_step"] / eager_ms_per_step if eager_ms_per_step > 0 else None,
                kernel_ms_over_graph_step_ms=dom["ms_per_step"] / (ms / args.steps) if ms > 0 else None,
                eager_ms_per_step=eager_ms_per_step, peak_source=pk["src"] + " (sustained bf16 cuBLAS)",
                other_kernels={k: v for k, v in prof.items() if k in ("gemm_all", "attention")},
                hbm_kernels=dict(peak_GBps=pk["hbm"], peak_source=pk["src"] + " (device copy bandwidth)",
                                 in_step={k: v for k, v in prof.items() if k in ("adam_polyak", "layernorm_bwd", "patch_embed", "patchify", "replay_gather")},
                                 in_step_how="CUDA events around every launch of the family inside the eager single-stream "
                                             "replica of the timed step (B=%d; includes the per-launch event gap), "
                                             "algorithmic bytes as stated in DESIGN.md §4" % B, **hbm))
    whole = dict(achieved_tflops=value * FLOP_PER_SAMPLE / 1e12, frac_of_peak=value * FLOP_PER_SAMPLE / 1e12 / pk["tf_sust"] / world)

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, cores, sample, _ = cpu_update_rate(12.0, B, 3, warmup=1)
        cpu = dict(value=rate, unit=UNIT, cores=cores, kind="port", sample=sample, act_p50_ms=cpu_act_latency(),
                   torch_eager_on_gpu=gpu_eager_rate(B, dev),
                   torch_eager_bf16_autocast_on_gpu=gpu_eager_rate(B, dev, autocast=True))

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="bf16" if args.precision == "bf16" else "f32", data="synthetic",
                config=dict(workload="full off-policy actor-critic update step, batch 256 per GPU, DGViT actor + "
                                     "Transformer critic (D=64, L=4, H=4, 65 tokens, MLP 2048), 1xB200 per rank",
                            batch_per_gpu=B, global_batch=B * world, precision=args.precision,
                            parallelism=f"dp{world}", replay_transitions=args.replay, cuda_graph=not args.no_graph,
                            l2="inputs gathered each step by random index from a %.2f GB device replay store (> 126 MB L2)"
                               % (ag.replay_buffer.obs.numel() * 4 / 1e9)),
                roofline=roof, whole_step=whole, cpu_baseline=cpu,
                value_sustained=sustained, dp_parity=dp_parity, c4=c4, c5=c5, next_rows=next_rows,
                e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d_bytes, d2h_bytes_per_step=16),
                act_latency=act, gpu_launches=int(launches), clocks=clocks, losses=losses)
    print(json.dumps(line), flush=True)
    teardown(ag, dist)


def teardown(ag, dist):
    """Captured graphs go before the communicator; a watchdog guarantees the process exits even if NCCL teardown stalls
    (the JSON line is already out)."""
    if dist is None:
        return
    threading.Timer(20.0, lambda: os._exit(0)).start()
    try:
        ag.close()
        dist.destroy_process_group()
    finally:
        os._exit(0)


def run_next_rows(dev):
    """SURVEY §8f "next" rows, measured through the agent API on one GPU (informational; B = 256, bf16):
       f1  SAC.learn with the CNN twin-Q critic (the reference's shipped default critic_type): the same single-call update;
       f3  the behaviour-cloning step as one C call (GoTPolicy.bc_step);
       f2  SAC.learn_guidence (agent + expert minibatch, guidance / engage imitation rows in the same fused update);
       f4  replay write path: store_transition one at a time (control loop) and a batched append (demonstration ingest)."""
    import numpy as np
    import dgvit_b200 as dg
    out = {}

    def rate(fn, n, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize(dev)
        return n / (time.perf_counter() - t0)
    try:
        a1 = dg.SAC(2, 2, "GaussianTransformer", "CNN", False, False, False, SEED, BUFFER_SIZE=4096, precision="bf16", device=dev,
                    use_cuda_graph=True, **HP, **PRESET)
        a1.replay_buffer.fill_synthetic(4096, seed=SEED)
        r_sync = rate(lambda: a1.learn(256), 10)
        r_async = rate(lambda: a1.learn_async(256), 20)
        out["f1_cnn_critic_learn"] = dict(value=256 * r_async, with_loss_readback=256 * r_sync, unit=UNIT,
                                          what="SAC.learn_async(256) / SAC.learn(256): QNetwork critic + DGViT actor, replay gather + "
                                               "the whole update in one C call (dgvit_sac_update, critic kind DGVIT_QNET), CUDA graph")
        a1.close()
        del a1
    except Exception as e:
        out["f1_cnn_critic_learn"] = dict(error=repr(e)[:200])
    try:
        pol = dg.GoTPolicy(2, 2, PRESET["block"], PRESET["head"], PRESET["l_f_size"]).to(dev)
        pol.precision = "bf16"
        g3 = torch.Generator(device=dev).manual_seed(SEED)
        rec = {}
        for bb in (32, 256):          # (32 = batch_size of vn/attention_imitating.py:96)
            img, goal = torch.rand(bb, 128, 160, device=dev, generator=g3), torch.rand(bb, 2, device=dev, generator=g3)
            act = torch.rand(bb, 2, device=dev, generator=g3) * 2 - 1
            rec["B%d" % bb] = bb * rate(lambda: pol.bc_step(img, goal, act), 20)
        out["f3_bc_step"] = dict(value=rec["B256"], at_batch_32=rec["B32"], unit=UNIT,
                                 what="GoTPolicy.bc_step: policy.sample -> RMSE -> backward -> clip_grad_norm_ -> Adam in one C call "
                                      "(dgvit_bc_step), eager calls, no host sync")
        del pol
    except Exception as e:
        out["f3_bc_step"] = dict(error=repr(e)[:200])
    try:
        a2 = dg.SAC(2, 2, "GaussianTransformer", "Transformer", False, False, True, SEED, BUFFER_SIZE=4096, buffer_size_expert=2048,
                    precision="bf16", device=dev, use_cuda_graph=True, **HP, **PRESET)
        a2.replay_buffer.fill_synthetic(4096, seed=SEED)
        a2.replay_buffer.engage_host[:4096:7] = 1.0
        a2.replay_buffer_expert.fill_synthetic(2048, seed=SEED + 1)
        out["f2_learn_guidence"] = dict(value=256 * rate(lambda: a2.learn_guidence(False, 256), 20, warm=8), unit=UNIT,
                                        what="SAC.learn_guidence(256): agent + expert minibatch, guidance / engage imitation rows (engage rows "
                                             "padded to a multiple of 32 with weight 0), assembly + one fused update per call, CUDA graph, "
                                             "losses read back every call")
        rs = np.random.RandomState(0)
        s0, s1 = rs.rand(128, 160, 1).astype(np.float32), rs.rand(128, 160, 1).astype(np.float32)
        one = lambda: a2.store_transition(s0, np.zeros(2, np.float32), np.zeros(2, np.float32), np.zeros(2, np.float32), 0.5, s1, 0.0, None, 0)
        obs = rs.rand(512, 128, 160).astype(np.float32)
        many = lambda: a2.replay_buffer.add(obs, np.zeros((512, 2), np.float32), np.zeros((512, 2), np.float32),
                                            np.zeros((512, 2), np.float32), np.zeros(512, np.float32), obs)
        out["f4_store_transition"] = dict(single_per_s=rate(one, 200), batched_per_s=512 * rate(many, 5, warm=1), unit="transitions/s",
                                          what="packed pinned record -> H2D -> one scatter kernel; host-side packing included")
        del a2
    except Exception as e:
        out["f2_f4"] = dict(error=repr(e)[:200])
    torch.cuda.empty_cache()
    return out


def run_extras(args, ag, dev, rank, world, dist, barrier):
    """Records beside the headline (all ranks take part):
       dp_parity  N > 1: one injected-noise fp32 update, data parallel over the N ranks, against a single-GPU replay of the
                  same GLOBAL batch on rank 0 (gradient arenas after the all-reduce, losses);
       c4         BASELINE config 4: global batch 4096 split N ways (strong scaling), samples/s and the all-reduce time;
       c5         BASELINE config 5: depth pre-processing (raw 1024x1280 -> 256x320 states written into the replay store) +
                  update of the wide / deep variant (D=128, 6 blocks, 6 heads, 257 tokens), global batch 4096 at N=8."""
    import dgvit_b200 as dg
    from dgvit_b200.parallel import shard_batch
    out = [None, None, None]

    def timed(fn, steps, warm):
        for _ in range(warm):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        return float(t.item()) / steps

    # ---- dp_parity
    if world > 1:
        try:
            Bg = 8 * world
            gen = torch.Generator().manual_seed(SEED + 21)            # same seed on every rank: the GLOBAL batch and noise
            rnd = lambda *sh: torch.rand(*sh, generator=gen)
            batch = dict(obs=rnd(Bg, 128 * 160), next_obs=rnd(Bg, 128 * 160), pobs=rnd(Bg, 2), next_pobs=rnd(Bg, 2),
                         act=rnd(Bg, 2) * 2 - 1, rew=torch.randn(Bg, 1, generator=gen) * 20, done=torch.zeros(Bg, 1))
            noise = {k: (rnd(Bg, 65, PRESET["l_f_size"]) > 0.1).float() for k in ("mask_a_next", "mask_ct", "mask_c", "mask_a", "mask_c_pi")}
            noise["eps_next"], noise["eps_pi"] = torch.randn(Bg, 2, generator=gen), torch.randn(Bg, 2, generator=gen)
            mk = lambda d: dg.SAC(2, 2, "GaussianTransformer", "Transformer", False, False, False, SEED, BUFFER_SIZE=8,
                                  precision="fp32", device=dev, distributed=d, **HP, **PRESET)
            cu = lambda d: {k: (v.to(torch.uint8) if k.startswith("mask") else v.reshape(v.shape[0], -1) if k in ("obs", "next_obs") else v).to(dev).contiguous()
                            for k, v in d.items() if v is not None}
            os.environ["DGVIT_DP_KEEP_REDUCED"] = "1"      # the fused all-reduce + Adam pass also stores the sums it read
            a_dp = mk(True)
            os.environ.pop("DGVIT_DP_KEEP_REDUCED", None)
            lb, ln = shard_batch(batch, world, rank), shard_batch(noise, world, rank)
            l_dp = a_dp.update_from_batch(cu(lb), cu(ln), global_batch=Bg).clone()
            torch.cuda.synchronize(dev)
            if rank == 0:
                a_1 = mk(False)
                l_1 = a_1.update_from_batch(cu(batch), cu(noise), global_batch=Bg, sample_offset=0).clone()
                torch.cuda.synchronize(dev)
                rel = {}
                red = a_dp.reduced_gradients()
                for nm, g_all, m_1 in (("critic_grads", red[0], a_1.critic), ("actor_grads", red[1], a_1.policy)):
                    n_used = int(m_1.layout().alpha_grad_slot) if nm == "actor_grads" else m_1._garena.numel()
                    g_dp, g_1 = g_all[:n_used], m_1._garena[:n_used]
                    rel[nm] = float((g_dp - g_1).abs().max() / g_1.abs().max().clamp_min(1e-30))
                rel["params"] = max(float((a_dp.policy._arena - a_1.policy._arena).abs().max()),
                                    float((a_dp.critic._arena - a_1.critic._arena).abs().max()))
                rel["losses"] = float((l_dp - l_1).abs().max() / l_1.abs().max().clamp_min(1e-30))
                worst = max(v for k, v in rel.items() if k != "params")
                out[0] = dict(max_abs=worst, ok=bool(worst < 1e-4 and rel["params"] < 5e-3), detail=rel, global_batch=Bg,
                              collective="fused into the Adam kernels (symmetric memory%s)" % (", NVLS multimem.ld_reduce" if a_dp._dp["multicast"] else ", peer loads")
                              if a_dp._dp is not None else "NCCL all-reduce between the phases",
                              precision="fp32", what="DP(%d) update vs single-GPU update of the same global batch: max |diff| "
                                                     "/ max |ref| of the all-reduced gradient arenas and of the losses" % world)
                del a_1
            del a_dp
        except Exception as e:
            out[0] = dict(ok=False, error=repr(e)[:300])
        barrier()

    # ---- c4: global 4096
    try:
        Bc = 4096 // world
        ms4 = timed(lambda: ag.learn_async(Bc), min(args.steps, 10), 3)
        rec = dict(global_batch=Bc * world, batch_per_gpu=Bc, ms_per_step=ms4, value=Bc * world / (ms4 / 1e3), unit=UNIT,
                   scaling="strong", whole_step_frac_of_peak=Bc * world / (ms4 / 1e3) * FLOP_PER_SAMPLE / 1e12 / peaks()["tf_sust"] / world)
        if world > 1:
            def ar():
                dist.all_reduce(ag.critic._garena)
                dist.all_reduce(ag.policy._garena)
            rec["nccl_allreduce_us_per_step"] = timed(ar, 20, 3) * 1e3      # what two NCCL all-reduces of the arenas cost
            rec["allreduce_bytes"] = 4 * (ag.critic._garena.numel() + ag.policy._garena.numel())
            rec["collective"] = "fused into the Adam kernels" if ag._dp is not None else "NCCL between the phases"
        out[1] = rec
    except Exception as e:
        out[1] = dict(error=repr(e)[:300])
    barrier()

    # ---- c5: depth pre-processing + wide / deep variant
    try:
        B5 = 4096 // world if world > 1 else 512
        n_aug = 64
        a5 = dg.SAC(2, 2, "GaussianTransformer", "Transformer", False, False, False, SEED, BUFFER_SIZE=2048, precision="bf16",
                    device=dev, distributed=world > 1, use_cuda_graph=True, image_size=(256, 320), block=6, head=6,
                    l_f_size=128, **HP)
        a5.replay_buffer.fill_synthetic(2048, seed=SEED + rank)
        raws = [torch.rand(n_aug, 1024, 1280, device=dev) * 10.0 for _ in range(2)]
        drng = torch.tensor([SEED, 0], dtype=torch.int64, device=dev)
        cur = [0]

        def step5():
            r0 = (cur[0] * n_aug) % 1920
            cur[0] += 1
            dg.depth_augment(raws[cur[0] & 1], rng_state=drng, out=a5.replay_buffer.obs[r0:r0 + n_aug])   # states -> store rows
            a5.learn_async(B5)
        ms5 = timed(step5, 5, 3)
        ms_aug = timed(lambda: dg.depth_augment(raws[0], rng_state=drng, out=a5.replay_buffer.obs[0:n_aug]), 5, 2)
        trunk5 = 20_971_520 + 6 * (75_792_384 + 50_725_632 + 50_725_632 + 25_264_128 + 269_484_032)
        out[2] = dict(batch_per_gpu=B5, global_batch=B5 * world, ms_per_step=ms5, value=B5 * world / (ms5 / 1e3), unit=UNIT,
                      depth_frames_per_step=n_aug, depth_aug_frames_per_s=n_aug * world / (ms_aug / 1e3),
                      model="D=128, 6 blocks, 6 heads, 257 tokens (256x320 frames), MLP 2048",
                      whole_step_frac_of_peak=B5 / (ms5 / 1e3) * 9 * trunk5 / 1e12 / peaks()["tf_sust"],
                      path="D=128 / 257 tokens: see DESIGN.md §4 for which kernels of this variant are tcgen05")
        del a5, raws
    except Exception as e:
        out[2] = dict(error=repr(e)[:300])
    barrier()
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="minibatch per GPU")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--replay", type=int, default=30000, help="replay store transitions (vn/config.yaml:17)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of replaying a CUDA graph")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--sustained-seconds", type=float, default=2.0)
    ap.add_argument("--no-extras", action="store_true", help="skip the dp_parity / c4 / c5 records")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
