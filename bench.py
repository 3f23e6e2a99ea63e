#!/usr/bin/env python
"""bench.py -- sequences/sec of the MegaCRN training step (12-step encoder + 12-step decoder,
forward + trainer loss + backward [+ one gradient all-reduce when N > 1]) on N B200s.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference        # the reference's CPU path (oracle port) on host cores

Workload = BASELINE.json configs[1]: METR-LA shape N=207, T_in=T_out=12, H=64, batch 64 per GPU
(weak scaling), synthetic inputs, random-init weights.  Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIGS = {   # name -> (Dims kwargs, per-GPU batch, T_in)
    "c2": (dict(num_nodes=207, horizon=12, rnn_units=64), 64, 12),
    "c3": (dict(num_nodes=325, horizon=12, rnn_units=64), 64, 12),
    "c4": (dict(num_nodes=1843, horizon=6, rnn_units=64), 32, 6),
    "c5": (dict(num_nodes=2841, horizon=12, rnn_units=128), 32, 12),
}


def fwd_flops(d, B, t_in):
    """Algorithmic forward FLOPs: identity blocks and hoisted T2 counted once."""
    N, H, D, M, dm = d.num_nodes, d.rnn_units, d.rnn_units + d.mem_dim, d.mem_num, d.mem_dim
    def agcn(C, O):
        return 2 * 4 * N * N * B * C + 2 * B * N * 6 * C * O
    enc = t_in * (agcn(d.input_dim + H, 2 * H) + agcn(d.input_dim + H, H))
    cd = d.output_dim + d.ycov_dim
    dec = d.horizon * (agcn(cd + D, 2 * D) + agcn(cd + D, D))
    misc = 4 * N ** 3 + 4 * N * N * dm + 4 * N * M * dm + 2 * B * N * H * dm + 4 * B * N * dm * M + 2 * d.horizon * B * N * D
    return enc + dec + misc


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        j = json.load(open(path))
        return dict(bf16_burst=j["bf16_tflops"], bf16_sustained=j["bf16_tflops_sustained"], hbm=j["hbm_gbs"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_threads():
    """Threads for the CPU arm.  torch's intra-op pool stops scaling on these op sizes well before a
    128-core host is full (measured: 128 threads -> 0.57 seq/s, 8 threads -> ~35 seq/s on the same
    workload), so the arm uses at most 16 and reports the number it used."""
    return max(1, min(os.cpu_count() or 1, int(os.environ.get("MCRN_CPU_THREADS", "16"))))


def cpu_reference_step(d, B, t_in, threads):
    """One training step of the reference's CPU path: the oracle port (torch CPU ops + autograd),
    same workload.  Returns a callable."""
    from oracle import megacrn_oracle as O
    torch.set_num_threads(threads)
    p = O.init_params(d, seed=0)
    x, y_cov, labels = O.synthetic_batch(d, B, t_in, seed=1234)
    flags = [True] * d.horizon

    def step():
        loss, _, _ = O.loss_and_grads(d, p, x, y_cov, labels, flags)
        return float(loss)
    return step


def run_reference_arm(args, d, B, t_in):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = cpu_threads()
    # bounded sample: pick the per-step batch so that warmup+steps stay within ~2 minutes
    probe = cpu_reference_step(d, 8, t_in, threads)
    probe()
    t0 = time.perf_counter(); probe(); per_seq = (time.perf_counter() - t0) / 8
    budget = 120.0 / max(1, args.steps + args.warmup)
    Bs = B
    while Bs > 8 and per_seq * Bs > budget:
        Bs //= 2
    step = cpu_reference_step(d, Bs, t_in, threads)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = Bs * args.steps / dt
    B_full, B = B, Bs
    line = {
        "impl": "reference", "metric": "sequences/sec (12-step enc+dec fwd+bwd)", "value": val, "unit": "sequences/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: N={d.num_nodes} T={t_in}/{d.horizon} H={d.rnn_units} batch={B_full}, "
                               f"train step fwd+loss+bwd on host CPU, {B} sequences per timed step"},
        "cpu_baseline": {"value": val, "unit": "sequences/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps of batch {B} (oracle port of the reference, torch CPU fp32)"},
        "e2e": {"value": val, "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=list(CONFIGS))
    ap.add_argument("--engine", default="default", choices=["default", "simt"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every kernel from the host instead of replaying CUDA graphs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)

    from oracle import megacrn_oracle as O   # Dims + synthetic inputs + (cpu_baseline leg only) the CPU port
    kw, B, t_in = CONFIGS[args.config]
    d = O.Dims(**kw)
    if args.impl == "reference":
        run_reference_arm(args, d, B, t_in)
        return

    import torch.distributed as dist
    from megacrn_b200 import MegaCRN, _abi
    from megacrn_b200.ddp import allreduce_gradients
    from megacrn_b200.train_step import GraphedTrainStep, train_step

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert world == args.gpus or world == 1, f"WORLD_SIZE={world} but --gpus {args.gpus}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _abi.load()
    lib.mcrn_set_engine(1 if args.engine == "simt" else 0)

    torch.manual_seed(0)
    np.random.seed(0)
    model = MegaCRN(d.num_nodes, d.input_dim, d.output_dim, d.horizon, d.rnn_units, mem_num=d.mem_num,
                    mem_dim=d.mem_dim).to(dev).train()
    hx, hy, hl = O.synthetic_batch(d, B, t_in, seed=1234 + rank)
    hx, hy, hl = hx.pin_memory(), hy.pin_memory(), hl.pin_memory()
    dx, dy, dl = hx.to(dev), hy.to(dev), hl.to(dev)
    params = list(model.parameters())
    loss_host = torch.zeros(1).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    stream = torch.cuda.current_stream(dev)

    graphed = None if args.no_graph else GraphedTrainStep(model, B, t_in)
    if graphed is not None:
        graphed.load(dx, dy, dl)

    def step_device():
        # public API: the drop-in module + the fused trainer loss; coin flips drawn on the host every step
        if graphed is not None:
            loss = graphed(batches_seen=0)
        else:
            for p in params:
                p.grad = None
            loss = train_step(model, dx, dy, dl, batches_seen=0)
        allreduce_gradients(params)
        return loss

    def step_e2e():
        if graphed is not None:
            graphed.load(hx, hy, hl)               # pinned host -> static device buffers
        else:
            dx.copy_(hx, non_blocking=True); dy.copy_(hy, non_blocking=True); dl.copy_(hl, non_blocking=True)
        loss = step_device()
        loss_host.copy_(loss, non_blocking=True)
        stream.synchronize()
        return loss_host

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        """K steps, each bracketed by CUDA events on the launching stream, L2 flushed between steps."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        l0 = lib.mcrn_launch_count() + (graphed.kernels_replayed if graphed is not None else 0)
        for s, e in evs:
            flush.zero_()
            s.record(stream)
            fn()
            e.record(stream)
        barrier()
        launches = lib.mcrn_launch_count() + (graphed.kernels_replayed if graphed is not None else 0) - l0
        total_ms = sum(s.elapsed_time(e) for s, e in evs)
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), launches

    for _ in range(args.warmup):
        step_device()
    step_e2e()
    barrier()
    with ClockSampler(local) as clocks:
        dev_ms, launches = timed(step_device, args.steps)
        e2e_ms, _ = timed(step_e2e, args.steps)
    final_loss = float(step_e2e().item())
    # secondary number: the same step followed by clip_grad_norm_(5.0) + Adam(lr 0.01, eps 1e-3), fused
    # (mcrn_adam_step); single GPU: inside the captured graph, data parallel: after the gradient all-reduce
    from megacrn_b200.optim import FusedClipAdam
    opt_ms = None
    try:
        opt = FusedClipAdam(model, lr=0.01, eps=1e-3, max_grad_norm=5.0)
        g2 = None if (args.no_graph or world > 1) else GraphedTrainStep(model, B, t_in, optimizer=opt)
        if g2 is not None:
            g2.load(dx, dy, dl)

        def step_opt():
            if g2 is not None:
                return g2(batches_seen=0)
            loss = step_device()
            opt.step()
            return loss
        for _ in range(3):
            step_opt()
        opt_ms, _ = timed(step_opt, args.steps)
    except Exception as e:      # the secondary number must never cost the headline
        print(f"[bench] optimizer-step measurement skipped: {e}", file=sys.stderr)
    # host-side cost of enqueueing one step (no device wait inside): tells CPU-bound from GPU-bound
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(3):
        step_device()
    enqueue_ms = (time.perf_counter() - t0) / 3 * 1e3
    torch.cuda.synchronize(dev)

    if rank == 0:
        peaks = measured_peaks()
        gB = B * world
        value = gB * args.steps / (dev_ms / 1e3)
        e2e_value = gB * args.steps / (e2e_ms / 1e3)
        # ---- roofline of the dominant kernel family: the fused AGCN kernels, timed LIVE with CUDA events on the launching
        # stream around every launch of a few extra (eager, L2-flushed) training steps (mcrn_kernel_timing) ----
        import ctypes as C
        N, H, D = d.num_nodes, d.rnn_units, d.rnn_units + d.mem_dim
        cd = d.output_dim + d.ycov_dim

        def agcn_flops(Cc, O):          # algorithmic FLOPs of one AGCN call (identity blocks counted once)
            return 2 * 4 * N * N * B * Cc + 2 * B * N * 6 * Cc * O
        lib.mcrn_kernel_timing(1)
        for _ in range(3):
            flush.zero_()
            for p_ in params:
                p_.grad = None
            train_step(model, dx, dy, dl, batches_seen=0)
        torch.cuda.synchronize(dev)
        lib.mcrn_kernel_timing(0)

        def kclass(bwd, hs, variant):
            return bwd * 8 + (4 if hs == 128 else 0) + variant

        def read(cls):
            ms, n = C.c_float(0), C.c_int(0)
            lib.mcrn_kernel_timing_read(cls, C.byref(ms), C.byref(n))
            return (ms.value / n.value if n.value else None), n.value
        kernels = []
        for name, cls, fl in (
                ("fwd decoder gate   agcn_fused_h_kernel<HS=D,O=2D>", kclass(0, D, 0), agcn_flops(cd + D, 2 * D)),
                ("fwd decoder update agcn_fused_h_kernel<HS=D,O=D>", kclass(0, D, 1), agcn_flops(cd + D, D)),
                ("fwd encoder gate   agcn_fused_h_kernel<HS=H,O=2H>", kclass(0, H, 0), agcn_flops(d.input_dim + H, 2 * H)),
                ("fwd encoder update agcn_fused_h_kernel<HS=H,O=H>", kclass(0, H, 1), agcn_flops(d.input_dim + H, H)),
                ("bwd decoder gate-AGCN   agcn_bwd_kernel<D> (dX only)", kclass(1, D, 1), agcn_flops(cd + D, 2 * D)),
                ("bwd decoder update-AGCN agcn_bwd_kernel<D> (dX only)", kclass(1, D, 0), agcn_flops(cd + D, D)),
                ("bwd encoder gate-AGCN   agcn_bwd_kernel<H> (dX only)", kclass(1, H, 1), agcn_flops(d.input_dim + H, 2 * H)),
                ("bwd encoder update-AGCN agcn_bwd_kernel<H> (dX only)", kclass(1, H, 0), agcn_flops(d.input_dim + H, H))):
            ms, n = read(cls)
            if ms:
                kernels.append({"kernel": name, "launches_timed": n, "us": 1e3 * ms, "gflop": fl / 1e9,
                                "tflops": fl / (ms * 1e-3) / 1e12})
        if kernels:
            top = kernels[0]
            k_ms, k_tflops, k_name, k_n = top["us"] / 1e3, top["tflops"], top["kernel"], top["launches_timed"]
        else:       # shapes the fused kernels do not take (hidden width not 64/128): the propagation GEMM, timed alone
            KS = 4
            M_, N_, K_ = KS * N, B * D, N
            ld = lib.mcrn_support_ld(N)
            a = torch.randn(M_, ld, device=dev); b = torch.randn(K_, N_, device=dev); c = torch.empty(M_, N_, device=dev)
            call = lambda: lib.mcrn_gemm(M_, N_, K_, a.data_ptr(), ld, 0, b.data_ptr(), N_, 0, c.data_ptr(), N_, 0, stream.cuda_stream)
            for _ in range(3):
                call()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            s.record(stream)
            for _ in range(20):
                call()
            e.record(stream)
            torch.cuda.synchronize(dev)
            k_ms = s.elapsed_time(e) / 20
            k_tflops = 2.0 * M_ * N_ * K_ / (k_ms * 1e-3) / 1e12
            k_name, k_n = f"propagation GEMM [{M_}x{K_}]x[{K_}x{N_}]", 20
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
        if os.path.exists(tpath) and args.config == "c2":
            traffic = json.load(open(tpath)).get("fwd_decoder_gate_dram_bytes_per_launch")
        step_flops = 3 * fwd_flops(d, B, t_in)
        line = {
            "metric": "sequences/sec (12-step enc+dec fwd+bwd)", "value": value, "unit": "sequences/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.engine == "simt" else "f16/tf32", "data": "synthetic",
            "config": {"workload": f"{args.config}: METR-LA-shaped N={d.num_nodes} T_in={t_in} T_out={d.horizon} "
                                   f"H={d.rnn_units} batch={B}/GPU, train step = forward + trainer loss + backward"
                                   + (" + 1 NCCL grad all-reduce" if world > 1 else ""),
                       "global_batch": gB, "parallelism": f"dp{world}", "engine": args.engine,
                       "launch": "eager" if args.no_graph else "cuda-graph replay (one graph per teacher-forcing pattern)",
                       "l2": "256 MiB memset between timed steps (flush) + 1.2 GB/step activation working set",
                       "final_loss": final_loss},
            "e2e": {"value": e2e_value, "unit": "sequences/s", "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": int(4 * (hx.numel() + hy.numel() + hl.numel())), "d2h_bytes_per_step": 4},
            "with_optimizer": None if opt_ms is None else {
                "value": gB * args.steps / (opt_ms / 1e3), "unit": "sequences/s", "ms_per_step": opt_ms / args.steps,
                "what": "step + fused clip_grad_norm_(5.0) + Adam(lr 0.01, eps 1e-3) (mcrn_adam_step)"},
            "gpu_launches": int(launches), "host_enqueue_ms_per_step": enqueue_ms,
            "clocks": clocks.summary(),
            "roofline": {"bound": "tensor", "kernel": k_name,
                         "achieved": k_tflops, "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
                         "frac": k_tflops / peaks["bf16_burst"], "traffic": traffic, "kernel_ms": k_ms,
                         "launches_timed": k_n,
                         "how": "algorithmic FLOPs of the AGCN call (SURVEY 8d: 2*4*N^2*B*C + 2*B*N*6C*O) / CUDA-event time around "
                                "each launch during 3 eager L2-flushed training steps inside this run",
                         "peak_source": peaks["source"] + "; bf16 burst (the fused forward computes in fp16 at the bf16 rate, fp32 accumulate)",
                         "kernels": kernels,
                         "step_tflops": step_flops / (dev_ms / args.steps * 1e-3) / 1e12,
                         "step_frac_of_sustained": step_flops / (dev_ms / args.steps * 1e-3) / 1e12 / peaks["bf16_sustained"]},
        }
        if not args.no_cpu_baseline:
            threads = cpu_threads()
            probe = cpu_reference_step(d, 8, t_in, threads)
            probe()
            t0 = time.perf_counter(); probe(); per_seq = (time.perf_counter() - t0) / 8
            Bs = B
            while Bs > 8 and per_seq * Bs > 5.0:
                Bs //= 2
            cstep = cpu_reference_step(d, Bs, t_in, threads)
            n_cpu = 3
            t0 = time.perf_counter()
            for _ in range(n_cpu):
                cstep()
            cdt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": Bs * n_cpu / cdt, "unit": "sequences/s", "cores": threads, "kind": "port",
                                    "sample": f"{n_cpu} steps of batch {Bs} of the same workload (oracle port of the "
                                              f"reference, torch CPU fp32, {threads} of {os.cpu_count()} host threads)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
