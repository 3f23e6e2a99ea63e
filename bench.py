#!/usr/bin/env python
"""bench.py -- sequences/sec of the MegaCRN training step (12-step encoder + 12-step decoder,
forward + trainer loss + backward [+ one gradient all-reduce when N > 1]) on N B200s.

    python bench.py --gpus 1 --steps 20 --warmup 3 [--config c2|c3|c4|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference        # the reference's CPU path (oracle port) on host cores

Default workload = BASELINE.json configs[1] (c2): METR-LA shape N=207, T_in=T_out=12, H=64, batch 64 per GPU
(weak scaling), synthetic inputs, random-init weights.  Prints ONE JSON line.

The GPU arm uses only the product package (megacrn_b200); oracle/ is imported by the `cpu_baseline` leg (CPU port timed
on the host cores, the same port on cuda:0 as the eager-PyTorch context row, and the loss check) and by `--impl reference`.
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

from megacrn_b200.workloads import CONFIGS, agcn_flops, config, forward_flops, synthetic_batch  # noqa: E402

METRIC = "sequences/sec (12-step enc+dec fwd+bwd)"
SCALER = dict(scaler_mean=54.0, scaler_std=20.0)          # synthetic StandardScaler of SURVEY.md 8(d)


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


# ---- the reference's path on CPU (oracle port) -----------------------------------------------------------------------
def cpu_threads():
    """Threads for the many-thread CPU leg.  torch's intra-op pool stops scaling on these op sizes well before a
    128-core host is full (round 1, same box class: 16 threads 67-75 seq/s, 128 threads 0.6 seq/s), so the leg uses
    min(cores, MCRN_CPU_THREADS=16) and reports the number; the reference trainer's own setting, 1 thread
    (model/traintest_MegaCRN.py:255-261), is timed beside it."""
    return max(1, min(os.cpu_count() or 1, int(os.environ.get("MCRN_CPU_THREADS", "16"))))


def cpu_reference_step(d, B, t_in, threads, device="cpu"):
    """One training step of the reference's path: the oracle port (torch ops + autograd), same workload."""
    from oracle import megacrn_oracle as O
    od = O.Dims(**{k: getattr(d, k) for k in ("num_nodes", "input_dim", "output_dim", "horizon", "rnn_units", "num_layers",
                                               "cheb_k", "ycov_dim", "mem_num", "mem_dim")})
    if device == "cpu":
        torch.set_num_threads(threads)
    p = {k: v.to(device) for k, v in O.init_params(od, seed=0).items()}
    x, y_cov, labels = (t.to(device) for t in synthetic_batch(d, B, t_in, seed=1234))
    flags = [True] * d.horizon

    def step():
        loss, _, _ = O.loss_and_grads(od, p, x, y_cov, labels, flags)
        return float(loss)
    return step


def time_cpu(d, t_in, threads, budget_s, B_full, n_steps):
    """Bounded sample: the per-step batch is halved until one step fits `budget_s`; returns (seq/s, batch, steps)."""
    probe = cpu_reference_step(d, min(8, B_full), t_in, threads)
    probe()
    t0 = time.perf_counter(); probe(); per_seq = (time.perf_counter() - t0) / min(8, B_full)
    Bs = B_full
    while Bs > 1 and per_seq * Bs > budget_s:
        Bs //= 2
    step = cpu_reference_step(d, Bs, t_in, threads)
    step()
    t0 = time.perf_counter()
    for _ in range(n_steps):
        step()
    dt = time.perf_counter() - t0
    return Bs * n_steps / dt, Bs, dt


def run_reference_arm(args, d, B, t_in, what):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = cpu_threads()
    n = max(1, args.steps)
    budget = 90.0 / max(1, n + max(1, args.warmup))
    val, Bs, dt = time_cpu(d, t_in, threads, budget, B, n)
    val1, Bs1, _ = time_cpu(d, t_in, 1, 8.0, min(B, 8), 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "sequences/s",
        "n_gpus": args.gpus, "steps": n, "warmup": args.warmup, "ms_per_step": 1e3 * dt / n,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: {what}, train step = forward + trainer loss + backward",
                   "where": f"host CPU, {Bs} sequences per timed step"},
        "cpu_baseline": {"value": val, "unit": "sequences/s", "cores": threads, "kind": "port",
                         "sample": f"{n} steps of batch {Bs} (oracle port of the reference, torch CPU fp32, {threads} threads)",
                         "value_1thread": val1,
                         "sample_1thread": f"1 step of batch {Bs1}, torch.set_num_threads(1) = the reference trainer's own setting "
                                           "(model/traintest_MegaCRN.py:255-261)"},
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
    ap.add_argument("--config", default="c2", choices=[c for c in CONFIGS if c != "c1"])
    ap.add_argument("--engine", default="default", choices=["default", "simt"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary numbers (optimizer, free-running, eval forward)")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every kernel from the host instead of replaying CUDA graphs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)

    d, B, t_in, what = config(args.config)
    if args.impl == "reference":
        run_reference_arm(args, d, B, t_in, what)
        return

    import torch.distributed as dist
    from megacrn_b200 import _abi
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
    model = d.build(dev).train()
    hx, hy, hl = synthetic_batch(d, B, t_in, seed=1234 + rank)
    hx, hy, hl = hx.pin_memory(), hy.pin_memory(), hl.pin_memory()
    dx, dy, dl = hx.to(dev), hy.to(dev), hl.to(dev)
    params = list(model.parameters())
    loss_host = torch.zeros(1).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    stream = torch.cuda.current_stream(dev)

    graphed = None if args.no_graph else GraphedTrainStep(model, B, t_in, **SCALER)
    if graphed is not None:
        graphed.load(dx, dy, dl)
    collectives_in_graph = graphed is not None and graphed.allreduce

    def step_device():
        # public API: the drop-in module + the fused trainer loss; coin flips drawn on the host every step
        if graphed is not None:
            loss = graphed(batches_seen=0)
        else:
            for p in params:
                p.grad = None
            loss = train_step(model, dx, dy, dl, batches_seen=0, **SCALER)
        if not collectives_in_graph:
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

    def timed(fn, steps, counters=()):
        """K steps, each bracketed by CUDA events on the launching stream, L2 flushed between steps."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        l0 = lib.mcrn_launch_count() + sum(c.kernels_replayed for c in counters)
        for s, e in evs:
            flush.zero_()
            s.record(stream)
            fn()
            e.record(stream)
        barrier()
        launches = lib.mcrn_launch_count() + sum(c.kernels_replayed for c in counters) - l0
        total_ms = sum(s.elapsed_time(e) for s, e in evs)
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), launches

    for _ in range(args.warmup):
        step_device()
    step_e2e()
    barrier()
    gcount = [graphed] if graphed is not None else []
    with ClockSampler(local) as clocks:
        dev_ms, launches = timed(step_device, args.steps, gcount)
        e2e_ms, _ = timed(step_e2e, args.steps, gcount)
    final_loss = float(step_e2e().item())

    # ---- secondary numbers (SURVEY 8d / VERDICT r1): optimizer inside the step, free-running decoder, eval forward ----
    extras = {}
    if not args.no_extras and not args.no_graph:
        from megacrn_b200.optim import FusedClipAdam
        try:
            opt = FusedClipAdam(model, lr=0.01, eps=1e-3, max_grad_norm=5.0)
            g2 = GraphedTrainStep(model, B, t_in, optimizer=opt, **SCALER)
            g2.load(dx, dy, dl)
            for _ in range(3):
                g2(batches_seen=0)
            ms, _ = timed(lambda: g2(batches_seen=0), args.steps)
            extras["with_optimizer"] = {"value": B * world * args.steps / (ms / 1e3), "unit": "sequences/s", "ms_per_step": ms / args.steps,
                                        "what": "step + gradient all-reduce (N > 1) + fused clip_grad_norm_(5.0) + Adam(lr 0.01, eps 1e-3); one graph at N = 1"}
            del g2
            # free-running decoder (no teacher forcing: what training looks like after ~10 k batches, tau(20000) = 0.083,
            # model/MegaCRN.py:146-147, :188-191): projection + input build + cell serialised per step
            free = [False] * d.horizon
            for _ in range(3):
                graphed(teacher_forcing=free)
                if not collectives_in_graph:
                    allreduce_gradients(params)

            def step_free():
                graphed(teacher_forcing=free)
                if not collectives_in_graph:
                    allreduce_gradients(params)
            ms, _ = timed(step_free, args.steps)
            extras["free_running"] = {"value": B * world * args.steps / (ms / 1e3), "unit": "sequences/s", "ms_per_step": ms / args.steps,
                                      "what": "same step with teacher_forcing = [False] * horizon (decoder feeds on its own predictions)"}
            # eval forward (model/traintest_MegaCRN.py:50-99): no labels, no coin flips, prologue reused across batches
            model.eval()
            with torch.no_grad():
                for _ in range(2):
                    model(dx, dy)
                torch.cuda.synchronize(dev)
                ge = torch.cuda.CUDAGraph()
                with torch.cuda.graph(ge):
                    model(dx, dy)
            for _ in range(3):
                ge.replay()
            ms, _ = timed(ge.replay, args.steps)
            extras["eval_forward"] = {"value": B * world * args.steps / (ms / 1e3), "unit": "sequences/s", "ms_per_step": ms / args.steps,
                                      "what": "model.eval() forward, torch.no_grad(), labels=None (evaluate(), traintest:50-99), graph replay"}
            model.train()
            del ge
        except Exception as e:      # the secondary numbers must never cost the headline
            print(f"[bench] secondary measurement skipped: {type(e).__name__}: {e}", file=sys.stderr)
            model.train()
    # host-side cost of enqueueing one step (no device wait inside): tells CPU-bound from GPU-bound
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(3):
        step_device()
    enqueue_ms = (time.perf_counter() - t0) / 3 * 1e3
    torch.cuda.synchronize(dev)

    # ---- roofline of the dominant kernel family: every fused AGCN kernel class timed LIVE with CUDA events on the launching
    # stream around each launch of a few extra (eager, L2-flushed) training steps (mcrn_kernel_timing); all ranks run the
    # steps (collectives), rank 0 reads the record ----
    import ctypes as C
    lib.mcrn_kernel_timing(1)
    for _ in range(3):
        flush.zero_()
        for p_ in params:
            p_.grad = None
        train_step(model, dx, dy, dl, batches_seen=0, **SCALER)
    torch.cuda.synchronize(dev)
    lib.mcrn_kernel_timing(0)

    if rank == 0:
        peaks = measured_peaks()
        gB = B * world
        value = gB * args.steps / (dev_ms / 1e3)
        e2e_value = gB * args.steps / (e2e_ms / 1e3)
        H, D = d.rnn_units, d.rnn_units + d.mem_dim
        cd = d.output_dim + d.ycov_dim

        def kclass(bwd, hs, variant):
            return bwd * 8 + (4 if hs == 128 else 0) + variant

        def read(cls):
            ms, n = C.c_float(0), C.c_int(0)
            lib.mcrn_kernel_timing_read(cls, C.byref(ms), C.byref(n))
            return (ms.value / n.value if n.value else None), n.value
        kernels = []
        for name, key, cls, fl in (
                ("fwd decoder gate   agcn_fused_h_kernel<HS=D,O=2D>", "fwd_dec_gate", kclass(0, D, 0), agcn_flops(d, B, cd + D, 2 * D)),
                ("fwd decoder update agcn_fused_h_kernel<HS=D,O=D>", "fwd_dec_update", kclass(0, D, 1), agcn_flops(d, B, cd + D, D)),
                ("fwd encoder gate   agcn_fused_h_kernel<HS=H,O=2H>", "fwd_enc_gate", kclass(0, H, 0), agcn_flops(d, B, d.input_dim + H, 2 * H)),
                ("fwd encoder update agcn_fused_h_kernel<HS=H,O=H>", "fwd_enc_update", kclass(0, H, 1), agcn_flops(d, B, d.input_dim + H, H)),
                ("bwd decoder gate-AGCN   agcn_bwd_h_kernel<D> (dX only)", "bwd_dec_gate", kclass(1, D, 1), agcn_flops(d, B, cd + D, 2 * D)),
                ("bwd decoder update-AGCN agcn_bwd_h_kernel<D> (dX only)", "bwd_dec_update", kclass(1, D, 0), agcn_flops(d, B, cd + D, D)),
                ("bwd encoder gate-AGCN   agcn_bwd_h_kernel<H> (dX only)", "bwd_enc_gate", kclass(1, H, 1), agcn_flops(d, B, d.input_dim + H, 2 * H)),
                ("bwd encoder update-AGCN agcn_bwd_h_kernel<H> (dX only)", "bwd_enc_update", kclass(1, H, 0), agcn_flops(d, B, d.input_dim + H, H))):
            ms, n = read(cls)
            if ms:
                kernels.append({"kernel": name, "key": key, "launches_timed": n, "us": 1e3 * ms, "gflop": fl / 1e9,
                                "tflops": fl / (ms * 1e-3) / 1e12, "frac": fl / (ms * 1e-3) / 1e12 / peaks["bf16_burst"]})
        # (classes coincide when H == D; keep the first of each)
        seen, uniq = set(), []
        for k in kernels:
            if (k["us"], k["launches_timed"]) not in seen:
                seen.add((k["us"], k["launches_timed"])); uniq.append(k)
        kernels = uniq
        if kernels:
            top = max(kernels, key=lambda k: k["us"])        # the kernel that dominates the step by time
            k_ms, k_tflops, k_name, k_n, k_key = top["us"] / 1e3, top["tflops"], top["kernel"], top["launches_timed"], top["key"]
        else:       # shapes the fused kernels do not take (hidden width not 64/128): the propagation GEMM, timed alone
            KS, N = 4, d.num_nodes
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
            k_name, k_n, k_key = f"propagation GEMM [{M_}x{K_}]x[{K_}x{N_}]", 20, "propagate"
        # DRAM bytes per launch of that kernel from this round's `ncu --set full` capture (profiles/r2_traffic.json: read +
        # write, with the capture command); null where no capture of this config / kernel exists
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath)).get(args.config, {})
            if k_key in tj:
                traffic = tj[k_key]["dram_bytes_read"] + tj[k_key]["dram_bytes_write"]
                traffic_src = tj[k_key].get("source")
        step_flops = 3 * forward_flops(d, B, t_in)
        line = {
            "metric": METRIC, "value": value, "unit": "sequences/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.engine == "simt" else "f16/tf32", "data": "synthetic",
            "config": {"workload": f"{args.config}: {what}, train step = forward + trainer loss + backward",
                       "collective": "1 NCCL gradient all-reduce (+ a 4-byte normaliser all-reduce) per step" if world > 1 else "none",
                       "global_batch": gB, "parallelism": f"dp{world}", "engine": args.engine,
                       "launch": "eager" if args.no_graph else "cuda-graph replay (one graph per teacher-forcing pattern"
                                 + ("; normaliser and gradient all-reduce issued around the replay)" if collectives_in_graph else ")"),
                       "l2": "256 MiB memset between timed steps (flush) + >1 GB/step activation working set",
                       "final_loss": final_loss},
            "e2e": {"value": e2e_value, "unit": "sequences/s", "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": int(4 * (hx.numel() + hy.numel() + hl.numel())), "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches), "host_enqueue_ms_per_step": enqueue_ms,
            "clocks": clocks.summary(),
            "roofline": {"bound": "tensor", "kernel": k_name,
                         "achieved": k_tflops, "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
                         "frac": k_tflops / peaks["bf16_burst"], "traffic": traffic, "traffic_source": traffic_src,
                         "kernel_ms": k_ms, "launches_timed": k_n,
                         "how": "headline = the fused AGCN kernel class with the longest launch; algorithmic FLOPs of the AGCN call "
                                "(SURVEY 8d: 2*4*N^2*B*C + 2*B*N*6C*O) / CUDA-event time around each launch during 3 eager "
                                "L2-flushed training steps inside this run",
                         "peak_source": peaks["source"] + "; bf16 burst (fp16 operands at the bf16 rate, fp32 accumulate)",
                         "kernels": kernels,
                         "step_tflops": step_flops / (dev_ms / args.steps * 1e-3) / 1e12,
                         "step_frac_of_sustained": step_flops / (dev_ms / args.steps * 1e-3) / 1e12 / peaks["bf16_sustained"]},
        }
        line.update(extras)
        if not args.no_cpu_baseline:
            from oracle import megacrn_oracle as O
            threads = cpu_threads()
            cval, cB, _ = time_cpu(d, t_in, threads, 5.0, B, 3)
            cval1, cB1, _ = time_cpu(d, t_in, 1, 8.0, min(B, 8), 1)
            line["cpu_baseline"] = {"value": cval, "unit": "sequences/s", "cores": threads, "kind": "port",
                                    "sample": f"3 steps of batch {cB} of the same workload (oracle port of the reference, torch CPU "
                                              f"fp32, {threads} of {os.cpu_count()} host threads)",
                                    "value_1thread": cval1,
                                    "sample_1thread": f"1 step of batch {cB1} with torch.set_num_threads(1), the reference trainer's "
                                                      "own setting (model/traintest_MegaCRN.py:255-261)"}
            # loss check: the step bench.py times vs the oracle on identical inputs, weights and flags
            Bc = min(B, 8)
            od = O.Dims(**{k: getattr(d, k) for k in ("num_nodes", "input_dim", "output_dim", "horizon", "rnn_units", "num_layers",
                                                       "cheb_k", "ycov_dim", "mem_num", "mem_dim")})
            sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
            flags = [True] * d.horizon
            torch.set_num_threads(threads)
            ref_loss, _, _ = O.loss_and_grads(od, sd, hx[:Bc], hy[:Bc], hl[:Bc], flags, **SCALER)
            for p_ in params:
                p_.grad = None
            got = float(train_step(model, dx[:Bc].contiguous(), dy[:Bc].contiguous(), dl[:Bc].contiguous(), teacher_forcing=flags,
                                   group=None, **SCALER).item()) if world == 1 else None
            if got is not None:
                rel = abs(got - float(ref_loss)) / abs(float(ref_loss))
                line["loss_check"] = {"gpu": got, "oracle": float(ref_loss), "rel_err": rel, "batch": Bc, "tol": 1e-3}
                assert rel < 1e-3, f"bench step loss {got} vs oracle {float(ref_loss)}"
            # context row: the same PyTorch port on cuda:0 (eager: cuBLAS / ATen kernels), fp32 and TF32 allowed
            try:
                eager = {}
                for name, tf32 in (("fp32", False), ("tf32", True)):
                    torch.backends.cuda.matmul.allow_tf32 = tf32
                    estep = cpu_reference_step(d, B, t_in, threads, device=dev)
                    estep(); estep()
                    torch.cuda.synchronize(dev)
                    t0 = time.perf_counter()
                    for _ in range(3):
                        estep()
                    torch.cuda.synchronize(dev)
                    eager[name] = B * 3 / (time.perf_counter() - t0)
                torch.backends.cuda.matmul.allow_tf32 = False
                line["gpu_eager_baseline"] = {"value_fp32": eager["fp32"], "value_tf32": eager["tf32"], "unit": "sequences/s",
                                              "what": f"the oracle port of the reference (plain PyTorch ops + autograd) on cuda:0, batch {B}, "
                                                      "eager, wall clock incl. host launch overhead: how the reference itself would run on this GPU"}
            except Exception as e:
                print(f"[bench] gpu eager baseline skipped: {type(e).__name__}: {e}", file=sys.stderr)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
