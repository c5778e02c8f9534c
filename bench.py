#!/usr/bin/env python
"""bench.py -- headline benchmark of the mini-stark prover hot path on B200.

Metric (BASELINE.json): "prove ms & LDE Melem/s, Goldilocks 2^22 rows x 32 cols, 1/2/4/8 B200".
  value   = LDE Melem/s: L*C output elements / device time of the batched coset LDE (coefficients
            resident in HBM -> evaluations resident in HBM, coset scaling included), blowup 4.
  e2e     = the same metric through the C-ABI host-buffer call ms_coset_lde_host (pinned host
            coefficients in, row-major host evaluations out; H2D + D2H inside the timed region).
  prove   = extra key: full Stark::prove of the synthetic AIR on the same shape (W=16, T=16), device
            resident trace -> proof bytes on the host, and host trace -> proof bytes.
A "step" is one LDE of the whole 2^22 x 32 batch.  With --gpus N every rank owns an independent
2^22 x 32 column shard (columns are independent: SURVEY.md 8e), no data-path collective: weak scaling.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GL = 0
P_GL = 2**64 - 2**32 + 1
SHIFT = 0x123456789ABCDEF % P_GL  # fixed coset offset for the stage benchmark (injected challenge)
# measured under ncu --set full for the headline shape (profiles/r01_e_ncu_ntt.txt):
# pass 1 1.211 + 4.257 GB, pass 2 4.297 + 4.268 GB (dram__bytes_read.sum + dram__bytes_write.sum)
NCU_TRAFFIC_BYTES = 14_033_000_000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-rows", type=int, default=22)
    ap.add_argument("--cols", type=int, default=32)
    ap.add_argument("--blowup", type=int, default=4)
    ap.add_argument("--no-prove", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--security-bits", type=int, default=100)
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,"
         "utilization.gpu")

    def __init__(self, device: int):
        self.device = device
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons, busy = [], [], set(), []
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            try:
                busy.append(float(f[9]) > 0)
            except (ValueError, IndexError):
                busy.append(True)
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        under_load = [v for v, b in zip(sm, busy) if b] or sm  # samples taken while the GPU was busy
        return {"sm_mhz": float(np.median(under_load)) if under_load else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "samples_under_load": len(under_load), "reasons": sorted(reasons)}


def workload_name(args):
    return (f"goldilocks coset-LDE 2^{args.log_rows} rows x {args.cols} cols blowup {args.blowup} per GPU "
            f"(BASELINE headline shape; coefficients -> evaluations on shift*<w_L>)")


def synth_coeffs(n, cols, seed):
    """poly-major [cols, n] canonical Goldilocks coefficients (NTT sweep treats all C
    columns as coefficient vectors)."""
    from tests.synth import synth_trace

    return np.ascontiguousarray(synth_trace(GL, n, cols, seed=seed).T)


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's own CPU algorithm for the path.  The reference is Rust and cannot be built in
    this image (no cargo/rustc), so this times the oracle port (oracle/liboracle.so: per-column
    radix-2 coset transforms + stride-C scatter, starks.rs:87-91) with every host thread on a bounded
    sample of the workload's columns."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O

    O.build()
    n, B = 1 << args.log_rows, args.blowup
    threads = os.cpu_count() or 1
    cols = min(args.cols, max(1, threads))
    coeffs = synth_coeffs(n, cols, 1)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        O.coset_lde(GL, coeffs, n * B, SHIFT, threads=threads)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    t = float(np.mean(times))
    val = n * B * cols / t / 1e6
    line = {
        "impl": "reference", "metric": "lde_melem_per_s", "value": val, "unit": "Melem/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(args), "sample": f"{cols} of {args.cols} columns per step"},
        "cpu_baseline": {"value": val, "unit": "Melem/s", "cores": min(threads, cols), "kind": "port",
                         "sample": f"{cols} columns x 2^{args.log_rows} -> 2^{args.log_rows + int(np.log2(B))} per step, oracle port "
                                   "(reference is Rust; no toolchain in the image)"},
        "e2e": {"value": val, "unit": "Melem/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    from ministark_b200 import Context
    from ministark_b200._lib import StarkParams

    dev = local_rank
    torch.cuda.set_device(dev)
    ctx = Context(GL, dev)
    n, C, B = 1 << args.log_rows, args.cols, args.blowup
    L = n * B
    peak, peak_src = peaks()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident LDE ------------------------------------------------------------------
    coeffs = ctx.to_device(synth_coeffs(n, C, 1 + rank))
    out = ctx.empty(C, L)
    for _ in range(args.warmup):
        ctx.coset_lde(coeffs, B, SHIFT, out=out)
    barrier()
    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()
    ctx.set_profiling(True)
    ctx.profile_collect()
    launches0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        ctx.coset_lde(coeffs, B, SHIFT, out=out)  # inputs (1 GiB) + outputs (4 GiB) exceed L2: no flush needed
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - launches0
    kern = ctx.profile_collect()
    ctx.set_profiling(False)
    # the sampler keeps running through the e2e section below (GPU busy throughout), so that the median is taken
    # over a dozen samples under load instead of the one or two that fit the 50 ms LDE region
    t = torch.tensor([ms_total], dtype=torch.float64, device=f"cuda:{dev}")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = world * L * C / (ms_step * 1e-3) / 1e6

    # ---- e2e: host buffers through ms_coset_lde_host -------------------------------------------
    e2e = None
    if not args.no_e2e:
        h_in = torch.from_numpy(synth_coeffs(n, C, 1 + rank).view(np.int64)).pin_memory()
        h_out = torch.empty((L, C), dtype=torch.int64).pin_memory()
        in_np, out_np = h_in.numpy().view(np.uint64), h_out.numpy().view(np.uint64)
        for _ in range(2):
            ctx.coset_lde_host(in_np, B, SHIFT, out=out_np)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ctx.coset_lde_host(in_np, B, SHIFT, out=out_np)  # synchronous: returns when the result is on the host
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{dev}")
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item()) / args.steps
        # spot check: the host result equals the device result
        assert (out_np[:4] == ctx.to_host(out[:, :4]).T).all() or rank != 0
        e2e = {"value": world * L * C / e2e_s / 1e6, "unit": "Melem/s", "ms_per_step": e2e_s * 1e3,
               "h2d_bytes_per_step": int(n * C * 8), "d2h_bytes_per_step": int(L * C * 8),
               "call": "ms_coset_lde_host (pinned host buffers, row-major evaluations out)"}
        del h_in, h_out

    clocks = sampler.stop() if rank == 0 else None  # window = the LDE and e2e regions (GPU busy throughout)
    # ---- full prove on the same shape -------------------------------------------------------------
    # N = 1: the plain prover.  N > 1: one replica per GPU with the trace tree and the LDE + its tree
    # sharded (ministark_b200/sharded.py): strong scaling of one proof, max over ranks.
    prove = None
    if not args.no_prove:
        from ministark_b200.sharded import SharedProofBuffer, stark_prove_sharded
        from tests.synth import synth_linear_matrix, synth_trace

        W = C // 2
        steps = n - 1
        trace_rm = synth_trace(GL, n, W, seed=0x5EED000000000000 + 1)
        m = synth_linear_matrix(GL, n, W)
        params = StarkParams(args.security_bits, B, steps, C, 2)
        bound = int(ctx.lib.ms_stark_proof_bound(GL, params, n, C))
        shared = None
        if world > 1:  # one shared host buffer: every rank downloads 1/world of the quotient polynomials
            shared = SharedProofBuffer(ctx, bound, dist)
            proof_buf = shared.array
        else:
            proof_buf = torch.empty(bound, dtype=torch.uint8).pin_memory().numpy()
        trace_cm = ctx.to_device(np.ascontiguousarray(trace_rm.T))
        torch.cuda.synchronize()
        ms_dev, ms_host, plen, stages = [], [], 0, None
        for i in range(3):
            barrier()
            t0 = time.perf_counter()
            if world > 1:
                plen = stark_prove_sharded(ctx, params, trace_cm, m, shared, dist)
            else:
                plen = ctx.stark_prove_device(params, trace_cm, m, proof_buf)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{dev}")
            if dist is not None:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            if i > 0:
                ms_dev.append(float(tt.item()) * 1e3)
                stages = ctx.last_timings()
        if world == 1:
            h_trace = torch.from_numpy(trace_rm.view(np.int64)).pin_memory().numpy().view(np.uint64)
            import ctypes as Cc

            for i in range(4):
                cap = Cc.c_uint64(proof_buf.size)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                rc = ctx.lib.ms_stark_prove(ctx.h, Cc.byref(params), h_trace.ctypes.data, n, W, m.ctypes.data, W,
                                            proof_buf.ctypes.data, Cc.byref(cap))
                dt = time.perf_counter() - t0
                ctx._check(rc)
                if i > 0:
                    ms_host.append(dt * 1e3)
        import hashlib

        prove = {"prove_ms": float(np.mean(ms_dev)), "prove_e2e_ms": float(np.median(ms_host)) if ms_host else None,
                 "prove_e2e_samples_ms": [round(v, 2) for v in ms_host],
                 "proof_bytes": plen, "proof_sha256": hashlib.sha256(proof_buf[:plen].tobytes()).hexdigest(),
                 "scaling": "strong (one proof; commitments and the proof download sharded over the ranks, FRI replicated)" if world > 1 else "single GPU",
                 "config": f"SynthLinear AIR W={W} T={W} (C={C}), N=2^{args.log_rows}, blowup {B}, security {args.security_bits} bits, binary trees",
                 "stages_ms": {k: round(v, 3) for k, v in (stages or [])}}
        if world > 1:
            prove["sharded"] = getattr(ctx, "last_sharded_stats", None)
        del trace_cm
        if shared is not None:
            del proof_buf
            shared.close()

    # ---- CPU baseline (rank 0): the oracle port, single thread, bounded sample ---------------------
    cpu = None
    if not args.no_cpu and rank == 0:
        from oracle import oracle as O

        O.build()
        sample_cols = 2
        cc = np.ascontiguousarray(synth_coeffs(n, C, 1)[:sample_cols])
        t0 = time.perf_counter()
        ref = O.coset_lde(GL, cc, L, SHIFT, threads=1)
        dt = time.perf_counter() - t0
        got = ctx.to_host(out[:sample_cols]) if rank == 0 else None
        assert (got.T == ref).all(), "device LDE differs from the oracle on the sampled columns"
        cpu = {"value": L * sample_cols / dt / 1e6, "unit": "Melem/s", "cores": 1, "kind": "port",
               "sample": f"{sample_cols} of {C} columns, 2^{args.log_rows} -> 2^{args.log_rows + int(np.log2(B))} (oracle/liboracle.so, "
                         f"single thread like the reference; {os.cpu_count()} host cores present); output compared bit-for-bit with the GPU's"}

    if rank == 0:
        bytes_alg = (n + L) * C * 8  # a4 algorithmic bytes = (N + L) * C * s per LDE
        lde_kernel_ms = sum(v[0] for k, v in kern.items()) / args.steps
        per_kernel = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps} for k, v in kern.items()}
        achieved = bytes_alg / (ms_step * 1e-3) / 1e9
        line = {
            "metric": "lde_melem_per_s", "value": value, "unit": "Melem/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(args),
                       "l2": "inputs (N*C*8) + outputs (L*C*8) exceed the 126 MB L2; no flush between iterations",
                       "parallelism": f"{world} independent column shards"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC_BYTES if (args.log_rows, C, B) == (22, 32, 4) else None,
                         "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, pass 1 + pass 2 "
                                           "(profiles/r01_e_ncu_ntt.txt)",
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_alg,
                         "kernel": "coset-LDE = k_ntt_fixed pass 1 + pass 2 (+ twiddle builders), one ms_coset_lde call; "
                                   "'launch' = that call, (N + L) * C * 8 algorithmic bytes",
                         "kernels_ms_per_step": per_kernel, "kernel_sum_ms_per_step": lde_kernel_ms},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        if prove:
            line["prove"] = prove
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
