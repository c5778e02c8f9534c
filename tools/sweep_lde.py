"""BASELINE config 4: batched coset-LDE sweep 2^16..2^24 rows x 64 cols, blowup 4, Goldilocks and BabyBear,
device-resident, CUDA events; prints one JSON line per point with the HBM-roofline fraction.

Single process: one GPU.  Under torchrun with W ranks: the 64 columns are sharded over the first g ranks for every
g in (1, 2, 4, 8) with g <= W (columns are independent: no data-path collective), time = max over the participating
ranks, fraction against g x the measured per-GPU HBM peak.

    python tools/sweep_lde.py [lo hi]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep_lde.py [lo hi]"""
import json, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ministark_b200 import Context
from ministark_b200.sharded import column_ranges
from ministark_b200.synth import synth_trace
peak = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs'] if os.path.exists('MEASURED_PEAKS.json') else 6650.0
cols, B = 64, 4
args = [a for a in sys.argv[1:]]
lo, hi = (int(args[0]), int(args[1])) if len(args) > 1 else (16, 24)
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
dist = None
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))


def max_over_ranks(v):
    t = torch.tensor([v], dtype=torch.float64, device=f"cuda:{local}")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


for field, name, s in ((0, 'goldilocks', 8), (1, 'babybear', 4)):
    ctx = Context(field, local)
    for logn in range(lo, hi + 1):
        n = 1 << logn; L = n * B
        for g in [x for x in (1, 2, 4, 8) if x <= world]:
            a, b = column_ranges(cols, g)[rank] if rank < g else (0, 0)
            ms = 0.0
            if b > a:
                base = synth_trace(field, min(n, 1 << 20), b - a, seed=logn)      # tile a 2^20-row block: content is irrelevant for timing
                coeffs = ctx.to_device(np.ascontiguousarray(np.tile(base, (n // base.shape[0], 1)).T))
                out = ctx.empty(b - a, L)
                for _ in range(2): ctx.coset_lde(coeffs, B, 12345, out=out)
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            if b > a:
                reps = 3 if logn >= 22 else 10
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps): ctx.coset_lde(coeffs, B, 12345, out=out)
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                del coeffs, out
                torch.cuda.empty_cache()
            ms = max_over_ranks(ms)
            alg = (n + L) * cols * s
            if rank == 0:
                print(json.dumps({"field": name, "log_rows": logn, "cols": cols, "blowup": B, "gpus": g, "ms": round(ms, 4),
                                  "melem_per_s": round(L * cols / ms / 1e3, 1), "alg_gbytes": round(alg / 1e9, 3),
                                  "achieved_gbs": round(alg / ms / 1e6, 1), "frac_of_measured_hbm_peak": round(alg / ms / 1e6 / (peak * g), 4)}), flush=True)
    ctx.close()
if dist is not None:
    dist.destroy_process_group()
