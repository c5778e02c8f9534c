"""BASELINE config 4: batched coset-LDE sweep 2^16..2^24 rows x 64 cols, blowup 4, Goldilocks and BabyBear,
device-resident, CUDA events; prints one JSON line per point with the HBM-roofline fraction."""
import json, sys, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from ministark_b200 import Context
from ministark_b200.synth import synth_trace
peak = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs'] if __import__('os').path.exists('MEASURED_PEAKS.json') else 6650.0
cols, B = 64, 4
lo, hi = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (16, 24)
for field, name, s in ((0, 'goldilocks', 8), (1, 'babybear', 4)):
    ctx = Context(field)
    for logn in range(lo, hi + 1):
        n = 1 << logn; L = n * B
        base = synth_trace(field, min(n, 1 << 20), cols, seed=logn)           # tile a 2^20-row block: content is irrelevant for timing
        coeffs = ctx.to_device(np.ascontiguousarray(np.tile(base, (n // base.shape[0], 1)).T))
        out = ctx.empty(cols, L)
        for _ in range(2): ctx.coset_lde(coeffs, B, 12345, out=out)
        torch.cuda.synchronize()
        reps = 3 if logn >= 22 else 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): ctx.coset_lde(coeffs, B, 12345, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        alg = (n + L) * cols * s
        print(json.dumps({"field": name, "log_rows": logn, "cols": cols, "blowup": B, "ms": round(ms, 4),
                          "melem_per_s": round(L * cols / ms / 1e3, 1), "alg_gbytes": round(alg / 1e9, 3),
                          "achieved_gbs": round(alg / ms / 1e6, 1), "frac_of_measured_hbm_peak": round(alg / ms / 1e6 / peak, 4)}), flush=True)
        del coeffs, out
        torch.cuda.empty_cache()
    ctx.close()
