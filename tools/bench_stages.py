"""Per-stage device timings (CUDA events) on the headline shape; used for A/B builds via MINISTARK_LIB."""
import sys, time, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from ministark_b200 import Context
from ministark_b200.synth import synth_trace
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 22
C = int(sys.argv[2]) if len(sys.argv) > 2 else 32
B = int(sys.argv[3]) if len(sys.argv) > 3 else 4
what = sys.argv[4].split(',') if len(sys.argv) > 4 else ['lde', 'intt', 'merkle', 'fri', 'open']
n = 1 << logn; L = n * B
ctx = Context(0)
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
res = {}
coeffs = ctx.to_device(np.ascontiguousarray(synth_trace(0, n, C, seed=1).T))
out = ctx.empty(C, L)
if 'lde' in what: res['lde'] = timeit(lambda: ctx.coset_lde(coeffs, B, 0x123456789ABCDEF, out=out))
else: ctx.coset_lde(coeffs, B, 0x123456789ABCDEF, out=out)
if 'intt' in what: res['intt_c/2'] = timeit(lambda: ctx.intt_columns(coeffs[: C // 2]))
if 'merkle' in what:
    res['lde_commit'] = timeit(lambda: ctx.merkle_commit(out, C, 2), reps=2)
    res['trace_commit'] = timeit(lambda: ctx.merkle_commit(coeffs[: C // 2], C, 2), reps=2)
if 'fri' in what:
    planes = ctx.zeros(2, n); planes[0] = coeffs[0]
    res['fri_commit_r0'] = timeit(lambda: ctx.fri_commit(planes, L, B), reps=2)
    planes[1] = coeffs[1]
    res['fri_commit_r0_ext'] = timeit(lambda: ctx.fri_commit(planes, L, B), reps=2)
if 'open' in what:
    z = np.array([[3, 5], [7, 11], [13, 17]], dtype=np.uint64)
    res['deep_open'] = timeit(lambda: ctx.deep_open(coeffs, z), reps=2)
print({k: round(v, 3) for k, v in res.items()})
