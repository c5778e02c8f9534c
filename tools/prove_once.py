import sys, time, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from ministark_b200 import Context
from ministark_b200._lib import StarkParams
from ministark_b200.synth import synth_linear_matrix, synth_trace
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 22
C = int(sys.argv[2]) if len(sys.argv) > 2 else 32
B = int(sys.argv[3]) if len(sys.argv) > 3 else 4
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
n = 1 << logn; W = C // 2
ctx = Context(0)
trace_rm = synth_trace(0, n, W, seed=0x5EED000000000001)
m = synth_linear_matrix(0, n, W)
params = StarkParams(100, B, n - 1, C, 2)
bound = int(ctx.lib.ms_stark_proof_bound(0, params, n, C))
proof_buf = torch.empty(bound, dtype=torch.uint8).pin_memory().numpy()
trace_cm = ctx.to_device(np.ascontiguousarray(trace_rm.T))
for i in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    plen = ctx.stark_prove_device(params, trace_cm, m, proof_buf)
    dt = time.perf_counter() - t0
    print('prove ms', dt * 1e3, 'proof bytes', plen, {k: round(v, 2) for k, v in ctx.last_timings()})
