"""ms_stark_prove (host trace -> proof bytes) several times, with per-stage timings (diagnostic)."""
import ctypes as Cc, sys, time, numpy as np, torch
sys.path.insert(0, '.')
from ministark_b200 import Context
from ministark_b200._lib import StarkParams
from tests.synth import synth_linear_matrix, synth_trace
n, W, B = 1 << 22, 16, 4
ctx = Context(0)
trace_rm = synth_trace(0, n, W, seed=0x5EED000000000001)
m = synth_linear_matrix(0, n, W)
params = StarkParams(100, B, n - 1, 2 * W, 2)
bound = int(ctx.lib.ms_stark_proof_bound(0, params, n, 2 * W))
proof_buf = torch.empty(bound, dtype=torch.uint8).pin_memory().numpy()
h_trace = torch.from_numpy(trace_rm.view(np.int64)).pin_memory().numpy().view(np.uint64)
for mode in ('pinned', 'pageable'):
    src = h_trace if mode == 'pinned' else trace_rm
    for i in range(4):
        cap = Cc.c_uint64(proof_buf.size)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        rc = ctx.lib.ms_stark_prove(ctx.h, Cc.byref(params), src.ctypes.data, n, W, m.ctypes.data, W, proof_buf.ctypes.data, Cc.byref(cap))
        dt = time.perf_counter() - t0
        ctx._check(rc)
        print(mode, i, round(dt * 1e3, 2), {k: round(v, 2) for k, v in ctx.last_timings()}, flush=True)
