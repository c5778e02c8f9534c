import sys, numpy as np
sys.path.insert(0, '.')
from ministark_b200 import Context
from oracle import oracle as O
import torch
ctx = Context(0)
for log_n in (12, 13):
    n = 1 << log_n
    rng = np.random.default_rng(1)
    tr = rng.integers(0, 2**62, size=(n, 1), dtype=np.uint64)
    want = O.trace_polys(0, tr)
    d = ctx.to_device(np.ascontiguousarray(tr.T))
    out = ctx.zeros(1, n)
    import ctypes as C
    rc = ctx.lib.ms_intt_columns(ctx.h, C.c_void_p(d.data_ptr()), n, n, 1, C.c_void_p(out.data_ptr()), n)
    print('rc', rc, ctx.lib.ms_last_error(ctx.h))
    ctx.sync(); torch.cuda.synchronize()
    got = ctx.to_host(out)
    print(log_n, (got == want).all(), got[0, :4], want[0, :4], (got==want).sum(), n)
