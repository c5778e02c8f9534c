"""torchrun script: sharded prove of the headline shape; prints per-rank timing + proof hash."""
import hashlib, os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '.')
from ministark_b200 import Context
from ministark_b200._lib import StarkParams
from ministark_b200.sharded import stark_prove_sharded
from tests.synth import synth_linear_matrix, synth_trace
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 22
C = int(sys.argv[2]) if len(sys.argv) > 2 else 32
B = int(sys.argv[3]) if len(sys.argv) > 3 else 4
k = int(sys.argv[4]) if len(sys.argv) > 4 else 2
rank, world, lr = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(lr)
d = None
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device(f'cuda:{lr}'))
    d = dist
n = 1 << logn; W = C // 2
ctx = Context(0, lr)
trace_rm = synth_trace(0, n, W, seed=0x5EED000000000001)
m = synth_linear_matrix(0, n, W)
params = StarkParams(100, B, n - 1, C, k)
bound = int(ctx.lib.ms_stark_proof_bound(0, params, n, C))
buf = torch.empty(bound, dtype=torch.uint8).pin_memory().numpy()
trace_cm = ctx.to_device(np.ascontiguousarray(trace_rm.T))
for i in range(3):
    if d: d.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ln = stark_prove_sharded(ctx, params, trace_cm, m, buf, d)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    if i == 2:
        print(f'rank {rank}/{world} prove ms {dt*1e3:.2f} proof {ln} sha {hashlib.sha256(buf[:ln].tobytes()).hexdigest()[:16]}', {a: round(b, 2) for a, b in ctx.last_timings()}, ctx.last_sharded_stats, flush=True)
if d: dist.destroy_process_group()
