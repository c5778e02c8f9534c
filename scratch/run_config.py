"""Full Stark::prove of one BASELINE.json configuration, single GPU or sharded under torchrun.
  python scratch/run_config.py LOGN C BLOWUP ARITY [SECURITY_BITS] [REPS]
Prints one JSON line (rank 0): prove ms (max over ranks, wall clock around the blocking call + device sync),
per-stage device timings of rank 0, the proof's length and sha256 (identical at every world size)."""
import hashlib, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from ministark_b200 import Context
from ministark_b200._lib import StarkParams
from ministark_b200.sharded import SharedProofBuffer, stark_prove_sharded
from tests.synth import synth_linear_matrix, synth_trace

logn, C, B, k = (int(x) for x in sys.argv[1:5])
sec = int(sys.argv[5]) if len(sys.argv) > 5 else 100
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
rank, world, lr = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(lr)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=torch.device(f'cuda:{lr}'))
n = 1 << logn
W = C // 2
ctx = Context(0, lr)
trace_rm = synth_trace(0, n, W, seed=0x5EED000000000001)
m = synth_linear_matrix(0, n, W)
params = StarkParams(sec, B, n - 1, C, k)
bound = int(ctx.lib.ms_stark_proof_bound(0, params, n, C))
shared_dl = world > 1 and os.environ.get('MINISTARK_DOWNLOAD', 'sharded') == 'sharded'
if shared_dl:
    shared = SharedProofBuffer(ctx, bound, dist)  # every rank downloads 1/world of the quotient polynomials
    buf = shared.array
else:
    buf = torch.empty(bound, dtype=torch.uint8).pin_memory().numpy() if rank == 0 else np.empty(bound, dtype=np.uint8)  # replicas only write the fixed part
trace_cm = ctx.to_device(np.ascontiguousarray(trace_rm.T))
del trace_rm
times, ln, stages = [], 0, None
for i in range(reps + 1):
    if dist: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if world > 1:
        ln = stark_prove_sharded(ctx, params, trace_cm, m, shared if shared_dl else buf, dist)
    else:
        ln = ctx.stark_prove_device(params, trace_cm, m, buf)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=f'cuda:{lr}')
    if dist: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if i > 0:
        times.append(float(t.item()) * 1e3)
        stages = ctx.last_timings()
if rank == 0:
    print(json.dumps({
        "config": f"goldilocks SynthLinear W={W} (C={C}) N=2^{logn} blowup {B} arity {k} security {sec}", "n_gpus": world,
        "prove_ms": round(float(np.mean(times)), 3), "prove_ms_min": round(float(np.min(times)), 3), "reps": reps,
        "proof_bytes": ln, "proof_sha256": hashlib.sha256(buf[:ln].tobytes()).hexdigest(),
        "stages_ms": {a: round(b, 3) for a, b in (stages or [])},
        "sharded": getattr(ctx, "last_sharded_stats", None) if world > 1 else None,
        "download": ("sharded over the ranks (shared host buffer)" if shared_dl else "rank 0") if world > 1 else "single GPU",
        "mem_gb": round(torch.cuda.max_memory_allocated() / 1e9, 2)}), flush=True)
if shared_dl:
    del buf
    shared.close()
if dist: dist.destroy_process_group()
ctx.close()
