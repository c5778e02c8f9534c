"""One sharded proof of the synthetic AIR on N GPUs (torchrun), timed: `torchrun --nproc-per-node N tools/prove_multi.py LOGN C BLOWUP REPS`.
Prints per-rep wall times (max over ranks), rank 0's stage times and the proof digest."""
import hashlib, os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ministark_b200 import Context
from ministark_b200._lib import StarkParams
from ministark_b200.sharded import SharedProofBuffer
from ministark_b200.synth import synth_linear_matrix
logn, C, B, reps = (int(x) for x in (sys.argv[1:5] + ["22", "32", "4", "5"][len(sys.argv) - 1:]))
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lr}"))
ctx = Context(0, lr)
uid = [Context.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
ctx.comm_init_nccl(uid[0], rank, world)
n, W = 1 << logn, C // 2
m = synth_linear_matrix(0, n, W)
params = StarkParams(100, B, n - 1, C, 2)
shared = SharedProofBuffer(ctx, int(ctx.lib.ms_stark_proof_bound(0, params, n, C)), dist)
trace_cm = ctx.trace_synth(n, W, seed=0x5EED000000000001)
times = []
for i in range(reps):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    plen = ctx.stark_prove_multi(params, trace_cm, m, shared.array, shared=True)
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{lr}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    times.append(round(float(t.item()) * 1e3, 2))
if rank == 0:
    print("prove ms", times, "proof", plen, hashlib.sha256(shared.array[:plen].tobytes()).hexdigest()[:16], {k: round(v, 2) for k, v in ctx.last_timings()}, flush=True)
shared.close(); ctx.comm_destroy(); dist.destroy_process_group(); ctx.close()
