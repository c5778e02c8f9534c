#!/bin/bash
N=4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
MINISTARK_DOWNLOAD=sharded $TR scratch/run_config.py 22 32 4 2 100 3 > gpurun_out/j15_hl_sharded_${N}gpu.json 2> gpurun_out/j15_hl_sharded_${N}gpu.err
tail -1 gpurun_out/j15_hl_sharded_${N}gpu.json | cut -c1-1500
