#!/bin/bash
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
MINISTARK_DOWNLOAD=sharded $TR scratch/run_config.py 22 32 4 2 100 3 > gpurun_out/j13_hl_sharded_${N}gpu.json 2> gpurun_out/j13_hl_sharded_${N}gpu.err
tail -1 gpurun_out/j13_hl_sharded_${N}gpu.json | cut -c1-1500
MINISTARK_DOWNLOAD=rank0 $TR scratch/run_config.py 22 32 4 2 100 3 > gpurun_out/j13_hl_rank0_${N}gpu.json 2> gpurun_out/j13_hl_rank0_${N}gpu.err
tail -1 gpurun_out/j13_hl_rank0_${N}gpu.json | cut -c1-1500
