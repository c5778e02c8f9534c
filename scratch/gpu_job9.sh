#!/bin/bash
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
python -m pytest tests/test_sharded.py tests/test_gpu_prove.py -m gpu -x -q 2>&1 | tail -4
for dl in sharded rank0; do
  MINISTARK_DOWNLOAD=$dl $TR scratch/run_config.py 22 32 4 2 100 3 > gpurun_out/j9_hl_${dl}_${N}gpu.json 2> gpurun_out/j9_hl_${dl}_${N}gpu.err
  tail -1 gpurun_out/j9_hl_${dl}_${N}gpu.json | cut -c1-1400; grep -i "error\|Traceback" gpurun_out/j9_hl_${dl}_${N}gpu.err | head -5
done
MINISTARK_DOWNLOAD=sharded $TR scratch/run_config.py 24 64 4 2 100 2 > gpurun_out/j9_c5a_sharded_${N}gpu.json 2> gpurun_out/j9_c5a_sharded_${N}gpu.err
tail -1 gpurun_out/j9_c5a_sharded_${N}gpu.json | cut -c1-1400
