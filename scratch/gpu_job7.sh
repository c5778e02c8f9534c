#!/bin/bash
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for mode in peer nccl; do
  MINISTARK_EXCHANGE=$mode $TR scratch/run_config.py 22 32 4 2 100 3 > gpurun_out/j7_hl_${mode}_${N}gpu.json 2> gpurun_out/j7_hl_${mode}_${N}gpu.err
  tail -1 gpurun_out/j7_hl_${mode}_${N}gpu.json | cut -c1-1300
done
MINISTARK_EXCHANGE=peer $TR scratch/run_config.py 24 64 4 2 100 2 > gpurun_out/j7_c5a_peer_${N}gpu.json 2> gpurun_out/j7_c5a_peer_${N}gpu.err
tail -1 gpurun_out/j7_c5a_peer_${N}gpu.json | cut -c1-1300
$TR bench.py --gpus $N > gpurun_out/j7_bench_${N}gpu.json 2> gpurun_out/j7_bench_${N}gpu.err
tail -1 gpurun_out/j7_bench_${N}gpu.json | cut -c1-400
