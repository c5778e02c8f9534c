#!/bin/bash
# N = 2: tests (incl. both exchange modes), sharded prove A/B peer vs nccl on headline + config 5a
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/j4_tests.log
cat gpurun_out/j4_tests.log
for mode in peer nccl; do
  MINISTARK_EXCHANGE=$mode python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scratch/run_config.py 22 32 4 2 100 3 > gpurun_out/j4_hl_${mode}_2gpu.json 2> gpurun_out/j4_hl_${mode}_2gpu.err
  tail -1 gpurun_out/j4_hl_${mode}_2gpu.json | cut -c1-1200
done
python scratch/bench_stages.py 22 32 4 merkle 2>&1 | tail -1
