#!/bin/bash
# configs 3 and 5 on one GPU, then refreshed ncu evidence for the current kernels
set -x
python scratch/run_config.py 20 16 8 4 100 3 > gpurun_out/cfg3_1gpu.json 2> gpurun_out/cfg3_1gpu.err
python scratch/run_config.py 24 64 4 8 100 2 > gpurun_out/cfg5_1gpu.json 2> gpurun_out/cfg5_1gpu.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ntt_fixed -s 8 -c 2 -o gpurun_out/r01_d_ntt -f python bench.py --steps 1 --warmup 3 --no-prove --no-e2e --no-cpu > gpurun_out/ncu_ntt.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_leaf_hash|k_node_hash" -c 3 -o gpurun_out/r01_d_merkle -f python scratch/bench_stages.py 22 32 4 merkle > gpurun_out/ncu_merkle.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_d_launches_lde.csv python bench.py --steps 2 --warmup 1 --no-prove --no-e2e --no-cpu > gpurun_out/ncu_l1.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_d_launches_prove.csv python scratch/prove_once.py 22 32 4 1 > gpurun_out/ncu_l2.log 2>&1
cat gpurun_out/cfg3_1gpu.json gpurun_out/cfg5_1gpu.json; tail -2 gpurun_out/cfg5_1gpu.err
