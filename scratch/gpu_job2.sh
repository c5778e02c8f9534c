#!/bin/bash
set -x
python -m pytest tests/test_gpu_stages.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/j2_tests.log
python bench.py --no-prove > gpurun_out/j2_bench.json 2> gpurun_out/j2_bench.err
python scratch/run_config.py 20 16 8 2 100 3 > gpurun_out/cfg3a_1gpu.json 2> gpurun_out/cfg3a_1gpu.err
python scratch/run_config.py 21 16 8 4 100 3 > gpurun_out/cfg3b_1gpu.json 2> gpurun_out/cfg3b_1gpu.err
python scratch/run_config.py 22 64 4 8 100 3 > gpurun_out/cfg5b_1gpu.json 2> gpurun_out/cfg5b_1gpu.err
python scratch/run_config.py 24 64 4 2 100 2 > gpurun_out/cfg5a_1gpu.json 2> gpurun_out/cfg5a_1gpu.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ntt_fixed -s 4 -c 2 -o gpurun_out/r01_d_ntt -f python bench.py --steps 1 --warmup 3 --no-prove --no-e2e --no-cpu > gpurun_out/ncu_ntt.log 2>&1
cat gpurun_out/j2_tests.log gpurun_out/cfg3a_1gpu.json gpurun_out/cfg3b_1gpu.json gpurun_out/cfg5b_1gpu.json gpurun_out/cfg5a_1gpu.json; tail -2 gpurun_out/cfg5a_1gpu.err
