set -x
timeout 300 python -m pytest tests/test_gpu_stages.py -m gpu -q -x -k "trace_generation or selftest" 2>&1 | tail -5
timeout 600 python bench.py > gpurun_out/r02_a_bench_1gpu.json 2> gpurun_out/r02_a_bench_1gpu.err; tail -3 gpurun_out/r02_a_bench_1gpu.err; head -c 1500 gpurun_out/r02_a_bench_1gpu.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_a_bench_ref.json 2> gpurun_out/r02_a_bench_ref.err; tail -3 gpurun_out/r02_a_bench_ref.err; head -c 600 gpurun_out/r02_a_bench_ref.json
# launch list of one whole prove (cold-cache, serialised: compare shares)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_a_launches_prove.csv python tools/prove_once.py 22 32 4 1 > gpurun_out/r02_a_prove_once.log 2>&1; tail -2 gpurun_out/r02_a_prove_once.log
# full capture of the LDE kernels (2 calls = 4 launches after warm-up) and of the Merkle kernels
timeout 600 ncu --set full --metrics smsp__thread_inst_executed.sum --clock-control none --import-source on -k regex:k_ntt -s 8 -c 4 -o gpurun_out/r02_a_ncu_lde python bench.py --steps 2 --warmup 4 --no-prove --no-e2e --no-cpu --no-extras > gpurun_out/r02_a_ncu_lde.log 2>&1; tail -2 gpurun_out/r02_a_ncu_lde.log
timeout 600 ncu --set full --metrics smsp__thread_inst_executed.sum --clock-control none --import-source on -k regex:"k_leaf_hash|k_node_hash" -c 3 -o gpurun_out/r02_a_ncu_merkle python tools/bench_stages.py 22 32 4 merkle > gpurun_out/r02_a_ncu_merkle.log 2>&1; tail -2 gpurun_out/r02_a_ncu_merkle.log
ls -la gpurun_out | tail -12
