# session 5 final 1-GPU evidence (r02_g): bench line, reference arm, prove launch list, full ncu captures of the LDE and Merkle kernels
set -x
timeout 900 python bench.py > gpurun_out/r02_g_bench_1gpu.json 2> gpurun_out/r02_g_bench_1gpu.err; tail -3 gpurun_out/r02_g_bench_1gpu.err; head -c 600 gpurun_out/r02_g_bench_1gpu.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_g_bench_ref.json 2> gpurun_out/r02_g_bench_ref.err; tail -3 gpurun_out/r02_g_bench_ref.err; head -c 300 gpurun_out/r02_g_bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_g_launches_prove.csv python tools/prove_once.py 22 32 4 1 > gpurun_out/r02_g_prove_once.log 2>&1; tail -1 gpurun_out/r02_g_prove_once.log | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_g_launches_lde.csv python bench.py --steps 2 --warmup 1 --no-prove --no-e2e --no-cpu --no-extras > gpurun_out/r02_g_launches_lde.log 2>&1
timeout 600 ncu --set full --metrics smsp__thread_inst_executed.sum --clock-control none --import-source on -k regex:k_ntt -s 8 -c 4 -o gpurun_out/r02_g_ncu_lde python bench.py --steps 2 --warmup 4 --no-prove --no-e2e --no-cpu --no-extras > gpurun_out/r02_g_ncu_lde.log 2>&1; tail -2 gpurun_out/r02_g_ncu_lde.log | cut -c1-300
timeout 600 ncu --set full --metrics smsp__thread_inst_executed.sum --clock-control none --import-source on -k regex:"k_leaf_hash|k_node_hash" -c 3 -o gpurun_out/r02_g_ncu_merkle python tools/bench_stages.py 22 32 4 merkle > gpurun_out/r02_g_ncu_merkle.log 2>&1; tail -2 gpurun_out/r02_g_ncu_merkle.log | cut -c1-300
ls -la gpurun_out | grep r02_g
