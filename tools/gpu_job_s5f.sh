# session 5, call f: 2^14-element tiles for large BabyBear transforms -- parity at the largest shapes, the sweep points they change, fresh ncu capture of the LDE
timeout 600 python -m pytest tests/test_gpu_scale.py tests/test_gpu_stages.py -m gpu -x -q -k "coset_lde or selftest or intt_columns or table_cache" 2>&1 | tail -3
timeout 300 python tools/sweep_lde.py 21 24 | tee gpurun_out/r02_h_lde_sweep_1gpu.jsonl | cut -c1-200
timeout 600 ncu --set full --metrics smsp__thread_inst_executed.sum --clock-control none --import-source on -k regex:k_ntt -s 8 -c 4 -o gpurun_out/r02_h_ncu_lde python bench.py --steps 2 --warmup 4 --no-prove --no-e2e --no-cpu --no-extras > gpurun_out/r02_h_ncu_lde.log 2>&1; tail -1 gpurun_out/r02_h_ncu_lde.log | cut -c1-200
