# last GPU calls of round 2 (budget: ~80 s).  Two calls: a short bench line on the final library, then the prover / verifier tests touched
# since the r02_h suite run (a '^' in a -k expression is a pytest syntax error: select the 2^14 x 8 golden shape by "14x8")
timeout 40 python bench.py --steps 3 --no-cpu --no-extras --no-e2e > gpurun_out/r02_i_bench_quick.json 2> gpurun_out/r02_i_bench_quick.err; tail -c 400 gpurun_out/r02_i_bench_quick.err; tail -c 300 gpurun_out/r02_i_bench_quick.json
timeout 33 python -m pytest tests/test_gpu_prove.py -x -q -k "degenerate or rejects or affine or (verifier and 14x8)" 2>&1 | tail -12
