# last GPU call of round 2 (budget: ~80 s): the prover / verifier tests touched since the r02_h suite run, then a short bench line
timeout 55 python -m pytest tests/test_gpu_prove.py -x -q -k "degenerate or e2e or synthetic or rejects or affine or gl_2^14x8_b4" 2>&1 | tail -4
timeout 40 python bench.py --steps 3 --no-cpu --no-extras --no-e2e > gpurun_out/r02_i_bench_quick.json 2> gpurun_out/r02_i_bench_quick.err; tail -c 400 gpurun_out/r02_i_bench_quick.err; tail -c 300 gpurun_out/r02_i_bench_quick.json
