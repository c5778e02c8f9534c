# session 5, call a: BabyBear product (double REDC) through the whole GPU suite + host topology / pinned-placement probe
timeout 300 python tools/numa_probe.py 1024 > gpurun_out/s5a_numa.log 2>&1; tail -12 gpurun_out/s5a_numa.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/s5a_tests.log
