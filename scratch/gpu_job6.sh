#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/j6_tests.log; cat gpurun_out/j6_tests.log
python bench.py > gpurun_out/j6_bench.json 2> gpurun_out/j6_bench.err; tail -2 gpurun_out/j6_bench.err
python -c "
import json
d=json.load(open('gpurun_out/j6_bench.json')); r=d['roofline']
print('LDE', d['value'], d['ms_per_step'], r['frac'], {k:round(v['ms_per_step'],3) for k,v in r['kernels_ms_per_step'].items()}); print('e2e', d['e2e']['ms_per_step']); print('prove', d['prove']['prove_ms'], d['prove']['stages_ms'], d['prove']['proof_sha256'][:16])"
python scratch/sweep_lde.py > gpurun_out/j6_sweep.log 2>&1; cat gpurun_out/j6_sweep.log | cut -c1-200
