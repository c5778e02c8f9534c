#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/j14_bench.json 2> gpurun_out/j14_bench.err; tail -2 gpurun_out/j14_bench.err
python -c "
import json
d=json.load(open('gpurun_out/j14_bench.json')); r=d['roofline']
print('LDE', d['value'], d['ms_per_step'], r['frac']); print('e2e', d['e2e']['ms_per_step'], 'cpu', d['cpu_baseline']['value']); print('prove', d['prove']['prove_ms'], d['prove']['prove_e2e_ms'], d['prove']['stages_ms']); print(d['clocks'], d['gpu_launches'])"
python -c "import __graft_entry__ as g; g.smoke()"
