#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/j11_bench.json 2> gpurun_out/j11_bench.err; tail -2 gpurun_out/j11_bench.err
python -c "
import json
d=json.load(open('gpurun_out/j11_bench.json')); r=d['roofline']
print('LDE', d['value'], d['ms_per_step'], r['frac']); print('e2e', d['e2e']['ms_per_step'], 'cpu', d['cpu_baseline']['value']); print('prove', d['prove']['prove_ms'], d['prove']['prove_e2e_ms'], d['prove']['stages_ms']); print(d['clocks'], d['gpu_launches'])"
for v in minb2; do MINISTARK_LIB=$PWD/scratch/ab/lib_$v.so python scratch/bench_stages.py 22 32 4 lde,intt 2>&1 | tail -1; done
python scratch/bench_stages.py 22 32 4 lde,intt 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_f_launches_prove.csv python scratch/prove_once.py 22 32 4 1 > gpurun_out/ncu_l3.log 2>&1
