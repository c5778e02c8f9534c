set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
B="python bench.py --no-prove --no-e2e --no-cpu --no-extras --steps 10"
$B | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('base tile13', d['ms_per_step'], d['roofline']['kernels_ms_per_step'])"
MINISTARK_NTT_TILE=12 $B | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tile12', d['ms_per_step'], d['roofline']['kernels_ms_per_step'])"
for v in minb2 minb4; do
  MINISTARK_LIB=ministark_b200/variants/lib_$v.so $B | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v tile13', d['ms_per_step'])"
  MINISTARK_LIB=ministark_b200/variants/lib_$v.so MINISTARK_NTT_TILE=12 $B | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v tile12', d['ms_per_step'])"
done
python tools/prove_once.py 22 32 4 3 | tail -2
MINISTARK_LDE_LINEARITY=0 python tools/prove_once.py 22 32 4 3 | tail -1
python tools/prove_once.py 24 64 4 3 | tail -2
python tools/sweep_lde.py 16 24 > gpurun_out/r02_c_lde_sweep_tile13.jsonl; MINISTARK_NTT_TILE=12 python tools/sweep_lde.py 16 24 > gpurun_out/r02_c_lde_sweep_tile12.jsonl
paste -d'|' <(cut -c1-200 gpurun_out/r02_c_lde_sweep_tile13.jsonl | python -c "import sys,json; [print(json.loads(l)['field'][:2], json.loads(l)['log_rows'], json.loads(l)['ms']) for l in sys.stdin]") <(python -c "import sys,json; [print(json.loads(l)['ms']) for l in open('gpurun_out/r02_c_lde_sweep_tile12.jsonl')]")
