# last check of the round on one GPU: the whole GPU suite, smoke(), and the default bench line (roofline.traffic / alu from the committed capture)
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_h_bench_1gpu.json 2> gpurun_out/r02_h_bench_1gpu.err; tail -2 gpurun_out/r02_h_bench_1gpu.err
python - <<PY
import json
s=open("gpurun_out/r02_h_bench_1gpu.json").read(); d=json.loads(s[s.index('{"metric'):])
print({k:d.get(k) for k in ("value","ms_per_step","prove_ms","gpu_launches")}); print(d["roofline"]["frac"], d["roofline"]["traffic"], d["roofline"]["alu"]["frac"], d["e2e"]["ms_per_step"], d["clocks"])
PY
