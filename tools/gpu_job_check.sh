set -x
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r02_d_bench_1gpu.json 2> gpurun_out/r02_d_bench_1gpu.err; tail -3 gpurun_out/r02_d_bench_1gpu.err
python - <<'PY'
import json
s=open("gpurun_out/r02_d_bench_1gpu.json").read(); d=json.loads(s[s.index('{"metric'):])
print({k:d.get(k) for k in ("value","ms_per_step","prove_ms","gpu_launches")}); print(d["roofline"]["alu"]); print(d["roofline"]["traffic"])
print(d["prove"]["stages_ms"]); print(d["cpu_baseline"]["prove"]); print({k:v.get("prove_ms") for k,v in d["baseline_configs"].items()})
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_d_bench_ref.json 2> gpurun_out/r02_d_bench_ref.err; tail -2 gpurun_out/r02_d_bench_ref.err; head -c 900 gpurun_out/r02_d_bench_ref.json
