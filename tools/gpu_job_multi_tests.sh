N=${1:-2}
timeout 900 python -m pytest tests/test_sharded.py tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/r02_e_bench_${N}gpu.json 2> gpurun_out/r02_e_bench_${N}gpu.err
python - <<PY
import json
s=open("gpurun_out/r02_e_bench_${N}gpu.json").read(); d=json.loads(s[s.index('{"metric'):])
print({k:d.get(k) for k in ("value","ms_per_step","prove_ms","n_gpus")}); print(d["e2e"]["ms_per_step"]); print(d["prove"]["stages_ms"]); print(d["prove"]["matches_oracle_digest"], d["prove"]["prove_samples_ms"])
PY
