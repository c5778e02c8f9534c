# usage: bash tools/gpu_job_multi.sh N   (N GPUs on one box)
N=${1:-2}
set -x
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "nccl" 2>&1 | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/r02_e_bench_${N}gpu.json 2> gpurun_out/r02_e_bench_${N}gpu.err
tail -5 gpurun_out/r02_e_bench_${N}gpu.err
python - <<PY
import json
s=open("gpurun_out/r02_e_bench_${N}gpu.json").read(); d=json.loads(s[s.index("{\"metric"):])
print({k:d.get(k) for k in ("value","ms_per_step","prove_ms","n_gpus")})
print(d.get("lde_weak")); print(d.get("e2e"))
print(d["prove"]["stages_ms"]); print(d["prove"]["matches_oracle_digest"], d["prove"]["prove_samples_ms"])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/prove_multi.py 24 64 4 4 2>&1 | grep "prove ms"
