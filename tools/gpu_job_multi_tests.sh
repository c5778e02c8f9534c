# multi-GPU evidence: the sharded / virtual-rank / NCCL-through-the-C-ABI tests on a box with N GPUs, then the bench line at N
N=${1:-2}
TAG=${2:-r02_g}
if [ "${3:-tests}" = "tests" ]; then timeout 900 python -m pytest tests/test_sharded.py tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -6; fi
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
tail -3 gpurun_out/${TAG}_bench_${N}gpu.err | cut -c1-300
python - <<PY
import json
s=open("gpurun_out/${TAG}_bench_${N}gpu.json").read(); d=json.loads(s[s.index('{"metric'):])
print({k:d.get(k) for k in ("value","ms_per_step","prove_ms","n_gpus")}); print(d["e2e"]["ms_per_step"]); print(d["prove"]["stages_ms"]); print(d["prove"]["matches_oracle_digest"], d["prove"]["prove_samples_ms"])
print(d.get("baseline_configs"))
PY
df -h /dev/shm | tail -1
