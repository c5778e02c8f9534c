N=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
numactl -H 2>/dev/null | head -4; nvidia-smi topo -m 2>/dev/null | head -12
echo "== all ranks download"; $T tools/prove_multi.py 22 32 4 5 2>&1 | grep "prove ms" | cut -c1-420
echo "== 4 ranks download"; MINISTARK_DL_MAX_RANKS=4 $T tools/prove_multi.py 22 32 4 5 2>&1 | grep "prove ms" | cut -c1-420
echo "== interleaved shm, all ranks"; MINISTARK_SHM_INTERLEAVE=1 $T tools/prove_multi.py 22 32 4 5 2>&1 | grep "prove ms" | cut -c1-420
echo "== interleaved shm, 4 ranks"; MINISTARK_SHM_INTERLEAVE=1 MINISTARK_DL_MAX_RANKS=4 $T tools/prove_multi.py 22 32 4 5 2>&1 | grep "prove ms" | cut -c1-420
