S=/usr/local/cuda/bin/compute-sanitizer
for tool in initcheck synccheck; do
  echo "== $tool: stages + prove"
  timeout 420 $S --tool $tool --error-exitcode 9 --print-limit 8 python -m pytest tests/test_gpu_stages.py tests/test_gpu_prove.py -m gpu -x -q -k "(coset_lde and not max and not host and not rejects) or merkle_commit_base or intt_columns or synthetic_air" 2>&1 | tail -12
  echo "rc=$?"
done
echo "== memcheck: virtual ranks (sharded prover on one GPU)"
timeout 420 $S --tool memcheck --error-exitcode 9 --print-limit 8 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "virtual_ranks_equal_single_gpu" 2>&1 | tail -8
echo "rc=$?"
