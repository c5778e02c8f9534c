# compute-sanitizer over the parity tests at small sizes: memcheck (out-of-bounds / misaligned / leaks of device memory)
# and racecheck (shared-memory hazards in the tile NTT, leaf-hash ring, tree-top and scan kernels)
S=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  echo "== $tool: stages"
  timeout 420 $S --tool $tool --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k "(coset_lde and not max and not host and not rejects) or merkle_commit_base or merkle_commit_ext or intt_columns or ntt_table_cache" 2>&1 | tail -6
  echo "rc=$?"
  echo "== $tool: prove + fri"
  timeout 420 $S --tool $tool --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_prove.py tests/test_gpu_poly_fri.py -m gpu -x -q -k "synthetic_air or fri_deep_and_fold or deep_open or fri_commit" 2>&1 | tail -6
  echo "rc=$?"
done
