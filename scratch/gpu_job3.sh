#!/bin/bash
for v in base add31 add31_fma12 add31_fma4 add15 add23; do
  echo "== $v"; MINISTARK_LIB=$PWD/scratch/ab/lib_$v.so python scratch/bench_stages.py 22 32 4 merkle,fri 2>&1 | tail -1
done > gpurun_out/j3_sha_ab.log 2>&1
cat gpurun_out/j3_sha_ab.log
