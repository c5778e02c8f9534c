for v in 8 9 10 12 24 14 26 11; do
  echo "== mask $v"; MINISTARK_LIB=ministark_b200/variants/lib_sha$v.so python tools/bench_stages.py 22 32 4 merkle,fri
done
