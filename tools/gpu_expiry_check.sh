#!/bin/bash
# E0 tensor-core kernel: the expiry tests, then throughput of both kernels
( timeout 600 python -m pytest tests -m gpu -x -q -k "expiry" 2>&1 | tail -12 )
for fp32 in "" 1; do
  echo -n "B200_DMZ_EXPIRY_FP32=$fp32: "
  B200_DMZ_EXPIRY_FP32=$fp32 SIDE_BENCH_CARDS=1024 timeout 300 python tools/gpu_side_bench.py 262144 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['expiry_digits']['per_s']), 'crops/s')"
done
