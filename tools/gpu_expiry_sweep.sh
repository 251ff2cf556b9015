#!/bin/bash
# best_expiry_seg throughput against cards per warp / chunk size
for cfg in "32 2048" "32 8192" "8 8192" "32 16384" "8 16384" "16 16384" "32 32768" "8 32768" "16 32768" "8 65536"; do set -- $cfg; cpw=$1; chunk=$2
  echo -n "cpw=$cpw chunk=$chunk: "
  B200_DMZ_EXPIRY_CARDS_PER_WARP=$cpw B200_DMZ_EXPIRY_CHUNK=$chunk SIDE_BENCH_CARDS=${SIDE_BENCH_CARDS:-65536} timeout 300 python tools/gpu_side_bench.py 65536 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['best_expiry_seg']['per_s']), 'cards/s', d['best_expiry_seg']['cards_with_groups'], 'with groups')"
done
