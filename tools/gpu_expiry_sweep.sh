#!/bin/bash
( timeout 600 python -m pytest tests -m gpu -x -q -k "expiry" 2>&1 | tail -4 )
for cfg in "32 8192 8192" "32 32768 65536"; do set -- $cfg; cpw=$1; chunk=$2; cards=$3
  echo -n "cpw=$cpw chunk=$chunk cards=$cards: "
  B200_DMZ_EXPIRY_CARDS_PER_WARP=$cpw B200_DMZ_EXPIRY_CHUNK=$chunk SIDE_BENCH_CARDS=$cards timeout 300 python tools/gpu_side_bench.py 65536 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['best_expiry_seg']['per_s']), 'cards/s', d['best_expiry_seg']['cards_with_groups'], 'with groups')"
done
