#!/bin/bash
# Quick GPU visit: GPU parity tests + device-resident bench lines for a list of env settings (one per argument).
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
: > gpurun_out/quick.jsonl
for cfg in "${@:-_=}"; do
  echo "== $cfg" | tee -a gpurun_out/quick.jsonl
  ( env $cfg timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({'value': d['value'], 'stages': {k: round(v['ms_per_100k_frames'], 2) for k, v in d['stages'].items()}}))
" ) | tee -a gpurun_out/quick.jsonl
done
