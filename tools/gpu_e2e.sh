#!/bin/bash
# e2e (host-buffer) throughput for a list of env settings (one per argument)
mkdir -p gpurun_out; : > gpurun_out/e2e.jsonl
for cfg in "${@:-_=}"; do
  echo "== $cfg" | tee -a gpurun_out/e2e.jsonl
  ( env $cfg timeout 300 python bench.py --steps 3 --warmup 3 --frames 16384 --no-cpu 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
e = d['e2e']
print(json.dumps({'value': d['value'], 'e2e': e['value'], 'h2d_per_frame': e['h2d_bytes_per_step'] / e['frames_per_step'], 'redos': e['full_frame_redos'], 'equal': e['records_equal_device_path']}))
" ) | tee -a gpurun_out/e2e.jsonl
done
