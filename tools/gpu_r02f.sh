#!/bin/bash
tag=${1:-r02f}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/${tag}_tests.txt
tail -25 gpurun_out/${tag}_tests.txt
( timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e > gpurun_out/${tag}_bench.json ) 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench.json'))
print('value', d['value'], {k:round(v['ms_per_100k_frames'],2) for k,v in d['stages'].items()})
print(json.dumps(d['cpu_baseline']['parity_vs_gpu_on_sample']))
PY
tail -3 gpurun_out/${tag}_bench.err
( B200_DMZ_VSEG_FP32=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/${tag}_bench_fp32.json ) 2> gpurun_out/${tag}_bench_fp32.err
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench_fp32.json'))
print('fp32 vseg: value', d['value'], {k:round(v['ms_per_100k_frames'],2) for k,v in d['stages'].items()})
PY
