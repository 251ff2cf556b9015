#!/bin/bash
tag=${1:-r02g}
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
( timeout 600 python bench.py --config categorize --steps 3 --warmup 3 --no-cpu > gpurun_out/${tag}_cat.json ) 2> gpurun_out/${tag}_cat.err
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_cat.json'))
print('categorize-only: value', d['value'], d['roofline']['frac'])
PY
tail -3 gpurun_out/${tag}_cat.err
