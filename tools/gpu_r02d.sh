#!/bin/bash
# tests + the three bench configs (short), device-resident only
tag=${1:-r02d}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/${tag}_tests.txt
tail -3 gpurun_out/${tag}_tests.txt
( timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/${tag}_bench.json ) 2> gpurun_out/${tag}_bench.err
( timeout 600 python bench.py --config categorize --steps 3 --warmup 3 --no-cpu > gpurun_out/${tag}_cat.json ) 2> gpurun_out/${tag}_cat.err
( timeout 900 python bench.py --config detect-sweep --steps 2 --warmup 3 --no-cpu > gpurun_out/${tag}_sweep.json ) 2> gpurun_out/${tag}_sweep.err
tail -c 1500 gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
tail -c 1500 gpurun_out/${tag}_cat.json; tail -3 gpurun_out/${tag}_cat.err
tail -c 2500 gpurun_out/${tag}_sweep.json; tail -3 gpurun_out/${tag}_sweep.err
