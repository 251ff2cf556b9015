#!/bin/bash
tag=${1:-r02e}
mkdir -p gpurun_out
./tools/microbench/ffma2 > gpurun_out/${tag}_ffma2.txt 2>&1; cat gpurun_out/${tag}_ffma2.txt
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/${tag}_tests.txt
tail -3 gpurun_out/${tag}_tests.txt
( timeout 900 python bench.py > gpurun_out/${tag}_bench.json ) 2> gpurun_out/${tag}_bench.err
tail -c 6000 gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
CMD="python bench.py --frames 4096 --steps 1 --warmup 1 --no-e2e --no-cpu --no-materialised"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > /dev/null 2> gpurun_out/${tag}_launches.err
ncu --set full --clock-control none --import-source on -k regex:"detect_strips|warp_rows|vseg_rows|categorize_kernel|hseg_kernel|digit_prep" -s 9 -c 9 -o gpurun_out/${tag}_prof -f $CMD > /dev/null 2> gpurun_out/${tag}_prof.err
ls -la gpurun_out/${tag}_prof.ncu-rep
