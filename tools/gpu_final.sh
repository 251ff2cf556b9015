#!/bin/bash
# One GPU-box visit for the record: tests, the three bench configs + the reference arm, launch list, full ncu capture of the step's kernels.
tag=${1:-r02z}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/${tag}_gpu.csv 2>&1
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > gpurun_out/${tag}_tests.txt; tail -3 gpurun_out/${tag}_tests.txt
( timeout 900 python bench.py > gpurun_out/${tag}_bench.json ) 2> gpurun_out/${tag}_bench.err; tail -2 gpurun_out/${tag}_bench.err
( timeout 300 python bench.py --impl reference > gpurun_out/${tag}_bench_ref.json ) 2> gpurun_out/${tag}_bench_ref.err
( timeout 600 python bench.py --config formats --steps 5 --warmup 3 > gpurun_out/${tag}_formats.json ) 2> gpurun_out/${tag}_formats.err
( timeout 300 python bench.py --impl reference --config formats --steps 3 --warmup 1 > gpurun_out/${tag}_formats_ref.json ) 2>> gpurun_out/${tag}_formats.err
bash tools/gpu_sanitize_formats.sh 2>&1 | tail -8
( timeout 600 python bench.py --config categorize > gpurun_out/${tag}_categorize.json ) 2> gpurun_out/${tag}_categorize.err
( timeout 900 python bench.py --config detect-sweep --steps 2 > gpurun_out/${tag}_detect_sweep.json ) 2> gpurun_out/${tag}_detect_sweep.err
( timeout 300 python tools/gpu_side_bench.py 262144 > gpurun_out/${tag}_side.json ) 2> gpurun_out/${tag}_side.err
CMD="python bench.py --frames 4096 --steps 1 --warmup 1 --no-e2e --no-cpu --no-materialised"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > /dev/null 2> gpurun_out/${tag}_launches.err
ncu --set full --clock-control none --import-source on -k regex:"detect_strips|warp_rows|vseg_mma|vseg_select|categorize_mma|hseg_kernel|digit_prep" -s 11 -c 11 -o gpurun_out/${tag}_prof -f $CMD > /dev/null 2> gpurun_out/${tag}_prof.err
ls -la gpurun_out/${tag}_prof.ncu-rep
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], {k:round(v['ms_per_100k_frames'],2) for k,v in d['stages'].items()})
PY
