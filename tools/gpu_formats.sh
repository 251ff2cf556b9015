#!/bin/bash
# formats rows: parity tests (formats + the two drop-in callers), the bench line, a launch list and one ncu --set full capture
tag=${TAG:-fmt}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_formats.py tests/test_gpu_parity.py -m gpu -x -q -k "formats or ycbcr or rgba or stencil or dropin" 2>&1 | tail -15 ) > gpurun_out/${tag}_tests.txt
tail -5 gpurun_out/${tag}_tests.txt
( timeout 600 python bench.py --config formats --steps 5 --warmup 3 > gpurun_out/${tag}_formats.json ) 2> gpurun_out/${tag}_formats.err
tail -3 gpurun_out/${tag}_formats.err
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_formats.json'))
print('value', d['value'], 'e2e', d['e2e'] and d['e2e']['value'], 'cpu', d['cpu_baseline'])
for k,v in d['legs'].items(): print(k, round(v['ms_per_launch'],3), 'ms', round(v['achieved_GBps']), 'GB/s', round(v['frac_of_hbm_peak'],3))
PY
( timeout 600 python bench.py --impl reference --config formats --steps 3 --warmup 1 > gpurun_out/${tag}_formats_ref.json ) 2>> gpurun_out/${tag}_formats.err
if [ -z "$NO_NCU" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ycbcr_to_rgb_kernel|rgba_to_r_kernel|stencil3_kernel" -c 7 -s 7 \
    -o gpurun_out/${tag}_formats_ncu -f python bench.py --config formats --steps 1 --warmup 1 --no-e2e --no-cpu --format-frames 2048 > gpurun_out/${tag}_ncu.log 2>&1
  tail -3 gpurun_out/${tag}_ncu.log
fi
