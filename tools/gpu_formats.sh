#!/bin/bash
# formats rows: parity tests (formats + the two drop-in callers), the bench line (+ the direct-store variant of the colour
# conversion), a launch list and one ncu --set full capture
tag=${TAG:-fmt}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_formats.py tests/test_gpu_parity.py -m gpu -x -q -k "formats or ycbcr or rgba or stencil or dropin" 2>&1 | tail -15 ) > gpurun_out/${tag}_tests.txt
tail -5 gpurun_out/${tag}_tests.txt
show() { python - "$1" <<'PY'
import json, sys
d=json.load(open(sys.argv[1]))
print('value', d['value'], 'e2e', d['e2e'] and d['e2e']['value'], 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'], d['cpu_baseline'] and d['cpu_baseline'].get('parity_vs_gpu_on_sample'))
for k,v in d['legs'].items(): print(' ', k, round(v['ms_per_launch'],3), 'ms', round(v['achieved_GBps']), 'GB/s', round(v['frac_of_hbm_peak'],3))
PY
}
( timeout 600 python bench.py --config formats --steps 5 --warmup 3 > gpurun_out/${tag}_formats.json ) 2> gpurun_out/${tag}_formats.err
tail -3 gpurun_out/${tag}_formats.err; show gpurun_out/${tag}_formats.json
[ -n "$WITH_DIRECT" ] && ( B200_DMZ_FORMATS_DIRECT=1 timeout 600 python bench.py --config formats --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/${tag}_formats_direct.json ) 2>> gpurun_out/${tag}_formats.err
[ -n "$WITH_DIRECT" ] && echo "== direct stores" && show gpurun_out/${tag}_formats_direct.json
( timeout 600 python bench.py --impl reference --config formats --steps 3 --warmup 1 > gpurun_out/${tag}_formats_ref.json ) 2>> gpurun_out/${tag}_formats.err
if [ -z "$NO_NCU" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ycbcr_to_rgb_kernel|rgba_to_r_kernel|stencil3_kernel" -c 7 \
    -o gpurun_out/${tag}_formats_ncu -f python bench.py --config formats --steps 1 --warmup 0 --no-e2e --no-cpu --format-frames 2048 > gpurun_out/${tag}_ncu.log 2>&1
  tail -3 gpurun_out/${tag}_ncu.log
fi
