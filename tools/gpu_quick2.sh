#!/bin/bash
# tests + device-resident bench line (stage times), optional env settings as arguments ("A=1 B=2" each)
tag=${TAG:-quick}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/${tag}_tests.txt; tail -4 gpurun_out/${tag}_tests.txt
for cfg in "${@:-_=}"; do
  ( env $cfg timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/${tag}_bench.json ) 2> gpurun_out/${tag}_bench.err
  echo "== $cfg"; python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench.json'))
print('value', round(d['value']), {k:round(v['ms_per_100k_frames'],2) for k,v in d['stages'].items()}, 'materialised', round(d['materialised']['value']) if d.get('materialised') else None, d['materialised']['stages_ms_per_100k_frames'].get('warp') if d.get('materialised') else None)
PY
  tail -2 gpurun_out/${tag}_bench.err
done
