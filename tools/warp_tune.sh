#!/bin/bash
# GPU-side sweep of the warp kernel's launch parameters (tile short side, rows per CTA, gather vs TMA tile), both card modes.
# Usage (on the GPU box): bash tools/warp_tune.sh > gpurun_out/warp_tune.txt
run() {  # label, env...
  label=$1; shift
  out=$(env "$@" python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu --card-mode $MODE 2>/dev/null | tail -1)
  python - "$label" "$MODE" <<PY
import json, sys
d = json.loads('''$out''')
st = d["stages"]
print("%-28s %-5s value %9.0f f/s  step %7.2f ms  warp %6.2f  detect %6.2f  vseg %6.2f  hseg %5.2f  cat %6.2f" % (
    sys.argv[1], sys.argv[2], d["value"], d["ms_per_step"], st.get("warp", {}).get("ms_per_100k_frames", 0), st["detect"]["ms_per_100k_frames"],
    st["vseg"]["ms_per_100k_frames"], st["hseg"]["ms_per_100k_frames"], st["categorize"]["ms_per_100k_frames"]))
PY
}
for MODE in full lazy; do
  run "default" X=1
  run "gather" B200_DMZ_WARP_GATHER=1
  run "tile64" B200_DMZ_WARP_TILE=64
  run "tile128" B200_DMZ_WARP_TILE=128
  run "tile128 rows 60" B200_DMZ_WARP_TILE=128 B200_DMZ_WARP_ROWS=60
  run "tile64 rows 16" B200_DMZ_WARP_TILE=64 B200_DMZ_WARP_ROWS=16
  run "tile128 segs 3" B200_DMZ_WARP_TILE=128 B200_DMZ_WARP_SEGS=3
done
