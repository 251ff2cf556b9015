#!/bin/bash
# One ncu pass over a short device-resident run (4096 frames per step): the launch list with per-launch device times, and a
# full-set capture of the hot kernels with source correlation.  Outputs under gpurun_out/ (copy summaries to profiles/).
# Usage (GPU box): bash tools/gpu_profile.sh <tag>
tag=${1:-r02}
CMD="python bench.py --frames 4096 --steps 1 --warmup 1 --no-e2e --no-cpu --card-mode ${CARD_MODE:-lazy}"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > /dev/null 2> gpurun_out/${tag}_launches.err
ncu --set full --clock-control none --import-source on -k regex:"detect_strips|warp_rows|vseg_rows|categorize_kernel|hseg_kernel|digit_prep" -s 9 -c 9 -o gpurun_out/${tag}_prof -f $CMD > /dev/null 2> gpurun_out/${tag}_prof.err
ls -la gpurun_out/${tag}_prof.ncu-rep
