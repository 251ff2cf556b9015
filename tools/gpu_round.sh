#!/bin/bash
# One GPU-box visit: gpu tests, bench (both arms), ncu launch list and full captures of the main kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 ) > gpurun_out/pytest_gpu.log
( timeout 600 python bench.py ${BENCH_ARGS} > gpurun_out/bench.json ) 2> gpurun_out/bench.err
( timeout 300 python bench.py --impl reference > gpurun_out/bench_ref.json ) 2> gpurun_out/bench_ref.err
if [ -z "$SKIP_NCU" ]; then
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --frames 8192 --no-e2e --no-cpu > gpurun_out/ncu_list_bench.json ) 2> gpurun_out/ncu_list.err
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:'detect_strips|warp_kernel|vseg_rows|hseg_kernel|categorize_kernel|finalize' -c 8 \
    -f -o gpurun_out/prof python bench.py --steps 1 --warmup 0 --frames 4096 --no-e2e --no-cpu > gpurun_out/ncu_full_bench.json ) 2> gpurun_out/ncu_full.err
fi
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json | head -c 3000; echo; cat gpurun_out/bench_ref.json | head -c 1500; tail -3 gpurun_out/bench.err
