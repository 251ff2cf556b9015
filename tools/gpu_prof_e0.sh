ncu --set full --clock-control none --import-source on -k regex:"expiry_mma" -s 1 -c 1 -o gpurun_out/r02y_e0_prof -f env SIDE_BENCH_CARDS=256 python tools/gpu_side_bench.py 65536 > /dev/null 2> gpurun_out/r02y_e0_prof.err
ls -la gpurun_out/r02y_e0_prof.ncu-rep
