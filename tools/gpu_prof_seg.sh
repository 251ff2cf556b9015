ncu --set full --clock-control none --import-source on -k regex:"expiry_groups|expiry_scharr|expiry_stripes|expiry_colsum" -s 4 -c 4 -o gpurun_out/r03e_seg_prof -f env SIDE_BENCH_CARDS=8192 python tools/gpu_side_bench.py 4096 > /dev/null 2> gpurun_out/r03e_seg_prof.err
ls -la gpurun_out/r03e_seg_prof.ncu-rep
