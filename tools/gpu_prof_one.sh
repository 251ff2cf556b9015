#!/bin/bash
# ncu --set full of one kernel (regex $1) during a short pipeline bench; report -> gpurun_out/$2.ncu-rep
k=$1; tag=$2; shift 2
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$k" -s ${SKIP:-2} -c ${COUNT:-1} -o gpurun_out/${tag} -f python bench.py --frames 4096 --steps 1 --warmup 1 --no-e2e --no-cpu --no-materialised "$@" > /dev/null 2> gpurun_out/${tag}.err
ls -la gpurun_out/${tag}.ncu-rep; tail -2 gpurun_out/${tag}.err
