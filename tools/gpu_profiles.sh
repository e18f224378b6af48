#!/bin/bash
# Produces the round's judged evidence under gpurun_out/ (copied into profiles/ afterwards):
#   bench line (default command), ncu launch list of the same command, one --set full capture of the
#   dominant kernel at the full 1M-window size (DRAM traffic per launch) with the details page.
R=${1:-r1}
mkdir -p gpurun_out
python bench.py > gpurun_out/${R}_bench_1gpu.json 2> gpurun_out/${R}_bench.err
tail -c 600 gpurun_out/${R}_bench_1gpu.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${R}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:poa_kernel -s 0 -c 1 -o gpurun_out/${R}_prof -f \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/${R}_ncu.log 2>&1
ncu -i gpurun_out/${R}_prof.ncu-rep --page details > gpurun_out/${R}_poa_kernel_Tc_1M_details.txt 2>/dev/null
tail -2 gpurun_out/${R}_ncu.log
