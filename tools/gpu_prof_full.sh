R=r2
ncu --set full --clock-control none --import-source on -k regex:poa_kernel -s 1 -c 1 -o gpurun_out/${R}_prof -f \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-compute-roofline > gpurun_out/${R}_ncu.log 2>&1
ncu -i gpurun_out/${R}_prof.ncu-rep --page details > gpurun_out/${R}_poa_kernel_Tc_1M_details.txt 2>/dev/null
ncu -i gpurun_out/${R}_prof.ncu-rep --page raw --csv > gpurun_out/${R}_raw.csv 2>/dev/null
