ncu --set full --clock-control none --import-source on -k regex:poa_kernel -s 0 -c 1 -o gpurun_out/prof -f \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-compute-roofline --stream data/captured/ecoli5mb_ctg1.inspect.gz data/captured/ecoli5mb_ctg2.inspect.gz > gpurun_out/prof_ecoli.log 2>&1
tail -2 gpurun_out/prof_ecoli.log
