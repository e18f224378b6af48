bash tools/sanitize.sh > /dev/null 2>&1
python bench.py > gpurun_out/r1_bench_1gpu.json 2>gpurun_out/bench.err
python bench.py --mix pipeline --windows 4000000 --steps 3 --warmup 2 --cpu-sample 100000 > gpurun_out/r1_bench_pipeline_mix.json 2>/dev/null
timeout 900 python tools/sweep.py --budget 1e11 --long > gpurun_out/r1_sweep.jsonl 2> gpurun_out/sweep.err
python tools/sweep_table.py gpurun_out/r1_sweep.jsonl > gpurun_out/r1_sweep.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r1_launches.log 2>&1
tail -c 200 gpurun_out/r1_bench_1gpu.json; grep -c True gpurun_out/r1_sweep.txt; grep -E "SUMMARY" gpurun_out/sanitizer_*.log
