#!/bin/bash
# Round evidence in one GPU call (outputs under gpurun_out/, copied into profiles/ afterwards):
# GPU parity tests, the default bench line, the ncu launch list of the same command, the pipeline-mix
# bench line, the window-shape sweep and a compute-sanitizer memcheck pass.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -2 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/r1_bench_1gpu.json 2>gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r1_launches.log 2>&1
python bench.py --mix pipeline --windows 4000000 --steps 3 --warmup 3 --cpu-sample 100000 > gpurun_out/r1_bench_pipeline_mix.json 2>/dev/null
timeout ${SWEEP_TIMEOUT:-150} python tools/sweep.py --budget ${SWEEP_BUDGET:-6e10} --long > gpurun_out/r1_sweep.jsonl 2> gpurun_out/sweep.err
python tools/sweep_table.py gpurun_out/r1_sweep.jsonl > gpurun_out/r1_sweep.txt
timeout 90 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_run.py > gpurun_out/sanitizer_memcheck.log 2>&1
tail -c 300 gpurun_out/r1_bench_1gpu.json; grep -c True gpurun_out/r1_sweep.txt; grep -E "SUMMARY|sanitize_run" gpurun_out/sanitizer_memcheck.log
