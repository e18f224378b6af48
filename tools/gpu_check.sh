#!/bin/bash
# Developer aid (run under gpurun): GPU parity tests, one bench line, one ncu source-level capture of the T0 kernel.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json
if [ "$1" = "ncu" ]; then
  ncu --set full --clock-control none --import-source on -k regex:poa_kernel -s 0 -c 1 -o gpurun_out/prof -f \
      python bench.py --steps 1 --warmup 0 --windows 100000 --no-cpu-baseline --no-e2e > gpurun_out/ncu.log 2>&1
  tail -3 gpurun_out/ncu.log
fi
