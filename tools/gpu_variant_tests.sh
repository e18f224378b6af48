#!/bin/bash
# Developer aid: run the GPU parity tests against build/variants/<name>.so (swapped in for the library).
name=$1; shift
cp hypo_b200/libhypo_b200.so /tmp/orig.so
cp build/variants/$name.so hypo_b200/libhypo_b200.so
python -m pytest tests/test_gpu_parity.py tests/test_gpu_tiers.py tests/test_gpu_host.py tests/test_window_stream.py -m gpu -x -q "$@" 2>&1 | tail -5
cp /tmp/orig.so hypo_b200/libhypo_b200.so
