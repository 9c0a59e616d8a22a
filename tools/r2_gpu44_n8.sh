#!/bin/bash
# Round 2 final at N = 8: multirank checks (bands through both exchanges, frame-parallel path; N = 2, 4, 8) and the bench line.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x > gpurun_out/r2_44_pytest_multirank.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_44_pytest_multirank.log
tail -3 gpurun_out/r2_44_pytest_multirank.log
bash tools/r2_gpu39_n8.sh 8
cp gpurun_out/r2_39_n8_bench.jsonl gpurun_out/r2_44_n8_bench.jsonl
