#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c10
CUDA_LAUNCH_BLOCKING=1 timeout 600 python -m pytest "tests/test_fashion_step_gpu.py" -m gpu -q --timeout 500 -x -k "oracle" > ${O}_pytest_fashion.log 2>&1; tail -30 ${O}_pytest_fashion.log | cut -c1-400
CUDA_LAUNCH_BLOCKING=1 MVAE_FUSED_SPLIT=0 timeout 600 python -m pytest "tests/test_fashion_step_gpu.py" -m gpu -q --timeout 500 -x -k "oracle" > ${O}_pytest_fashion_nosplit.log 2>&1; tail -5 ${O}_pytest_fashion_nosplit.log | cut -c1-400
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest "tests/test_fashion_step_gpu.py::test_fashion_step_matches_oracle_fp64[512]" -m gpu -q --timeout 500 -x > ${O}_memcheck_f512.log 2>&1; grep -m 20 -n "Invalid\|ERROR SUMMARY\|at \|by thread\|Address" ${O}_memcheck_f512.log | head -40
