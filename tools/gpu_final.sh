#!/bin/bash
mkdir -p gpurun_out
( timeout 200 python -m pytest tests -m gpu -q --timeout 120 ) > gpurun_out/final_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/final_pytest.log; tail -3 gpurun_out/final_pytest.log
timeout 100 python bench.py > gpurun_out/final_bench_mnist.json 2> gpurun_out/final_bench_mnist.err
tail -c 600 gpurun_out/final_bench_mnist.json
timeout 60 python __graft_entry__.py --smoke > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log
