#!/bin/bash
# round-2 call 1: baseline of the small per-GPU batches on ONE GPU (where does an 8-GPU strong-scaling step spend time?)
# + compute-sanitizer over the hand-rolled synchronisation (chain GEMM, PoE, BCE, one MNIST step)
mkdir -p gpurun_out
O=gpurun_out/r2c1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > ${O}_smi.txt
export MVAE_TIMES_MIN_MS=0.004
for wl in "mnist 512" "mnist 1024" "fashion 512" "fashion 1024" "fashion 4096"; do
  set -- $wl
  timeout 200 python tools/gemm_times.py $1 $2 > ${O}_times_$1_$2.txt 2>&1
done
timeout 200 python bench.py --global-batch 512 --steps 200 --no-cpu-baseline > ${O}_bench_mnist_b512.json 2> ${O}_bench_mnist_b512.err
timeout 200 python bench.py --workload fashion --global-batch 512 --steps 50 --no-cpu-baseline > ${O}_bench_fashion_b512.json 2> ${O}_bench_fashion_b512.err
timeout 200 python bench.py --workload fashion --steps 30 --no-cpu-baseline > ${O}_bench_fashion_b4096.json 2> ${O}_bench_fashion_b4096.err
# sanitizer: memcheck on kernels + chain + a step; racecheck/synccheck on the chain GEMM tests (small sizes)
export MVAE_SANITIZER=1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gemm_chain_gpu.py tests/test_kernels_gpu.py -m gpu -x -q --timeout 800 > ${O}_memcheck.log 2>&1; echo "memcheck rc=$?" >> ${O}_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gemm_chain_gpu.py -m gpu -x -q --timeout 800 > ${O}_racecheck.log 2>&1; echo "racecheck rc=$?" >> ${O}_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_gemm_chain_gpu.py -m gpu -x -q --timeout 500 > ${O}_synccheck.log 2>&1; echo "synccheck rc=$?" >> ${O}_synccheck.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/profile_step.py step --steps 2 --batch 256 > ${O}_memcheck_step.log 2>&1; echo "memcheck step rc=$?" >> ${O}_memcheck_step.log
tail -3 ${O}_memcheck.log ${O}_racecheck.log ${O}_synccheck.log ${O}_memcheck_step.log
tail -12 ${O}_times_fashion_512.txt
