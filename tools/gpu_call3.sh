#!/bin/bash
mkdir -p gpurun_out
# baseline path still green?
( timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_gemm_chain_gpu.py tests/test_mnist_step_gpu.py -m gpu -q --timeout 120 -x ) > gpurun_out/c3_pytest_base.log 2>&1
echo "base rc=$?"; tail -3 gpurun_out/c3_pytest_base.log
# pair kernel
( MVAE_PAIR=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 -x -k "linear" ) > gpurun_out/c3_pytest_pair_k.log 2>&1
echo "pair kernels rc=$?"; tail -15 gpurun_out/c3_pytest_pair_k.log
( MVAE_PAIR=1 timeout 600 python -m pytest tests/test_gemm_chain_gpu.py tests/test_mnist_step_gpu.py -m gpu -q --timeout 120 -x ) > gpurun_out/c3_pytest_pair_c.log 2>&1
echo "pair chain rc=$?"; tail -15 gpurun_out/c3_pytest_pair_c.log
rm -f gpurun_out/c3_timeline.txt
for k in fwd dgrad wgrad; do
  echo "=== 3xTF32 pair $k" >> gpurun_out/c3_timeline.txt
  MVAE_PAIR=1 timeout 120 python tools/timeline.py 1 $k >> gpurun_out/c3_timeline.txt 2>&1
done
cat gpurun_out/c3_timeline.txt
MVAE_PAIR=1 timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/c3_bench_pair.json 2> gpurun_out/c3_bench_pair.err
tail -3 gpurun_out/c3_bench_pair.err
for f in gpurun_out/c3_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["kernel_breakdown_ms"])
except Exception as e: print("ERR", e)
PY
done
