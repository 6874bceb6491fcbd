#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_gemm_chain_gpu.py tests/test_mnist_step_gpu.py -m gpu -q --timeout 120 -x ) > gpurun_out/ab_pytest_base.log 2>&1
echo "base rc=$?"; tail -2 gpurun_out/ab_pytest_base.log
( MVAE_PAIR=1 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_gemm_chain_gpu.py tests/test_mnist_step_gpu.py -m gpu -q --timeout 120 -x ) > gpurun_out/ab_pytest_pair.log 2>&1
echo "pair rc=$?"; tail -2 gpurun_out/ab_pytest_pair.log
python tools/tile_gantt.py 4096 > gpurun_out/ab_gantt.txt 2>&1; cat gpurun_out/ab_gantt.txt
python tools/chain_times.py 4096
for pair in 0 1; do
MVAE_PAIR=$pair timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/ab_bench_pair$pair.json 2> gpurun_out/ab_bench_pair$pair.err
tail -2 gpurun_out/ab_bench_pair$pair.err
done
for f in gpurun_out/ab_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["clocks"])
except Exception as e: print("ERR", e)
PY
done
