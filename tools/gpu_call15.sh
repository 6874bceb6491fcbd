#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_mnist_step_gpu.py tests/test_modules_gpu.py tests/test_dp_gpu.py -m gpu -q --timeout 240 ) > gpurun_out/c15_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/c15_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/c15_bench_n2.json 2> gpurun_out/c15_bench_n2.err
tail -3 gpurun_out/c15_bench_n2.err
timeout 300 python bench.py --gpus 1 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/c15_bench_n1.json 2> gpurun_out/c15_bench_n1.err
timeout 300 python bench.py --gpus 1 --global-batch 2048 --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/c15_bench_n1_b2048.json 2> gpurun_out/c15_bench_n1_b2048.err
for f in gpurun_out/c15_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches","n_gpus")}, d["e2e"]["value"], d.get("kernel_breakdown_ms"))
except Exception as e: print("ERR", e)
PY
done
