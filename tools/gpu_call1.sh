#!/bin/bash
# GPU session 1: full GPU test suite, chain on/off benches, small-batch benches, main-loop timelines
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -q --timeout 240 ) > gpurun_out/c1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
tail -5 gpurun_out/c1_pytest.log
timeout 300 python bench.py --steps 200 --warmup 5 > gpurun_out/c1_bench_chain.json 2> gpurun_out/c1_bench_chain.err
MVAE_CHAIN=0 timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/c1_bench_nochain.json 2> gpurun_out/c1_bench_nochain.err
timeout 200 python bench.py --global-batch 512 --steps 400 --warmup 5 --no-cpu-baseline > gpurun_out/c1_bench_b512_chain.json 2> gpurun_out/c1_b512.err
MVAE_CHAIN=0 timeout 200 python bench.py --global-batch 512 --steps 400 --warmup 5 --no-cpu-baseline > gpurun_out/c1_bench_b512_nochain.json 2>> gpurun_out/c1_b512.err
for k in fwd dgrad wgrad; do
  echo "=== 3xTF32 $k" >> gpurun_out/c1_timeline.txt
  timeout 120 python tools/timeline.py 1 $k >> gpurun_out/c1_timeline.txt 2>&1
done
for f in gpurun_out/c1_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["kernel_breakdown_ms"])
except Exception as e: print("ERR", e)
PY
done
