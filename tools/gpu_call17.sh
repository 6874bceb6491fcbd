#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_dp_p2p_gpu.py tests/test_dp_gpu.py -m gpu -q --timeout 200 -x ) > gpurun_out/c17_pytest.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/c17_pytest.log | cut -c1-300
for mode in p2p nccl; do
MVAE_DP=$mode timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/c17_bench_n2_$mode.json 2> gpurun_out/c17_bench_n2_$mode.err
tail -2 gpurun_out/c17_bench_n2_$mode.err | cut -c1-300
done
for f in gpurun_out/c17_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches","n_gpus")}, d["e2e"]["value"], d["config"].get("loss_last"))
except Exception as e: print("ERR", e)
PY
done
