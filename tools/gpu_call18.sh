#!/bin/bash
N=$1
mkdir -p gpurun_out
for mode in p2p nccl; do
MVAE_DP=$mode timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/c18_bench_n${N}_$mode.json 2> gpurun_out/c18_bench_n${N}_$mode.err
tail -2 gpurun_out/c18_bench_n${N}_$mode.err | cut -c1-300
done
for f in gpurun_out/c18_bench_n${N}_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches","n_gpus")}, d["e2e"]["value"], d["config"].get("loss_last"))
except Exception as e: print("ERR", e)
PY
done
