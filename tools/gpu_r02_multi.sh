#!/bin/bash
# multi-GPU validation of the data-parallel exchange (run with gpurun --gpus N): protocol tests for every world size the box
# can run, then bench.py at N GPUs with the NCCL exchange and with the fused peer-memory kernel
N=${1:-4}
mkdir -p gpurun_out
O=gpurun_out/r2m${N}
nvidia-smi topo -m > ${O}_topo.txt 2>&1
( time timeout 900 python -m pytest tests/test_dp_p2p_gpu.py tests/test_dp_gpu.py -m gpu -q --timeout 600 -rs ) > ${O}_pytest.log 2>&1; tail -12 ${O}_pytest.log
for mode in nccl auto; do
  MVAE_DP=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 50 --warmup 5 > ${O}_bench_${mode}.json 2> ${O}_bench_${mode}.err
  echo "bench $mode rc=$?"; tail -3 ${O}_bench_${mode}.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > ${O}_bench_reference.json 2> ${O}_bench_reference.err; echo "reference arm rc=$?"; cut -c1-300 ${O}_bench_reference.json
python - "$N" <<'PY'
import json, sys
n = sys.argv[1]
for mode in ("nccl", "auto"):
    f = f"gpurun_out/r2m{n}_bench_{mode}.json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(mode, "N=", d["n_gpus"], "mnist", round(d["value"]), round(d["ms_per_step"] * 1e3, 1), "us; e2e", round(d["e2e"]["value"]), d["config"]["exchange"][:40], "loss", d["config"]["loss_last"])
        for k, v in d.get("extra", {}).items():
            print("   ", k, round(v["value"]), round(v["ms_per_step"] * 1e3, 1), "us; e2e", round(v["e2e"]["value"]), "loss", v["loss_last"])
        print("    breakdown", d["kernel_breakdown_ms"])
    except Exception as e:
        print(mode, "ERR", e)
PY
