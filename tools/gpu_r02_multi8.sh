#!/bin/bash
# 8-GPU validation: the world-8 protocol tests, then bench.py at 8 GPUs with the fused peer-memory exchange (default) and NCCL
mkdir -p gpurun_out
O=gpurun_out/r2m8
nvidia-smi topo -m > ${O}_topo.txt 2>&1
( time timeout 600 python -m pytest tests/test_dp_p2p_gpu.py -m gpu -q --timeout 500 -rs -k "8 or mismatch" ) > ${O}_pytest.log 2>&1; tail -8 ${O}_pytest.log
for mode in p2p nccl; do
  MVAE_DP=$mode timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 50 --warmup 5 > ${O}_bench_${mode}.json 2> ${O}_bench_${mode}.err
  echo "bench $mode rc=$?"; tail -2 ${O}_bench_${mode}.err
done
python - <<'PY'
import json
for mode in ("p2p", "nccl"):
    f = f"gpurun_out/r2m8_bench_{mode}.json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(mode, "N=", d["n_gpus"], "mnist", round(d["value"]), round(d["ms_per_step"] * 1e3, 1), "us; e2e", round(d["e2e"]["value"]), d["config"]["exchange"][:40], "loss", d["config"]["loss_last"])
        for k, v in d.get("extra", {}).items():
            print("   ", k, round(v["value"]), round(v["ms_per_step"] * 1e3, 1), "us; e2e", round(v["e2e"]["value"]), "loss", v["loss_last"])
        print("    breakdown", d["kernel_breakdown_ms"])
    except Exception as e:
        print(mode, "ERR", e)
PY
