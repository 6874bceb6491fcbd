#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c4
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 ) > ${O}_pytest.log 2>&1
echo "pytest rc=$?" >> ${O}_pytest.log; tail -8 ${O}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > ${O}_bench.json 2> ${O}_bench.err; echo "bench rc=$?"; tail -3 ${O}_bench.err
timeout 300 python bench.py --global-batch 512 --steps 100 --no-cpu-baseline > ${O}_bench_b512.json 2> ${O}_bench_b512.err
MVAE_FUSED_SPLIT=0 timeout 300 python bench.py --global-batch 512 --steps 100 --no-cpu-baseline > ${O}_bench_b512_nosplit.json 2> ${O}_bench_b512_nosplit.err
timeout 300 python bench.py --workload fashion --global-batch 512 --steps 50 --no-cpu-baseline > ${O}_bench_f512.json 2> ${O}_bench_f512.err
MVAE_FUSED_SPLIT=0 timeout 300 python bench.py --workload fashion --global-batch 512 --steps 50 --no-cpu-baseline > ${O}_bench_f512_nosplit.json 2> ${O}_bench_f512_nosplit.err
export MVAE_TIMES_MIN_MS=0.003
timeout 200 python tools/gemm_times.py mnist 4096 > ${O}_times_mnist_4096.txt 2>&1
timeout 200 python tools/gemm_times.py fashion 4096 > ${O}_times_fashion_4096.txt 2>&1
timeout 200 python tools/gemm_times.py fashion 512 > ${O}_times_fashion_512.txt 2>&1
( cd tools/micro && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../multimodal_vae_public_b200/csrc tma_issue.cu -o tma_issue -lcuda 2>&1 | tail -2; timeout 120 ./tma_issue ) > ${O}_tma_issue.txt 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/r2c4_bench.json","gpurun_out/r2c4_bench_b512.json","gpurun_out/r2c4_bench_b512_nosplit.json","gpurun_out/r2c4_bench_f512.json","gpurun_out/r2c4_bench_f512_nosplit.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, d.get("e2e",{}).get("value"), d["roofline"]["frac"])
        print("  ", d["kernel_breakdown_ms"])
        for k in ("roofline_hbm_poe_fwd","roofline_hbm_poe_bwd"):
            if k in d: print("  ", k, d[k]["frac"], d[k]["frac_algorithmic"], d[k]["avg_launch_ms"])
        if "extra" in d: print("  extra", {k:(v["value"], v["ms_per_step"], v["kernel_breakdown_ms"]) for k,v in d["extra"].items()})
    except Exception as e: print(f, "ERR", e)
PY
tail -22 ${O}_times_mnist_4096.txt; cat ${O}_times_fashion_4096.txt | tail -45; cat ${O}_tma_issue.txt
