#!/bin/bash
# Round-2 evidence in one gpurun session (1 GPU): full GPU test suite, smoke(), the bench line of every workload and of the
# reference arm, the ncu launch list of a bench step, `ncu --set full` captures of the dominant GEMM launches and of the fused
# PoE / BCE kernels at roofline size.  Summaries are copied to profiles/ by hand afterwards (gpurun_out/ is scratch).
mkdir -p gpurun_out
O=gpurun_out/ev5
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 ) > ${O}_pytest.log 2>&1
echo "pytest rc=$?" >> ${O}_pytest.log; tail -4 ${O}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; tail -2 ${O}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > ${O}_bench_mnist.json 2> ${O}_bench_mnist.err; tail -2 ${O}_bench_mnist.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > ${O}_bench_reference.json 2> ${O}_bench_reference.err
timeout 400 python bench.py --workload fashion --steps 30 > ${O}_bench_fashion.json 2> ${O}_bench_fashion.err
timeout 400 python bench.py --workload celeba --steps 20 > ${O}_bench_celeba.json 2> ${O}_bench_celeba.err
timeout 400 python bench.py --workload celeba19 --steps 10 > ${O}_bench_celeba19.json 2> ${O}_bench_celeba19.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${O}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph --no-extras > ${O}_ncu_bench.log 2>&1
# (GEMM chain capture: profiles/r02_gemm_chain_raw_key_metrics.txt, unchanged kernels)
# (PoE capture: profiles/r02_poe_v2_raw_key_metrics.txt)
# (BCE capture: profiles/r02_bce_raw_key_metrics.txt, unchanged kernel)

tail -3 ${O}_bench_celeba19.err
for f in ${O}_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d.get("roofline",{}).get("frac"), d.get("cpu_baseline",{}).get("value"), (d.get("gpu_eager_baseline") or {}).get("value"))
except Exception as e: print("ERR", e)
PY
done
