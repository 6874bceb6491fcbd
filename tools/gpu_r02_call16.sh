#!/bin/bash
# A/B session 16: B_lo by TMA for narrow (N <= 64) problems only (MVAE_PRESPLIT=narrow) on the conv flavours; per-call
# times of the FashionMNIST step at the 8-GPU per-rank batch
mkdir -p gpurun_out
O=gpurun_out/r2c16
timeout 300 python tools/diag_subpixel.py 8192 5 > ${O}_diag.txt 2>&1; grep -v "^  \(epi\|split\)" ${O}_diag.txt | cut -c1-400
( MVAE_PRESPLIT=narrow timeout 900 python -m pytest tests/test_fashion_step_gpu.py tests/test_celeba_step_gpu.py tests/test_celeba19_step_gpu.py tests/test_implicit_conv_gpu.py -m gpu -q -x --timeout 600 ) > ${O}_pytest.log 2>&1
echo "pytest rc=$?" >> ${O}_pytest.log; tail -4 ${O}_pytest.log
for tag in "base:MVAE_PRESPLIT=0" "narrow:MVAE_PRESPLIT=narrow"; do
  name=${tag%%:*}; envs=${tag#*:}
  env $envs timeout 300 python bench.py --workload fashion --steps 30 --warmup 5 --no-cpu-baseline > ${O}_fashion_${name}.json 2> ${O}_fashion_${name}.err
  env $envs timeout 300 python bench.py --workload fashion --global-batch 512 --steps 60 --warmup 5 --no-cpu-baseline > ${O}_f512_${name}.json 2> ${O}_f512_${name}.err
  env $envs timeout 300 python bench.py --workload celeba --steps 10 --warmup 5 --no-cpu-baseline > ${O}_celeba_${name}.json 2> ${O}_celeba_${name}.err
  env $envs timeout 300 python bench.py --workload celeba19 --steps 5 --warmup 3 --no-cpu-baseline > ${O}_celeba19_${name}.json 2> ${O}_celeba19_${name}.err
done
timeout 200 python tools/gemm_times.py fashion 512 > ${O}_times_fashion_512.txt 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c16_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("r2c16_")[1], {k:d.get(k) for k in ("value","ms_per_step")}, round(d["e2e"]["value"]), {k:v for k,v in list(d["kernel_breakdown_ms"].items())[:4]})
    except Exception as e: print(f, "ERR", e)
PY
cut -c1-180 ${O}_times_fashion_512.txt
