#!/bin/bash
# A/B session 15: PoE fast path with .ftz MUFU forms (tests + roofline-size numbers), diagnosis of the N = 64 sub-pixel GEMMs
mkdir -p gpurun_out
O=gpurun_out/r2c15
( time timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_golden_kats_gpu.py tests/test_mnist_step_gpu.py tests/test_modules_gpu.py -m gpu -q --timeout 600 ) > ${O}_pytest.log 2>&1
echo "pytest rc=$?" >> ${O}_pytest.log; tail -6 ${O}_pytest.log
timeout 300 python tools/diag_subpixel.py 8192 5 > ${O}_diag.txt 2>&1; cat ${O}_diag.txt
MVAE_DBG_EPI=1 timeout 300 python tools/diag_subpixel.py 8192 5 > ${O}_diag_noepi.txt 2>&1; head -12 ${O}_diag_noepi.txt
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-extras > ${O}_mnist.json 2> ${O}_mnist.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c15_mnist.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step")}, d["kernel_breakdown_ms"])
for k in ("roofline_hbm_poe_fwd","roofline_hbm_poe_bwd"): print(k, {q:d[k][q] for q in ("avg_launch_ms","frac","frac_algorithmic")})
PY
