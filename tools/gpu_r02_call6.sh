#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c6
( timeout 600 python -m pytest tests/test_conv_small_gpu.py tests/test_mnist_step_gpu.py tests/test_fashion_step_gpu.py -m gpu -q --timeout 600 -x ) > ${O}_pytest.log 2>&1; tail -5 ${O}_pytest.log
timeout 300 python tools/profile_conv_small.py 4096 > ${O}_conv_small_times.txt 2>&1; cat ${O}_conv_small_times.txt
timeout 300 python tools/profile_conv_small.py 512 > ${O}_conv_small_times_512.txt 2>&1; cat ${O}_conv_small_times_512.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_cin|convT_cout" -s 12 -c 4 -o ${O}_conv_small -f python tools/profile_conv_small.py 4096 > ${O}_ncu_conv.log 2>&1
ncu -i ${O}_conv_small.ncu-rep --page details --csv > ${O}_conv_small_details.csv 2>/dev/null
python - <<'PY'
import csv, collections
rows = list(csv.DictReader(open("gpurun_out/r2c6_conv_small_details.csv")))
want = ("Duration", "DRAM Throughput", "Memory Throughput", "Compute (SM) Throughput", "Registers Per Thread", "Achieved Occupancy",
        "Executed Ipc Active", "Issue Slots Busy", "L1/TEX Hit Rate", "No Eligible", "Mem Busy", "Executed Instructions", "Mem Pipes Busy")
by = collections.OrderedDict()
for r in rows:
    k = (r["ID"], r["Kernel Name"][:40])
    if r["Metric Name"] in want:
        by.setdefault(k, []).append(f'{r["Metric Name"]}={r["Metric Value"]}{r["Metric Unit"]}')
for k, v in by.items():
    print(k, "; ".join(v))
PY
timeout 300 python bench.py --workload fashion --steps 20 --warmup 5 --no-cpu-baseline > ${O}_bench_fashion.json 2> ${O}_bench_fashion.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c6_bench_fashion.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, d["roofline"]["frac"], d["kernel_breakdown_ms"])
PY
