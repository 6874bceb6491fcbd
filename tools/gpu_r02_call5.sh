#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c5
( cd tools/micro && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 im2col_probe.cu -o im2col_probe -lcuda 2>&1 | grep -v Wno-dep | tail -2; timeout 120 ./im2col_probe ) > ${O}_im2col_probe.txt 2>&1
cat ${O}_im2col_probe.txt
timeout 300 python tools/profile_conv_small.py 4096 > ${O}_conv_small_times.txt 2>&1; cat ${O}_conv_small_times.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_cin|convT_cout" -s 12 -c 4 -o ${O}_conv_small -f python tools/profile_conv_small.py 4096 > ${O}_ncu_conv.log 2>&1
ncu -i ${O}_conv_small.ncu-rep --page details --csv > ${O}_conv_small_details.csv 2>/dev/null
python - <<'PY'
import csv, collections
rows = list(csv.DictReader(open("gpurun_out/r2c5_conv_small_details.csv")))
want = ("Duration", "DRAM Throughput", "Memory Throughput", "Compute (SM) Throughput", "Registers Per Thread", "Achieved Occupancy", "Theoretical Occupancy",
        "Executed Ipc Active", "Issue Slots Busy", "L1/TEX Hit Rate", "L2 Hit Rate", "No Eligible", "Block Limit Registers", "Block Limit Shared Mem",
        "Mem Busy", "Max Bandwidth", "Shared Memory Configuration Size", "Dynamic Shared Memory Per Block", "Warp Cycles Per Issued Instruction", "Executed Instructions")
by = collections.OrderedDict()
for r in rows:
    k = (r["ID"], r["Kernel Name"][:40])
    if r["Metric Name"] in want:
        by.setdefault(k, []).append(f'{r["Metric Name"]}={r["Metric Value"]}{r["Metric Unit"]}')
for k, v in by.items():
    print(k, "; ".join(v))
PY
( time timeout 900 python -m pytest tests/test_gemm_chain_gpu.py tests/test_golden_kats_gpu.py tests/test_mnist_step_gpu.py tests/test_conv_small_gpu.py tests/test_kernels_gpu.py -m gpu -q --timeout 600 ) > ${O}_pytest.log 2>&1; tail -5 ${O}_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > ${O}_bench.json 2> ${O}_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c5_bench.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, d["roofline"]["frac"], d["kernel_breakdown_ms"])
PY
