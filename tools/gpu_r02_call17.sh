#!/bin/bash
# ncu --set full of the N = 64 sub-pixel GEMM (and its dense / tf32 yardsticks): what paces the k-blocks?
mkdir -p gpurun_out
O=gpurun_out/r2c17
DIAG_ONLY=subpixel_b,view64,dense64,dense128,subpixel_tf32 timeout 600 ncu --set full --clock-control none -k regex:gemm_kernel -c 15 -o ${O}_subpixel -f python tools/diag_subpixel.py 8192 > ${O}_ncu.log 2>&1
tail -3 ${O}_ncu.log
ncu -i ${O}_subpixel.ncu-rep --page raw --csv > ${O}_subpixel_raw.csv 2>/dev/null
ls -la ${O}_*
