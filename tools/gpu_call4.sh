#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/c4_timeline.txt
for pair in 0 1; do for k in fwd wgrad; do
  echo "=== 3xTF32 pair=$pair $k" >> gpurun_out/c4_timeline.txt
  MVAE_PAIR=$pair timeout 120 python tools/timeline.py 1 $k >> gpurun_out/c4_timeline.txt 2>&1
done; done
cat gpurun_out/c4_timeline.txt
