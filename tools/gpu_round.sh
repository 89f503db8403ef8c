#!/bin/bash
# One gpurun call: parity tests, timing, ncu launch list.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.txt
tail -25 gpurun_out/pytest_gpu.txt
for f in 0 1; do timeout 120 python tools/time_merge.py --cfg C2 --fused $f 2>&1 | tail -3; done | tee gpurun_out/time_merge.txt
timeout 120 python tools/time_merge.py --cfg C3 --fused 1 2>&1 | tail -1 | tee -a gpurun_out/time_merge.txt
timeout 120 python tools/time_merge.py --cfg C4 --fused 1 2>&1 | tail -1 | tee -a gpurun_out/time_merge.txt
timeout 120 python tools/time_merge.py --cfg C2 --fused 1 --calls 4 2>&1 | tail -1 | tee -a gpurun_out/time_merge.txt
timeout 120 python tools/time_merge.py --cfg C2 --fused 0 --calls 4 2>&1 | tail -1 | tee -a gpurun_out/time_merge.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_fused.csv python tools/time_merge.py --cfg C2 --fused 1 --iters 1 > gpurun_out/ncu_launch.log 2>&1
grep -E "k_fused|k_links|status" gpurun_out/launches_fused.csv | tail -8 | cut -c60-130,330-
