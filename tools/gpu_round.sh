#!/bin/bash
# One gpurun call: parity tests, timing, ncu launch list + full profile of the read-once kernel.  Outputs under gpurun_out/.
TAG=${1:-x}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_cuda_parity.py -x -q -k "fused" 2>&1 | tail -15 > gpurun_out/pytest_fused_$TAG.txt
tail -8 gpurun_out/pytest_fused_$TAG.txt
for c in C2 C3 C4; do timeout 120 python tools/time_merge.py --cfg $c --fused 1 2>&1 | tail -1; done | tee gpurun_out/time_merge_$TAG.txt
timeout 120 python tools/time_merge.py --cfg C2 --fused 0 2>&1 | tail -1 | tee -a gpurun_out/time_merge_$TAG.txt
timeout 120 python tools/time_merge.py --cfg C2 --fused 1 --calls 4 2>&1 | tail -1 | tee -a gpurun_out/time_merge_$TAG.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_$TAG.txt
tail -8 gpurun_out/pytest_gpu_$TAG.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$TAG.csv python tools/time_merge.py --cfg C2 --fused 1 --iters 1 > gpurun_out/ncu_launch.log 2>&1
grep -E "k_fused" gpurun_out/launches_$TAG.csv | tail -3 | awk -F'","' '{print $5, $(NF)}'
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused_merge -s 3 -c 1 -o gpurun_out/prof_fused_$TAG python tools/time_merge.py --cfg C2 --fused 1 --iters 2 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
