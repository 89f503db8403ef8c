#!/bin/bash
# One gpurun call: parity tests, probes, timing, ncu launch list.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt | tail -15
timeout 120 ./tools/microbench > gpurun_out/microbench.txt 2>&1; cat gpurun_out/microbench.txt
timeout 120 python tools/probe_topk.py > gpurun_out/probe_topk.txt 2>&1; cat gpurun_out/probe_topk.txt
for f in 0 1; do timeout 300 python tools/time_merge.py --cfg C2 --fused $f 2>&1 | tail -3; done | tee gpurun_out/time_merge.txt
timeout 300 python tools/time_merge.py --cfg C3 --fused 0 2>&1 | tail -1 | tee -a gpurun_out/time_merge.txt
timeout 300 python tools/time_merge.py --cfg C2 --fused 0 --calls 4 2>&1 | tail -1 | tee -a gpurun_out/time_merge.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python tools/time_merge.py --cfg C2 --fused 0 --iters 1 > gpurun_out/ncu_launch.log 2>&1
tail -30 gpurun_out/launches.csv | cut -c1-250
