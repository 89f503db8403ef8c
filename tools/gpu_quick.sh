#!/bin/bash
# Short gpurun call while tuning the read-once kernel: its parity cases, timing, one full ncu capture.
# Every step has its own short timeout: a kernel that hangs costs one step, not the call.
TAG=${1:-x}
mkdir -p gpurun_out
timeout 60 python tools/time_merge.py --cfg C2 --fused 1 2>&1 | tail -1 | tee gpurun_out/time_merge_$TAG.txt
if ! grep -q "ff_merge_layer" gpurun_out/time_merge_$TAG.txt; then echo "C2 did not finish: stopping"; exit 1; fi
for bl in "576 2304" "1024 2048" "1152 3456" "2048 8192" "2304 4608" "4608 9216" "2048 1000000"; do set -- $bl; echo -n "band=$1 lag=$2 "; FF_FUSED_BAND=$1 FF_FUSED_LAG=$2 timeout 60 python tools/time_merge.py --cfg C2 --fused 1 2>&1 | tail -1; done | tee -a gpurun_out/time_merge_$TAG.txt
timeout 300 python -m pytest tests/test_cuda_parity.py tests/test_cuda_large.py tests/test_hooks_gpu.py -x -q -k "fused or single_pass" 2>&1 | tail -4 | tee gpurun_out/pytest_fused_$TAG.txt
for c in C3 C4; do timeout 60 python tools/time_merge.py --cfg $c --fused 1 2>&1 | tail -1; done | tee -a gpurun_out/time_merge_$TAG.txt
timeout 60 python tools/time_merge.py --cfg C2 --fused 1 --calls 4 2>&1 | tail -1 | tee -a gpurun_out/time_merge_$TAG.txt
timeout 60 python tools/time_merge.py --cfg C2 --fused 0 2>&1 | tail -1 | tee -a gpurun_out/time_merge_$TAG.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_fused_merge -s 3 -c 1 -o gpurun_out/prof_fused_$TAG python tools/time_merge.py --cfg C2 --fused 1 --iters 2 > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
