#!/bin/bash
# Short gpurun call while tuning the read-once kernel: its parity cases, timing, one full ncu capture.
# Every step has its own short timeout: a kernel that hangs costs one step, not the call.
TAG=${1:-x}
mkdir -p gpurun_out
timeout 60 python tools/time_merge.py --cfg C2 --fused 1 2>&1 | tail -1 | tee gpurun_out/time_merge_$TAG.txt
if ! grep -q "ff_merge_layer" gpurun_out/time_merge_$TAG.txt; then echo "C2 did not finish: stopping"; exit 1; fi
for bl in $SWEEP; do b=${bl%:*}; l=${bl#*:}; echo -n "band=$b lag=$l "; FF_FUSED_BAND=$b FF_FUSED_LAG=$l timeout 60 python tools/time_merge.py --cfg C2 --fused 1 2>&1 | tail -1; done | tee -a gpurun_out/time_merge_$TAG.txt
timeout 300 python -m pytest tests/test_cuda_parity.py tests/test_cuda_large.py tests/test_hooks_gpu.py -x -q -k "fused or single_pass" 2>&1 | tail -4 | tee gpurun_out/pytest_fused_$TAG.txt
for c in C3 C4; do timeout 60 python tools/time_merge.py --cfg $c --fused 1 2>&1 | tail -1; done | tee -a gpurun_out/time_merge_$TAG.txt
timeout 60 python tools/time_merge.py --cfg C2 --fused 1 --calls 4 2>&1 | tail -1 | tee -a gpurun_out/time_merge_$TAG.txt
timeout 60 python tools/time_merge.py --cfg C2 --fused 0 2>&1 | tail -1 | tee -a gpurun_out/time_merge_$TAG.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_fused_merge -s 3 -c 1 -o gpurun_out/prof_fused_$TAG python tools/time_merge.py --cfg C2 --fused 1 --iters 2 > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
if [ -n "$NCU2" ]; then
FF_FUSED_LAG=1000000 timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_fused_merge -s 3 -c 1 -o gpurun_out/prof_fused_${TAG}_2phase python tools/time_merge.py --cfg C2 --fused 1 --iters 2 > gpurun_out/ncu_full2.log 2>&1
tail -1 gpurun_out/ncu_full2.log
fi
