#!/bin/bash
# Short gpurun call while tuning the frame-pipelined kernel: parity, timing against the multi-kernel path, stamps, ncu.
# Every step has its own short timeout: a kernel that hangs costs one step, not the call.
TAG=${1:-x}
mkdir -p gpurun_out
timeout 120 python tools/time_merge.py --cfg C2 2>&1 | tail -3 | tee gpurun_out/frame_time_$TAG.txt
if ! grep -q "ff_merge_layer" gpurun_out/frame_time_$TAG.txt; then echo "C2 did not finish: stopping"; exit 1; fi
timeout 400 python -m pytest tests/test_cuda_parity.py -x -q 2>&1 | tail -6 | tee gpurun_out/frame_pytest_$TAG.txt
FF_NO_FRAME=1 timeout 60 python tools/time_merge.py --cfg C2 2>&1 | tail -1 | tee -a gpurun_out/frame_time_$TAG.txt
for c in C3 C4; do timeout 60 python tools/time_merge.py --cfg $c 2>&1 | tail -1; done | tee -a gpurun_out/frame_time_$TAG.txt
for s in $STAGES; do echo -n "stages=$s "; FF_FRAME_STAGES=$s timeout 60 python tools/time_merge.py --cfg C2 2>&1 | tail -1; done | tee -a gpurun_out/frame_time_$TAG.txt
FF_LIB_PATH=$PWD/framefusion_b200/variants/libff_trace.so timeout 120 python tools/frame_trace.py --cfg C2 2>&1 | tail -40 | tee gpurun_out/frame_trace_$TAG.txt
if [ -n "$NCU" ]; then
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_frame_merge -s 3 -c 1 -o gpurun_out/prof_frame_$TAG python tools/time_merge.py --cfg C2 --iters 2 > gpurun_out/ncu_frame.log 2>&1
tail -1 gpurun_out/ncu_frame.log
fi
