#!/bin/bash
# Round-end measurements beyond tools/final_round.sh: bench lines of C3 / C4, the roofline sweep, ncu --set full of the
# frame-pipelined kernel and of the multi-kernel path, the stamps of one traced launch.  Outputs under gpurun_out/.
TAG=${1:-final}
mkdir -p gpurun_out
for c in C3 C4; do timeout 300 python bench.py --workload $c --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_${c}_$TAG.json; cut -c1-300 gpurun_out/bench_${c}_$TAG.json; done
timeout 600 python tools/sweep_roofline.py 2>&1 | tee gpurun_out/sweep_$TAG.jsonl | tail -50 | cut -c1-200
FF_LIB_PATH=$PWD/framefusion_b200/variants/libff_trace.so timeout 120 python tools/frame_trace.py --cfg C2 > gpurun_out/frame_trace_$TAG.txt 2>&1; tail -12 gpurun_out/frame_trace_$TAG.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_frame_merge -s 3 -c 1 -o gpurun_out/prof_frame_$TAG python tools/time_merge.py --cfg C2 --iters 2 > gpurun_out/ncu_frame.log 2>&1; tail -1 gpurun_out/ncu_frame.log
