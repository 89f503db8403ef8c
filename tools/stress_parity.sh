#!/bin/bash
# repeats the fused parity cases; stops at the first failure and keeps its report
N=${1:-10}
for i in $(seq 1 $N); do
  timeout 300 python -m pytest tests/test_cuda_parity.py -x -q -k "fused" > gpurun_out/stress_last.txt 2>&1 || { echo "FAILED in round $i"; grep -E "FAILED|Error|differ" gpurun_out/stress_last.txt | head; cp gpurun_out/stress_last.txt gpurun_out/stress_fail_$i.txt; }
done
echo "stress done: $N rounds"; tail -1 gpurun_out/stress_last.txt
