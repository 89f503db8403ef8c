#!/bin/bash
# DRAM bytes, L2 hit rate and duration of the read-once kernel against the lag, for the default library and a variant
# built with $VARIANT (one ncu metrics pass per point; durations under ncu are cold-cache, for comparison only)
mkdir -p gpurun_out
LIBS=""
if [ -n "$VARIANT" ]; then
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 $VARIANT -Xcompiler -fPIC -shared -I include -o gpurun_out/libff_variant.so framefusion_b200/csrc/ff_api.cu 2>/dev/null
  LIBS="gpurun_out/libff_variant.so"
fi
for lib in "" $LIBS; do
for l in $SWEEP; do
  echo -n "lib=${lib:-default} lag=$l: "
  FF_LIB_PATH=$lib FF_FUSED_LAG=$l timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_fused_merge -s 3 -c 1 --csv python tools/time_merge.py --cfg C2 --fused 1 --iters 2 2>/dev/null | grep k_fused | awk -F'","' '{printf "%s=%s ", $(NF-2), $NF}' ; echo
  FF_LIB_PATH=$lib FF_FUSED_LAG=$l timeout 60 python tools/time_merge.py --cfg C2 --fused 1 2>&1 | tail -1
done; done | tee gpurun_out/l2_lag_probe.txt
