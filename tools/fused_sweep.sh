#!/bin/bash
# Timing of the read-once kernel over its knobs: "sframes:ctas:lag:prefetch" tuples in $SWEEP; DRAM bytes for those in $NCU
mkdir -p gpurun_out
for t in $SWEEP; do IFS=: read f c l pf <<< "$t"
  echo -n "sframes=$f ctas=$c lag=$l pf=$pf: "
  FF_FUSED_SFRAMES=$f FF_FUSED_CTAS=$c FF_FUSED_LAG=$l timeout 60 python tools/time_merge.py --cfg ${CFG:-C2} --fused 1 2>&1 | tail -1 | sed 's/|.*//'
done | tee gpurun_out/fused_sweep.txt
for t in $NCU; do IFS=: read f c l pf <<< "$t"
  echo -n "sframes=$f ctas=$c lag=$l pf=$pf: "
  FF_FUSED_SFRAMES=$f FF_FUSED_CTAS=$c FF_FUSED_LAG=$l timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_fused_merge -s 3 -c 1 --csv python tools/time_merge.py --cfg ${CFG:-C2} --fused 1 --iters 2 2>/dev/null | grep k_fused | awk -F'","' '{printf "%s=%s ", $(NF-2), $NF}' ; echo
done | tee -a gpurun_out/fused_sweep.txt
