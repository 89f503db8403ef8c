#!/bin/bash
# times merge call #0 of C2 with every library under framefusion_b200/variants/ (and the default one)
mkdir -p gpurun_out
TAG=${1:-x}
echo -n "default: "; timeout 60 python tools/time_merge.py --cfg C2 2>&1 | tail -1 | cut -c1-90
for l in framefusion_b200/variants/libff_*.so; do echo -n "$(basename $l): "; FF_LIB_PATH=$PWD/$l timeout 60 python tools/time_merge.py --cfg C2 2>&1 | tail -1 | cut -c1-90; done
