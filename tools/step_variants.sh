#!/bin/bash
# whole bench step (ms_per_step, kernel_us of call #0, roofline fraction) with the default library and every variant
for l in "" framefusion_b200/variants/libff_*.so; do echo -n "step [$l]: "; FF_LIB_PATH=${l:+$PWD/$l} timeout 200 python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_us'], d['roofline']['frac'])"; done
