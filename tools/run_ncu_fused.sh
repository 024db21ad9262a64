#!/bin/bash
# one full-set ncu capture of the fused fine-pass kernel (1 GPU)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_fused -s 7 -c 1 -o gpurun_out/prof_fused -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/prof_fused.ncu-rep
