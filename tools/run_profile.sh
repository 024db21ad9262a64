#!/bin/bash
# ncu launch list + full capture of the fused kernel + in-kernel cycle breakdown (1 GPU).
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_power_cap --format=csv,noheader,nounits 2>&1 | head -3
timeout 120 python tools/prof_breakdown.py 2>&1 | tee gpurun_out/breakdown.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_fused -s 6 -c 2 -o gpurun_out/prof_fused -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/
