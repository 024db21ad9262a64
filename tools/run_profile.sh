#!/bin/bash
# GPU tests + bench + ncu launch list + one full capture of the fused fine-pass kernel (1 GPU).
# usage: tools/run_profile.sh <tag>
TAG=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv,noheader,nounits 2>&1 | head -1
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/gpu_tests_$TAG.log; tail -3 gpurun_out/gpu_tests_$TAG.log
timeout 300 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json
timeout 120 python tools/prof_breakdown.py 2>&1 | tee gpurun_out/breakdown_$TAG.txt | tail -8
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_fused -s 7 -c 1 -o gpurun_out/prof_fused_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out/ | tail -12
