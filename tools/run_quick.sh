#!/bin/bash
# GPU tests + bench + cycle breakdown (no ncu), 1 GPU.  usage: tools/run_quick.sh <tag>
TAG=${1:-q}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/gpu_tests_$TAG.log; tail -3 gpurun_out/gpu_tests_$TAG.log
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; python -c "
import json;d=json.load(open('gpurun_out/bench_$TAG.json'));print('value %.1f M/s  e2e %.1f M/s  frac %.3f  kernel_ms %.4f clocks %s'%(d['value']/1e6,d['e2e']['value']/1e6,d['roofline']['frac'],d['roofline']['kernel_ms'],d['clocks']))"
timeout 120 python tools/prof_breakdown.py 2>&1 | tee gpurun_out/breakdown_$TAG.txt | head -16
