#!/bin/bash
# TMEM drain-bandwidth matrix (tools/tmem_bw_probe.cu), 1 GPU
mkdir -p gpurun_out
{
for mma in 0 1; do
  for warps in 4 8; do
    for shape in 16 32 64 33 100 101; do
      timeout 30 tools/tmem_bw_probe $shape $warps $mma 400
    done
  done
done
} 2>&1 | tee gpurun_out/tmem_probe.log
