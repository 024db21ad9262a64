#!/bin/bash
# Runs the UMMA probe matrix; each variant in its own process so one bad
# encoding cannot poison the CUDA context of the next.
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
nproc; free -g | head -2
for mode in 0 1 2; do
  for N in 64 128 256; do
    timeout 30 tools/umma_probe $mode $N 128 0 1 0
  done
done
timeout 30 tools/umma_probe 1 128 256 0 1 128     # TS, K=256, D at column offset 128
timeout 30 tools/umma_probe 2 128 256 1 1 256     # mixed, bf16
timeout 30 tools/umma_probe 0 256 64 0 1 0
# issue-rate measurements (reps of K=256 -> 16 MMAs per rep)
for mode in 0 1; do
  for N in 128 256; do
    timeout 30 tools/umma_probe $mode $N 256 0 200 0
  done
done
} 2>&1 | tee gpurun_out/probe.log
