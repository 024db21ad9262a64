#!/bin/bash
# issue-rate matrix: TS N=128 K=256, commit cadence / accumulator alternation / whole-chip
mkdir -p gpurun_out
{
for ce in 0 4 1; do
  for alt in 0 1; do
    for grid in 1 148; do
      timeout 30 tools/umma_probe 1 128 256 0 200 0 $grid $ce $alt | tail -1
    done
  done
done
timeout 30 tools/umma_probe 0 128 256 0 200 0 148 4 0 | tail -1
timeout 30 tools/umma_probe 1 256 256 0 200 0 148 4 0 | tail -1
} 2>&1 | tee gpurun_out/probe2.log
