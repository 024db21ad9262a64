#!/bin/bash
# is the MMA rate data/power dependent?  random mantissas, 1 SM vs whole chip, long run
mkdir -p gpurun_out
{
for rnd in 0 1; do
  for grid in 1 148; do
    timeout 30 tools/umma_probe 1 128 256 0 2000 0 $grid 4 0 384 $rnd | grep "^probe" | sed -e "s/max_abs_err.*cycles/cycles/"
  done
done
} 2>&1 | tee gpurun_out/probe4.log
