#!/bin/bash
# does mbarrier polling by other warps slow the MMA stream?
mkdir -p gpurun_out
{
for pollers in 0 4 8 16; do
  timeout 30 tools/umma_probe 1 128 256 0 2000 0 1 4 0 384 0 $pollers | grep "^probe" | sed -e "s/max_abs_err.*cycles/pollers=$pollers cycles/"
done
} 2>&1 | tee gpurun_out/probe5.log
