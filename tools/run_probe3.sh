#!/bin/bash
# does the TS MMA rate depend on where A and D sit in TMEM?
mkdir -p gpurun_out
{
for ad in "384 0" "0 128" "128 0" "256 384" "0 256" "128 384" "256 0" "0 384"; do
  set -- $ad
  timeout 30 tools/umma_probe 1 128 256 0 200 $2 1 4 0 $1 | tail -1
done
} 2>&1 | tee gpurun_out/probe3.log
