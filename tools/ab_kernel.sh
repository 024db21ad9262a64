#!/bin/bash
# A/B timing of two builds of the library in one GPU session, alternating.  usage: tools/ab_kernel.sh A.so B.so
for i in 1 2 3; do
  for lib in "$@"; do CRNERF_B200_LIB=$lib timeout 100 python tools/time_kernel.py 40; done
done
