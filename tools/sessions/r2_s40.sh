#!/bin/bash
# round 2, GPU session 40: gradient kernel occupancy re-check (2 CTAs x 128 threads @250 registers = shipped, 3 @168, 4 @128)
set -u
O=gpurun_out
for rep in 1 2; do
  echo "--- [shipped: 2 CTAs @250]" >> $O/r2_s40_grad.log; timeout 300 python tools/grad_probe.py 2>&1 | grep "value + 2" >> $O/r2_s40_grad.log
  for v in g3 g4; do echo "--- [$v]" >> $O/r2_s40_grad.log; MFS_B200_LIB=$PWD/ab/libmfs_$v.so timeout 300 python tools/grad_probe.py 2>&1 | grep "value + 2" >> $O/r2_s40_grad.log; done
done
cat $O/r2_s40_grad.log
