#!/bin/bash
# round 2, GPU session 1: parity of the v2 1-D kernel, same-box A/B of the register / occupancy variants, smoke, quick bench
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/r2_s1_gpu.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_filter1d.py tests/test_gpu_quadrature.py -x -q -m gpu > $O/r2_s1_pytest_1d.log 2>&1
echo "pytest exit $?" >> $O/r2_s1_pytest_1d.log
tail -5 $O/r2_s1_pytest_1d.log
for v in zs4 zs6 zr4 zr5; do
  MFS_B200_LIB=$PWD/ab/libmfs_$v.so timeout 300 python tools/ab_cases.py $v --quick >> $O/r2_s1_ab.log 2>&1
done
timeout 600 python tools/ab_cases.py default >> $O/r2_s1_ab.log 2>&1
cat $O/r2_s1_ab.log
timeout 600 python __graft_entry__.py smoke > $O/r2_s1_smoke.log 2>&1; echo "smoke exit $?" >> $O/r2_s1_smoke.log; tail -8 $O/r2_s1_smoke.log
timeout 900 python bench.py --steps 2 --warmup 3 --secondary-scale 0.1 > $O/r2_s1_bench.json 2> $O/r2_s1_bench.err; echo "bench exit $?"; tail -3 $O/r2_s1_bench.err; head -c 6000 $O/r2_s1_bench.json
