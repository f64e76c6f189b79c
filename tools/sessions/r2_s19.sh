#!/bin/bash
# round 2, GPU session 19: 1-D step barrier -- per-instance policy build (tests), N sweep with / without the barrier, gradient kernel with the barrier (64- and 128-thread CTAs)
set -u
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests/test_gpu_filter1d.py tests/test_gpu_bench_configs.py tests/test_gpu_gradients.py -q -m gpu --maxfail=10 --timeout 300 > $O/r2_s19_pytest.log 2>&1
echo "policy-build pytest exit $?"; tail -3 $O/r2_s19_pytest.log
for v in g1 g2; do
  MFS_B200_LIB=$PWD/ab/libmfs_$v.so timeout 900 python -m pytest tests/test_gpu_gradients.py -q -m gpu --maxfail=10 --timeout 300 > $O/r2_s19_pytest_$v.log 2>&1
  echo "$v pytest exit $?"; tail -1 $O/r2_s19_pytest_$v.log
done
for rep in 1 2; do
  echo "--- [policy]" >> $O/r2_s19_grad.log; timeout 300 python tools/grad_probe.py >> $O/r2_s19_grad.log 2>&1
  for v in g1 g2; do echo "--- [$v]" >> $O/r2_s19_grad.log; MFS_B200_LIB=$PWD/ab/libmfs_$v.so timeout 300 python tools/grad_probe.py >> $O/r2_s19_grad.log 2>&1; done
done
cat $O/r2_s19_grad.log
timeout 600 python tools/ab_cases.py policy --sweep >> $O/r2_s19_sweep.log 2>&1
MFS_B200_LIB=$PWD/ab/libmfs_bar1.so timeout 600 python tools/ab_cases.py barrier --sweep >> $O/r2_s19_sweep.log 2>&1
MFS_B200_LIB=$PWD/ab/libmfs_bar0.so timeout 600 python tools/ab_cases.py none --sweep >> $O/r2_s19_sweep.log 2>&1
timeout 600 python tools/ab_cases.py policy --sweep >> $O/r2_s19_sweep.log 2>&1
cat $O/r2_s19_sweep.log
