#!/bin/bash
# round 2, GPU session 4: fused update+predict (raw Benes) -- parity, same-box A/B, full-size bench, ncu launch list + capture
set -u
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -q -m gpu --maxfail=8 > $O/r2_s4_pytest.log 2>&1
echo "pytest exit $?" >> $O/r2_s4_pytest.log
tail -8 $O/r2_s4_pytest.log
MFS_B200_LIB=$PWD/ab/libmfs_nofuse.so timeout 600 python tools/ab_cases.py nofuse --quick > $O/r2_s4_ab.log 2>&1
timeout 600 python tools/ab_cases.py fused >> $O/r2_s4_ab.log 2>&1
MFS_B200_LIB=$PWD/ab/libmfs_nofuse.so timeout 600 python tools/ab_cases.py nofuse --quick >> $O/r2_s4_ab.log 2>&1
cat $O/r2_s4_ab.log
timeout 600 python tools/exactness_report.py > $O/r2_exactness_report.md 2> $O/r2_s4_exact.err; tail -16 $O/r2_exactness_report.md
timeout 1500 python bench.py > $O/r2_bench_line_v1.json 2> $O/r2_s4_bench.err; echo "bench exit $?"; tail -3 $O/r2_s4_bench.err; head -c 12000 $O/r2_bench_line_v1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_reference_arm.json 2>> $O/r2_s4_bench.err; cat $O/r2_bench_reference_arm.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary > $O/r2_s4_ncu_bench.log 2>&1
timeout 600 python tools/profile_case.py 8 303104 100 raw full > $O/r2_s4_profile_case.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter1d_kernel -c 1 -o $O/r2_filter1d_N8_v8 -f python tools/profile_case.py 8 303104 100 raw full > $O/r2_s4_ncu1.log 2>&1
tail -2 $O/r2_s4_profile_case.log
