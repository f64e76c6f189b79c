#!/bin/bash
# round 2, GPU session 5: final 1-D build -- full GPU suite, default bench line, DRAM traffic at the bench batch, launch list, ncu capture
set -u
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -q -m gpu --maxfail=8 > $O/r2_s5_pytest.log 2>&1
echo "pytest exit $?" >> $O/r2_s5_pytest.log
tail -4 $O/r2_s5_pytest.log
timeout 600 python __graft_entry__.py smoke > $O/r2_s5_smoke.log 2>&1; echo "smoke exit $?" >> $O/r2_s5_smoke.log; tail -7 $O/r2_s5_smoke.log
timeout 1500 python bench.py > $O/r2_bench_line_v2.json 2> $O/r2_s5_bench.err; echo "bench exit $?"; tail -3 $O/r2_s5_bench.err; head -c 3000 $O/r2_bench_line_v2.json
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"filter1d_kernel|nan_fill_kernel" -c 17 --csv --log-file $O/r2_traffic_bench_batch.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --no-secondary > $O/r2_s5_ncu_traffic.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary > $O/r2_s5_ncu_bench.log 2>&1
timeout 600 python tools/profile_case.py 8 303104 100 raw full > $O/r2_s5_profile_case.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter1d_kernel -c 1 -o $O/r2_filter1d_N8_v9 -f python tools/profile_case.py 8 303104 100 raw full > $O/r2_s5_ncu1.log 2>&1
tail -2 $O/r2_s5_profile_case.log
timeout 600 python tools/time_profile.py > $O/r2_time_profile_N_sweep.md 2> $O/r2_s5_tp.err; tail -20 $O/r2_time_profile_N_sweep.md
