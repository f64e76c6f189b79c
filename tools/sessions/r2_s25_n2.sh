#!/bin/bash
# round 2, 2-GPU session: e2e legs with the ranks bound to their GPU's NUMA-local CPUs (bench.py bind_to_gpu_numa), topology recorded
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/r2_s25_topo.log 2>&1
lscpu | grep -iE "numa|socket|^cpu\(s\)" >> $O/r2_s25_topo.log 2>&1
cat $O/r2_s25_topo.log
NCCL_DEBUG=VERSION timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-secondary --no-cpu > $O/r2_s25_bench_n2.json 2> $O/r2_s25_bench.err
echo "bench exit $?"; tail -3 $O/r2_s25_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_s25_bench_n2.json') if l.startswith('{')][-1])
print('N=2 value', d['value']); print(d['e2e'].get('host_numa_binding'))
for k,v in d['e2e_modes'].items(): print(k, v['value'], v['pcie_gb_per_s_per_gpu'])
PY
timeout 900 python bench.py --no-secondary --no-cpu > $O/r2_s25_bench_n1.json 2>> $O/r2_s25_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_s25_bench_n1.json') if l.startswith('{')][-1])
print('N=1 value', d['value']); print(d['e2e'].get('host_numa_binding'))
for k,v in d['e2e_modes'].items(): print(k, v['value'], v['pcie_gb_per_s_per_gpu'])
PY
