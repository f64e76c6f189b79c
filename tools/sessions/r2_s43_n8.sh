#!/bin/bash
# round 2, 8-GPU session with the final library: the bench under torchrun exactly as the driver launches it (all legs)
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/r2_s43_gpus.log
free -g | head -2 >> $O/r2_s43_gpus.log
NCCL_DEBUG=VERSION timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 > $O/r2_bench_line_v8_n8.json 2> $O/r2_s43_bench.err
echo "bench exit $?"; tail -3 $O/r2_s43_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_line_v8_n8.json') if l.startswith('{')][-1])
print('N=8 value', d['value'], d['roofline']['frac'])
s=d['secondary']
print('theta', s['theta_grid']['value'], {k:v for k,v in s['theta_grid'].items() if 'argmin' in k})
print('nd', {k:v['value'] for k,v in s['nd_filter']['transitions'].items()})
print('grid', s['grid_filter']['value'], s['grid_filter']['power_operator']['value'])
for k,v in d['e2e_modes'].items(): print(k, v['value'], v['pcie_gb_per_s_per_gpu'], v['batch_per_gpu'])
PY
