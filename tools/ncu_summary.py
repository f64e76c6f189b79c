"""Summarise ncu output for profiles/: `ncu_summary.py rep <file.ncu-rep> <title>` prints the metric table of one
capture (read with `ncu -i ... --page raw --csv`), `ncu_summary.py launches <launches.csv> <title>` aggregates a
`--metrics gpu__time_duration.sum` launch list per kernel."""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEEP = ('Kernel Name', 'Block Size', 'Grid Size', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__time_duration.sum',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__registers_per_thread',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum',
        'lts__t_sectors_srcunit_tex_op_write.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio')
STALL = 'smsp__average_warps_issue_stalled_'


def rep(path, title):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    names, units, vals = rows[0], rows[1], rows[2]
    print(f'# {title}\n')
    print(f'source: `{path}` (ncu --set full --clock-control none --import-source on), read with `ncu -i ... --page raw --csv`\n')
    print('| metric | unit | value |\n|---|---|---|')
    for n, u, v in zip(names, units, vals):
        if n in KEEP or (n.startswith(STALL) and n.endswith('_per_issue_active.ratio')):
            print(f'| {n} | {u} | {v} |')


def launches(path, title):
    txt = open(path).read()
    start = txt.index('"ID"')
    agg = OrderedDict()
    for r in csv.DictReader(io.StringIO(txt[start:])):
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit']
        ms = v * {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'nsecond': 1e-6, 'ms': 1., 'msecond': 1., 'second': 1e3, 's': 1e3}[unit]
        k = r['Kernel Name'][:80]
        a = agg.setdefault(k, [0, 0.])
        a[0] += 1
        a[1] += ms
    total = sum(a[1] for a in agg.values())
    print(f'# {title}\n')
    print(f'source: `{path}`. Per-launch times are cold-cache and serialised; what must agree with the bench is the SHARE.\n')
    print(f'Total device time of all {sum(a[0] for a in agg.values())} launches: {total:.1f} ms.\n')
    print('| kernel | launches | total ms | mean ms | share |\n|---|---|---|---|---|')
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f'| `{k}` | {n} | {ms:.2f} | {ms / n:.3f} | {100 * ms / total:.2f} % |')


if __name__ == '__main__':
    {'rep': rep, 'launches': launches}[sys.argv[1]](sys.argv[2], sys.argv[3])
