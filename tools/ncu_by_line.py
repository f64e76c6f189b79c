"""Join the SASS page of an ncu report (per-instruction execution counts, stall samples, local-memory sectors) with the
line table of the matching cubin (nvdisasm -g), by instruction index, and print dynamic totals per source line.
usage: python tools/ncu_by_line.py report.ncu-rep kernel.cubin [top] [kernel-name-substring]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, cubin = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
want = sys.argv[4] if len(sys.argv) > 4 else ''
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = next(r for r in rows if r and r[0] == 'Address')
body = [r for r in rows if r and r[0].startswith('0x')]
col = {n: i for i, n in enumerate(hdr)}
dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
lines, cur, inside = [], None, True
for line in dis.splitlines():
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', line)
    if m:
        inside = want in m.group(1)
        continue
    if re.match(r'\s*\.section\s', line):
        inside = False
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)', line)
    if m:
        lines.append((cur, m.group(2)))
assert len(lines) == len(body), (len(lines), len(body))
agg = collections.defaultdict(lambda: collections.Counter())
for (src, op), r in zip(lines, body):
    ie = int(r[col['Instructions Executed']])
    a = agg[src]
    a['inst'] += ie
    base = op.split('.')[0]
    if base in ('DFMA', 'DMUL', 'DADD', 'DSETP'):
        a['fp64'] += ie
    if base == 'MUFU':
        a['mufu'] += ie
    if base in ('LDL', 'STL'):
        a['local'] += ie
    a['samples'] += int(r[col['# Samples']])
    a['stall_wait'] += int(r[col['stall_wait']])
    a['stall_long_sb'] += int(r[col['stall_long_sb']])
    a['stall_no_inst'] += int(r[col['stall_no_inst']])
tot = collections.Counter()
for a in agg.values():
    tot.update(a)
print('TOTAL', dict(tot))
for src, a in sorted(agg.items(), key=lambda kv: -kv[1]['inst'])[:top]:
    print(f'{src[0]}:{src[1]:<4d} inst {100 * a["inst"] / tot["inst"]:5.1f}%  fp64 {100 * a["fp64"] / max(1, tot["fp64"]):5.1f}%  '
          f'local {100 * a["local"] / max(1, tot["local"]):5.1f}%  samples {100 * a["samples"] / tot["samples"]:5.1f}%  '
          f'wait {a["stall_wait"]}  long_sb {a["stall_long_sb"]}  no_inst {a["stall_no_inst"]}  (inst {a["inst"]:.3g} fp64 {a["fp64"]:.3g} local {a["local"]:.3g})')
