"""Static SASS instruction counts per source line (needs -lineinfo): which source lines the instructions of a kernel
come from, split by opcode class.  usage: python tools/sass_lines.py file.cubin [kernel-substring] [min-count]"""
import collections
import re
import subprocess
import sys

txt = subprocess.run(['nvdisasm', '-g', '-c', sys.argv[1]], capture_output=True, text=True).stdout
pat = sys.argv[2] if len(sys.argv) > 2 else ''
minc = int(sys.argv[3]) if len(sys.argv) > 3 else 20
cur, func = None, None
per = collections.defaultdict(collections.Counter)
for line in txt.splitlines():
    m = re.match(r'\s*\.text\.(\S+):', line)
    if m:
        func = m.group(1)
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
    if m and func and pat in func and cur:
        op = m.group(2)
        cls = 'fp64' if op in ('DFMA', 'DMUL', 'DADD', 'DSETP') else 'mufu' if op == 'MUFU' else \
              'local' if op in ('LDL', 'STL') else 'mem' if op in ('LDG', 'STG', 'LDS', 'STS', 'LD', 'ST') else 'other'
        per[cur][cls] += 1
tot = collections.Counter()
for k in sorted(per):
    c = per[k]
    tot.update(c)
    if sum(c.values()) >= minc:
        print(f'{k[0]}:{k[1]:<5d} total {sum(c.values()):5d}  ' + '  '.join(f'{a} {c[a]}' for a in ('fp64', 'mufu', 'local', 'mem', 'other')))
print('ALL', dict(tot))
