"""Static SASS statistics of an object/shared library: instructions and opcode mix per kernel."""
import collections
import re
import subprocess
import sys

out = subprocess.run(['cuobjdump', '-sass', sys.argv[1]], capture_output=True, text=True).stdout
pat = sys.argv[2] if len(sys.argv) > 2 else ''
name, stats = None, {}
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = m.group(1)
        stats[name] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
    if m and name:
        stats[name][m.group(2)] += 1
for name, c in stats.items():
    if pat in name:
        tot = sum(c.values())
        fp64 = sum(v for k, v in c.items() if k in ('DFMA', 'DMUL', 'DADD', 'DSETP', 'MUFU'))
        print(f'{name}: {tot} instrs, fp64-ish {fp64} ({100 * fp64 / max(tot, 1):.0f}%)', dict(c.most_common(8)))
