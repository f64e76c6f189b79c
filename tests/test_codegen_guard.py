"""Code-generation guard (no GPU needed: nvcc cross-compiles to PTX).

The QL eigen-solve of the 1-D kernel must keep its (d, e, z) arrays in REGISTERS.  Round 1 found that the optimiser
had merged the N-1 unrolled copies of the chase's early-exit block into one block with a run-time index; that single
dynamic access demoted d[] and e[] to local memory for the whole kernel (2 loads + 2 stores per rotation) and cost
15 % of the headline throughput (profiles/r1_ql_regs_ab.log).  In PTX the symptom is a local depot
larger than the out-of-line slow paths' mailboxes and `ld.local` / `st.local` with register offsets inside the kernel
body, so this test compiles one order to PTX and checks for exactly that."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, 'mfs_b200', 'csrc')
NVCC = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'


@pytest.mark.skipif(not os.path.exists(NVCC), reason='nvcc not available')
def test_ql_arrays_stay_in_registers(tmp_path):
    N = 5
    ptx = tmp_path / 'filter1d.ptx'
    subprocess.run([NVCC, '-gencode', 'arch=compute_100a,code=compute_100a', '-O3', '-std=c++17',
                    '--expt-relaxed-constexpr', f'-DMFS_N={N}', '-ptx', os.path.join(CSRC, 'filter1d_inst.cu'),
                    '-o', str(ptx)], check=True, capture_output=True)
    text = ptx.read_text()
    entries = re.split(r'\n(?=\.visible \.entry |\.func |\.weak \.func |\.visible \.func )', text)
    kernels = [e for e in entries if e.startswith('.visible .entry') and 'filter1d_kernel' in e.split('(')[0]]
    assert len(kernels) >= 6                       # (raw, central, scaled) x transition kinds
    mailbox = 4 * N * 8                            # slow-path mailboxes: QL continuation (3N doubles) + LDL route share it
    for k in kernels:
        name = k.split('(')[0]
        depots = [int(m) for m in re.findall(r'__local_depot\d+\[(\d+)\]', k)]
        assert sum(depots) <= mailbox, (name, depots)
        # the mailboxes are addressed with constant offsets; a register-indexed local access means an array was demoted
        dyn = [ln for ln in k.splitlines() if re.search(r'(ld|st)\.local', ln) and re.search(r'\[%rd\d+\]', ln)
               and 'depot' not in ln]
        base_regs = set(re.findall(r'add\.u64\s+(%rd\d+), %SPL, \d+', k)) | set(re.findall(r'cvta\.local\.u64\s+(%rd\d+)', k)) \
            | set(re.findall(r'cvta\.to\.local\.u64\s+(%rd\d+)', k))   # mailbox addresses formed from %SP + constant
        dyn = [ln for ln in dyn if not any(f'[{r}]' in ln for r in base_regs)]
        # two instances of the quadrature (first-step / literal recursion, and the per-step update), each reading its
        # slow-path mailboxes (QL continuation 2N, LDL route 2N) back with constant offsets
        assert len(re.findall(r'ld\.local', k)) <= 8 * N, name
        assert not dyn, (name, dyn[:3])
