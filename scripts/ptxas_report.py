"""Registers / spills per kernel from the ptxas logs next to the built library (vos_e_sam_b200/lib/obj/*.ptxas.log)."""
import glob
import os
import re
import subprocess
import sys

here = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'vos_e_sam_b200', 'lib', 'obj')
pat = sys.argv[1] if len(sys.argv) > 1 else ''
for path in sorted(glob.glob(os.path.join(here, '*.ptxas.log'))):
    cur, sp = None, ''
    for ln in open(path).read().split('\n'):
        m = re.search(r"Compiling entry function '(\S+)'", ln)
        if m:
            cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r'\(anonymous namespace\)::|vosmem::', '', cur)[:100]
        if 'spill' in ln:
            sp = ln.strip()
        if 'Used' in ln and cur:
            if re.search(pat, cur):
                print(os.path.basename(path)[:-10], '|', cur, '|', re.search(r'Used (\d+) registers', ln).group(1), '|', sp)
            cur = None
