#!/usr/bin/env python
"""Opcode summary of the in-tree sm_100a library (no GPU needed): per kernel, the SASS mnemonics that prove the
Blackwell-native paths (TMA bulk copies `UBLKCP`, their completion `ACQBULK`/`UTMACMDFLUSH`, programmatic
dependent launch `ACQBULK`..., system-scope accesses of the host mailbox) and the static instruction count.

    python tools/sass_summary.py > profiles/r02_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'bsuite_b200', 'libbsuite_b200.so')
WATCH = ('UBLKCP', 'UTMACMDFLUSH', 'ACQBULK', 'UTMASTG', 'UTMALDG', 'SYNCS', 'MEMBAR', 'ATOMG', 'REDG', 'LDG', 'STG',
         'STS', 'LDS', 'IMAD', 'DFMA', 'DMUL', 'DADD', 'MUFU', 'SHFL', 'BAR', 'CCTL', 'HMMA', 'UTCHMMA')


def main():
  archs = subprocess.run(['cuobjdump', '-lelf', LIB], capture_output=True, text=True).stdout
  print('embedded cubins:', sorted(set(re.findall(r'sm_\d+a?', archs))))
  sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
  kernels, name = collections.OrderedDict(), None
  for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
      name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
      kernels[name] = collections.Counter()
      continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
    if m and name:
      kernels[name][m.group(1)] += 1
  total = collections.Counter()
  print(f'{len(kernels)} kernels; per kernel: static instructions, then watched opcodes')
  for name, ops in kernels.items():
    short = re.sub(r'bsb::|\(bsb::EnvParams, bsb::LaunchArgs\)|void ', '', name)
    watched = ' '.join(f'{op}={ops[op]}' for op in WATCH if ops.get(op))
    print(f'{short[:78]:78s} {sum(ops.values()):6d}  {watched}')
    total.update(ops)
  print('library totals:', ' '.join(f'{op}={total[op]}' for op in WATCH if total.get(op)))


if __name__ == '__main__':
  sys.exit(main())
