#!/usr/bin/env python
"""SASS-level profile of one launch of an ncu report: executed warp instructions by opcode and the hottest instructions.

    python tools/ncu_hot_sass.py gpurun_out/ncu_r02a_bandit_0.ncu-rep [launch-index]
"""
import collections
import csv
import subprocess
import sys


def main():
  report, which = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0
  raw = subprocess.run(['ncu', '-i', report, '--page', 'source', '--csv', '--print-source', 'sass'],
                       capture_output=True, text=True).stdout
  launches, cur = [], None
  for row in csv.reader(raw.splitlines()):
    if row and row[0] == 'Kernel Name':
      cur = {'kernel': row[1], 'rows': [], 'header': None}
      launches.append(cur)
    elif cur is not None and cur['header'] is None:
      cur['header'] = row
    elif cur is not None and row:
      cur['rows'].append(row)
  launch = launches[which]
  h = launch['header']
  i_src, i_exec, i_samp = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
  by_op, total = collections.Counter(), 0
  for r in launch['rows']:
    n = int(r[i_exec] or 0)
    op = r[i_src].split()[0] if not r[i_src].strip().startswith('@') else r[i_src].split()[1]
    by_op[op.split('.')[0]] += n
    total += n
  print(launch['kernel'][:100], 'static instrs', len(launch['rows']), 'executed warp-instrs', total)
  print('by opcode:', ', '.join(f'{op}={n} ({100.0 * n / total:.1f}%)' for op, n in by_op.most_common(14)))
  print('hottest by stall samples:')
  for r in sorted(launch['rows'], key=lambda r: -int(r[i_samp] or 0))[:12]:
    print(f'  samples={r[i_samp]:>6} exec={r[i_exec]:>9}  {r[i_src].strip()[:90]}')


if __name__ == '__main__':
  main()
