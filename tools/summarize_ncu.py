#!/usr/bin/env python
"""Compact per-launch summary of `ncu --set full` reports (needs `ncu` on PATH; no GPU).

    python tools/summarize_ncu.py gpurun_out/ncu_r02a_*.ncu-rep --out profiles/r02_family_ncu_metrics.csv

One row per captured launch: duration, DRAM bytes, executed warp instructions, issue / pipe utilisation, resident
warps, registers, shared memory, the three largest issue-stall reasons (cycles stalled per issued instruction).
"""
import argparse
import csv
import os
import re
import subprocess

METRICS = [
    ('us', 'gpu__time_duration.sum'),
    ('dram_rd_MB', 'dram__bytes_read.sum'), ('dram_wr_MB', 'dram__bytes_write.sum'),
    ('tma_st_MB', 'l1tex__m_l1tex2xbar_write_bytes_mem_global_op_tma_st.sum'),
    ('warp_inst', 'smsp__inst_executed.sum'),
    ('issue_pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
    ('sm_thr_pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed'),
    ('fmaheavy_pct', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed'),
    ('alu_pct', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed'),
    ('fp64_pct', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'),
    ('lsu_pct', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed'),
    ('dram_pct', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
    ('warps_per_sm', 'sm__warps_active.avg.per_cycle_active'),
    ('regs', 'launch__registers_per_thread'), ('grid', 'launch__grid_size'), ('block', 'launch__block_size'),
    ('smem_dyn_B', 'launch__shared_mem_per_block_dynamic'), ('waves', 'launch__waves_per_multiprocessor'),
]
SCALE = {'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3, 'byte': 1e-6, 'ms': 1e3, 'us': 1.0, 'ns': 1e-3, 'second': 1e6}


def rows_of(report):
  raw = subprocess.run(['ncu', '-i', report, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
  rows = list(csv.reader(raw.splitlines()))
  header, units = rows[0], rows[1]
  stalls = [(i, h) for i, h in enumerate(header)
            if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio') and 'selected' not in h]
  out = []
  for r in rows[2:]:
    kernel = r[header.index('Kernel Name')]
    short = re.sub(r'^void |bsb::', '', kernel)[:90]
    row = {'report': os.path.basename(report), 'kernel': short}
    for name, key in METRICS:
      if key in header:
        i = header.index(key)
        v = float(r[i].replace(',', '')) if r[i] not in ('', 'n/a') else float('nan')
        if name.endswith('_MB') or name == 'us':
          v *= SCALE.get(units[i], 1.0)
        row[name] = round(v, 3)
    top = sorted(((float(r[i]), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]) for i, h in stalls), reverse=True)[:3]
    row['top_stalls'] = ' '.join(f'{n}={v:.2f}' for v, n in top)
    out.append(row)
  return out


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('reports', nargs='+')
  ap.add_argument('--out', default=None)
  args = ap.parse_args()
  rows = [row for rep in args.reports for row in rows_of(rep)]
  cols = ['report', 'kernel'] + [n for n, _ in METRICS] + ['top_stalls']
  if args.out:
    with open(args.out, 'w', newline='') as fh:
      w = csv.DictWriter(fh, fieldnames=cols)
      w.writeheader()
      w.writerows(rows)
  for row in rows:
    print(row['report'][9:34].ljust(26), ' '.join(f"{n}={row.get(n)}" for n, _ in METRICS[:17]), '|', row['top_stalls'])


if __name__ == '__main__':
  main()
