#!/usr/bin/env python
"""Extracts the counters DESIGN.md / bench.py cite from an `ncu --set full` report (needs `ncu` on PATH; no GPU).

    python tools/extract_ncu.py gpurun_out/prof.ncu-rep profiles/r01_deep_sea_bulk_ncu_metrics.csv
"""
import csv
import subprocess
import sys

KEEP = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_write.sum.per_second',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
    'launch__waves_per_multiprocessor', 'l1tex__m_l1tex2xbar_write_bytes_mem_global_op_tma_st.sum',
    'l1tex__m_l1tex2xbar_write_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_sectors_srcunit_tex_op_write.sum', 'smsp__inst_executed.sum', 'smsp__inst_executed_op_tma_st.sum',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
]


def main():
  report, out = sys.argv[1], sys.argv[2]
  raw = subprocess.run(['ncu', '-i', report, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
  rows = list(csv.reader(raw.splitlines()))
  header, units, launches = rows[0], rows[1], rows[2:]
  with open(out, 'w', newline='') as fh:
    w = csv.writer(fh)
    w.writerow(['metric', 'unit'] + [f'launch_{i}' for i in range(len(launches))])
    w.writerow(['kernel', ''] + [r[header.index('Kernel Name')] for r in launches])
    for key in KEEP:
      if key in header:
        i = header.index(key)
        w.writerow([key, units[i]] + [r[i] for r in launches])
  print(open(out).read())


if __name__ == '__main__':
  main()
