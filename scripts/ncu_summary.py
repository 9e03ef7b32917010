"""Turns `ncu -i report.ncu-rep --page raw --csv` into the markdown table kept under profiles/.

  ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv
  python scripts/ncu_summary.py /tmp/raw.csv > profiles/rNN_chain_ncu_summary.md

One column per distinct kernel (first launch of each), the metrics the roofline discussion uses.
"""
import csv
import json
import re
import sys

METRICS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
    'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
]


def short(name):
  m = re.search(r'(\w+)<([^>]*)>', name)
  return f'{m.group(1)}<{m.group(2)}>' if m else name.split('(')[0]


def main(path):
  rows = list(csv.reader(open(path)))
  hdr, units, data = rows[0], rows[1], rows[2:]
  kcol = hdr.index('Kernel Name')
  seen, cols = set(), []
  for r in data:
    k = short(r[kcol])
    if k not in seen:
      seen.add(k)
      cols.append((k, r))
  print('| metric | unit | ' + ' | '.join(k for k, _ in cols) + ' |')
  print('|---|---|' + '---|' * len(cols))
  traffic = {}
  for m in METRICS:
    if m not in hdr:
      continue
    i = hdr.index(m)
    print(f'| {m} | {units[i]} | ' + ' | '.join(r[i] for _, r in cols) + ' |')
  ir, iw = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
  scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
  for k, r in cols:
    traffic[k] = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
  print()
  print('DRAM traffic per launch (read + write, bytes): ' + json.dumps(traffic))


if __name__ == '__main__':
  main(sys.argv[1])
