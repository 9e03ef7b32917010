"""Registers / spills per kernel from a `-Xptxas -v` log (csrc/Makefile writes them next to the objects).

  python scripts/ptxas_summary.py /tmp/cfd_b200_build/poisson_2d.ptxas.log [substring ...]
"""
import re
import subprocess
import sys


def main(path, *filters):
  txt = open(path).read()
  pat = re.compile(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes "
                   r"spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", re.S)
  names, rows = [], []
  for name, stack, ss, sl, regs in pat.findall(txt):
    names.append(name)
    rows.append((regs, stack, ss, sl))
  dem = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout.split('\n')
  for d, (regs, stack, ss, sl) in zip(dem, rows):
    d = re.sub(r'\(.*', '', d.replace('(anonymous namespace)::', '')).replace('void cfd::', '')
    if filters and not any(f in d for f in filters):
      continue
    print(f'{d:70s} regs {regs:>3s} stack {stack:>4s} spill st/ld {ss}/{sl}')


if __name__ == '__main__':
  main(*sys.argv[1:])
