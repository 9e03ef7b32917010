"""The C/OpenMP restatement (oracle/cfd_oracle_c.c + scipy FFT = the CPU baseline) against the
NumPy oracle and the reference's golden vectors.  CPU only."""
import numpy as np
import pytest

import cfd_oracle
import cpu_baseline
import golden_util as gu


@pytest.mark.parametrize('name', ['k2d_64x32', 'd2d_128', 'd2d_48x36'])
def test_c_oracle_matches_golden(name):
  rec = gu.load(name)
  forcing = gu.oracle_forcing(rec)
  const = lin = None
  for kind, arg in (forcing.terms if forcing else ()):
    if kind == 'const':
      const = arg
    elif kind == 'linear':
      lin = arg
  cs = cpu_baseline.CpuStep(rec['shape'], rec['h'], rec['dt'], rec['density'], rec['viscosity'],
                            const, lin, workers=2)
  u, v = rec['v0_0'].copy(), rec['v0_1'].copy()
  n = rec['nsteps'][-1]
  for _ in range(n):
    u, v = cs.step(u, v)
  assert gu.rel_l2(u, rec[f'f32_v{n}_0']) < 1e-5
  assert gu.rel_l2(v, rec[f'f32_v{n}_1']) < 1e-5
