"""Helpers shared by the parity tests: load tests/golden fixtures (generated from the reference
by oracle/gen_golden.py) and turn their forcing spec into oracle / product descriptors."""
import ast
import glob
import os

import numpy as np

import cfd_oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
  z = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
  rec = {k: z[k] for k in z.files}
  rec['shape'] = tuple(int(s) for s in rec['shape'])
  rec['domain'] = tuple((float(a), float(b)) for a, b in rec['domain'])
  for k in ('density', 'viscosity', 'dt', 'smag_cs'):
    if k in rec:
      rec[k] = float(rec[k])
  if 'forcing_spec' in rec:
    rec['forcing_spec'] = ast.literal_eval(str(rec['forcing_spec']))
  if 'nsteps' in rec:
    rec['nsteps'] = [int(n) for n in rec['nsteps']]
  if 'stepper' in rec:
    rec['stepper'] = str(rec['stepper'])
  rec['ndim'] = len(rec['shape'])
  rec['h'] = cfd_oracle.grid_step(rec['shape'], rec['domain'])
  return rec


def step_cases():
  return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, '*.npz'))
                if not os.path.basename(p).startswith(('proj', 'imp', 'post')))


def implicit_cases():
  return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, 'imp*.npz')))


def oracle_forcing(rec, dtype=np.float32):
  terms = []
  for kind, arg in (rec['forcing_spec'] or []):
    if kind == 'kolmogorov':
      terms.append(('const', cfd_oracle.kolmogorov_field(rec['shape'], rec['domain'], dtype=dtype, **arg)))
    elif kind == 'taylor_green':
      terms.append(('const', cfd_oracle.taylor_green_field(rec['shape'], dtype=dtype, **arg)))
    elif kind == 'linear':
      terms.append(('linear', arg))
  if rec['smag_cs'] >= 0:
    terms.append(('smagorinsky', rec['smag_cs']))
  return cfd_oracle.Forcing(tuple(terms)) if terms else None


def rel_l2(a, b):
  a = np.asarray(a, np.float64)
  b = np.asarray(b, np.float64)
  return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


TABLEAUS = {
    'forward_euler': ([], [1]),
    'midpoint_rk2': ([[1 / 2]], [0, 1]),
    'heun_rk2': ([[1]], [1 / 2, 1 / 2]),
    'classic_rk4': ([[1 / 2], [0, 1 / 2], [0, 0, 1]], [1 / 6, 1 / 3, 1 / 3, 1 / 6]),
}
