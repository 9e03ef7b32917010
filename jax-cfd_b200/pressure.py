"""jax_cfd/base/pressure.py: projection by fast diagonalisation (periodic, rfft path)."""
from __future__ import annotations

from . import grids
from ._engine import NativeProjection


def solve_fast_diag(v, q0=None, pressure_bc=None, implementation=None) -> grids.GridArray:
  """pressure.py:115-157: q = pinv(divergence(v)).  `q0` is unused, like in the reference."""
  del q0, pressure_bc
  if implementation not in (None, 'rfft'):
    raise NotImplementedError('only the default rfft implementation exists on the B200 path')
  _, q = NativeProjection()(v, return_q=True)
  return q


def projection(v, solve=solve_fast_diag):
  """pressure.py:181-198."""
  if solve is not solve_fast_diag:
    raise NotImplementedError('only pressure.solve_fast_diag is implemented on the B200 path')
  return NativeProjection()(v)
