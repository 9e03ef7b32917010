"""jax_cfd/base/pressure.py: projection by fast diagonalisation (periodic)."""
from __future__ import annotations

import functools
from typing import Callable, Optional

from . import _lib
from . import grids
from ._engine import NativeProjection


def solve_fast_diag(v, q0=None, pressure_bc=None, implementation: Optional[str] = None) -> grids.GridArray:
  """pressure.py:115-157: q = pinv(divergence(v)).  `q0` is unused, like in the reference.
  `implementation`: None (rfft when every axis is a power of two, else matmul -- the reference's
  fallback for odd last axes, fast_diagonalization.py:101-108), 'rfft', 'fft' or 'matmul'."""
  del q0, pressure_bc
  _, q = NativeProjection(implementation=implementation)(v, return_q=True)
  return q


def implementation_of(solve: Callable) -> Optional[str]:
  """The `implementation` a pressure_solve argument selects: pressure.solve_fast_diag itself or a
  functools.partial of it (how the reference's callers pick 'matmul', e.g. pressure_test.py)."""
  if solve is solve_fast_diag:
    return None
  if isinstance(solve, functools.partial) and solve.func is solve_fast_diag and not solve.args and set(
      solve.keywords) <= {'implementation'}:
    impl = solve.keywords.get('implementation')
    _lib.implementation_code(impl)
    return impl
  raise NotImplementedError('only pressure.solve_fast_diag (optionally functools.partial(..., '
                            'implementation=...)) is implemented on the B200 path')


def projection(v, solve: Callable = solve_fast_diag):
  """pressure.py:181-198."""
  return NativeProjection(implementation=implementation_of(solve))(v)
