"""jax_cfd/base/advection.py: the pieces on the hot path."""
from __future__ import annotations

from . import grids


def advect_van_leer_using_limiters(c, v, dt):
  """advection.py:387-395.  On the B200 path Van-Leer advection only exists fused inside the
  explicit-terms kernel (csrc/explicit_2d.cu); use equations.navier_stokes_explicit_terms."""
  raise NotImplementedError('advection is fused into the explicit-terms kernel on the B200 path; '
                            'call equations.navier_stokes_explicit_terms(...) instead')


def stable_time_step(max_velocity: float, max_courant_number: float, grid: grids.Grid) -> float:
  """advection.py:398-416."""
  dx = min(grid.step)
  return max_courant_number * dx / max_velocity
