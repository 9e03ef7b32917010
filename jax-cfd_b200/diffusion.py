"""jax_cfd/base/diffusion.py: the pieces on the hot path."""
from __future__ import annotations

from . import grids


def diffuse(c, nu):
  """diffusion.py:35-37 -- accepted as the `diffuse=` argument of the equation builders (it is
  the only supported choice); the Laplacian itself is fused into the explicit-terms kernel."""
  raise NotImplementedError('diffusion is fused into the explicit-terms kernel on the B200 path')


def stable_time_step(viscosity: float, grid: grids.Grid) -> float:
  """diffusion.py:40-57."""
  if viscosity == 0:
    return float('inf')
  dx = min(grid.step)
  ndim = grid.ndim
  return dx ** 2 / (viscosity * 2 ** ndim)
