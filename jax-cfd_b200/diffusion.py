"""jax_cfd/base/diffusion.py: the pieces on the hot path."""
from __future__ import annotations

from typing import Optional

from . import _engine
from . import array_utils
from . import boundaries
from . import fast_diagonalization
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


_ops = {}


def _diffusion_op(grid: grids.Grid, nu: float, dt: float, implementation: Optional[str]):
  key = (grid, float(nu), float(dt), implementation)
  op = _ops.get(key)
  if op is None:
    def func(x):  # diffusion.py:175-177
      dt_nu_x = (dt * nu) * x
      return dt_nu_x / (1 - dt_nu_x)
    bc = boundaries.periodic_boundary_conditions(grid.ndim)
    laplacians = array_utils.laplacian_matrix_w_boundaries(grid, grid.cell_center, bc)
    op = _ops[key] = fast_diagonalization.transform(func, laplacians, 'float32', hermitian=True,
                                                    circulant=True, implementation=implementation)
  return op


def solve_fast_diag(v, nu: float, dt: float, implementation: Optional[str] = None):
  """diffusion.py:166-212 (periodic): u + (1 - nu dt lap)^-1 (nu dt lap) u per component, the
  second term by fast diagonalisation with func(x) = dt nu x / (1 - dt nu x) -- on the device:
  one table-driven transform (csrc/poisson_2d.cu x-line kernel in table mode) and one axpy per
  component."""
  if not boundaries.has_all_periodic_boundary_conditions(*v):
    raise NotImplementedError('only periodic boundary conditions are implemented on the B200 path')
  grid = grids.consistent_grid(*v)
  op = _diffusion_op(grid, nu, dt, implementation)
  w = tuple(grids.GridVariable(grids.GridArray(op(u.data), u.offset, u.grid), u.bc) for u in v)
  return _engine.axpy(v, [w], [1.0])  # u + op(u), one fused kernel per component
