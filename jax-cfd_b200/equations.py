"""jax_cfd/base/equations.py: the Navier-Stokes equation builders on the hot path."""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np

from . import _engine
from . import _lib
from . import advection
from . import diffusion
from . import grids
from . import pressure
from . import time_stepping


def stable_time_step(max_velocity: float, max_courant_number: float, viscosity: float,
                     grid: grids.Grid, implicit_diffusion: bool = False) -> float:
  """equations.py:45-59."""
  dt = advection.stable_time_step(max_velocity, max_courant_number, grid)
  if not implicit_diffusion:
    diffusion_dt = diffusion.stable_time_step(viscosity, grid)
    if diffusion_dt < dt:
      raise ValueError(f'stable time step for diffusion is smaller than '
                       f'the chosen timestep: {diffusion_dt} vs {dt}')
  return dt


def dynamic_time_step(v, max_courant_number: float, viscosity: float, grid: grids.Grid,
                      implicit_diffusion: bool = False) -> float:
  """equations.py:62-70; the max of sum(u^2) is a fused device reduction."""
  v_max = float(np.sqrt(_engine.diagnostics(v)['max_speed_sq']))
  return stable_time_step(v_max, max_courant_number, viscosity, grid, implicit_diffusion)


def _check_terms(convect, diffuse):
  if convect is not None:
    raise NotImplementedError(
        'the B200 path implements the default convection (Van-Leer limited, '
        'advection.advect_van_leer_using_limiters) only; pass convect=None')
  if diffuse is not diffusion.diffuse:
    raise NotImplementedError('the B200 path implements diffuse=diffusion.diffuse only')


def navier_stokes_explicit_terms(density: float, viscosity: float, dt: float, grid: grids.Grid,
                                 convect=None, diffuse=diffusion.diffuse,
                                 forcing=None) -> Callable:
  """equations.py:77-116: conv + (nu/rho) lap + forcing/rho as ONE kernel."""
  _check_terms(convect, diffuse)
  return _engine.NativeExplicitTerms(grid, dt, density, viscosity, _engine.as_forcing(forcing))


def _check_grid(grid: grids.Grid):
  """Build-time validation (the reference raises at trace time too): what the kernels take."""
  if grid.ndim not in (2, 3):
    raise NotImplementedError(f'the B200 path implements 2-D and 3-D grids; got ndim={grid.ndim}')


def semi_implicit_navier_stokes(density: float, viscosity: float, dt: float, grid: grids.Grid,
                                convect=None, diffuse=diffusion.diffuse,
                                pressure_solve: Callable = pressure.solve_fast_diag,
                                forcing=None,
                                time_stepper: Callable = time_stepping.forward_euler) -> Callable:
  """equations.py:120-151.  Returns `step_fn(v) -> v'` with the same GridVariable in/out types.

  Supported (anything else raises NotImplementedError at build time, there is no fallback):
  all-periodic boundaries, offsets == grid.cell_faces, float32, any grid shape with axes of at
  least 4 cells (power-of-two axes take the line-FFT pressure solve, other shapes the matmul one),
  convect=None, diffuse=diffusion.diffuse, pressure_solve=pressure.solve_fast_diag (or a
  functools.partial of it selecting `implementation`), forcing from `forcings.*`; time_stepper any
  of time_stepping.{forward_euler, midpoint_rk2, heun_rk2, classic_rk4} (or a custom tableau
  through navier_stokes_rk).  Not differentiable.
  """
  _check_terms(convect, diffuse)
  _check_grid(grid)
  impl = pressure.implementation_of(pressure_solve)
  f = _engine.as_forcing(forcing)
  explicit_terms = _engine.NativeExplicitTerms(grid, dt, density, viscosity, f, implementation=impl)
  ode = time_stepping.ExplicitNavierStokesODE(explicit_terms, _engine.NativeProjection(grid, impl))
  ode.native_projection = True
  ode.fused_step = _engine.NativeStep(grid, dt, density, viscosity, f, implementation=impl)
  return time_stepper(ode, dt)


def implicit_diffusion_navier_stokes(density: float, viscosity: float, dt: float, grid: grids.Grid,
                                     convect=None, diffusion_solve: Callable = diffusion.solve_fast_diag,
                                     pressure_solve: Callable = pressure.solve_fast_diag,
                                     forcing=None) -> Callable:
  """equations.py:154-195: v* = v + dt (conv + forcing / rho); v = P(v*); then the implicit
  diffusion solve (1 - nu dt lap)^-1 per component (diffusion.py:166-212).  On the device: the fused
  step with the explicit Laplacian switched off (one stencil kernel + the pressure projection),
  then one table-driven fast-diagonalisation transform and one axpy per component.  Like the
  reference, `viscosity` (not viscosity / density) multiplies the Laplacian here (equations.py:192)."""
  if convect is not None:
    raise NotImplementedError('the B200 path implements the default Van-Leer convection only; pass convect=None')
  if diffusion_solve is not diffusion.solve_fast_diag:
    raise NotImplementedError('the B200 path implements diffusion_solve=diffusion.solve_fast_diag only')
  _check_grid(grid)
  impl = pressure.implementation_of(pressure_solve)
  f = _engine.as_forcing(forcing)
  inviscid = _engine.NativeStep(grid, dt, density, None, f, implementation=impl)

  def navier_stokes_step(v):
    return diffusion.solve_fast_diag(inviscid(v), viscosity, dt)

  return navier_stokes_step
