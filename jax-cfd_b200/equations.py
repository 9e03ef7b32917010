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


def semi_implicit_navier_stokes(density: float, viscosity: float, dt: float, grid: grids.Grid,
                                convect=None, diffuse=diffusion.diffuse,
                                pressure_solve: Callable = pressure.solve_fast_diag,
                                forcing=None,
                                time_stepper: Callable = time_stepping.forward_euler) -> Callable:
  """equations.py:120-151.  Returns `step_fn(v) -> v'` with the same GridVariable in/out types.

  Supported (anything else raises NotImplementedError at build time, there is no fallback):
  all-periodic boundaries, offsets == grid.cell_faces, float32, power-of-two grid,
  convect=None, diffuse=diffusion.diffuse, pressure_solve=pressure.solve_fast_diag, forcing from
  `forcings.*`; time_stepper any of time_stepping.{forward_euler, midpoint_rk2, heun_rk2,
  classic_rk4} (or a custom tableau through navier_stokes_rk).  Not differentiable.
  """
  _check_terms(convect, diffuse)
  if pressure_solve is not pressure.solve_fast_diag:
    raise NotImplementedError('the B200 path implements pressure_solve=pressure.solve_fast_diag only')
  f = _engine.as_forcing(forcing)
  explicit_terms = _engine.NativeExplicitTerms(grid, dt, density, viscosity, f)
  ode = time_stepping.ExplicitNavierStokesODE(explicit_terms, _engine.NativeProjection(grid))
  ode.native_projection = True
  ode.fused_step = _engine.NativeStep(grid, dt, density, viscosity, f)
  return time_stepper(ode, dt)
