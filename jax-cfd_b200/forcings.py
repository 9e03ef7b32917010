"""Forcing terms: jax_cfd/base/forcings.py, as descriptors the fused stencil kernel consumes.

Each factory returns a `ForcingFn` that (a) can be called like the reference's (`forcing(v)` ->
tuple of GridArray, evaluated on the host for inspection) and (b) carries the tables / scalars
the CUDA kernel needs, so nothing is evaluated per step on the host.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from . import grids
from ._engine import FieldTerm, ForcingFn, LinearTerm, SeparableTerm


def kolmogorov_forcing(grid: grids.Grid, scale: float = 1, k: int = 2, swap_xy: bool = False,
                       offsets: Optional[Tuple[Tuple[float, ...], ...]] = None) -> ForcingFn:
  """forcings.py:63-104: f_0 = scale * sin(k * y) on the u-face mesh (swap_xy: f_1 = scale *
  sin(k * x) on the v-face mesh).  The profile is evaluated once, in float32, with the same
  expression the reference evaluates eagerly (forcings.py:88-89, grids.py:600-602)."""
  if grid.ndim not in (2, 3):
    raise NotImplementedError
  if offsets is None:
    offsets = grid.cell_faces
  d = grid.ndim
  profiles = [[None] * d for _ in range(d)]
  has = [False] * d
  scales = [0.0] * d
  f32 = np.float32
  if swap_xy:
    x = grid.axes(offsets[1])[0]
    profiles[1][0] = np.sin((f32(k) * x).astype(f32)).astype(f32)
    has[1], scales[1] = True, float(scale)
  else:
    y = grid.axes(offsets[0])[1]
    profiles[0][1] = np.sin((f32(k) * y).astype(f32)).astype(f32)
    has[0], scales[0] = True, float(scale)
  return ForcingFn([SeparableTerm(grid, profiles, scales, has, grid.cell_faces)], grid)


def taylor_green_forcing(grid: grids.Grid, scale: float = 1, k: int = 2) -> ForcingFn:
  """forcings.py:35-60 with validation_problems.py:72-74,90-104: the Taylor-Green velocity of a
  (0, 2 pi)^2 grid of shape grid.shape[:2], re-labelled onto `grid`; constant along z in 3-D."""
  if grid.ndim not in (2, 3):
    raise NotImplementedError
  d = grid.ndim
  tg = grids.Grid(grid.shape[:2], domain=((0., 2. * np.pi), (0., 2. * np.pi)))
  f32 = np.float32
  xu, yu = tg.axes(tg.cell_faces[0])
  xv, yv = tg.axes(tg.cell_faces[1])
  kk = f32(k)
  profiles = [[None] * d for _ in range(d)]
  profiles[0][0] = np.cos(kk * xu).astype(f32)       # scale(=1) * cos(kx x) * sin(ky y)
  profiles[0][1] = np.sin(kk * yu).astype(f32)
  profiles[1][0] = (-np.sin(kk * xv)).astype(f32)    # -scale * sin(kx x) * cos(ky y)
  profiles[1][1] = np.cos(kk * yv).astype(f32)
  has = [True, True] + [False] * (d - 2)
  scales = [float(scale), float(scale)] + [0.0] * (d - 2)
  return ForcingFn([SeparableTerm(grid, profiles, scales, has, grid.cell_faces)], grid)


def linear_forcing(grid, coefficient: float) -> ForcingFn:
  """forcings.py:107-113."""
  return ForcingFn([LinearTerm(coefficient)], grid)


def no_forcing(grid) -> ForcingFn:
  """forcings.py:116-122 (0 * u: contributes nothing)."""
  return ForcingFn([], grid)


def constant_forcing(grid: grids.Grid, arrays) -> ForcingFn:
  """Any velocity-independent forcing, given as one array (or None) per component on
  grid.cell_faces -- what a constant reference ForcingFn returns."""
  arrays = [None if a is None else (a.data if isinstance(a, grids.GridArray) else a) for a in arrays]
  return ForcingFn([FieldTerm(grid, arrays, grid.cell_faces)], grid)


def sum_forcings(*forcings: ForcingFn) -> ForcingFn:
  """forcings.py:125-129: terms are summed left to right."""
  terms = []
  grid = None
  for f in forcings:
    if not isinstance(f, ForcingFn):
      raise NotImplementedError('sum_forcings accepts forcings built by this module only')
    terms.extend(f.terms)
    grid = grid or f.grid
  return ForcingFn(terms, grid)


FORCING_FUNCTIONS = dict(kolmogorov=kolmogorov_forcing, taylor_green=taylor_green_forcing)


def simple_turbulence_forcing(grid: grids.Grid, constant_magnitude: float = 0,
                              constant_wavenumber: int = 2, linear_coefficient: float = 0,
                              forcing_type: str = 'kolmogorov') -> ForcingFn:
  """forcings.py:136-178: linear THEN constant."""
  linear_force = linear_forcing(grid, linear_coefficient)
  constant_force_fn = FORCING_FUNCTIONS.get(forcing_type)
  if constant_force_fn is None:
    raise ValueError('Unknown `forcing_type`. '
                     f'Expected one of {list(FORCING_FUNCTIONS.keys())}; got {forcing_type}.')
  constant_force = constant_force_fn(grid, constant_magnitude, constant_wavenumber)
  return sum_forcings(linear_force, constant_force)
