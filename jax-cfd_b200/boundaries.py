"""Boundary conditions: the periodic branch of jax_cfd/base/boundaries.py.

The B200 path is periodic-only (the reference's advection raises NotImplementedError for anything
else too, advection.py:108-110).  Non-periodic constructors are not provided.
"""
from __future__ import annotations

import dataclasses
from typing import Optional, Sequence, Tuple

import numpy as np

from . import grids


class BCType:
  """boundaries.py:33-36."""
  PERIODIC = 'periodic'
  DIRICHLET = 'dirichlet'
  NEUMANN = 'neumann'


@dataclasses.dataclass(init=False, frozen=True)
class ConstantBoundaryConditions(grids.BoundaryConditions):
  """boundaries.py:44-530 restricted to `periodic` types."""
  types: Tuple[Tuple[str, str], ...]
  bc_values: Tuple[Tuple[Optional[float], Optional[float]], ...]

  def __init__(self, types: Sequence[Tuple[str, str]], values):
    types = tuple(tuple(t) for t in types)
    values = tuple(tuple(v) for v in values)
    object.__setattr__(self, 'bc_values', values)
    object.__setattr__(self, 'types', types)

  def shift(self, u: grids.GridArray, offset: int, axis: int, mode=None) -> grids.GridArray:
    """out[i] = in[(i + offset) mod N], offset += k  (boundaries.py:81-104, 193-199)."""
    if self.types[axis][0] != BCType.PERIODIC:
      raise NotImplementedError('only periodic boundaries are implemented on the B200 path')
    data = np.asarray(u.data)
    ax = axis + (data.ndim - u.grid.ndim)  # tolerate leading batch dims (grids.py:59-63)
    new_offset = tuple(o + offset if i == axis else o for i, o in enumerate(u.offset))
    return grids.GridArray(np.roll(data, -offset, axis=ax), new_offset, u.grid)

  def trim_boundary(self, u: grids.GridArray) -> grids.GridArray:
    return u

  def impose_bc(self, u: grids.GridArray) -> grids.GridVariable:
    """No-op for periodic boundaries (boundaries.py:514-530)."""
    return grids.GridVariable(u, self)


class HomogeneousBoundaryConditions(ConstantBoundaryConditions):
  """boundaries.py:533-553."""

  def __init__(self, types: Sequence[Tuple[str, str]]):
    ndim = len(types)
    super().__init__(types, ((0.0, 0.0),) * ndim)


def periodic_boundary_conditions(ndim: int) -> ConstantBoundaryConditions:
  """boundaries.py:556-559."""
  return HomogeneousBoundaryConditions(((BCType.PERIODIC, BCType.PERIODIC),) * ndim)


def is_periodic_boundary_conditions(c: grids.GridVariable, axis: int) -> bool:
  """boundaries.py:681-685."""
  return c.bc.types[axis][0] == BCType.PERIODIC


def has_all_periodic_boundary_conditions(*arrays: grids.GridVariable) -> bool:
  """boundaries.py:688-694."""
  for a in arrays:
    for axis in range(a.grid.ndim):
      if not is_periodic_boundary_conditions(a, axis):
        return False
  return True


def get_pressure_bc_from_velocity(v) -> HomogeneousBoundaryConditions:
  """boundaries.py:724-735 (periodic velocity -> periodic pressure)."""
  if not has_all_periodic_boundary_conditions(*v):
    raise NotImplementedError('only periodic boundaries are implemented on the B200 path')
  return periodic_boundary_conditions(v[0].grid.ndim)
