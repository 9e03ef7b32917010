"""Host-side mirror of the reference's grid / field types for the periodic staggered path.

Same names and meaning as jax_cfd/base/grids.py (Grid :499-675, GridArray :40-108, GridVariable
:265-395, control_volume_offsets :448-454, consistent_offset/grid :461-482), so user code and
tests written against the reference read the same here.  `data` may be a numpy array (host; the
step then runs through the host-buffer entry point) or a device array (DeviceArray, torch CUDA
tensor, anything with __cuda_array_interface__).  Metadata (offset, grid, bc) is static, exactly
like the aux data of the reference's pytrees.
"""
from __future__ import annotations

import dataclasses
import numbers
import operator
from typing import Any, Callable, Optional, Sequence, Tuple, Union

import numpy as np

Array = Any


class InconsistentOffsetError(Exception):
  """grids.py:458."""


class InconsistentGridError(Exception):
  """grids.py:473."""


class InconsistentBoundaryConditionsError(Exception):
  """grids.py:486."""


@dataclasses.dataclass(init=False, frozen=True)
class Grid:
  """Arakawa C-grid description (grids.py:499-675)."""
  shape: Tuple[int, ...]
  step: Tuple[float, ...]
  domain: Tuple[Tuple[float, float], ...]

  def __init__(self, shape, step=None, domain=None):
    shape = tuple(operator.index(s) for s in shape)
    object.__setattr__(self, 'shape', shape)
    if step is not None and domain is not None:
      raise TypeError('cannot provide both step and domain')
    if domain is not None:
      if isinstance(domain, (int, float)):
        domain = ((0, domain),) * len(shape)
      else:
        if len(domain) != len(shape):
          raise ValueError(f'length of domain does not match ndim: {len(domain)} != {len(shape)}')
        for bounds in domain:
          if len(bounds) != 2:
            raise ValueError(f'domain is not sequence of pairs of numbers: {domain}')
      domain = tuple((float(lo), float(up)) for lo, up in domain)
    else:
      if step is None:
        step = 1
      if isinstance(step, numbers.Number):
        step = (step,) * len(shape)
      elif len(step) != len(shape):
        raise ValueError(f'length of step does not match ndim: {len(step)} != {len(shape)}')
      domain = tuple((0.0, float(s * n)) for s, n in zip(step, shape))
    object.__setattr__(self, 'domain', domain)
    object.__setattr__(self, 'step', tuple((up - lo) / n for (lo, up), n in zip(domain, shape)))

  @property
  def ndim(self) -> int:
    return len(self.shape)

  @property
  def cell_center(self) -> Tuple[float, ...]:
    return self.ndim * (0.5,)

  @property
  def cell_faces(self) -> Tuple[Tuple[float, ...], ...]:
    d = self.ndim
    offsets = (np.eye(d) + np.ones([d, d])) / 2.
    return tuple(tuple(float(o) for o in row) for row in offsets)

  def stagger(self, v):
    return tuple(GridArray(u, o, self) for u, o in zip(v, self.cell_faces))

  def center(self, v):
    if isinstance(v, (tuple, list)):
      return type(v)(GridArray(u, self.cell_center, self) for u in v)
    return GridArray(v, self.cell_center, self)

  def axes(self, offset=None):
    """float32 coordinates evaluated the way x64-disabled JAX evaluates grids.py:600-602."""
    if offset is None:
      offset = self.cell_center
    if len(offset) != self.ndim:
      raise ValueError(f'unexpected offset length: {len(offset)} vs {self.ndim}')
    f32 = np.float32
    out = []
    for (lo, _), o, n, h in zip(self.domain, offset, self.shape, self.step):
      a = np.arange(n, dtype=np.int32).astype(f32)
      a = (a + f32(o)).astype(f32)
      a = (a * f32(h)).astype(f32)
      out.append((f32(lo) + a).astype(f32))
    return tuple(out)

  def fft_axes(self):
    return tuple(np.fft.fftfreq(n, d=s) for n, s in zip(self.shape, self.step))

  def rfft_axes(self):
    return tuple(np.fft.fftfreq(n, d=s) for n, s in zip(self.shape[:-1], self.step[:-1])) + (
        np.fft.rfftfreq(self.shape[-1], d=self.step[-1]),)

  def mesh(self, offset=None):
    return tuple(np.meshgrid(*self.axes(offset), indexing='ij'))

  def rfft_mesh(self):
    return tuple(np.meshgrid(*self.rfft_axes(), indexing='ij'))

  def eval_on_mesh(self, fn: Callable, offset=None) -> 'GridArray':
    if offset is None:
      offset = self.cell_center
    return GridArray(fn(*self.mesh(offset)), offset, self)


@dataclasses.dataclass
class GridArray(np.lib.mixins.NDArrayOperatorsMixin):
  """Data with an alignment offset and a grid (grids.py:40-108)."""
  data: Array
  offset: Tuple[float, ...]
  grid: Grid

  @property
  def dtype(self):
    return self.data.dtype

  @property
  def shape(self):
    return tuple(self.data.shape)

  _HANDLED = (numbers.Number, np.ndarray)

  def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
    """Arithmetic for HOST (numpy) data, with the reference's offset/grid consistency checks."""
    for x in inputs:
      if not isinstance(x, self._HANDLED + (GridArray,)):
        return NotImplemented
    if method != '__call__':
      return NotImplemented
    arrays = []
    for x in inputs:
      if isinstance(x, GridArray):
        if not isinstance(x.data, np.ndarray):
          raise TypeError('elementwise arithmetic on device-resident GridArrays is not provided; '
                          'copy to host with np.asarray(x.data)')
        arrays.append(x.data)
      else:
        arrays.append(x)
    result = ufunc(*arrays, **kwargs)
    gas = [x for x in inputs if isinstance(x, GridArray)]
    offset = consistent_offset(*gas)
    grid = consistent_grid(*gas)
    if isinstance(result, tuple):
      return tuple(GridArray(r, offset, grid) for r in result)
    return GridArray(result, offset, grid)


GridArrayVector = Tuple[GridArray, ...]


class BoundaryConditions:
  """grids.py:140-262 (abstract)."""
  types: Tuple[Tuple[str, str], ...]

  def shift(self, u, offset, axis, mode=None):
    raise NotImplementedError

  def impose_bc(self, u):
    raise NotImplementedError


@dataclasses.dataclass
class GridVariable:
  """GridArray + boundary conditions (grids.py:265-395)."""
  array: GridArray
  bc: BoundaryConditions

  def __post_init__(self):
    if not isinstance(self.array, GridArray):
      raise ValueError(f'Expected array type to be GridArray, got {type(self.array)}')
    if len(self.bc.types) != self.grid.ndim:
      raise ValueError('Incompatible dimension between grid and bc, grid dimension = '
                       f'{self.grid.ndim}, bc dimension = {len(self.bc.types)}')

  @property
  def dtype(self):
    return self.array.dtype

  @property
  def shape(self):
    return self.array.shape

  @property
  def data(self):
    return self.array.data

  @property
  def offset(self):
    return self.array.offset

  @property
  def grid(self):
    return self.array.grid

  def shift(self, offset: int, axis: int, mode=None) -> GridArray:
    return self.bc.shift(self.array, offset, axis, mode)

  def trim_boundary(self) -> GridArray:
    return self.array  # periodic: every point is interior (grids.py:372-386)

  def impose_bc(self) -> 'GridVariable':
    return self.bc.impose_bc(self.array)


GridVariableVector = Tuple[GridVariable, ...]


def averaged_offset(*arrays) -> Tuple[float, ...]:
  """grids.py:439-445."""
  return tuple(float(o) for o in np.mean([a.offset for a in arrays], axis=0))


def control_volume_offsets(c) -> Tuple[Tuple[float, ...], ...]:
  """Offsets of the faces of the control volume centred on `c` (grids.py:448-454)."""
  return tuple(tuple(o + .5 if i == j else o for i, o in enumerate(c.offset))
               for j in range(len(c.offset)))


def consistent_offset(*arrays) -> Tuple[float, ...]:
  """grids.py:461-469."""
  offsets = {a.offset for a in arrays}
  if len(offsets) != 1:
    raise InconsistentOffsetError(f'arrays do not have a unique offset: {offsets}')
  return offsets.pop()


def consistent_grid(*arrays) -> Grid:
  """grids.py:476-482."""
  grids = {a.grid for a in arrays}
  if len(grids) != 1:
    raise InconsistentGridError(f'arrays do not have a unique grid: {grids}')
  return grids.pop()


def unique_boundary_conditions(*arrays) -> BoundaryConditions:
  """grids.py:489-496."""
  bcs = {a.bc for a in arrays}
  if len(bcs) != 1:
    raise InconsistentBoundaryConditionsError(f'arrays do not have unique bc: {bcs}')
  return bcs.pop()


def applied(func):
  """grids.py:401-418: lift an array function to GridArrays (host data)."""
  def wrapper(*args, **kwargs):
    gas = [a for a in args if isinstance(a, GridArray)]
    offset = consistent_offset(*gas)
    grid = consistent_grid(*gas)
    raw = [a.data if isinstance(a, GridArray) else a for a in args]
    return GridArray(func(*raw, **kwargs), offset, grid)
  return wrapper
