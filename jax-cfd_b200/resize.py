"""jax_cfd/base/resize.py (the flux-preserving coarse-graining used by `trajectory(post_process=...)`
when training data is generated) and the 2-D vorticity of
jax_cfd/data/xarray_utils.py:155-163, evaluated on the device next to the step."""
from __future__ import annotations

import ctypes
from typing import Sequence, Tuple

import numpy as np

from . import _lib
from . import grids
from ._lib import DeviceArray, check, lib


def _as_device(a) -> Tuple[object, bool]:
  """(device array, came_from_host): host arrays are uploaded -- there is no CPU implementation."""
  if _lib.is_device_array(a):
    return a, False
  return DeviceArray.from_numpy(np.ascontiguousarray(a, np.float32)), True


def downsample_staggered_velocity_component(u, direction: int, factor: int, ndim: int = None):
  """resize.py:38-74: keep the faces of `u` (velocities in `direction`) that lie on a coarse face and
  average the factor**(ndim-1) of them that tile it.  `u`: (..., *grid shape); leading axes are
  batch / time.  Device arrays in -> device array out; NumPy in -> NumPy out."""
  shape = tuple(u.shape)
  nd = len(shape) if ndim is None else int(ndim)
  if nd not in (2, 3) or len(shape) < nd:
    raise NotImplementedError('downsampling is implemented for 2-D and 3-D grids')
  gshape, lead = shape[-nd:], shape[:-nd]
  if any(n % factor for j, n in enumerate(gshape) if j != direction):
    raise ValueError(f'`block_size` must divide `array.shape`;got {factor}, {gshape}.')  # array_utils.py:159
  dev, from_host = _as_device(u)
  out = DeviceArray(lead + tuple(n // factor for n in gshape), device=_lib.device_of(dev))
  cshape = (ctypes.c_int64 * nd)(*gshape)
  batch = int(np.prod(lead, dtype=np.int64)) if lead else 1
  stream = _lib.stream_of(dev)
  check(lib().cfd_downsample_component(stream, _lib.device_ptr(dev), out.ptr, nd, cshape, batch, int(direction),
                                       int(factor)))
  return out.numpy() if from_host else out


def downsample_staggered_velocity(source_grid: grids.Grid, destination_grid: grids.Grid, velocity: Sequence):
  """resize.py:187-222: every component `j` is coarse-grained along its own direction `j`; GridVariable
  / GridArray inputs come back on `destination_grid` with their offset (and bc) unchanged."""
  factor = destination_grid.step[0] / source_grid.step[0]
  assert destination_grid.domain == source_grid.domain
  assert round(factor) == factor, factor
  f = int(round(factor))
  result = []
  for j, u in enumerate(velocity):
    if isinstance(u, (grids.GridVariable, grids.GridArray)):
      if u.grid != source_grid:
        raise grids.InconsistentGridError(
            f'source_grid for downsampling is {source_grid}, but u is defined on {u.grid}')
      arr = grids.GridArray(downsample_staggered_velocity_component(u.data, j, f, source_grid.ndim), u.offset,
                            destination_grid)
      result.append(grids.GridVariable(arr, u.bc) if isinstance(u, grids.GridVariable) else arr)
    else:
      result.append(downsample_staggered_velocity_component(u, j, f, source_grid.ndim))
  return tuple(result)


def vorticity_2d(v: Sequence):
  """(D+_x v - D+_y u) at offset (1, 1), periodic (data/xarray_utils.py:155-163); `v` = (u, v)
  GridVariables / GridArrays of one 2-D grid.  Returns a GridArray."""
  u, w = v
  grid = grids.consistent_grid(u, w)
  if grid.ndim != 2:
    raise ValueError('vorticity_2d needs a 2-D velocity field')
  du, hu = _as_device(u.data)
  dw, hw = _as_device(w.data)
  lead = tuple(du.shape)[:-2]
  out = DeviceArray(tuple(du.shape), device=_lib.device_of(du))
  cshape = (ctypes.c_int64 * 2)(*grid.shape)
  batch = int(np.prod(lead, dtype=np.int64)) if lead else 1
  check(lib().cfd_vorticity_2d(_lib.stream_of(du), _lib.device_ptr(du), _lib.device_ptr(dw), out.ptr, cshape, batch,
                               float(grid.step[0]), float(grid.step[1])))
  data = out.numpy() if (hu and hw) else out
  return grids.GridArray(data, (1.0, 1.0), grid)
