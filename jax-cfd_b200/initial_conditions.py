"""jax_cfd/base/initial_conditions.py: wrap_variables and filtered_velocity_field ("next" row)."""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np

from . import _engine
from . import _lib
from . import boundaries
from . import filter_utils
from . import grids


def wrap_variables(var, grid: grids.Grid, bcs, offsets=None, batch_dim: bool = False):
  """initial_conditions.py:40-57."""
  if offsets is None:
    offsets = grid.cell_faces
  return tuple(bc.impose_bc(grids.GridArray(u, o, grid)) for u, o, bc in zip(var, offsets, bcs))


def _log_normal_pdf(x, mode, variance=.25):
  """initial_conditions.py:60-64."""
  mean = np.log(mode) + variance
  with np.errstate(divide='ignore', invalid='ignore'):
    logx = np.log(x)
    return np.exp(-(mean - logx) ** 2 / 2 / variance - logx)


def filtered_velocity_field(rng_key, grid: grids.Grid, maximum_velocity: float = 1,
                            peak_wavenumber: float = 3, iterations: int = 3):
  """initial_conditions.py:71-121.  `rng_key` is an int seed (numpy RandomState stands in for
  jax.random: same distribution, different bits).  Everything after the noise runs on the device:
  the spectral filter (filter_utils.py:32-42) is one table-driven transform per component, and each
  project-and-normalise iteration is the device projection, the fused max-speed reduction and one
  scaling kernel; the state never returns to the host (device arrays are returned)."""
  rs = np.random.RandomState(int(rng_key))

  def spectral_density(k):
    return _log_normal_pdf(k, peak_wavenumber) / k ** (grid.ndim - 1)

  comps = []
  for _ in range(grid.ndim):
    noise = _lib.DeviceArray.from_numpy(rs.standard_normal(grid.shape).astype(np.float32))
    comps.append(filter_utils.filter(spectral_density, noise, grid,
                                     cache_key=('log_normal', float(peak_wavenumber))))
  bcs = [boundaries.periodic_boundary_conditions(grid.ndim)] * grid.ndim
  v = wrap_variables(comps, grid, bcs)
  project = _engine.NativeProjection(grid)
  for _ in range(iterations):
    v = project(v)
    vmax = float(np.sqrt(_engine.diagnostics(v)['max_speed_sq']))  # initial_conditions.py:67-68
    v = _engine.scale(v, maximum_velocity, vmax)                    # maximum_velocity * u / vmax
  return v
