"""jax_cfd/base/initial_conditions.py: wrap_variables and filtered_velocity_field ("next" row)."""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np

from . import _engine
from . import boundaries
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
  jax.random: same distribution, different bits).  The spectral filter (filter_utils.py:32-42)
  runs once on the host; the project-and-normalise iterations use the device projection and the
  fused max-speed reduction."""
  rs = np.random.RandomState(int(rng_key))
  freqs = np.meshgrid(*[2 * np.pi * np.fft.fftfreq(n, s) for n, s in zip(grid.shape, grid.step)],
                      indexing='ij')
  k = np.sqrt(sum(f ** 2 for f in freqs))
  with np.errstate(divide='ignore', invalid='ignore'):
    filt = np.where(k > 0, _log_normal_pdf(k, peak_wavenumber) / k ** (grid.ndim - 1), 0.0)
  comps = []
  for _ in range(grid.ndim):
    noise = rs.standard_normal(grid.shape)
    comps.append(np.fft.ifftn(np.fft.fftn(noise) * filt).real.astype(np.float32))
  bcs = [boundaries.periodic_boundary_conditions(grid.ndim)] * grid.ndim
  v = wrap_variables(comps, grid, bcs)
  project = _engine.NativeProjection(grid)
  for _ in range(iterations):
    v = project(v)
    vmax = float(np.sqrt(_engine.diagnostics(v)['max_speed_sq']))
    v = tuple(grids.GridVariable(grids.GridArray(
        (np.float32(maximum_velocity) * np.asarray(u.data) / np.float32(vmax)).astype(np.float32),
        u.offset, u.grid), u.bc) for u in v)
  return v
