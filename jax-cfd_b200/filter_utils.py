"""jax_cfd/base/filter_utils.py: spectral filtering on the device."""
from __future__ import annotations

from typing import Callable

import numpy as np

from . import _engine
from . import grids

_filters = {}


def _real_fourier_basis(n: int):
  """Orthonormal real Fourier basis of a periodic axis (columns: constant, cos / sin pairs of
  wavenumber index m = 1 .. ceil(n/2) - 1, alternating vector for even n) and each column's m."""
  i = np.arange(n)
  cols, ms = [np.full(n, 1 / np.sqrt(n))], [0]
  for m in range(1, (n + 1) // 2):
    ang = 2 * np.pi * ((m * i) % n) / n
    cols += [np.sqrt(2 / n) * np.cos(ang), np.sqrt(2 / n) * np.sin(ang)]
    ms += [m, m]
  if n % 2 == 0:
    cols.append(np.where(i % 2, -1.0, 1.0) / np.sqrt(n))
    ms.append(n // 2)
  return np.stack(cols, axis=1), np.array(ms)


def _filter_op(spectral_density: Callable, grid: grids.Grid, key):
  op = _filters.get(key) if key is not None else None
  if op is not None:
    return op
  with np.errstate(divide='ignore', invalid='ignore'):
    if _engine.fft_shape_ok(grid.shape):
      # |k| on the rfftn layout: fftfreq on the leading axes, rfftfreq on the last (the filter only
      # depends on |k|, so the half spectrum carries all of it: filter_utils.py:40-42)
      freqs = [2 * np.pi * np.fft.fftfreq(n, s) for n, s in zip(grid.shape[:-1], grid.step[:-1])]
      freqs.append(2 * np.pi * np.fft.rfftfreq(grid.shape[-1], grid.step[-1]))
      k = np.sqrt(sum(f ** 2 for f in np.meshgrid(*freqs, indexing='ij')))
      filters = np.where(k > 0, spectral_density(k), 0.0)
      op = _engine.NativeTransform(grid, filters)
    else:
      # other shapes: the same multiplier in the real Fourier basis, applied by the matmul kernels
      bases = [_real_fourier_basis(n) for n in grid.shape]
      freqs = [2 * np.pi * ms / (n * s) for (_, ms), n, s in zip(bases, grid.shape, grid.step)]
      k = np.sqrt(sum(f ** 2 for f in np.meshgrid(*freqs, indexing='ij')))
      filters = np.where(k > 0, spectral_density(k), 0.0)
      op = _engine.NativeTransform(grid, filters, eigvecs=[b for b, _ in bases])
  if key is not None:
    _filters[key] = op
  return op


def filter(spectral_density: Callable, array, grid: grids.Grid, cache_key=None):  # pylint: disable=redefined-builtin
  """filter_utils.py:32-42: ifftn(fftn(array) * where(|k| > 0, spectral_density(|k|), 0)).real, as one
  table-driven device transform (`array`: float32, numpy or device, shape (..., *grid.shape))."""
  key = None if cache_key is None else (grid, cache_key)
  return _filter_op(spectral_density, grid, key)(array)
