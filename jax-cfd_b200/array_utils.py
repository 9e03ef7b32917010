"""jax_cfd/base/array_utils.py: the operator builders the fast-diagonalisation callers use."""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from . import boundaries
from . import grids


def laplacian_matrix(size: int, step: float) -> np.ndarray:
  """array_utils.py:168-173: periodic 1-D Laplacian, circulant column [-2, 1, 0, ..., 0, 1] / step^2
  (dense; the CUDA pressure solve never builds it -- its eigenvalues are formed analytically)."""
  column = np.zeros(size)
  column[0] = -2 / step ** 2
  column[1] = column[-1] = 1 / step ** 2
  idx = (np.arange(size)[:, None] - np.arange(size)[None, :]) % size  # scipy.linalg.circulant
  return column[idx]


def laplacian_matrix_w_boundaries(grid: grids.Grid, offset, bc) -> List[np.ndarray]:
  """array_utils.py:246-290 for the boundary conditions this path supports (all periodic: the
  circulant operators are returned unchanged, array_utils.py:268-270)."""
  del offset
  if not isinstance(bc, boundaries.ConstantBoundaryConditions) or any(
      t != (boundaries.BCType.PERIODIC, boundaries.BCType.PERIODIC) for t in bc.types):
    raise NotImplementedError('only periodic boundary conditions are implemented on the B200 path')
  return [laplacian_matrix(n, s) for n, s in zip(grid.shape, grid.step)]
