"""jax_cfd/base/fast_diagonalization.py: functions of separable linear operators on the device.

`transform(func, operators, dtype, ...)` returns `apply(rhs)` computing
F(A (x) I + I (x) B) rhs = (X_A (x) X_B) func(L_A (+) L_B) (X_A^-1 (x) X_B^-1) rhs  (fast_diagonalization.py:28-126).
The eigen-decomposition and `func` run once on the host in float64 exactly like the reference's
trace-time code; `apply` runs the CUDA kernels:
  'rfft' / 'fft'  the shared-memory line FFTs (csrc/poisson_2d.cu, poisson_3d.cu) with the diagonal
                  read from a table; circulant operators whose func(eigenvalues) is real
                  (symmetric operators), every axis a power of two;
  'matmul'        eigenvector products along each axis on the FP64 tensor cores
                  (csrc/transform_matmul.cu); any hermitian operators, any shape.
Default: 'rfft' where its shape constraint holds, else 'matmul' (the reference falls back the same way
when the last axis is odd, fast_diagonalization.py:107-108).
"""
from __future__ import annotations

import functools
from typing import Callable, Optional, Sequence

import numpy as np

from . import _engine
from . import _lib
from . import grids


def _is_circulant(op: np.ndarray) -> bool:
  n = op.shape[0]
  idx = (np.arange(n)[:, None] - np.arange(n)[None, :]) % n
  return bool(np.array_equal(op, op[:, 0][idx]))


def transform(func: Callable[[np.ndarray], np.ndarray], operators: Sequence[np.ndarray], dtype,
              *, hermitian: bool = False, circulant: bool = False,
              implementation: Optional[str] = None, precision=None) -> Callable:
  """fast_diagonalization.py:28-126.  `precision` is accepted for signature parity: the matmul
  products always run in float64 (FP64 DMMA), above `Precision.HIGHEST`."""
  del precision
  operators = [np.asarray(op) for op in operators]
  if any(op.ndim != 2 or op.shape[0] != op.shape[1] for op in operators):
    raise ValueError('operators are not all square matrices. Shapes are '
                     + ', '.join(str(op.shape) for op in operators))
  if np.dtype(dtype) != np.float32:
    raise NotImplementedError(f'the B200 path computes in float32; got dtype={dtype}')
  if len(operators) == 1:
    # one axis: A = 0 (x) I + I (x) A on a grid with a leading axis of one cell
    inner = transform(func, [np.zeros((1, 1)), operators[0]], dtype, hermitian=hermitian,
                      circulant=circulant, implementation=implementation)
    def apply_1d(rhs):
      if isinstance(rhs, np.ndarray):
        return inner(rhs[None])[0]
      view = _lib.DeviceArray((1,) + tuple(rhs.shape), np.float32, ptr=_lib.device_ptr(rhs), owner=rhs,
                              device=_lib.device_of(rhs))
      out = inner(view)
      return _lib.DeviceArray(tuple(rhs.shape), np.float32, ptr=out.ptr, owner=out, device=out.device)
    return apply_1d
  if len(operators) != 2 and len(operators) != 3:
    raise NotImplementedError('fast diagonalisation is implemented for up to 3 operators')
  shape = tuple(op.shape[0] for op in operators)
  if implementation is None:
    implementation = 'rfft' if (circulant and _engine.fft_shape_ok(shape)) else 'matmul'
  grid = grids.Grid(shape)  # only the shape matters: the plan's own Laplacian tables are not used

  if implementation == 'matmul':
    if not hermitian:
      raise ValueError('non-hermitian operators not yet supported with implementation="matmul"')
    eigenvalues, eigenvectors = zip(*map(np.linalg.eigh, operators))
    summed = functools.reduce(np.add.outer, eigenvalues)
    diagonals = np.asarray(func(summed))
    if diagonals.shape != summed.shape:
      raise ValueError('output shape from func() does not match input shape: '
                       f'{diagonals.shape} vs {summed.shape}')
    _engine.check_implementation(grid, 'matmul')
    return _engine.NativeTransform(grid, diagonals, eigvecs=eigenvectors)

  if implementation in ('fft', 'rfft'):
    if not circulant:
      raise ValueError(f'non-circulant operators not yet supported with implementation="{implementation}"')
    if implementation == 'rfft' and shape[-1] % 2:
      raise ValueError('implementation="rfft" currently requires an even size for the last axis')
    # https://en.wikipedia.org/wiki/Circulant_matrix#Eigenvectors_and_eigenvalues
    eigenvalues = ([np.fft.fft(op[:, 0]) for op in operators[:-1]] + [np.fft.rfft(operators[-1][:, 0])])
    summed = functools.reduce(np.add.outer, eigenvalues)
    diagonals = np.asarray(func(summed))
    if diagonals.shape != summed.shape:
      raise ValueError('output shape from func() does not match input shape: '
                       f'{diagonals.shape} vs {summed.shape}')
    scale = max(1.0, float(np.abs(diagonals).max()))
    if np.iscomplexobj(diagonals) and float(np.abs(diagonals.imag).max()) > 1e-6 * scale:
      raise NotImplementedError(
          'the line-FFT kernels take a real diagonal (symmetric circulant operators); use '
          'implementation="matmul" for hermitian operators')
    if not _engine.fft_shape_ok(shape):
      raise NotImplementedError(
          f'implementation="{implementation}" runs on the radix-2 line-FFT kernels: every axis a power of '
          f'two >= 16 (>= 32 on the last axis); got {shape}.  Hermitian operators of any shape: "matmul".')
    return _engine.NativeTransform(grid, np.real(diagonals))

  raise ValueError(f'invalid implementation: {implementation}')


def pseudoinverse(operators: Sequence[np.ndarray], dtype, *, hermitian: bool = False,
                  circulant: bool = False, implementation: Optional[str] = None, precision=None,
                  cutoff: Optional[float] = None) -> Callable:
  """fast_diagonalization.py:228-266: eigenvalues below `cutoff` (default 10 * eps) are discarded."""
  if cutoff is None:
    cutoff = 10 * np.finfo(np.dtype(dtype)).eps

  def func(v):
    with np.errstate(divide='ignore', invalid='ignore'):
      return np.where(abs(v) > cutoff, 1 / v, 0)

  return transform(func, operators, dtype, hermitian=hermitian, circulant=circulant,
                   implementation=implementation, precision=precision)
