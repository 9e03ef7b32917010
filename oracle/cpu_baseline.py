"""Multi-threaded CPU implementation of the step built from the oracle.  TEST INFRASTRUCTURE.

`CpuStep` = oracle/cfd_oracle_c.c (OpenMP stencils) + scipy.fft.rfft2/irfft2 on all cores with the
oracle's own pseudo-inverse diagonal (cfd_oracle.pinv_diagonals).  bench.py times it as the
`cpu_baseline` / `--impl reference` arm (kind "port": the reference's jitted JAX CPU path cannot be
run because JAX is absent from this image); tests/test_oracle_c.py checks it against the NumPy
oracle and the golden vectors.
"""
import ctypes
import os

import numpy as np
import scipy.fft

import cfd_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, '_build', 'libcfd_oracle.so')
_f = ctypes.c_float
_p = ctypes.c_void_p
_i = ctypes.c_int


def load():
  if not os.path.exists(SO):
    import subprocess
    subprocess.run(['make', '-C', HERE], check=True)
  lib = ctypes.CDLL(SO)
  lib.oracle_explicit_2d.argtypes = [_p, _p, _p, _p, _i, _i, _f, _f, _f, _f, _f, _i, _f, _f, _p, _p, _i, _f]
  lib.oracle_divergence_2d.argtypes = [_p, _p, _p, _i, _i, _f, _f]
  lib.oracle_correct_2d.argtypes = [_p, _p, _p, _p, _p, _i, _i, _f, _f]
  for f in (lib.oracle_explicit_2d, lib.oracle_divergence_2d, lib.oracle_correct_2d):
    f.restype = None
  lib.oracle_set_threads.argtypes = [_i]
  lib.oracle_set_threads.restype = None
  return lib


class CpuStep:
  def __init__(self, shape, h, dt, density=1.0, viscosity=None, const_force=None, linear=None,
               workers=None):
    self.lib = load()
    self.shape, self.h, self.dt = tuple(shape), h, dt
    self.density, self.viscosity = density, viscosity
    self.const_force = None if const_force is None else tuple(
        None if f is None else np.ascontiguousarray(f, np.float32) for f in const_force)
    self.linear = linear
    self.workers = workers or os.cpu_count()
    # not OMP_NUM_THREADS: torchrun sets it to 1 for its workers
    self.lib.oracle_set_threads(int(self.workers))
    self.diag = cfd_oracle.pinv_diagonals(shape, h, np.float32)
    nx, ny = shape
    self.us = np.empty(shape, np.float32)
    self.vs = np.empty(shape, np.float32)
    self.rhs = np.empty(shape, np.float32)

  def step(self, u, v, uo=None, vo=None):
    nx, ny = self.shape
    hx, hy = self.h
    uo = np.empty_like(u) if uo is None else uo
    vo = np.empty_like(v) if vo is None else vo
    fu = fv = None
    if self.const_force is not None:
      fu, fv = (None if f is None else f.ctypes.data for f in self.const_force)
    nu = 0.0 if self.viscosity is None else self.viscosity / self.density
    self.lib.oracle_explicit_2d(u.ctypes.data, v.ctypes.data, self.us.ctypes.data, self.vs.ctypes.data,
                                nx, ny, self.dt, self.dt / hx, self.dt / hy, hx, hy,
                                0 if self.viscosity is None else 1, nu, self.density, fu, fv,
                                0 if self.linear is None else 1, 0.0 if self.linear is None else self.linear)
    self.lib.oracle_divergence_2d(self.us.ctypes.data, self.vs.ctypes.data, self.rhs.ctypes.data,
                                  nx, ny, hx, hy)
    spec = scipy.fft.rfft2(self.rhs, workers=self.workers)
    spec *= self.diag
    q = scipy.fft.irfft2(spec, s=self.shape, workers=self.workers).astype(np.float32, copy=False)
    self.lib.oracle_correct_2d(self.us.ctypes.data, self.vs.ctypes.data, q.ctypes.data,
                               uo.ctypes.data, vo.ctypes.data, nx, ny, hx, hy)
    self.q = q
    return uo, vo
