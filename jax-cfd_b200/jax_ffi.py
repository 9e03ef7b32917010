"""`jax.ffi` registration of the B200 step (import-guarded: JAX is absent from this image).

With JAX installed and `csrc/xla_ffi_shim.cc` built (`libcfd_b200_xla.so`), `register()` makes the
four handlers of the shim available as XLA custom calls and `semi_implicit_navier_stokes` /
`projection` below return traceable functions with the reference's pytree types, composable with
`jax.jit`, `jax.vmap` (leading batch axes) and `funcutils.repeated`.

Everything that defines the equation travels as typed scalar ATTRIBUTES of the custom call
(`step_attrs`), never as host pointers: the calls serialise, so the persistent compilation cache and
multi-process XLA work.  Forcing tables (separable profiles, constant fields) travel as operands.
`step_attrs` and `forcing_operands` are plain NumPy and are exercised by tests/test_host_logic.py
without JAX; the handler bodies are exercised by tests/test_xla_ffi_shim.py against a stand-in for
the XLA FFI header.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional, Tuple

import numpy as np

from . import _engine
from . import _lib
from . import grids as my_grids

try:  # pragma: no cover - not executable in this image
  import jax
  HAVE_JAX = True
except ImportError:
  jax = None
  HAVE_JAX = False

_XLA_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'lib', 'libcfd_b200_xla.so')
TARGETS = {'b200cfd_step2d': 'B200CfdStep2D', 'b200cfd_step3d': 'B200CfdStep3D',
           'b200cfd_project2d': 'B200CfdProject2D', 'b200cfd_project3d': 'B200CfdProject3D'}
_registered = False


def step_attrs(grid: my_grids.Grid, dt: float, density: float, viscosity: Optional[float],
               forcing: Optional[_engine.ForcingFn], nsteps: int = 1, convect_dt: Optional[float] = None,
               implementation: int = _lib.IMPL_AUTO) -> Dict[str, object]:
  """The attribute dictionary of `b200cfd_step{2,3}d` (names and types of CFD_STEP_ATTRS in
  csrc/xla_ffi_shim.cc; same content as struct cfd_params, include/cfd_b200.h)."""
  nd = grid.ndim
  terms = () if forcing is None else forcing.terms
  if len(terms) > _engine.MAX_TERMS:
    raise NotImplementedError(f'at most {_engine.MAX_TERMS} forcing terms are supported')
  if len({t.kind for t in terms}) != len(terms):
    raise NotImplementedError('each kind of forcing term may appear once in a sum')
  linear_coef = smag_cs = 0.0
  sep_scale = [0.0] * nd
  sep_mask = sep_has = field_mask = 0
  for t in terms:
    if t.kind == _engine.FORCE_SEPARABLE:
      for a in range(nd):
        sep_has |= (1 << a) if t.has[a] else 0
        sep_scale[a] = float(t.scales[a])
        for j in range(nd):
          if t.profiles[a][j] is not None:
            sep_mask |= 1 << (a * nd + j)
    elif t.kind == _engine.FORCE_FIELD:
      for a in range(nd):
        field_mask |= (1 << a) if t.arrays[a] is not None else 0
    elif t.kind == _engine.FORCE_LINEAR:
      linear_coef = t.coef
    elif t.kind == _engine.FORCE_SMAGORINSKY:
      smag_cs = t.cs
  return dict(step=np.asarray(grid.step, np.float64), nsteps=np.int32(nsteps),
              implementation=np.int32(implementation), dt=np.float64(dt),
              convect_dt=np.float64(0.0 if convect_dt is None else convect_dt),
              density=np.float64(density),
              viscosity=np.float64(-1.0 if viscosity is None else viscosity),  # < 0: viscosity=None
              linear_coef=np.float64(linear_coef), smagorinsky_cs=np.float64(smag_cs),
              terms=np.asarray([t.kind for t in terms], np.int32),
              sep_scale=np.asarray(sep_scale, np.float64), sep_mask=np.int64(sep_mask),
              sep_has=np.int64(sep_has), field_mask=np.int64(field_mask))


def forcing_operands(grid: my_grids.Grid, forcing: Optional[_engine.ForcingFn]) -> Tuple[np.ndarray, np.ndarray]:
  """(sep_prof, field) operands of the step handlers: the profiles present in `sep_mask` as rows of
  a (count, max N) array in (component, axis) order, and the constant fields present in
  `field_mask` stacked as (count, *grid.shape); empty arrays when the term is absent."""
  nd, nmax = grid.ndim, max(grid.shape)
  rows, fields = [], []
  for t in (() if forcing is None else forcing.terms):
    if t.kind == _engine.FORCE_SEPARABLE:
      for a in range(nd):
        for j in range(nd):
          p = t.profiles[a][j]
          if p is not None:
            row = np.zeros(nmax, np.float32)
            row[:grid.shape[j]] = np.asarray(p, np.float32)
            rows.append(row)
    elif t.kind == _engine.FORCE_FIELD:
      fields += [np.asarray(a, np.float32) for a in t.arrays if a is not None]
  sep = np.stack(rows) if rows else np.zeros((0, nmax), np.float32)
  fld = np.stack(fields) if fields else np.zeros((0,) + tuple(grid.shape), np.float32)
  return sep, fld


def register():  # pragma: no cover
  global _registered
  if not HAVE_JAX:
    raise ImportError('jax is not installed: the jax.ffi binding cannot be registered '
                      '(use the ctypes-driven jax_cfd_b200.equations API instead)')
  if _registered:
    return
  lib = ctypes.cdll.LoadLibrary(_XLA_LIB)
  for target, symbol in TARGETS.items():
    jax.ffi.register_ffi_target(target, jax.ffi.pycapsule(getattr(lib, symbol)), platform='CUDA')
  _registered = True


def semi_implicit_navier_stokes(density, viscosity, dt, grid, forcing=None, nsteps=1):  # pragma: no cover
  """Traceable drop-in for jax_cfd.base.equations.semi_implicit_navier_stokes (equations.py:120);
  `nsteps` > 1 runs funcutils.repeated(step_fn, nsteps) inside one custom call (lazy projection)."""
  if not HAVE_JAX:
    raise ImportError('jax is not installed')
  from jax_cfd.base import grids as ref_grids  # the reference's own pytree types
  register()
  g = my_grids.Grid(grid.shape, domain=grid.domain)
  f = _engine.as_forcing(forcing)
  attrs = step_attrs(g, dt, density, viscosity, f, nsteps)
  sep, fld = (jax.numpy.asarray(a) for a in forcing_operands(g, f))
  target = 'b200cfd_step2d' if g.ndim == 2 else 'b200cfd_step3d'

  def step_fn(v):
    outs = jax.ffi.ffi_call(
        target, tuple(jax.ShapeDtypeStruct(u.data.shape, np.float32) for u in v),
        vmap_method='broadcast_all')(*[u.data for u in v], sep, fld, **attrs)
    return tuple(ref_grids.GridVariable(ref_grids.GridArray(a, u.offset, u.grid), u.bc)
                 for a, u in zip(outs, v))

  return step_fn


def projection(v, implementation: int = _lib.IMPL_AUTO):  # pragma: no cover
  """Traceable drop-in for jax_cfd.base.pressure.projection with solve_fast_diag
  (pressure.py:181-198); returns (projected velocity, q)."""
  if not HAVE_JAX:
    raise ImportError('jax is not installed')
  from jax_cfd.base import grids as ref_grids
  register()
  grid = v[0].grid
  target = 'b200cfd_project2d' if grid.ndim == 2 else 'b200cfd_project3d'
  shapes = tuple(jax.ShapeDtypeStruct(u.data.shape, np.float32) for u in v)
  outs = jax.ffi.ffi_call(target, shapes + (shapes[0],), vmap_method='broadcast_all')(
      *[u.data for u in v], step=np.asarray(grid.step, np.float64), implementation=np.int32(implementation))
  vp = tuple(ref_grids.GridVariable(ref_grids.GridArray(a, u.offset, u.grid), u.bc)
             for a, u in zip(outs[:-1], v))
  return vp, outs[-1]
