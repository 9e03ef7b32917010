"""`jax.ffi` registration of the B200 step (import-guarded: JAX is absent from this image).

With JAX installed and `csrc/xla_ffi_shim.cc` built, `register()` makes `b200cfd_step2d` available
as an XLA custom call and `semi_implicit_navier_stokes` below returns a traceable `step_fn` with
the reference's pytree types, composable with `jax.jit` and `funcutils.repeated`.
"""
from __future__ import annotations

import ctypes
import os

try:  # pragma: no cover - not executable in this image
  import jax
  HAVE_JAX = True
except ImportError:
  jax = None
  HAVE_JAX = False

_XLA_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'lib', 'libcfd_b200_xla.so')


def register():  # pragma: no cover
  if not HAVE_JAX:
    raise ImportError('jax is not installed: the jax.ffi binding cannot be registered '
                      '(use the ctypes-driven jax_cfd_b200.equations API instead)')
  lib = ctypes.cdll.LoadLibrary(_XLA_LIB)
  jax.ffi.register_ffi_target('b200cfd_step2d', jax.ffi.pycapsule(lib.B200CfdStep2D),
                              platform='CUDA')


def semi_implicit_navier_stokes(density, viscosity, dt, grid, forcing=None):  # pragma: no cover
  """Traceable drop-in for jax_cfd.base.equations.semi_implicit_navier_stokes (equations.py:120)."""
  if not HAVE_JAX:
    raise ImportError('jax is not installed')
  import numpy as np
  from jax_cfd.base import grids as ref_grids  # the reference's own pytree types
  from . import _engine
  from . import grids as my_grids
  register()
  g = my_grids.Grid(grid.shape, domain=grid.domain)
  params, keep = _engine.make_params(g, dt, density, viscosity, _engine.as_forcing(forcing))
  plans = {}

  def step_fn(v):
    batch = int(np.prod(v[0].data.shape[:-g.ndim], dtype=np.int64)) if v[0].data.ndim > g.ndim else 1
    plan = plans.setdefault(batch, _engine.get_plan(g, batch))
    outs = jax.ffi.ffi_call(
        'b200cfd_step2d',
        tuple(jax.ShapeDtypeStruct(u.data.shape, np.float32) for u in v),
        vmap_method='broadcast_all')(
            *[u.data for u in v], plan=np.int64(plan.handle.value),
            params=np.int64(ctypes.addressof(params)), nsteps=np.int32(1))
    return tuple(ref_grids.GridVariable(ref_grids.GridArray(a, u.offset, u.grid), u.bc)
                 for a, u in zip(outs, v))

  step_fn._keepalive = (params, keep, plans)
  return step_fn
