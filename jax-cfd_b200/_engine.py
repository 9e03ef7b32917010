"""Plan cache, parameter marshalling and the native callables behind the reference-shaped API."""
from __future__ import annotations

import ctypes
import threading
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib
from . import boundaries
from . import grids
from ._lib import (FORCE_FIELD, FORCE_LINEAR, FORCE_SEPARABLE, FORCE_SMAGORINSKY, MAX_DIM,
                   MAX_TERMS, CfdError, DeviceArray, Params, check, lib)

_plans = {}
_plans_lock = threading.Lock()


class Plan:
  """Owns one cfd_plan (tables + workspace) for a (grid, batch, device)."""

  def __init__(self, grid: grids.Grid, batch: int, device: int = 0, implementation: int = _lib.IMPL_AUTO):
    _lib.require_device()
    self.grid, self.batch, self.device = grid, batch, device
    shape = (ctypes.c_int64 * grid.ndim)(*grid.shape)
    step = (ctypes.c_double * grid.ndim)(*grid.step)
    handle = ctypes.c_void_p()
    check(lib().cfd_plan_create_impl(ctypes.byref(handle), grid.ndim, shape, step, batch, device,
                                     implementation))
    self.handle = handle
    self.implementation = lib().cfd_plan_implementation(handle)  # IMPL_RFFT or IMPL_MATMUL

  def __del__(self):
    h = getattr(self, 'handle', None)
    try:
      if h and _lib is not None and _lib._lib is not None:
        _lib._lib.cfd_plan_destroy(h)
    except Exception:  # interpreter shutdown
      pass
    self.handle = None


def get_plan(grid: grids.Grid, batch: int = 1, device: Optional[int] = None, stream=None,
             implementation: int = _lib.IMPL_AUTO) -> Plan:
  """The plan (tables + ONE workspace) for this grid on `device` (default: the current device).
  A plan's workspace may only be used by work ordered on one stream, so plans are cached per
  stream: callers that enqueue on different streams get independent workspaces."""
  if device is None:
    device = _lib.current_device()
  key = (grid.shape, grid.step, batch, device, stream, implementation)
  with _plans_lock:
    p = _plans.get(key)
    if p is None:
      p = _plans[key] = Plan(grid, batch, device, implementation)
    return p


def plan_for(grid: grids.Grid, batch: int, arrays, stream=None, implementation: int = _lib.IMPL_AUTO) -> Plan:
  """Plan on the device that owns `arrays` (the first device array; host arrays -> current)."""
  for a in arrays:
    if _lib.is_device_array(a):
      return get_plan(grid, batch, _lib.device_of(a), stream, implementation)
  return get_plan(grid, batch, None, stream, implementation)


def fft_shape_ok(shape) -> bool:
  """The shapes the radix-2 line-FFT kernels take (cfd_plan_create_impl, CFD_IMPL_RFFT)."""
  shape = tuple(shape)
  pow2 = all(n >= 16 and n & (n - 1) == 0 for n in shape)
  mid_ok = all(n <= 1 << 14 for n in shape[1:-1])
  return (pow2 and shape[-1] >= 32 and shape[-1] <= 1 << 15 and mid_ok and
          shape[0] <= (1 << 15 if len(shape) == 2 else 1 << 14))


def check_implementation(grid: grids.Grid, implementation) -> int:
  """Build-time validation of (grid, implementation): the errors cfd_plan_create_impl would raise at
  the first call, raised when the step function is built."""
  code = _lib.implementation_code(implementation)
  if code == _lib.IMPL_RFFT and not fft_shape_ok(grid.shape):
    raise NotImplementedError(
        f'implementation={implementation!r} needs every axis a power of two >= 16 (>= 32 on the last '
        f'axis); got {grid.shape}.  Use implementation=None or "matmul".')
  if (code == _lib.IMPL_MATMUL or not fft_shape_ok(grid.shape)) and any(n > 4096 for n in grid.shape):
    raise NotImplementedError('the matmul implementation takes axes of up to 4096 cells')
  return code


def clear_plans():
  with _plans_lock:
    _plans.clear()


# ----------------------------------------------------------------------------- forcing terms
class Term:
  kind = 0


class SeparableTerm(Term):
  """f_a = (prod_j prof[a][j][i_j]) * scale[a]; prof entries None = ones; has[a] False = zero."""
  kind = FORCE_SEPARABLE

  def __init__(self, grid, profiles, scales, has, offsets):
    self.grid, self.profiles, self.scales, self.has, self.offsets = grid, profiles, scales, has, offsets
    self._dev = None

  def device_tables(self):
    if self._dev is None:
      self._dev = [[None if p is None else DeviceArray.from_numpy(np.asarray(p, np.float32))
                    for p in comp] for comp in self.profiles]
    return self._dev

  def host(self, v):
    out = []
    for a in range(self.grid.ndim):
      if not self.has[a]:
        out.append(np.zeros(self.grid.shape, np.float32))
        continue
      val = None
      for j, p in enumerate(self.profiles[a]):
        if p is None:
          continue
        shp = [1] * self.grid.ndim
        shp[j] = -1
        pj = np.asarray(p, np.float32).reshape(shp)
        val = pj if val is None else (val * pj).astype(np.float32)
      val = np.ones((1,) * self.grid.ndim, np.float32) if val is None else val
      val = (val * np.float32(self.scales[a])).astype(np.float32)
      out.append(np.ascontiguousarray(np.broadcast_to(val, self.grid.shape)))
    return tuple(out)


class FieldTerm(Term):
  kind = FORCE_FIELD

  def __init__(self, grid, arrays, offsets):
    self.grid, self.arrays, self.offsets = grid, arrays, offsets
    self._dev = None

  def device_fields(self):
    if self._dev is None:
      self._dev = [None if a is None else
                   (a if _lib.is_device_array(a) else DeviceArray.from_numpy(np.asarray(a, np.float32)))
                   for a in self.arrays]
    return self._dev

  def host(self, v):
    return tuple(np.zeros(self.grid.shape, np.float32) if a is None else np.asarray(a, np.float32)
                 for a in self.arrays)


class LinearTerm(Term):
  kind = FORCE_LINEAR

  def __init__(self, coef):
    self.coef = float(coef)

  def host(self, v):
    return tuple((np.float32(self.coef) * np.asarray(u.data)).astype(np.float32) for u in v)


class SmagorinskyTerm(Term):
  kind = FORCE_SMAGORINSKY

  def __init__(self, cs):
    self.cs = float(cs)

  def host(self, v):
    raise NotImplementedError('the Smagorinsky acceleration is evaluated on the device only')


class ForcingFn:
  """A sum of supported forcing terms, callable like the reference's ForcingFn
  (forcings.py:31): `forcing(v) -> tuple of GridArray` (host evaluation, for inspection)."""

  def __init__(self, terms: Sequence[Term], grid: Optional[grids.Grid] = None):
    self.terms = tuple(terms)
    self.grid = grid

  def __call__(self, v):
    grid = v[0].grid
    total = None
    for t in self.terms:
      vals = t.host(v)
      total = vals if total is None else tuple((a + b).astype(np.float32) for a, b in zip(total, vals))
    if total is None:
      total = tuple(np.zeros(grid.shape, np.float32) for _ in v)
    return tuple(grids.GridArray(a, u.offset, grid) for a, u in zip(total, v))


def as_forcing(forcing) -> Optional[ForcingFn]:
  if forcing is None:
    return None
  if isinstance(forcing, ForcingFn):
    return forcing
  raise NotImplementedError(
      'the B200 step fuses forcing into the stencil kernel and only accepts forcings built by '
      'jax_cfd_b200.forcings (kolmogorov / taylor_green / linear / constant / sum) or '
      'subgrid_models.explicit_smagorinsky_navier_stokes; arbitrary Python ForcingFn callables '
      'cannot be traced without JAX')


def make_params(grid: grids.Grid, dt: float, density: float, viscosity: Optional[float],
                forcing: Optional[ForcingFn], convect_dt: Optional[float] = None):
  """Fills struct cfd_params; returns (params, keepalive).  `convect_dt`: the dt bound into the
  convection term by the equation builder (equations.py:127-128) when it differs from the time
  stepper's `dt`."""
  p = Params()
  keep = []
  p.dt = float(dt)
  p.convect_dt = 0.0 if convect_dt is None else float(convect_dt)
  p.density = float(density)
  p.has_viscosity = 0 if viscosity is None else 1
  p.viscosity = 0.0 if viscosity is None else float(viscosity)
  terms = () if forcing is None else forcing.terms
  if len(terms) > MAX_TERMS:
    raise NotImplementedError(f'at most {MAX_TERMS} forcing terms are supported')
  seen = set()
  p.n_terms = len(terms)
  for i, t in enumerate(terms):
    if t.kind in seen:
      raise NotImplementedError('each kind of forcing term may appear once in a sum')
    seen.add(t.kind)
    p.term_kind[i] = t.kind
    if t.kind == FORCE_SEPARABLE:
      tabs = t.device_tables()
      keep.append(tabs)
      for a in range(grid.ndim):
        p.has_sep[a] = 1 if t.has[a] else 0
        p.sep_scale[a] = float(t.scales[a])
        for j in range(grid.ndim):
          p.sep_prof[a][j] = None if tabs[a][j] is None else tabs[a][j].ptr
    elif t.kind == FORCE_FIELD:
      fields = t.device_fields()
      keep.append(fields)
      for a in range(grid.ndim):
        p.field[a] = None if fields[a] is None else _lib.device_ptr(fields[a])
    elif t.kind == FORCE_LINEAR:
      p.linear_coef = t.coef
    elif t.kind == FORCE_SMAGORINSKY:
      p.smagorinsky_cs = t.cs
  return p, keep


# ----------------------------------------------------------------------------- validation
def validate_velocity(v, grid: Optional[grids.Grid] = None) -> Tuple[grids.Grid, int, Tuple[int, ...], bool]:
  """Checks the inputs the fused path accepts; returns (grid, batch, lead_shape, on_device)."""
  if not isinstance(v, (tuple, list)) or not v:
    raise TypeError('velocity must be a tuple of GridVariable')
  for u in v:
    if not isinstance(u, grids.GridVariable):
      raise TypeError(f'expected GridVariable, got {type(u)}')
  g = grids.consistent_grid(*v)
  if grid is not None and g != grid:
    raise grids.InconsistentGridError('velocity is defined on a different grid than the step function')
  if len(v) != g.ndim:
    raise ValueError(f'expected {g.ndim} velocity components, got {len(v)}')
  if not boundaries.has_all_periodic_boundary_conditions(*v):
    raise NotImplementedError('Non-periodic boundary conditions are not implemented.')
  for u, face in zip(v, g.cell_faces):
    if tuple(u.offset) != tuple(face):
      raise grids.InconsistentOffsetError(
          f'velocity components must live on grid.cell_faces; got {u.offset}, expected {face}')
  shapes = {tuple(u.data.shape) for u in v}
  if len(shapes) != 1:
    raise ValueError(f'velocity components have different shapes: {shapes}')
  shape = shapes.pop()
  if shape[len(shape) - g.ndim:] != g.shape:
    raise ValueError(f'data shape {shape} does not end with grid shape {g.shape}')
  lead = shape[:len(shape) - g.ndim]
  batch = int(np.prod(lead, dtype=np.int64)) if lead else 1
  dev = [_lib.is_device_array(u.data) for u in v]
  if any(dev) and not all(dev):
    raise TypeError('velocity components must be all host or all device arrays')
  for u in v:
    dt = np.dtype(str(u.data.dtype).replace('torch.', ''))
    if dt != np.float32:
      raise TypeError(f'the B200 path computes in float32; got {u.data.dtype}')
  if all(dev):
    # the kernels read raw base pointers: strided views (slices, transposes) would be misread
    for u in v:
      if not _lib.is_c_contiguous(u.data):
        raise ValueError('device arrays must be C-contiguous (got a strided view; make it '
                         'contiguous first)')
    if len({_lib.device_of(u.data) for u in v}) != 1:
      raise ValueError('velocity components live on different CUDA devices')
  return g, batch, lead, all(dev)


def rewrap(v, datas):
  return tuple(grids.GridVariable(grids.GridArray(d, u.offset, u.grid), u.bc) for d, u in zip(datas, v))


# ----------------------------------------------------------------------------- native callables
class NativeStep:
  """step_fn of semi_implicit_navier_stokes with forward Euler (equations.py:120-151)."""

  def __init__(self, grid, dt, density, viscosity, forcing: Optional[ForcingFn],
               convect_dt: Optional[float] = None, implementation=None):
    self.grid, self.dt, self.density, self.viscosity, self.forcing = grid, dt, density, viscosity, forcing
    self.convect_dt = convect_dt
    self.implementation = implementation
    self.impl_code = check_implementation(grid, implementation)
    self._params = None
    self._keep = None
    self.last_q = None

  def with_time_step(self, time_step: float) -> 'NativeStep':
    """The same equation advanced with another stepper dt; the convection term keeps the dt it was
    built with (time_stepping.py:59-106 vs equations.py:127-128)."""
    if time_step == self.dt and self.convect_dt is None:
      return self
    base = self.dt if self.convect_dt is None else self.convect_dt
    return NativeStep(self.grid, time_step, self.density, self.viscosity, self.forcing,
                      convect_dt=None if base == time_step else base, implementation=self.implementation)

  def params(self):
    if self._params is None:
      self._params, self._keep = make_params(self.grid, self.dt, self.density, self.viscosity,
                                             self.forcing, self.convect_dt)
    return self._params

  def __call__(self, v):
    return self.advance(v, 1)

  def advance(self, v, nsteps: int, return_q: bool = False):
    """`nsteps` steps (funcutils.repeated semantics)."""
    grid, batch, lead, on_dev = validate_velocity(v, self.grid)
    if nsteps == 0:
      return tuple(v)
    stream = _lib.stream_of(v[0].data) if on_dev else None
    plan = plan_for(grid, batch, [u.data for u in v], stream, self.impl_code)
    params = self.params()
    if not on_dev:
      ins = [np.ascontiguousarray(u.data, dtype=np.float32) for u in v]
      outs = [np.empty_like(a) for a in ins]
      q = np.empty_like(ins[0]) if return_q else None
      check(lib().cfd_step_host(plan.handle, _lib.ptr_array([a.ctypes.data for a in ins]),
                                _lib.ptr_array([a.ctypes.data for a in outs]),
                                None if q is None else q.ctypes.data, nsteps,
                                ctypes.byref(params)))
      res = rewrap(v, outs)
      return (res, q) if return_q else res
    a = [u.data for u in v]
    b = [_lib.empty_like(u.data) for u in v]
    q = _lib.empty_like(v[0].data) if return_q else None
    if nsteps == 1:
      check(lib().cfd_step(plan.handle, stream, _lib.ptr_array(a), _lib.ptr_array(b),
                           None if q is None else _lib.device_ptr(q), ctypes.byref(params)))
      res = rewrap(v, b)
      return (res, q) if return_q else res
    # ping-pong without touching the caller's input: first step a -> b, then b <-> c
    c = [_lib.empty_like(u.data) for u in v]
    check(lib().cfd_step(plan.handle, stream, _lib.ptr_array(a), _lib.ptr_array(b), None,
                         ctypes.byref(params)))
    in_c = ctypes.c_int(0)
    if return_q and nsteps > 1:
      check(lib().cfd_repeated(plan.handle, stream, _lib.ptr_array(b), _lib.ptr_array(c), nsteps - 2,
                               ctypes.byref(params), ctypes.byref(in_c)))
      src, dst = (c, b) if in_c.value else (b, c)
      check(lib().cfd_step(plan.handle, stream, _lib.ptr_array(src), _lib.ptr_array(dst),
                           _lib.device_ptr(q), ctypes.byref(params)))
      return rewrap(v, dst), q
    check(lib().cfd_repeated(plan.handle, stream, _lib.ptr_array(b), _lib.ptr_array(c), nsteps - 1,
                             ctypes.byref(params), ctypes.byref(in_c)))
    return rewrap(v, c if in_c.value else b)


def _to_device_tuple(v):
  return [u.data if _lib.is_device_array(u.data) else DeviceArray.from_numpy(
      np.ascontiguousarray(u.data, np.float32)) for u in v]


def _from_device(datas, like_host: bool):
  return [d.numpy() if like_host and isinstance(d, DeviceArray) else d for d in datas]


class NativeExplicitTerms:
  """explicit_terms of navier_stokes_explicit_terms (equations.py:77-116) on the device."""

  def __init__(self, grid, dt, density, viscosity, forcing, implementation=None):
    self.grid = grid
    self._step = NativeStep(grid, dt, density, viscosity, forcing, implementation=implementation)

  def __call__(self, v):
    grid, batch, lead, on_dev = validate_velocity(v, self.grid)
    ins = _to_device_tuple(v)
    outs = [_lib.empty_like(x) for x in ins]
    stream = _lib.stream_of(ins[0])
    plan = plan_for(grid, batch, ins, stream, self._step.impl_code)
    check(lib().cfd_explicit_terms(plan.handle, stream, _lib.ptr_array(ins), _lib.ptr_array(outs),
                                   ctypes.byref(self._step.params())))
    if not on_dev:
      check(lib().cfd_stream_sync(stream))
    return rewrap(v, _from_device(outs, not on_dev))


class NativeProjection:
  """pressure.projection with solve_fast_diag (pressure.py:181-198)."""

  def __init__(self, grid=None, implementation=None):
    self.grid = grid
    self.implementation = implementation
    self.impl_code = (_lib.implementation_code(implementation) if grid is None
                      else check_implementation(grid, implementation))

  def __call__(self, v, return_q=False):
    grid, batch, lead, on_dev = validate_velocity(v, self.grid)
    if self.grid is None:
      check_implementation(grid, self.implementation)
    ins = _to_device_tuple(v)
    outs = [_lib.empty_like(x) for x in ins]
    q = _lib.empty_like(ins[0]) if return_q else None
    stream = _lib.stream_of(ins[0])
    plan = plan_for(grid, batch, ins, stream, self.impl_code)
    check(lib().cfd_project(plan.handle, stream, _lib.ptr_array(ins), _lib.ptr_array(outs),
                            None if q is None else _lib.device_ptr(q)))
    if not on_dev:
      check(lib().cfd_stream_sync(stream))
    res = rewrap(v, _from_device(outs, not on_dev))
    if return_q:
      qd = _from_device([q], not on_dev)[0]
      return res, grids.GridArray(qd, grid.cell_center, grid)
    return res


def axpy(v, ks, coefs):
  """u0 + sum_j coef_j k_j on the device (stage combination of navier_stokes_rk)."""
  grid, batch, lead, on_dev = validate_velocity(v)
  x = _to_device_tuple(v)
  ys = [_to_device_tuple(k) for k in ks]
  outs = [_lib.empty_like(a) for a in x]
  stream = _lib.stream_of(x[0])
  plan = plan_for(grid, batch, x, stream)
  n = len(ys)
  ypp = (ctypes.POINTER(ctypes.c_void_p) * max(n, 1))()
  keep = []
  for i, y in enumerate(ys):
    arr = _lib.ptr_array(y)
    keep.append(arr)
    ypp[i] = ctypes.cast(arr, ctypes.POINTER(ctypes.c_void_p))
  cf = (ctypes.c_double * max(n, 1))(*[float(c) for c in coefs])
  check(lib().cfd_axpy(plan.handle, stream, _lib.ptr_array(x), n, ypp, cf, _lib.ptr_array(outs)))
  if not on_dev:
    check(lib().cfd_stream_sync(stream))
  return rewrap(v, _from_device(outs, not on_dev))


def diagnostics(v):
  grid, batch, lead, on_dev = validate_velocity(v)
  x = _to_device_tuple(v)
  d = _lib.Diag()
  stream = _lib.stream_of(x[0])
  plan = plan_for(grid, batch, x, stream)
  check(lib().cfd_diagnostics(plan.handle, stream, _lib.ptr_array(x), ctypes.byref(d)))
  return dict(kinetic_energy=d.kinetic_energy, enstrophy=d.enstrophy, max_abs_div=d.max_abs_div,
              max_speed_sq=d.max_speed_sq)


def scale(v, numer: float, denom: float):
  """numer * u / denom per component on the device, float32, in that order
  (initial_conditions.py:118-121)."""
  grid, batch, lead, on_dev = validate_velocity(v)
  x = _to_device_tuple(v)
  outs = [_lib.empty_like(a) for a in x]
  stream = _lib.stream_of(x[0])
  plan = plan_for(grid, batch, x, stream)
  check(lib().cfd_scale(plan.handle, stream, _lib.ptr_array(x), float(numer), float(denom), _lib.ptr_array(outs)))
  if not on_dev:
    check(lib().cfd_stream_sync(stream))
  return rewrap(v, _from_device(outs, not on_dev))


class NativeTransform:
  """`fast_diagonalization.transform` on one real float32 field of a periodic grid: out = X diag X^-1 in.

  rfft: `diag` = func(eigenvalue sums) in rfftn layout (N0, [N1,] N_last/2+1), real.
  matmul: `eigvecs[j]` (N_j, N_j) float64 with eigenvectors in columns, `diag` of the grid shape."""

  def __init__(self, grid: grids.Grid, diag: np.ndarray, eigvecs=None):
    self.grid = grid
    self.eigvecs = None if eigvecs is None else [np.ascontiguousarray(v, np.float64) for v in eigvecs]
    if eigvecs is None:
      nlast = grid.shape[-1] // 2 + 1
      if tuple(diag.shape) != grid.shape[:-1] + (nlast,):
        raise ValueError(f'diagonal shape {diag.shape} does not match the rfftn layout of {grid.shape}')
      # line layout: (N_last/2+1, [N1,] N0) -- the layout the x-line kernels read
      self.diag = np.ascontiguousarray(np.transpose(np.asarray(diag, np.float32)))
      self.impl = _lib.IMPL_RFFT
    else:
      if tuple(diag.shape) != grid.shape:
        raise ValueError(f'diagonal shape {diag.shape} does not match the grid shape {grid.shape}')
      # the reference narrows `diagonals` to the data dtype (fast_diagonalization.py:143)
      self.diag = np.ascontiguousarray(np.asarray(diag, np.float32).astype(np.float64))
      self.impl = _lib.IMPL_MATMUL
    self._dev = {}

  def _tables(self, device):
    t = self._dev.get(device)
    if t is None:
      prev = _lib.current_device()
      check(lib().cfd_set_device(device))
      try:
        d = DeviceArray.from_numpy(self.diag)
        vs = vts = None
        if self.eigvecs is not None:
          vs = [DeviceArray.from_numpy(v) for v in self.eigvecs]
          vts = [DeviceArray.from_numpy(np.ascontiguousarray(v.T)) for v in self.eigvecs]
      finally:
        check(lib().cfd_set_device(prev))
      t = self._dev[device] = (d, vs, vts)
    return t

  def __call__(self, rhs):
    """rhs: array of shape (..., *grid.shape), float32, numpy or device.  Returns the same kind."""
    on_dev = _lib.is_device_array(rhs)
    shape = tuple(rhs.shape)
    nd = self.grid.ndim
    if shape[len(shape) - nd:] != self.grid.shape:
      raise ValueError(f'rhs.shape={shape} does not match shape={self.grid.shape}')
    if np.dtype(str(rhs.dtype).replace('torch.', '')) != np.float32:
      raise ValueError(f'rhs.dtype={rhs.dtype} does not match dtype=float32')
    if on_dev and not _lib.is_c_contiguous(rhs):
      raise ValueError('device arrays must be C-contiguous')
    batch = int(np.prod(shape[:len(shape) - nd], dtype=np.int64)) if len(shape) > nd else 1
    x = rhs if on_dev else DeviceArray.from_numpy(np.ascontiguousarray(rhs, np.float32))
    out = _lib.empty_like(x)
    stream = _lib.stream_of(x)
    plan = plan_for(self.grid, batch, [x], stream, self.impl)
    d, vs, vts = self._tables(plan.device)
    if self.impl == _lib.IMPL_RFFT:
      check(lib().cfd_transform_rfft(plan.handle, stream, _lib.device_ptr(x), _lib.device_ptr(out), d.ptr))
    else:
      check(lib().cfd_transform_matmul(plan.handle, stream, _lib.device_ptr(x), _lib.device_ptr(out),
                                       _lib.ptr_array(vs), _lib.ptr_array(vts), d.ptr))
    if not on_dev:
      check(lib().cfd_stream_sync(stream))
      return out.numpy()
    return out
