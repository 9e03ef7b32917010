"""CPU oracle for the jax-cfd staggered-grid FVM time step.  TEST INFRASTRUCTURE ONLY.

This file is a NumPy restatement of the algorithm behind the `step_fn` returned by
`jax_cfd.base.equations.semi_implicit_navier_stokes` (periodic, float32/float64).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu-baseline / reference arm
may import it -- as the checker, never as the thing measured or shipped.  The product
(`jax_cfd_b200`) never imports anything under `oracle/`.

Parity status: PINNED.  (1) The reference's own test files run unmodified under
`oracle/jax_shim` (oracle/run_reference_tests.sh: boundaries 125/125, grids 47/47,
finite_differences 19/19, fast_diagonalization 20/20, forcings 8/8, interpolation 18/18 on
path, pressure fast-diag 11/11, advection using_limiters, subgrid_models, equations fast_diag).
(2) `oracle/gen_golden.py` runs the reference's own `semi_implicit_navier_stokes` from
/root/reference under that stand-in and commits input/output vectors to `tests/golden/`;
`tests/test_oracle_golden.py` checks this restatement against them.

Every function cites the reference file:line it restates (paths relative to
/root/reference/jax_cfd/base/).  Conventions: periodic arrays, axis 0 slowest;
`shift(a, k, ax)[i] = a[(i + k) mod N]`.
"""
from __future__ import annotations

import dataclasses
from typing import Optional, Sequence, Tuple

import numpy as np


# --------------------------------------------------------------------------- indexing
def shift(a: np.ndarray, k: int, axis: int) -> np.ndarray:
  """`GridVariable.shift` for periodic bc: grids.py:331-349 -> boundaries.py:81-104,193-199.

  out[i] = in[(i + k) mod N] (KAT boundaries_test.py:170-209: [11,12,13,14] shift +1 ->
  [12,13,14,11]).
  """
  return np.roll(a, -k, axis=axis)


def cell_faces(ndim: int) -> Tuple[Tuple[float, ...], ...]:
  """`Grid.cell_faces`: grids.py:567-572."""
  offsets = (np.eye(ndim) + np.ones([ndim, ndim])) / 2.
  return tuple(tuple(float(o) for o in offset) for offset in offsets)


def grid_step(shape: Sequence[int], domain: Sequence[Tuple[float, float]]) -> Tuple[float, ...]:
  """`Grid.step`: grids.py:553-555 (python floats)."""
  return tuple((float(up) - float(lo)) / n for (lo, up), n in zip(domain, shape))


def grid_axes(shape, domain, offset, dtype=np.float32):
  """`Grid.axes`: grids.py:600-602, evaluated as x64-disabled jax does: int32 arange plus
  weak python floats -> float32 at every step."""
  step = grid_step(shape, domain)
  out = []
  for (lo, _), o, n, h in zip(domain, offset, shape, step):
    a = np.arange(n, dtype=np.int32).astype(dtype)
    a = (a + dtype(o)).astype(dtype)
    a = (a * dtype(h)).astype(dtype)
    a = (dtype(lo) + a).astype(dtype)
    out.append(a)
  return tuple(out)


# --------------------------------------------------------------------------- forcing
@dataclasses.dataclass
class Forcing:
  """Descriptor of the forcing terms the fused path supports, summed left to right.

  terms: sequence of ('const', tuple_of_arrays) | ('linear', coef) | ('smagorinsky', cs)
  (forcings.py:63-104 / 35-60 const fields, 107-113 linear, 125-129 sum;
  subgrid_models.py:188-213 appends the Smagorinsky acceleration last).
  """
  terms: tuple = ()


def kolmogorov_field(shape, domain, scale=1.0, k=2, swap_xy=False, dtype=np.float32):
  """Constant field of `kolmogorov_forcing`: forcings.py:63-104.

  f_0 = scale * sin(k * y) on the u-face mesh (offset cell_faces[0]), others zero; with
  swap_xy f_1 = scale * sin(k * x) on the v-face mesh.
  """
  ndim = len(shape)
  faces = cell_faces(ndim)
  out = [np.zeros(shape, dtype) for _ in range(ndim)]
  if swap_xy:
    x = grid_axes(shape, domain, faces[1], dtype)[0]
    prof = (dtype(scale) * np.sin((dtype(k) * x).astype(dtype)).astype(dtype)).astype(dtype)
    out[1] = np.ascontiguousarray(np.broadcast_to(prof.reshape((-1,) + (1,) * (ndim - 1)), shape), dtype)
  else:
    y = grid_axes(shape, domain, faces[0], dtype)[1]
    prof = (dtype(scale) * np.sin((dtype(k) * y).astype(dtype)).astype(dtype)).astype(dtype)
    out[0] = np.ascontiguousarray(np.broadcast_to(prof.reshape((1, -1) + (1,) * (ndim - 2)), shape), dtype)
  return tuple(out)


def taylor_green_field(shape, scale=1.0, k=2, dtype=np.float32):
  """Constant field of `taylor_green_forcing`: forcings.py:35-60, evaluating
  validation_problems.py:72-74,90-104 at t=0 on its own (0, 2pi)^2 grid of shape[:2]:
  u = cos(kx x) sin(ky y) at offset (1, .5); v = -sin(kx x) cos(ky y) at offset (.5, 1)."""
  n2 = tuple(shape[:2])
  dom = ((0.0, 2 * np.pi),) * 2
  xu, yu = grid_axes(n2, dom, (1.0, 0.5), dtype)
  xv, yv = grid_axes(n2, dom, (0.5, 1.0), dtype)
  kk = dtype(k)
  u = np.cos(kk * xu)[:, None] * np.sin(kk * yu)[None, :]
  v = -np.sin(kk * xv)[:, None] * np.cos(kk * yv)[None, :]
  u = (u.astype(dtype) * dtype(scale)).astype(dtype)
  v = (v.astype(dtype) * dtype(scale)).astype(dtype)
  if len(shape) == 2:
    return (u, v)
  u3 = np.ascontiguousarray(np.broadcast_to(u[..., None], shape), dtype)
  v3 = np.ascontiguousarray(np.broadcast_to(v[..., None], shape), dtype)
  return (u3, v3, np.zeros(shape, dtype))


# --------------------------------------------------------------------------- advection
def _safe_div(x, y):
  """interpolation.py:224-226."""
  return x / np.where(y != 0, y, 1)


def _van_leer(r):
  """interpolation.py:229-231."""
  return np.where(r > 0, _safe_div(2 * r, 1 + r), 0.0).astype(r.dtype)


def advect_van_leer(c, v, a, dt, h):
  """-(div of TVD-limited flux) for component `c = v[a]`.

  advection.py:387-395 -> 81-116 -> 34-78 with interpolation.py:67-93 (linear),
  96-155 (upwind), 158-221 (lax_wendroff), 234-303 (apply_tvd_limiter) and
  finite_differences.py:136-143,95-102 (divergence by backward differences).
  """
  ndim = len(v)
  total = None
  for j in range(ndim):
    # interpolation.py:57-62 with offset delta 1/2 along axis `a`: weights (.5, .5)
    U = 0.5 * v[j] + 0.5 * shift(v[j], +1, a)
    cL, cR, cRR = shift(c, -1, j), shift(c, +1, j), shift(c, +2, j)
    pos = U > 0
    low = np.where(pos, c, cR)                                   # interpolation.py:147-151
    C = (dt / h[j]) * U                                          # interpolation.py:210
    high = np.where(pos, c + 0.5 * (1 - C) * (cR - c),
                    cR - 0.5 * (1 + C) * (cR - c))               # interpolation.py:211-217
    r_pos = _safe_div(c - cL, cR - c)                            # interpolation.py:287-289
    r_neg = _safe_div(cRR - cR, cR - c)
    phi = np.where(pos, _van_leer(r_pos), _van_leer(r_neg))      # interpolation.py:290-296
    face = low - (low - high) * phi                              # interpolation.py:297
    F = face * U                                                 # advection.py:73
    d = (F - shift(F, -1, j)) / h[j]                             # finite_differences.py:95-102
    total = d if total is None else total + d                    # finite_differences.py:143 (sum)
  return -total                                                  # advection.py:78


def laplacian(c, h):
  """finite_differences.py:127-133: scales formed in the field dtype."""
  scales = np.square(1 / np.array(h, dtype=c.dtype))
  result = -2 * c * np.sum(scales)
  for ax in range(c.ndim):
    result = result + (shift(c, -1, ax) + shift(c, +1, ax)) * scales[ax]
  return result


def smagorinsky_acceleration(v, h, cs):
  """`evm_model` with `smagorinsky_viscosity`: subgrid_models.py:101-134, 40-98."""
  d = len(v)
  dt_ = v[0].dtype.type
  fwd = lambda u, ax: (shift(u, +1, ax) - u) / h[ax]             # finite_differences.py:117-124
  s = [[0.5 * (fwd(v[i], j) + fwd(v[j], i)) for j in range(d)] for i in range(d)]

  def to_center(x, i, j):                                        # subgrid_models.py:88-89 via linear
    if i == j:
      return shift(x, -1, i)                                     # integer offset delta: pure shift
    for ax in sorted((i, j)):
      x = 0.5 * shift(x, -1, ax) + 0.5 * x
    return x

  def from_center(x, i, j):                                      # subgrid_models.py:94-97
    if i == j:
      return shift(x, +1, i)
    for ax in sorted((i, j)):
      x = 0.5 * x + 0.5 * shift(x, +1, ax)
    return x

  sc = [[to_center(s[i][j], i, j) for j in range(d)] for i in range(d)]
  tr = None
  for i in range(d):                                             # np.trace(S.dot(S))
    acc = None
    for k in range(d):
      t = sc[i][k] * sc[k][i]
      acc = t if acc is None else acc + t
    tr = acc if tr is None else tr + acc
  cutoff = np.prod(np.array(h)) ** (1 / d)                       # subgrid_models.py:91
  nu_t = (dt_((cs * cutoff) ** 2) * np.sqrt(2 * tr)).astype(v[0].dtype)
  out = []
  for i in range(d):
    div = None
    for j in range(d):
      tau = -2. * from_center(nu_t, i, j) * s[i][j]              # subgrid_models.py:130
      t = (tau - shift(tau, -1, j)) / h[j]
      div = t if div is None else div + t
    out.append(-div)                                             # subgrid_models.py:131-134
  return tuple(out)


def explicit_terms(v, dt, h, nu_over_rho, inv_rho_forcing: Optional[Forcing], density=1.0):
  """`navier_stokes_explicit_terms`: equations.py:77-116 (order: conv + diffusion + forcing/rho)."""
  d = len(v)
  out = []
  conv = [advect_van_leer(v[a], v, a, dt, h) for a in range(d)]
  diff = [nu_over_rho * laplacian(v[a], h) for a in range(d)] if nu_over_rho is not None else None
  force = None
  if inv_rho_forcing is not None and inv_rho_forcing.terms:
    force = [0] * d                                              # equations.py:41-42: python sum()
    for kind, arg in inv_rho_forcing.terms:
      if kind == 'const':
        term = arg
      elif kind == 'linear':
        term = tuple(arg * u for u in v)                         # forcings.py:111-113
      elif kind == 'smagorinsky':
        term = smagorinsky_acceleration(v, h, arg)
      else:
        raise ValueError(kind)
      force = [f + t for f, t in zip(force, term)]
  for a in range(d):
    t = conv[a]
    if diff is not None:
      t = t + diff[a]
    if force is not None:
      t = t + force[a] / density
    out.append(t)
  return tuple(out)


# --------------------------------------------------------------------------- projection
def pinv_diagonals(shape, h, dtype=np.float32):
  """Diagonal of the pseudo-inverse in rfftn layout.

  fast_diagonalization.py:199-225 (`_circulant_rfft_transform`) with the operators of
  array_utils.py:168-173 (circulant column [-2, 1, 0, ..., 0, 1] / h^2, f64) and the cutoff
  rule of fast_diagonalization.py:257-262: |L| > 10 * eps(dtype) ? 1/L : 0, evaluated in f64,
  then narrowed to the complex type matching `dtype` as x64-disabled jax does (line 214).
  """
  eig = []
  for ax, (n, step) in enumerate(zip(shape, h)):
    col = np.zeros(n)
    col[0] = -2 / step ** 2
    col[1] = col[-1] = 1 / step ** 2
    eig.append(np.fft.rfft(col) if ax == len(shape) - 1 else np.fft.fft(col))
  import functools
  summed = functools.reduce(np.add.outer, eig)
  cutoff = 10 * np.finfo(dtype).eps
  with np.errstate(divide='ignore', invalid='ignore'):
    diag = np.where(abs(summed) > cutoff, 1 / summed, 0)
  return diag.astype(np.complex64 if np.dtype(dtype) == np.float32 else np.complex128)


def divergence(v, h):
  """finite_differences.py:136-143."""
  tot = None
  for j, u in enumerate(v):
    t = (u - shift(u, -1, j)) / h[j]
    tot = t if tot is None else tot + t
  return tot


def solve_pressure(v, h, diag=None):
  """`solve_fast_diag` periodic branch: pressure.py:115-157; fast_diagonalization.py:223."""
  rhs = divergence(v, h)
  if diag is None:
    diag = pinv_diagonals(rhs.shape, h, np.float32)
  q = np.fft.irfftn(diag * np.fft.rfftn(rhs), s=rhs.shape, axes=tuple(range(rhs.ndim)))
  return q.astype(rhs.dtype)


def projection(v, h, diag=None):
  """`pressure.projection`: pressure.py:181-198 -> (v - forward_difference(q), q)."""
  q = solve_pressure(v, h, diag)
  out = tuple(u - (shift(q, +1, j) - q) / h[j] for j, u in enumerate(v))
  return out, q


# --------------------------------------------------------------------------- the step
def step(v, dt, h, density=1.0, viscosity: Optional[float] = None,
         forcing: Optional[Forcing] = None, diag=None, return_q=False, return_ustar=False):
  """One forward-Euler projection step: time_stepping.py:88-104,109-118 with
  equations.py:120-151:  v' = P(v + dt * F(v))."""
  nu = None if viscosity is None else viscosity / density        # equations.py:107
  k0 = explicit_terms(v, dt, h, nu, forcing, density)
  ustar = tuple(u + dt * (1 * k) for u, k in zip(v, k0))         # time_stepping.py:101
  vnew, q = projection(ustar, h, diag)
  res = [vnew]
  if return_q:
    res.append(q)
  if return_ustar:
    res.append(ustar)
  return res[0] if len(res) == 1 else tuple(res)


def diffusion_diagonals(shape, h, nu, dt, dtype=np.float32):
  """Diagonal of diffusion.solve_fast_diag's transform in rfftn layout: diffusion.py:175-177
  func(x) = dt nu x / (1 - dt nu x) on the summed circulant eigenvalues (array_utils.py:168-173,
  fast_diagonalization.py:199-214), narrowed to complex64 for float32 data like x64-disabled jax."""
  eig = []
  for ax, (n, step) in enumerate(zip(shape, h)):
    col = np.zeros(n)
    col[0] = -2 / step ** 2
    col[1] = col[-1] = 1 / step ** 2
    eig.append(np.fft.rfft(col) if ax == len(shape) - 1 else np.fft.fft(col))
  import functools
  summed = functools.reduce(np.add.outer, eig)
  dt_nu_x = (dt * nu) * summed
  diag = dt_nu_x / (1 - dt_nu_x)
  return diag.astype(np.complex64 if np.dtype(dtype) == np.float32 else np.complex128)


def implicit_diffusion_step(v, dt, h, density=1.0, viscosity=0.0, forcing: Optional[Forcing] = None,
                            diag=None, ddiag=None):
  """`implicit_diffusion_navier_stokes`: equations.py:154-195 with diffusion.solve_fast_diag
  (diffusion.py:166-212; periodic; even last axis -> the rfft implementation):
  v* = v + (conv + forcing / rho) dt;  v = P(v*);  v = v + irfftn(D rfftn(v))."""
  k0 = explicit_terms(v, dt, h, None, forcing, density)          # no explicit diffusion
  ustar = tuple(u + k * dt for u, k in zip(v, k0))               # equations.py:186-188
  vp, _ = projection(ustar, h, diag)
  if ddiag is None:
    ddiag = diffusion_diagonals(vp[0].shape, h, viscosity, dt, vp[0].dtype)  # nu = viscosity (equations.py:192)
  axes = tuple(range(vp[0].ndim))
  return tuple(u + np.fft.irfftn(ddiag * np.fft.rfftn(u), s=u.shape, axes=axes).astype(u.dtype) for u in vp)


def rk_step(v, dt, h, tableau_a, tableau_b, density=1.0, viscosity=None, forcing=None, diag=None):
  """`navier_stokes_rk`: time_stepping.py:59-106 (projection after every stage)."""
  nu = None if viscosity is None else viscosity / density
  F = lambda u: explicit_terms(u, dt, h, nu, forcing, density)
  P = lambda u: projection(u, h, diag)[0]
  n = len(tableau_b)
  u = [None] * n
  k = [None] * n
  u[0] = v
  k[0] = F(v)
  for i in range(1, n):
    us = tuple(v[c] + dt * sum(tableau_a[i - 1][j] * k[j][c] for j in range(i) if tableau_a[i - 1][j])
               for c in range(len(v)))
    u[i] = P(us)
    k[i] = F(u[i])
  us = tuple(v[c] + dt * sum(tableau_b[j] * k[j][c] for j in range(n) if tableau_b[j])
             for c in range(len(v)))
  return P(us)


# --------------------------------------------------------------------------- diagnostics
def downsample_staggered_velocity_component(u, direction, factor):
  """resize.py:72-74: slice_along_axis(u, direction, slice(factor - 1, None, factor)) then
  block_reduce(w, block_size, mean) with block 1 along `direction` and `factor` elsewhere
  (array_utils.py:136-166)."""
  u = np.asarray(u)
  sl = [slice(None)] * u.ndim
  sl[direction] = slice(factor - 1, None, factor)
  w = u[tuple(sl)]
  new_shape, axes = [], []
  for j, s_ in enumerate(w.shape):
    b = 1 if j == direction else factor
    if s_ % b:
      raise ValueError('`block_size` must divide `array.shape`')
    new_shape += [s_ // b, b]
    axes.append(2 * j + 1)
  return w.reshape(new_shape).mean(axis=tuple(axes), dtype=w.dtype if w.dtype.kind == 'f' else None)


def vorticity_2d(u, v, dx, dy):
  """data/xarray_utils.py:155-163: (roll(v, -1, x) - v) / dx - (roll(u, -1, y) - u) / dy."""
  dv_dx = (np.roll(v, -1, axis=-2) - v) / v.dtype.type(dx)
  du_dy = (np.roll(u, -1, axis=-1) - u) / u.dtype.type(dy)
  return dv_dx - du_dy


def diagnostics(v, h):
  """Mean kinetic energy, mean enstrophy (2-D), max |div|, max speed^2.

  data/xarray_utils.py:155-188 (KE = |v|^2/2 on raw staggered samples, enstrophy = w^2/2 with
  w = D+_x v - D+_y u), finite_differences.py:136-143, equations.py:68 (max of sum u_a^2)."""
  v64 = [np.asarray(u, np.float64) for u in v]
  ke = 0.5 * sum(np.mean(u * u) for u in v64)
  div = divergence(v64, h)
  out = {'kinetic_energy': float(ke), 'max_div': float(np.abs(div).max()),
         'max_speed_sq': float(sum(u * u for u in v64).max())}
  if len(v) == 2:
    w = (shift(v64[1], +1, 0) - v64[1]) / h[0] - (shift(v64[0], +1, 1) - v64[0]) / h[1]
    out['enstrophy'] = float(0.5 * np.mean(w * w))
  return out


# --------------------------------------------------------------------------- initial conditions
def filtered_velocity_field(seed, shape, domain, maximum_velocity=1.0, peak_wavenumber=3.0,
                            iterations=3, dtype=np.float32):
  """Restatement of initial_conditions.py:71-121 / filter_utils.py:32-42 with
  numpy.random.RandomState instead of jax.random (threefry is unavailable: the distribution
  matches, the bits do not)."""
  ndim = len(shape)
  h = grid_step(shape, domain)
  rs = np.random.RandomState(seed)

  def spectral_density(k):                                       # initial_conditions.py:60-64,94-95
    variance = .25
    mean = np.log(peak_wavenumber) + variance
    with np.errstate(divide='ignore', invalid='ignore'):
      logk = np.log(k)
      return np.exp(-(mean - logk) ** 2 / 2 / variance - logk) / k ** (ndim - 1)

  freqs = np.meshgrid(*[2 * np.pi * np.fft.fftfreq(n, s) for n, s in zip(shape, h)],
                      indexing='ij')                             # filter_utils.py:25-29
  kmag = np.sqrt(sum(f ** 2 for f in freqs))
  with np.errstate(divide='ignore', invalid='ignore'):
    filt = np.where(kmag > 0, spectral_density(kmag), 0.0)       # filter_utils.py:38
  v = []
  for _ in range(ndim):
    noise = rs.standard_normal(shape)
    v.append(np.fft.ifftn(np.fft.fftn(noise) * filt).real)
  diag = pinv_diagonals(shape, h, np.float64)
  for _ in range(iterations):                                    # initial_conditions.py:112-121
    v, _ = projection(tuple(v), h, diag)
    vmax = np.sqrt(max(1e-300, float(sum(u * u for u in v).max())))
    v = [maximum_velocity * u / vmax for u in v]
  return tuple(np.ascontiguousarray(u, dtype=dtype) for u in v)
