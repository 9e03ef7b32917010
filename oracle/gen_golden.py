#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own code.  TEST INFRASTRUCTURE.

Runs only in the build container: it imports `jax_cfd.base` unmodified from /root/reference
with the NumPy stand-in for jax/tree_math (oracle/jax_shim) first on sys.path.  The outputs
(small input/output vectors) are committed so that the oracle restatement and the CUDA path
can be checked against the reference on the GPU box, where /root/reference does not exist.

  python oracle/gen_golden.py            # rewrites tests/golden/*.npz

Each fixture stores: the inputs (v0_*), parameters, and the reference's outputs after
`n` steps (v{n}_*), the reference's u* and q for the first step, in float32 ("f32": what
x64-disabled jax computes) and, for the same inputs, float64 ("f64": the gold the f32
noise floor is measured against).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'jax_shim'))
sys.path.insert(1, '/root/reference')
sys.path.insert(2, HERE)

import numpy as np  # noqa: E402
import jax  # noqa: E402  (the stand-in)
import jax_cfd.base as cfd  # noqa: E402  (the reference, unmodified)
import cfd_oracle  # noqa: E402  (only for the seeded initial condition)

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


def wrap(grid, arrays):
  bc = cfd.boundaries.periodic_boundary_conditions(grid.ndim)
  return tuple(cfd.grids.GridVariable(cfd.grids.GridArray(np.asarray(a), o, grid), bc)
               for a, o in zip(arrays, grid.cell_faces))


def build_forcing(grid, spec):
  if spec is None:
    return None
  fs = []
  for kind, arg in spec:
    if kind == 'kolmogorov':
      fs.append(cfd.forcings.kolmogorov_forcing(grid, **arg))
    elif kind == 'taylor_green':
      fs.append(cfd.forcings.taylor_green_forcing(grid, **arg))
    elif kind == 'linear':
      fs.append(cfd.forcings.linear_forcing(grid, arg))
    else:
      raise ValueError(kind)
  return cfd.forcings.sum_forcings(*fs) if len(fs) > 1 else fs[0]


def run_case(name, shape, domain, seed, vmax, peak_k, density, viscosity, dt, forcing_spec,
             smag_cs, nsteps, stepper='forward_euler'):
  grid = cfd.grids.Grid(shape, domain=domain)
  v0 = cfd_oracle.filtered_velocity_field(seed, shape, domain, vmax, peak_k, dtype=np.float64)
  rec = dict(shape=np.array(shape), domain=np.array(domain, dtype=np.float64), seed=seed,
             density=density, viscosity=viscosity, dt=dt, smag_cs=-1.0 if smag_cs is None else smag_cs,
             forcing_spec=repr(forcing_spec), nsteps=np.array(nsteps), stepper=stepper)
  for i, a in enumerate(v0):
    rec[f'v0_{i}'] = a.astype(np.float32)
  for prec in ('f32', 'f64'):
    jax.config.update('jax_enable_x64', prec == 'f64')
    dtype = np.float32 if prec == 'f32' else np.float64
    v = wrap(grid, [a.astype(np.float32).astype(dtype) for a in v0])
    forcing = build_forcing(grid, forcing_spec)
    ts = getattr(cfd.time_stepping, stepper)
    kw = dict(density=density, viscosity=viscosity, dt=dt, grid=grid, forcing=forcing,
              time_stepper=ts)
    if smag_cs is not None:
      kw.pop('dt')
      step = cfd.subgrid_models.explicit_smagorinsky_navier_stokes(dt=dt, cs=smag_cs, **kw)
      viscosity_fn = __import__('functools').partial(
          cfd.subgrid_models.smagorinsky_viscosity, dt=dt, cs=smag_cs)
      smag = __import__('functools').partial(cfd.subgrid_models.evm_model,
                                             viscosity_fn=viscosity_fn)
      full_forcing = smag if forcing is None else cfd.forcings.sum_forcings(forcing, smag)
    else:
      step = cfd.equations.semi_implicit_navier_stokes(**kw)
      full_forcing = forcing
    # first-step internals: u* and q straight from the reference's building blocks
    explicit = cfd.equations.navier_stokes_explicit_terms(
        density=density, viscosity=viscosity, dt=dt, grid=grid, forcing=full_forcing)
    k0 = explicit(v)
    ustar = tuple(cfd.grids.GridVariable(
        cfd.grids.GridArray(u.data + dt * k.data, u.offset, grid), u.bc) for u, k in zip(v, k0))
    q = cfd.pressure.solve_fast_diag(ustar)
    if stepper == 'forward_euler':
      for i, a in enumerate(ustar):
        rec[f'{prec}_ustar_{i}'] = np.asarray(a.data)
      rec[f'{prec}_q'] = np.asarray(q.data)
    cur = v
    done = 0
    for n in nsteps:
      for _ in range(n - done):
        cur = step(cur)
      done = n
      for i, a in enumerate(cur):
        assert a.data.dtype == dtype, (a.data.dtype, dtype)
        assert a.offset == grid.cell_faces[i]
        rec[f'{prec}_v{n}_{i}'] = np.asarray(a.data)
    if forcing_spec is not None and prec == 'f32':
      f = build_forcing(grid, [s for s in forcing_spec if s[0] != 'linear'] or None)
      if f is not None:
        for i, a in enumerate(f(v)):
          rec[f'f32_constforce_{i}'] = np.asarray(a.data)
  jax.config.update('jax_enable_x64', False)
  path = os.path.join(OUT, name + '.npz')
  np.savez_compressed(path, **rec)
  print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


def run_projection_case(name, shape, domain, seed):
  grid = cfd.grids.Grid(shape, domain=domain)
  rs = np.random.RandomState(seed)
  arrays = [rs.standard_normal(shape).astype(np.float32) for _ in shape]
  rec = dict(shape=np.array(shape), domain=np.array(domain, dtype=np.float64))
  for i, a in enumerate(arrays):
    rec[f'v0_{i}'] = a
  for prec in ('f32', 'f64'):
    jax.config.update('jax_enable_x64', prec == 'f64')
    dtype = np.float32 if prec == 'f32' else np.float64
    v = wrap(grid, [a.astype(dtype) for a in arrays])
    q = cfd.pressure.solve_fast_diag(v)
    vp = cfd.pressure.projection(v)
    rec[f'{prec}_q'] = np.asarray(q.data)
    for i, a in enumerate(vp):
      rec[f'{prec}_proj_{i}'] = np.asarray(a.data)
  jax.config.update('jax_enable_x64', False)
  path = os.path.join(OUT, name + '.npz')
  np.savez_compressed(path, **rec)
  print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


def run_implicit_case(name, shape, domain, seed, density, viscosity, dt, forcing_spec, nsteps):
  """equations.implicit_diffusion_navier_stokes (equations.py:154-195) run by the reference."""
  grid = cfd.grids.Grid(shape, domain=domain)
  v0 = cfd_oracle.filtered_velocity_field(seed, shape, domain, 1.0, 2, dtype=np.float64)
  rec = dict(shape=np.array(shape), domain=np.array(domain, dtype=np.float64), seed=seed,
             density=density, viscosity=viscosity, dt=dt, smag_cs=-1.0,
             forcing_spec=repr(forcing_spec), nsteps=np.array(nsteps), stepper='implicit_diffusion')
  for i, a in enumerate(v0):
    rec[f'v0_{i}'] = a.astype(np.float32)
  for prec in ('f32', 'f64'):
    jax.config.update('jax_enable_x64', prec == 'f64')
    dtype = np.float32 if prec == 'f32' else np.float64
    v = wrap(grid, [a.astype(np.float32).astype(dtype) for a in v0])
    step = cfd.equations.implicit_diffusion_navier_stokes(
        density=density, viscosity=viscosity, dt=dt, grid=grid, forcing=build_forcing(grid, forcing_spec))
    cur, done = v, 0
    for n in nsteps:
      for _ in range(n - done):
        cur = step(cur)
      done = n
      for i, a in enumerate(cur):
        assert a.data.dtype == dtype, (a.data.dtype, dtype)
        rec[f'{prec}_v{n}_{i}'] = np.asarray(a.data)
  jax.config.update('jax_enable_x64', False)
  path = os.path.join(OUT, name + '.npz')
  np.savez_compressed(path, **rec)
  print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


def run_post_case(name, shape, seed, factors):
  """resize.downsample_staggered_velocity (resize.py:187-222) of a random staggered field, every
  component along its own direction, for each coarse-graining factor."""
  from jax_cfd.base import resize
  nd = len(shape)
  domain = tuple((0.0, 2 * np.pi) for _ in shape)
  grid = cfd.grids.Grid(shape, domain=domain)
  rs = np.random.RandomState(seed)
  v0 = [rs.standard_normal(shape).astype(np.float32) for _ in range(nd)]
  rec = {'shape': np.asarray(shape), 'domain': np.asarray(domain), 'factors': np.asarray(factors)}
  for i, a in enumerate(v0):
    rec[f'v0_{i}'] = a
  for f in factors:
    dst = cfd.grids.Grid(tuple(n // f for n in shape), domain=domain)
    out = resize.downsample_staggered_velocity(grid, dst, wrap(grid, v0))
    for i, u in enumerate(out):
      assert u.offset == grid.cell_faces[i] and u.grid == dst
      rec[f'down{f}_{i}'] = np.asarray(u.data, np.float32)
  path = os.path.join(OUT, name + '.npz')
  np.savez_compressed(path, **rec)
  print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


def cases():
  """name -> thunk.  New fixtures are appended at the end; running without arguments rewrites every
  fixture (bit-identical on re-generation), `gen_golden.py NAME...` only the named ones."""
  two_pi = 2 * np.pi
  d2 = ((0.0, two_pi), (0.0, two_pi))
  d3 = d2 + ((0.0, two_pi),)
  kolm = [('kolmogorov', dict(scale=1.0, k=4)), ('linear', -0.1)]
  c = {}

  def case(name, *a, **k):
    c[name] = lambda: run_case(name, *a, **k)

  def pcase(name, *a, **k):
    c[name] = lambda: run_projection_case(name, *a, **k)

  # K64x32: Kolmogorov (paper config: scale 1, k 4, linear -0.1, nu 1e-3), non-square pow2 grid
  case('k2d_64x32', (64, 32), d2, 1, 3.0, 3, 1.0, 1e-3, 0.5 * (two_pi / 64) / 3.0, kolm, None, [1, 10])
  # demo.ipynb-like decaying turbulence (no forcing), square
  case('d2d_128', (128, 128), d2, 2, 2.0, 3, 1.0, 1e-3, 0.5 * (two_pi / 128) / 2.0, None, None, [1, 20])
  # density != 1, anisotropic domain, swap_xy kolmogorov
  case('k2d_32x64_rho', (32, 64), ((0.0, two_pi), (0.0, 2 * two_pi)), 3, 1.5, 2, 2.0, 5e-3,
       0.01, [('linear', 0.05), ('kolmogorov', dict(scale=0.5, k=2, swap_xy=True))], None, [1, 5])
  # non power-of-two grid
  case('d2d_48x36', (48, 36), d2, 4, 1.0, 3, 1.0, 1e-2, 0.02, None, None, [1, 3])
  # inviscid (viscosity=None is allowed by equations.py:106) with taylor-green forcing
  case('tg2d_32', (32, 32), d2, 5, 1.0, 2, 1.0, 1e-2, 0.02,
       [('taylor_green', dict(scale=0.7, k=2))], None, [1, 4])
  # 3-D with Smagorinsky closure (config #5 in miniature)
  case('s3d_16', (16, 16, 16), d3, 6, 1.0, 2, 1.0, 1.0 / 1600, 0.05, None, 0.2, [1, 5])
  case('s3d_16x8x32_kolm', (16, 8, 32), d3, 7, 1.0, 2, 1.0, 1e-2, 0.03,
       [('kolmogorov', dict(scale=1.0, k=2)), ('linear', -0.1)], 0.17, [1, 3])
  case('d3d_16', (16, 16, 16), d3, 8, 1.0, 2, 1.0, 1e-2, 0.05, None, None, [1, 3])
  # 3-D cases at sizes the CUDA line FFTs accept (axes >= 16, last axis >= 32)
  case('s3d_16x16x32', (16, 16, 32), d3, 14, 1.0, 2, 1.0, 1.0 / 1600, 0.04,
       [('kolmogorov', dict(scale=1.0, k=2)), ('linear', -0.1)], 0.2, [1, 3])
  case('d3d_32x16x32', (32, 16, 32), d3, 15, 1.0, 2, 1.0, 1e-2, 0.04, None, None, [1, 3])
  case('tg3d_16x32x32', (16, 32, 32), d3, 16, 1.0, 2, 1.0, 5e-3, 0.03,
       [('taylor_green', dict(scale=0.5, k=1))], 0.15, [1, 2])
  # RK steppers ("next" row f2)
  case('rk4_2d_32', (32, 32), d2, 9, 1.0, 2, 1.0, 1e-2, 0.04, kolm, None, [1, 3], stepper='classic_rk4')
  case('rk2_2d_32', (32, 32), d2, 9, 1.0, 2, 1.0, 1e-2, 0.04, kolm, None, [1, 3], stepper='midpoint_rk2')
  pcase('proj2d_64x32', (64, 32), d2, 11)
  pcase('proj3d_16x8x32', (16, 8, 32), d3, 12)
  pcase('proj3d_32x16x64', (32, 16, 64), d3, 17)
  pcase('proj2d_step1_30x20', (30, 20), ((0.0, 30.0), (0.0, 20.0)), 13)
  # ---- round 2 ----
  # Smagorinsky closure in 2-D (subgrid_models.py is dimension agnostic; subgrid_models_test.py:117-217)
  case('s2d_64x32', (64, 32), d2, 21, 2.0, 3, 1.0, 1e-3, 0.01, kolm, 0.2, [1, 4])
  case('s2d_100', (100, 100), d2, 22, 1.0, 3, 1.0, 1e-3, 0.01, None, 0.2, [1, 3])
  # grids that are not powers of two, odd axes (fast_diagonalization.py:101-108: matmul fallback)
  case('d2d_100', (100, 100), d2, 23, 1.0, 3, 1.0, 1e-2, 0.02, kolm, None, [1, 3])
  case('d2d_33x27', (33, 27), d2, 24, 1.0, 2, 1.0, 1e-2, 0.02, None, None, [1, 3])
  case('d3d_20x24x36', (20, 24, 36), d3, 25, 1.0, 2, 1.0, 1e-2, 0.04, None, None, [1, 2])
  case('s3d_24x20x12', (24, 20, 12), d3, 26, 1.0, 2, 1.0, 1e-2, 0.04, [('linear', -0.1)], 0.2, [1, 2])
  # implicit diffusion (equations.py:154-195; what the ML configs step with)
  c['imp2d_64x32'] = lambda: run_implicit_case('imp2d_64x32', (64, 32), d2, 31, 1.5, 5e-2, 0.02, kolm, [1, 4])
  c['imp2d_48x36'] = lambda: run_implicit_case('imp2d_48x36', (48, 36), d2, 32, 1.0, 1e-2, 0.02, None, [1, 3])
  c['imp3d_16x16x32'] = lambda: run_implicit_case('imp3d_16x16x32', (16, 16, 32), d3, 33, 1.0, 2e-2, 0.03,
                                                  [('linear', -0.1)], [1, 2])
  # flux-preserving coarse-graining of trajectories (resize.py:38-74, 187-222; "next" row f4)
  c['post2d_32x48'] = lambda: run_post_case('post2d_32x48', (32, 48), 41, [2, 4])
  c['post3d_8x12x16'] = lambda: run_post_case('post3d_8x12x16', (8, 12, 16), 42, [2, 4])
  return c


def main(argv):
  os.makedirs(OUT, exist_ok=True)
  table = cases()
  names = argv or list(table)
  for n in names:
    table[n]()


if __name__ == '__main__':
  main(sys.argv[1:])
