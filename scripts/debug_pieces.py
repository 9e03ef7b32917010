"""Debug helper: error of each stage vs the oracle for several sizes (run on the GPU box)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
  sys.path.insert(0, p)
import numpy as np
import jax_cfd_b200 as cfd
import cfd_oracle
import golden_util as gu

def wrap(grid, arrays):
  bc = cfd.boundaries.periodic_boundary_conditions(grid.ndim)
  return tuple(cfd.grids.GridVariable(cfd.grids.GridArray(cfd.DeviceArray.from_numpy(np.ascontiguousarray(a, np.float32)), o, grid), bc)
               for a, o in zip(arrays, grid.cell_faces))

shapes = [tuple(int(x) for x in s.split('x')) for s in (sys.argv[1:] or ['64x32', '32x64', '128x128', '16x256', '256x2048', '512x512'])]
for shape in shapes:
  dom = ((0.0, 2 * np.pi),) * len(shape)
  grid = cfd.grids.Grid(shape, domain=dom)
  h = grid.step
  v0 = cfd_oracle.filtered_velocity_field(1, shape, dom, 3.0, 3)
  dt = 0.5 * min(h) / 3.0
  nu = 1e-3
  forcing = cfd.forcings.sum_forcings(cfd.forcings.kolmogorov_forcing(grid, k=4), cfd.forcings.linear_forcing(grid, -0.1))
  of = cfd_oracle.Forcing((('const', cfd_oracle.kolmogorov_field(shape, dom, 1.0, 4)), ('linear', -0.1)))
  f = cfd.equations.navier_stokes_explicit_terms(1.0, nu, dt, grid, forcing=forcing)
  got = [np.asarray(u.data) for u in f(wrap(grid, v0))]
  want = cfd_oracle.explicit_terms(v0, dt, h, nu, of, 1.0)
  e_exp = [gu.rel_l2(a, b) for a, b in zip(got, want)]
  # no forcing / no viscosity variants
  f2 = cfd.equations.navier_stokes_explicit_terms(1.0, None, dt, grid)
  got2 = [np.asarray(u.data) for u in f2(wrap(grid, v0))]
  want2 = cfd_oracle.explicit_terms(v0, dt, h, None, None, 1.0)
  e_adv = [gu.rel_l2(a, b) for a, b in zip(got2, want2)]
  rs = np.random.RandomState(0)
  w = [rs.standard_normal(shape).astype(np.float32) for _ in shape]
  vp, q = cfd._engine.NativeProjection(grid)(wrap(grid, w), return_q=True)
  wp, wq = cfd_oracle.projection(tuple(w), h)
  e_q = gu.rel_l2(np.asarray(q.data), wq)
  e_p = [gu.rel_l2(np.asarray(a.data), b) for a, b in zip(vp, wp)]
  step = cfd.equations.semi_implicit_navier_stokes(1.0, nu, dt, grid, forcing=forcing)
  s1, q1 = step.advance(wrap(grid, v0), 1, return_q=True)
  ws, wq1 = cfd_oracle.step(v0, dt, h, 1.0, nu, of, return_q=True)
  e_s = [gu.rel_l2(np.asarray(a.data), b) for a, b in zip(s1, ws)]
  print(shape, 'explicit', ['%.1e' % e for e in e_exp], 'adv-only', ['%.1e' % e for e in e_adv],
        'q', '%.1e' % e_q, 'proj', ['%.1e' % e for e in e_p], 'step', ['%.1e' % e for e in e_s],
        'q1 %.1e' % gu.rel_l2(np.asarray(q1), wq1), flush=True)
  # chained steps (lazy projection) vs oracle
  vn = cfd.funcutils.repeated(step, 4)(wrap(grid, v0))
  wn = v0
  for _ in range(4):
    wn = cfd_oracle.step(wn, dt, h, 1.0, nu, of)
  print('   chain4', ['%.1e' % gu.rel_l2(np.asarray(a.data), b) for a, b in zip(vn, wn)], flush=True)
