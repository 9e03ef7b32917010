"""Parity of the CUDA path (through the C ABI) against the oracle and the reference's golden
vectors.  Bar: per-step relative L2 <= 1e-5 on every velocity component
and on q.  All tests here need a GPU."""
import numpy as np
import pytest

import cfd_oracle
import golden_util as gu

pytestmark = pytest.mark.gpu

TOL = 1e-5  # the north star's per-step relative L2 bar (float32)


@pytest.fixture(scope='module')
def cfd():
  import jax_cfd_b200 as m
  m._lib.require_device()
  return m


def make_forcing(cfd, grid, rec):
  fs = []
  for kind, arg in (rec['forcing_spec'] or []):
    if kind == 'kolmogorov':
      fs.append(cfd.forcings.kolmogorov_forcing(grid, **arg))
    elif kind == 'taylor_green':
      fs.append(cfd.forcings.taylor_green_forcing(grid, **arg))
    elif kind == 'linear':
      fs.append(cfd.forcings.linear_forcing(grid, arg))
  if not fs:
    return None
  return cfd.forcings.sum_forcings(*fs) if len(fs) > 1 else fs[0]


def wrap(cfd, grid, arrays, device=True):
  bc = cfd.boundaries.periodic_boundary_conditions(grid.ndim)
  conv = (lambda a: cfd.DeviceArray.from_numpy(np.ascontiguousarray(a, np.float32))) if device else (
      lambda a: np.ascontiguousarray(a, np.float32))
  return tuple(cfd.grids.GridVariable(cfd.grids.GridArray(conv(a), o, grid), bc)
               for a, o in zip(arrays, grid.cell_faces))


def to_np(v):
  return [np.asarray(u.data) for u in v]


def build_step(cfd, rec, grid, stepper=None):
  kw = dict(density=rec['density'], viscosity=rec['viscosity'], dt=rec['dt'], grid=grid,
            forcing=make_forcing(cfd, grid, rec))
  if stepper is not None:
    kw['time_stepper'] = getattr(cfd.time_stepping, stepper)
  if rec['smag_cs'] >= 0:
    dt = kw.pop('dt')
    return cfd.subgrid_models.explicit_smagorinsky_navier_stokes(dt=dt, cs=rec['smag_cs'], **kw)
  return cfd.equations.semi_implicit_navier_stokes(**kw)


GOLDEN_2D = ['k2d_64x32', 'd2d_128', 'k2d_32x64_rho', 'tg2d_32']
GOLDEN_3D = ['s3d_16x16x32', 'd3d_32x16x32', 'tg3d_16x32x32']


@pytest.mark.parametrize('name', GOLDEN_2D + GOLDEN_3D)
def test_step_matches_reference_golden(cfd, name):
  rec = gu.load(name)
  grid = cfd.grids.Grid(rec['shape'], domain=rec['domain'])
  step = build_step(cfd, rec, grid)
  v = wrap(cfd, grid, [rec[f'v0_{i}'] for i in range(rec['ndim'])])
  # first step, with q
  v1, q = step.advance(v, 1, return_q=True)
  for i, a in enumerate(to_np(v1)):
    assert gu.rel_l2(a, rec[f'f32_v1_{i}']) < TOL
    assert gu.rel_l2(a, rec[f'f64_v1_{i}']) < TOL
  assert gu.rel_l2(np.asarray(q), rec['f32_q']) < TOL
  # all recorded step counts through repeated()
  for n in rec['nsteps']:
    vn = cfd.funcutils.repeated(step, n)(v)
    for i, a in enumerate(to_np(vn)):
      err = gu.rel_l2(a, rec[f'f32_v{n}_{i}'])
      assert err < TOL, (name, n, i, err)  # measured: <= 2e-7 at n = 20


@pytest.mark.parametrize('name', ['rk4_2d_32', 'rk2_2d_32'])
def test_rk_steppers_match_reference_golden(cfd, name):
  rec = gu.load(name)
  grid = cfd.grids.Grid(rec['shape'], domain=rec['domain'])
  step = build_step(cfd, rec, grid, stepper=rec['stepper'])
  v = wrap(cfd, grid, [rec[f'v0_{i}'] for i in range(2)])
  done = 0
  for n in rec['nsteps']:
    for _ in range(n - done):
      v = step(v)
    done = n
    for i, a in enumerate(to_np(v)):
      assert gu.rel_l2(a, rec[f'f32_v{n}_{i}']) < TOL * n


@pytest.mark.parametrize('name', ['proj2d_64x32', 'proj3d_32x16x64'])
def test_projection_matches_reference_golden(cfd, name):
  rec = gu.load(name)
  grid = cfd.grids.Grid(rec['shape'], domain=rec['domain'])
  v = wrap(cfd, grid, [rec[f'v0_{i}'] for i in range(rec['ndim'])])
  vp = cfd.pressure.projection(v)
  q = cfd.pressure.solve_fast_diag(v)
  assert q.offset == grid.cell_center
  assert gu.rel_l2(np.asarray(q.data), rec['f32_q']) < TOL
  for i, a in enumerate(to_np(vp)):
    assert gu.rel_l2(a, rec[f'f32_proj_{i}']) < TOL
    assert vp[i].offset == grid.cell_faces[i]
  # pressure_test.py:59-110: div(projection(v)) ~ 0
  div = cfd_oracle.divergence(to_np(vp), rec['h'])
  assert np.abs(div).max() < 1e-3  # white-noise input of magnitude ~3/h


# The thin shapes run every FFT kernel variant at its full-size line length (x lines of 4096 / 8192
# points: 32 points per thread, two CTAs per SM; 16384: one 1024-thread CTA; 32768: the split
# transform; rows of 8192 / 16384 reals) at a cell count the oracle finishes in seconds.
@pytest.mark.parametrize('shape,nsteps', [((256, 256), 10), ((512, 1024), 3), ((1024, 64), 3),
                                         ((16, 32), 5), ((2048, 2048), 1), ((8192, 64), 3),
                                         ((4096, 128), 3), ((16384, 32), 2), ((32768, 32), 2),
                                         ((64, 8192), 3), ((32, 16384), 2)])
def test_step_matches_oracle(cfd, shape, nsteps):
  dom = ((0.0, 2 * np.pi), (0.0, 2 * np.pi))
  grid = cfd.grids.Grid(shape, domain=dom)
  v0 = cfd_oracle.filtered_velocity_field(42, shape, dom, 3.0, 4)
  dt = 0.5 * min(grid.step) / 3.0
  nu = 1e-3 if max(shape) <= 512 else 1e-4
  forcing = cfd.forcings.sum_forcings(cfd.forcings.kolmogorov_forcing(grid, k=4),
                                      cfd.forcings.linear_forcing(grid, -0.1))
  step = cfd.equations.semi_implicit_navier_stokes(1.0, nu, dt, grid, forcing=forcing)
  of = cfd_oracle.Forcing((('const', cfd_oracle.kolmogorov_field(shape, dom, 1.0, 4)),
                           ('linear', -0.1)))
  got, q = step.advance(wrap(cfd, grid, v0), nsteps, return_q=True)
  want = v0
  for n in range(nsteps):
    want, wq = cfd_oracle.step(want, dt, grid.step, 1.0, nu, of, return_q=True)
  for a, b in zip(to_np(got), want):
    assert gu.rel_l2(a, b) < TOL
  if max(shape) < 64 * min(shape):
    assert gu.rel_l2(np.asarray(q), wq) < TOL
  else:
    # Cells stretched 128:1 and more: q itself is ill-conditioned in float32 (the reference's own
    # f32 run is 6e-6 away from its f64 run at 16384x32), so q is held to the float32 reference's
    # distance from the float64 answer instead; the velocities above keep the 1e-5 bar.
    w64 = [a.astype(np.float64) for a in v0]
    of64 = cfd_oracle.Forcing((('const', [a.astype(np.float64) for a in
                                          cfd_oracle.kolmogorov_field(shape, dom, 1.0, 4)]),
                               ('linear', -0.1)))
    for n in range(nsteps):
      w64, q64 = cfd_oracle.step(w64, dt, grid.step, 1.0, nu, of64, return_q=True)
    assert q64.dtype == np.float64
    floor = gu.rel_l2(wq, q64)
    assert gu.rel_l2(np.asarray(q), q64) < max(TOL, 3 * floor)
  # divergence-free residual (equations_test.py:99: max|div| stays small); it is rounding noise
  # divided by h, so on the very fine thin grids it is held to the reference's own residual
  ref_div = np.abs(cfd_oracle.divergence(want, grid.step)).max()
  assert np.abs(cfd_oracle.divergence(to_np(got), grid.step)).max() < max(2e-3, 3 * ref_div)


@pytest.mark.parametrize('shape', [(8192, 64), (32768, 64)])
def test_paired_spectrum_layout_is_bitwise_equal_to_plain(cfd, shape, monkeypatch):
  """The pair-interleaved spectrum layout T[ky/2][x][ky&1] (chosen by the plan for rows of >= 16384
  reals; forced here on a thin grid) changes where the numbers live, not the numbers: x lines of
  8192 points (stride-2 line kernel) and of 32768 points (4-CTA cluster kernel with the DSMEM
  radix-2 exchange, vs the 2-CTA cluster of the plain layout)."""
  from jax_cfd_b200 import _engine
  dom = ((0.0, 2 * np.pi), (0.0, 2 * np.pi))
  grid = cfd.grids.Grid(shape, domain=dom)
  v0 = cfd_oracle.filtered_velocity_field(5, shape, dom, 3.0, 4)
  dt = 0.5 * min(grid.step) / 3.0
  forcing = cfd.forcings.sum_forcings(cfd.forcings.kolmogorov_forcing(grid, k=4),
                                      cfd.forcings.linear_forcing(grid, -0.1))

  def run():
    with _engine._plans_lock:
      _engine._plans.clear()
    step = cfd.equations.semi_implicit_navier_stokes(1.0, 1e-4, dt, grid, forcing=forcing)
    got, q = step.advance(wrap(cfd, grid, v0), 3, return_q=True)
    return to_np(got) + [np.asarray(q)]

  try:
    monkeypatch.setenv('CFD_T_PAIRED', '0')
    plain = run()
    monkeypatch.setenv('CFD_T_PAIRED', '1')
    paired = run()
  finally:
    with _engine._plans_lock:
      _engine._plans.clear()
  for a, b in zip(plain, paired):
    np.testing.assert_array_equal(a, b)


def test_host_and_device_paths_agree_bitwise(cfd):
  rec = gu.load('k2d_64x32')
  grid = cfd.grids.Grid(rec['shape'], domain=rec['domain'])
  step = build_step(cfd, rec, grid)
  arrays = [rec[f'v0_{i}'] for i in range(2)]
  dev = to_np(cfd.funcutils.repeated(step, 4)(wrap(cfd, grid, arrays, device=True)))
  host_v = cfd.funcutils.repeated(step, 4)(wrap(cfd, grid, arrays, device=False))
  for a, u in zip(dev, host_v):
    assert isinstance(u.data, np.ndarray)
    np.testing.assert_array_equal(a, u.data)
  loop = wrap(cfd, grid, arrays, device=True)
  for _ in range(4):
    loop = step(loop)
  for a, b in zip(dev, to_np(loop)):
    np.testing.assert_array_equal(a, b)


def test_batched_members_equal_single_runs(cfd):
  shape = (64, 128)
  dom = ((0.0, 2 * np.pi), (0.0, 2 * np.pi))
  grid = cfd.grids.Grid(shape, domain=dom)
  members = [cfd_oracle.filtered_velocity_field(s, shape, dom, 2.0, 3) for s in range(3)]
  forcing = cfd.forcings.sum_forcings(cfd.forcings.kolmogorov_forcing(grid, k=4),
                                      cfd.forcings.linear_forcing(grid, -0.1))
  step = cfd.equations.semi_implicit_navier_stokes(1.0, 1e-3, 0.01, grid, forcing=forcing)
  batched = [np.stack([m[c] for m in members]) for c in range(2)]
  out_b = to_np(cfd.funcutils.repeated(step, 3)(wrap(cfd, grid, batched)))
  for s, m in enumerate(members):
    out_s = to_np(cfd.funcutils.repeated(step, 3)(wrap(cfd, grid, m)))
    for c in range(2):
      np.testing.assert_array_equal(out_b[c][s], out_s[c])


def test_explicit_terms_match_oracle(cfd):
  rec = gu.load('k2d_64x32')
  grid = cfd.grids.Grid(rec['shape'], domain=rec['domain'])
  f = cfd.equations.navier_stokes_explicit_terms(rec['density'], rec['viscosity'], rec['dt'], grid,
                                                 forcing=make_forcing(cfd, grid, rec))
  arrays = [rec[f'v0_{i}'] for i in range(2)]
  got = to_np(f(wrap(cfd, grid, arrays)))
  want = cfd_oracle.explicit_terms(tuple(arrays), rec['dt'], rec['h'],
                                   rec['viscosity'] / rec['density'], gu.oracle_forcing(rec),
                                   rec['density'])
  for a, b in zip(got, want):
    assert gu.rel_l2(a, b) < TOL


def test_diagnostics_match_oracle(cfd):
  rec = gu.load('d2d_128')
  grid = cfd.grids.Grid(rec['shape'], domain=rec['domain'])
  arrays = [rec[f'f32_v20_{i}'] for i in range(2)]
  got = cfd.diagnostics(wrap(cfd, grid, arrays))
  want = cfd_oracle.diagnostics(arrays, rec['h'])
  assert abs(got['kinetic_energy'] - want['kinetic_energy']) < 1e-6 * want['kinetic_energy']
  assert abs(got['enstrophy'] - want['enstrophy']) < 1e-5 * want['enstrophy']
  assert abs(got['max_speed_sq'] - want['max_speed_sq']) < 1e-6 * want['max_speed_sq']
  assert abs(got['max_abs_div'] - want['max_div']) < 1e-4


def test_full_size_properties_2048(cfd):
  """BASELINE config #2 size: size-independent properties (no oracle needed)."""
  shape = (2048, 2048)
  dom = ((0.0, 2 * np.pi), (0.0, 2 * np.pi))
  grid = cfd.grids.Grid(shape, domain=dom)
  rs = np.random.RandomState(0)
  k = np.arange(shape[0]) * (2 * np.pi / shape[0])
  u = (np.sin(3 * k)[:, None] * np.cos(5 * k)[None, :] + 0.1 * rs.standard_normal(shape)).astype(np.float32)
  v = (np.cos(2 * k)[:, None] * np.sin(7 * k)[None, :] + 0.1 * rs.standard_normal(shape)).astype(np.float32)
  vv = wrap(cfd, grid, [u, v])
  p1 = cfd.pressure.projection(vv)
  a1 = to_np(p1)
  # (1) divergence-free, (2) idempotent, (3) mean (momentum) preserved, (4) linear
  # float32 cancellation floor: eps * |u| / h ~ 1e-4 of the input divergence (oracle: 1.2e-4)
  assert np.abs(cfd_oracle.divergence(a1, grid.step)).max() < 5e-4 * np.abs(cfd_oracle.divergence([u, v], grid.step)).max()
  a2 = to_np(cfd.pressure.projection(p1))
  for x, y in zip(a1, a2):
    assert gu.rel_l2(y, x) < 5e-5  # float32 floor for this noisy field (oracle: ~1.5e-5)
  for x, y in zip([u, v], a1):
    assert abs(float(x.mean(dtype=np.float64)) - float(y.mean(dtype=np.float64))) < 1e-6
  a3 = to_np(cfd.pressure.projection(wrap(cfd, grid, [2 * u, 2 * v])))
  for x, y in zip(a1, a3):
    assert gu.rel_l2(y, 2 * x) < 1e-6
  # full step: stays divergence free, momentum conserved without forcing (equations_test.py:84-163)
  step = cfd.equations.semi_implicit_navier_stokes(1.0, 1e-4, 0.2 * grid.step[0], grid)
  s = to_np(cfd.funcutils.repeated(step, 5)(p1))
  assert np.abs(cfd_oracle.divergence(s, grid.step)).max() < 1e-2
  for x, y in zip(a1, s):
    assert abs(float(x.mean(dtype=np.float64)) - float(y.mean(dtype=np.float64))) < 1e-5


def test_unsupported_options_raise(cfd):
  grid = cfd.grids.Grid((64, 64), domain=((0, 1), (0, 1)))
  with pytest.raises(NotImplementedError):
    cfd.equations.semi_implicit_navier_stokes(1.0, 1e-3, 0.01, grid, convect=lambda v: v)
  with pytest.raises(NotImplementedError):
    cfd.equations.semi_implicit_navier_stokes(1.0, 1e-3, 0.01, grid, forcing=lambda v: v)
  step = cfd.equations.semi_implicit_navier_stokes(1.0, 1e-3, 0.01, grid)
  bc = cfd.boundaries.periodic_boundary_conditions(2)
  bad = tuple(cfd.grids.GridVariable(cfd.grids.GridArray(np.zeros((64, 64), np.float32), (0.5, 0.5), grid), bc)
              for _ in range(2))
  with pytest.raises(cfd.grids.InconsistentOffsetError):
    step(bad)
  g1 = cfd.grids.Grid((64,), domain=((0, 1),))
  with pytest.raises(NotImplementedError):
    cfd.equations.semi_implicit_navier_stokes(1.0, 1e-3, 0.01, g1)  # 1-D grids


@pytest.mark.parametrize('shape,cs', [((64, 32, 64), 0.2), ((32, 64, 128), None)])
def test_step_3d_matches_oracle(cfd, shape, cs):
  """Config #5 in miniature: Taylor-Green-like field, Smagorinsky closure, vs the oracle."""
  dom = ((0.0, 2 * np.pi),) * 3
  grid = cfd.grids.Grid(shape, domain=dom)
  v0 = cfd_oracle.filtered_velocity_field(7, shape, dom, 1.0, 2)
  dt = 0.5 * min(grid.step) / 1.0
  nu = 1.0 / 1600
  lin = cfd.forcings.linear_forcing(grid, 0.05)
  if cs is not None:
    step = cfd.subgrid_models.explicit_smagorinsky_navier_stokes(
        dt=dt, cs=cs, forcing=lin, density=1.0, viscosity=nu, grid=grid)
    of = cfd_oracle.Forcing((('linear', 0.05), ('smagorinsky', cs)))
  else:
    step = cfd.equations.semi_implicit_navier_stokes(1.0, nu, dt, grid, forcing=lin)
    of = cfd_oracle.Forcing((('linear', 0.05),))
  got, q = step.advance(wrap(cfd, grid, v0), 2, return_q=True)
  want = v0
  for _ in range(2):
    want, wq = cfd_oracle.step(want, dt, grid.step, 1.0, nu, of, return_q=True)
  for a, b in zip(to_np(got), want):
    assert gu.rel_l2(a, b) < TOL
  assert gu.rel_l2(np.asarray(q), wq) < TOL
  assert np.abs(cfd_oracle.divergence(to_np(got), grid.step)).max() < 1e-3
  d = cfd.diagnostics(got)
  wd = cfd_oracle.diagnostics(to_np(got), grid.step)
  assert abs(d['kinetic_energy'] - wd['kinetic_energy']) < 1e-6 * wd['kinetic_energy']
  assert abs(d['max_speed_sq'] - wd['max_speed_sq']) < 1e-6 * wd['max_speed_sq']


def test_slab_stepper_world1_equals_single_gpu_path(cfd):
  """The distributed code path (slab sources, peer-aware x lines, correct with q-next) with one
  rank must reproduce the ordinary path bit for bit."""
  shape = (128, 256)
  dom = ((0.0, 2 * np.pi), (0.0, 2 * np.pi))
  grid = cfd.grids.Grid(shape, domain=dom)
  v0 = cfd_oracle.filtered_velocity_field(3, shape, dom, 3.0, 3)
  dt = 0.5 * min(grid.step) / 3.0
  forcing = cfd.forcings.sum_forcings(cfd.forcings.kolmogorov_forcing(grid, k=4),
                                      cfd.forcings.linear_forcing(grid, -0.1))
  st = cfd.distributed.SlabStepper(grid, dt, 1.0, 1e-3, forcing, rank=0, world=1, device=0,
                                   exchange=lambda b: [b])
  st.load(list(v0))
  st.advance(3)
  st.advance(2)
  outs, q = st.store(want_q=True)
  step = cfd.equations.semi_implicit_navier_stokes(1.0, 1e-3, dt, grid, forcing=forcing)
  ref, rq = step.advance(wrap(cfd, grid, v0), 5, return_q=True)
  for a, b in zip(outs, to_np(ref)):
    np.testing.assert_array_equal(a.numpy(), b)
  np.testing.assert_array_equal(q.numpy(), np.asarray(rq))
  st.close()


def test_slab_stepper_3d_world1_equals_single_gpu_path(cfd):
  """The 3-D distributed code path (slab plane sources in the marching kernels, peer-aware x lines,
  divergence / correction with the neighbour plane) with one rank reproduces the ordinary path bit
  for bit -- Smagorinsky closure included (config #5)."""
  shape = (32, 32, 64)
  dom = ((0.0, 2 * np.pi),) * 3
  grid = cfd.grids.Grid(shape, domain=dom)
  v0 = cfd_oracle.filtered_velocity_field(7, shape, dom, 1.0, 2)
  dt = 0.5 * min(grid.step) / 1.0
  nu = 1.0 / 1600
  lin = cfd.forcings.linear_forcing(grid, 0.05)
  forcing = cfd.forcings.sum_forcings(lin, cfd._engine.ForcingFn([cfd._engine.SmagorinskyTerm(0.2)]))
  st = cfd.distributed.SlabStepper(grid, dt, 1.0, nu, forcing, rank=0, world=1, device=0,
                                   exchange=lambda b: [b])
  st.load(list(v0))
  st.advance(2)
  st.advance(1)
  outs, q = st.store(want_q=True)
  step = cfd.subgrid_models.explicit_smagorinsky_navier_stokes(
      dt=dt, cs=0.2, forcing=lin, density=1.0, viscosity=nu, grid=grid)
  ref, rq = step.advance(wrap(cfd, grid, v0), 3, return_q=True)
  for a, b in zip(outs, to_np(ref)):
    np.testing.assert_array_equal(a.numpy(), b)
  np.testing.assert_array_equal(q.numpy(), np.asarray(rq))
  st.close()


def test_filtered_velocity_field_is_divergence_free_with_requested_speed(cfd):
  """initial_conditions.py:71-121 ("next" row f3): projection + fused max-speed reduction."""
  grid = cfd.grids.Grid((256, 256), domain=((0.0, 2 * np.pi), (0.0, 2 * np.pi)))
  v = cfd.initial_conditions.filtered_velocity_field(0, grid, maximum_velocity=2.0, peak_wavenumber=3)
  arrs = to_np(v)
  speed = np.sqrt((arrs[0] ** 2 + arrs[1] ** 2).max())
  assert abs(speed - 2.0) < 1e-3
  assert np.abs(cfd_oracle.divergence(arrs, grid.step)).max() < 2e-3
  for u, o in zip(v, grid.cell_faces):
    assert u.offset == o and u.data.dtype == np.float32
  # dynamic_time_step (equations.py:62-70) uses the same reduction
  dt = cfd.equations.dynamic_time_step(v, 0.5, 1e-3, grid)
  assert abs(dt - 0.5 * grid.step[0] / speed) < 1e-6


def test_trajectory_on_device_matches_repeated(cfd):
  """funcutils.trajectory (funcutils.py:95-126) with inner repeated steps, device-resident state."""
  rec = gu.load('k2d_64x32')
  grid = cfd.grids.Grid(rec['shape'], domain=rec['domain'])
  step = build_step(cfd, rec, grid)
  v = wrap(cfd, grid, [rec[f'v0_{i}'] for i in range(2)])
  inner = cfd.funcutils.repeated(step, 5)
  final, traj = cfd.funcutils.trajectory(inner, 2, post_process=lambda s: s)(v)
  assert traj[0].data.shape == (2, 64, 32)
  for i, a in enumerate(to_np(final)):
    assert gu.rel_l2(a, rec[f'f32_v10_{i}']) < TOL * 5
    np.testing.assert_array_equal(traj[i].data[-1], a)


def test_rk4_3d_runs_and_stays_divergence_free(cfd):
  """equations_test.py:101-128 style: RK stepper in 3-D, momentum and divergence invariants."""
  shape = (32, 32, 32)
  dom = ((0.0, 2 * np.pi),) * 3
  grid = cfd.grids.Grid(shape, domain=dom)
  v0 = cfd_oracle.filtered_velocity_field(5, shape, dom, 1.0, 2)
  step = cfd.equations.semi_implicit_navier_stokes(1.0, 1e-2, 0.02, grid,
                                                   time_stepper=cfd.time_stepping.classic_rk4)
  v = wrap(cfd, grid, v0)
  for _ in range(3):
    v = step(v)
  want = v0
  a, b = gu.TABLEAUS['classic_rk4']
  for _ in range(3):
    want = cfd_oracle.rk_step(want, 0.02, grid.step, a, b, 1.0, 1e-2, None)
  for x, y in zip(to_np(v), want):
    assert gu.rel_l2(x, y) < TOL
  assert np.abs(cfd_oracle.divergence(to_np(v), grid.step)).max() < 1e-4


def test_thousand_step_statistics_track_the_oracle(cfd):
  """North star: kinetic energy, enstrophy and the divergence-free residual tracked over 1000
  steps.  Kolmogorov 128^2 (paper forcing), CUDA path (chained, lazy projection) vs the float32
  oracle on the same initial condition.  Stated tolerances: while the two trajectories are still
  the same realisation (first 200 steps) fields agree to 1e-4 rel-L2 and KE / enstrophy to 1e-5
  relative; afterwards chaotic divergence of trajectories is expected, so only the statistics are
  compared: KE within 2 %, enstrophy within 5 % at every checkpoint up to step 1000, and
  max|div v| < 2e-3 throughout (equations_test.py:99,127 uses the same bar)."""
  shape = (128, 128)
  dom = ((0.0, 2 * np.pi), (0.0, 2 * np.pi))
  grid = cfd.grids.Grid(shape, domain=dom)
  v0 = cfd_oracle.filtered_velocity_field(11, shape, dom, 3.0, 4)
  nu = 1e-3
  dt = cfd.equations.stable_time_step(7.0, 0.5, nu, grid)
  forcing = cfd.forcings.sum_forcings(cfd.forcings.kolmogorov_forcing(grid, k=4),
                                      cfd.forcings.linear_forcing(grid, -0.1))
  step = cfd.equations.semi_implicit_navier_stokes(1.0, nu, dt, grid, forcing=forcing)
  of = cfd_oracle.Forcing((('const', cfd_oracle.kolmogorov_field(shape, dom, 1.0, 4)), ('linear', -0.1)))
  diag = cfd_oracle.pinv_diagonals(shape, grid.step, np.float32)
  chunk = 100
  advance = cfd.funcutils.repeated(step, chunk)
  v = wrap(cfd, grid, v0)
  w = v0
  rows = []
  for k in range(1, 11):
    v = advance(v)
    for _ in range(chunk):
      w = cfd_oracle.step(w, dt, grid.step, 1.0, nu, of, diag=diag)
    d = cfd.diagnostics(v)
    wd = cfd_oracle.diagnostics(w, grid.step)
    err = max(gu.rel_l2(a, b) for a, b in zip(to_np(v), w))
    rows.append((k * chunk, d['kinetic_energy'], wd['kinetic_energy'], d['enstrophy'], wd['enstrophy'],
                 d['max_abs_div'], err))
    assert d['max_abs_div'] < 2e-3
    assert abs(d['kinetic_energy'] - wd['kinetic_energy']) < 2e-2 * wd['kinetic_energy'], rows[-1]
    assert abs(d['enstrophy'] - wd['enstrophy']) < 5e-2 * wd['enstrophy'], rows[-1]
    assert err < 1e-4, rows[-1]  # measured: 1.6e-7 (step 100) ... 7.5e-7 (step 1000)
    if k <= 2:
      assert abs(d['kinetic_energy'] - wd['kinetic_energy']) < 1e-5 * wd['kinetic_energy']
      assert abs(d['enstrophy'] - wd['enstrophy']) < 1e-5 * wd['enstrophy']
  print('\nstep  KE(cuda)  KE(oracle)  Z(cuda)  Z(oracle)  max|div|  field rel-L2')
  for r in rows:
    print('%5d %.6f %.6f %.5f %.5f %.2e %.2e' % r)


def test_stepper_time_step_separate_from_convection_dt(cfd):
  """time_stepper=lambda ode, dt: forward_euler(ode, dt / 2): the Courant number inside
  Lax-Wendroff keeps the builder's dt (equations.py:127-128), the update uses dt / 2."""
  rec = gu.load('k2d_64x32')
  grid = cfd.grids.Grid(rec['shape'], domain=rec['domain'])
  dt = rec['dt']
  step = cfd.equations.semi_implicit_navier_stokes(
      rec['density'], rec['viscosity'], dt, grid, forcing=make_forcing(cfd, grid, rec),
      time_stepper=lambda ode, t: cfd.time_stepping.forward_euler(ode, t / 2))
  arrays = tuple(rec[f'v0_{i}'] for i in range(2))
  got = to_np(step(wrap(cfd, grid, arrays)))
  k0 = cfd_oracle.explicit_terms(arrays, dt, rec['h'], rec['viscosity'] / rec['density'],
                                 gu.oracle_forcing(rec), rec['density'])
  ustar = tuple(u + np.float32(dt / 2) * k for u, k in zip(arrays, k0))
  want, _ = cfd_oracle.projection(ustar, rec['h'])
  for a, b in zip(got, want):
    assert gu.rel_l2(a, b) < TOL
  plain = to_np(cfd.equations.semi_implicit_navier_stokes(
      rec['density'], rec['viscosity'], dt, grid, forcing=make_forcing(cfd, grid, rec))(wrap(cfd, grid, arrays)))
  assert gu.rel_l2(got[0], plain[0]) > 1e-4  # and it really is a different step


# ---------------------------------------------------------------------------------------------
# Round 2: grids of any shape (matmul fast diagonalisation + one-thread-per-cell stencils), the
# Smagorinsky closure in 2-D, implicit diffusion, the transform API, the device initial condition.

GOLDEN_ANY_SHAPE = ['d2d_48x36', 'd2d_100', 'd2d_33x27', 's2d_64x32', 's2d_100', 'd3d_20x24x36',
                    's3d_24x20x12', 's3d_16', 'd3d_16', 's3d_16x8x32_kolm']


@pytest.mark.parametrize('name', GOLDEN_ANY_SHAPE)
def test_step_matches_reference_golden_any_shape(cfd, name):
  """The reference's own test grids: 48x36, 100^2, odd axes (33x27: the reference's matmul
  fallback, fast_diagonalization.py:107-108), small 3-D grids, Smagorinsky in 2-D
  (subgrid_models_test.py:117-217)."""
  rec = gu.load(name)
  grid = cfd.grids.Grid(rec['shape'], domain=rec['domain'])
  step = build_step(cfd, rec, grid)
  v = wrap(cfd, grid, [rec[f'v0_{i}'] for i in range(rec['ndim'])])
  v1, q = step.advance(v, 1, return_q=True)
  for i, a in enumerate(to_np(v1)):
    assert gu.rel_l2(a, rec[f'f32_v1_{i}']) < TOL
    assert gu.rel_l2(a, rec[f'f64_v1_{i}']) < TOL
  assert gu.rel_l2(np.asarray(q), rec['f32_q']) < TOL
  for n in rec['nsteps']:
    vn = cfd.funcutils.repeated(step, n)(v)
    for i, a in enumerate(to_np(vn)):
      assert gu.rel_l2(a, rec[f'f32_v{n}_{i}']) < TOL, (name, n, i)


@pytest.mark.parametrize('name', ['proj3d_16x8x32', 'proj2d_step1_30x20'])
def test_projection_matches_reference_golden_any_shape(cfd, name):
  rec = gu.load(name)
  grid = cfd.grids.Grid(rec['shape'], domain=rec['domain'])
  v = wrap(cfd, grid, [rec[f'v0_{i}'] for i in range(rec['ndim'])])
  vp = cfd.pressure.projection(v)
  q = cfd.pressure.solve_fast_diag(v)
  assert gu.rel_l2(np.asarray(q.data), rec['f32_q']) < TOL
  for i, a in enumerate(to_np(vp)):
    assert gu.rel_l2(a, rec[f'f32_proj_{i}']) < TOL


@pytest.mark.parametrize('shape', [(64, 32), (16, 16, 32)])
def test_matmul_and_rfft_implementations_agree(cfd, shape):
  """fast_diagonalization.py:90-125: the same pseudo-inverse through the FP64 tensor-core matmul
  path and through the line FFTs (pressure_test.py runs solve_fast_diag with both)."""
  import functools
  dom = ((0.0, 2 * np.pi),) * len(shape)
  grid = cfd.grids.Grid(shape, domain=dom)
  rs = np.random.RandomState(3)
  v = wrap(cfd, grid, [rs.standard_normal(shape).astype(np.float32) for _ in shape])
  q_fft = np.asarray(cfd.pressure.solve_fast_diag(v, implementation='rfft').data)
  q_alias = np.asarray(cfd.pressure.solve_fast_diag(v, implementation='fft').data)
  q_mm = np.asarray(cfd.pressure.solve_fast_diag(v, implementation='matmul').data)
  np.testing.assert_array_equal(q_fft, q_alias)
  assert gu.rel_l2(q_mm, q_fft) < 1e-6
  want = cfd_oracle.solve_pressure([np.asarray(u.data) for u in v], grid.step)
  assert gu.rel_l2(q_mm, want) < 1e-6
  p_mm = to_np(cfd.pressure.projection(v, functools.partial(cfd.pressure.solve_fast_diag, implementation='matmul')))
  p_fft = to_np(cfd.pressure.projection(v))
  for a, b in zip(p_mm, p_fft):
    assert gu.rel_l2(a, b) < 1e-6
  # a whole step through the matmul solve
  step_mm = cfd.equations.semi_implicit_navier_stokes(
      1.0, 1e-2, 0.01, grid, pressure_solve=functools.partial(cfd.pressure.solve_fast_diag, implementation='matmul'))
  step_fft = cfd.equations.semi_implicit_navier_stokes(1.0, 1e-2, 0.01, grid)
  for a, b in zip(to_np(cfd.funcutils.repeated(step_mm, 3)(v)), to_np(cfd.funcutils.repeated(step_fft, 3)(v))):
    assert gu.rel_l2(a, b) < 1e-6


def _apply_operators(operators, x):
  """fast_diagonalization_test.py:30-41."""
  out = 0
  for axis, m in enumerate(operators):
    out = out + np.moveaxis(np.tensordot(m, x, axes=(1, axis)), 0, axis)
  return out


def test_fast_diagonalization_api_matches_reference_tests(cfd):
  """fast_diagonalization_test.py:47-130 against the device transforms."""
  fd = cfd.fast_diagonalization
  rs = np.random.RandomState(0)
  # test_random_1d_matmul
  a = rs.randn(3, 3)
  a = (a + a.T).astype(np.float32)
  b = rs.randn(3).astype(np.float32)
  got = fd.pseudoinverse([a], b.dtype, hermitian=True, implementation='matmul')(b)
  np.testing.assert_allclose(got, np.linalg.solve(a.astype(np.float64), b), atol=1e-5)
  # test_identity_nd (matmul; 2-D and 3-D)
  for ndim in (1, 2, 3):
    bb = rs.randn(*(2, 4, 6)[:ndim]).astype(np.float32)
    ops = [np.eye(2), 2 * np.eye(4), 3 * np.eye(6)][:ndim]
    got = fd.pseudoinverse(ops, bb.dtype, hermitian=True, circulant=True, implementation='matmul')(bb)
    np.testing.assert_allclose(got, bb / sum(range(1, 1 + ndim)), rtol=1e-5, atol=1e-5)
  # test_poisson_2d_matmul: non-periodic (Dirichlet-like) and periodic operators
  for px, py in ((False, False), (False, True), (True, True)):
    a1 = np.array([[-2, 1, 0, px], [1, -2, 1, 0], [0, 1, -2, 1], [px, 0, 1, -2]], np.float32)
    a2 = np.array([[-2, 1, py], [1, -2, 1], [py, 1, -2]], np.float32)
    bb = np.random.RandomState(0).randn(4, 3).astype(np.float32)
    x = fd.pseudoinverse([a1, a2], bb.dtype, hermitian=True)(bb)
    want = bb - bb.mean() if (px and py) else bb
    np.testing.assert_allclose(_apply_operators([a1, a2], x), want, atol=1e-5)
  # test_poisson_2d_fft at sizes the line-FFT kernels take, both names
  for impl in ('fft', 'rfft'):
    ops = [cfd.array_utils.laplacian_matrix(16, 1.0), cfd.array_utils.laplacian_matrix(32, 0.5)]
    bb = np.random.RandomState(0).randn(16, 32).astype(np.float32)
    x = fd.pseudoinverse(ops, bb.dtype, circulant=True, implementation=impl)(bb)
    np.testing.assert_allclose(_apply_operators(ops, x), bb - bb.mean(), atol=2e-5)
  with pytest.raises(NotImplementedError):
    fd.pseudoinverse([cfd.array_utils.laplacian_matrix(4, 1.0), cfd.array_utils.laplacian_matrix(6, 1.0)],
                     np.float32, circulant=True, implementation='rfft')
  with pytest.raises(ValueError):
    fd.pseudoinverse([np.eye(3)], np.float32, implementation='matmul')  # non-hermitian flag


@pytest.mark.parametrize('shape', [(64, 32), (48, 36), (16, 16, 32)])
def test_implicit_diffusion_matches_oracle(cfd, shape):
  """equations.implicit_diffusion_navier_stokes (equations.py:154-195) with diffusion.solve_fast_diag
  (diffusion.py:166-212), against the NumPy restatement."""
  nd = len(shape)
  dom = ((0.0, 2 * np.pi),) * nd
  grid = cfd.grids.Grid(shape, domain=dom)
  v0 = cfd_oracle.filtered_velocity_field(21, shape, dom, 1.0, 2)
  dt, nu, rho = 0.02, 5e-2, 1.5
  forcing = cfd.forcings.linear_forcing(grid, -0.1)
  step = cfd.equations.implicit_diffusion_navier_stokes(rho, nu, dt, grid, forcing=forcing)
  got = wrap(cfd, grid, v0)
  for _ in range(3):
    got = step(got)
  want = v0
  of = cfd_oracle.Forcing((('linear', -0.1),))
  for _ in range(3):
    want = cfd_oracle.implicit_diffusion_step(want, dt, grid.step, rho, nu, of)
  for a, b in zip(to_np(got), want):
    assert gu.rel_l2(a, b) < TOL


@pytest.mark.parametrize('name', ['imp2d_64x32', 'imp2d_48x36', 'imp3d_16x16x32'])
def test_implicit_diffusion_matches_reference_golden(cfd, name):
  rec = gu.load(name)
  grid = cfd.grids.Grid(rec['shape'], domain=rec['domain'])
  step = cfd.equations.implicit_diffusion_navier_stokes(
      rec['density'], rec['viscosity'], rec['dt'], grid, forcing=make_forcing(cfd, grid, rec))
  v = wrap(cfd, grid, [rec[f'v0_{i}'] for i in range(rec['ndim'])])
  done = 0
  for n in rec['nsteps']:
    for _ in range(n - done):
      v = step(v)
    done = n
    for i, a in enumerate(to_np(v)):
      assert gu.rel_l2(a, rec[f'f32_v{n}_{i}']) < TOL, (name, n, i)
      assert gu.rel_l2(a, rec[f'f64_v{n}_{i}']) < TOL, (name, n, i)


def test_filtered_velocity_field_matches_oracle(cfd):
  """initial_conditions.filtered_velocity_field (initial_conditions.py:71-121) with the spectral
  filter and the normalisation on the device, against the float64 restatement on the same noise."""
  for shape, dom in (((256, 256), ((0.0, 2 * np.pi),) * 2), ((48, 36), ((0.0, 2 * np.pi), (0.0, np.pi))),
                     ((32, 16, 32), ((0.0, 2 * np.pi),) * 3)):
    grid = cfd.grids.Grid(shape, domain=dom)
    v = cfd.initial_conditions.filtered_velocity_field(7, grid, maximum_velocity=2.0, peak_wavenumber=3)
    want = cfd_oracle.filtered_velocity_field(7, shape, dom, 2.0, 3.0)
    for u, w, o in zip(v, want, grid.cell_faces):
      assert u.offset == o and not isinstance(u.data, np.ndarray)  # stays on the device
      assert gu.rel_l2(np.asarray(u.data), w) < 2e-5, shape


def test_unsupported_shapes_and_implementations_raise_at_build_time(cfd):
  grid = cfd.grids.Grid((48, 64), domain=((0, 1), (0, 1)))
  import functools
  with pytest.raises(NotImplementedError):  # rfft requested on a grid the radix-2 kernels do not take
    cfd.equations.semi_implicit_navier_stokes(
        1.0, 1e-3, 0.01, grid, pressure_solve=functools.partial(cfd.pressure.solve_fast_diag, implementation='rfft'))
  with pytest.raises(ValueError):
    cfd.pressure.solve_fast_diag((), implementation='cg')
  with pytest.raises(NotImplementedError):
    cfd.equations.semi_implicit_navier_stokes(1.0, 1e-3, 0.01, grid, pressure_solve=lambda v: v)


@pytest.mark.parametrize('name', ['post2d_32x48', 'post3d_8x12x16'])
def test_downsample_matches_reference_golden(cfd, name):
  """resize.downsample_staggered_velocity on the device ("next" row f4) vs the reference's own output;
  offsets and the coarse grid as in resize.py:204-216."""
  rec = gu.load(name)
  nd = rec['ndim']
  grid = cfd.grids.Grid(rec['shape'], domain=rec['domain'])
  v = wrap(cfd, grid, [rec[f'v0_{i}'] for i in range(nd)])
  for f in rec['factors']:
    f = int(f)
    dst = cfd.grids.Grid(tuple(n // f for n in rec['shape']), domain=rec['domain'])
    out = cfd.resize.downsample_staggered_velocity(grid, dst, v)
    for i, u in enumerate(out):
      assert u.grid == dst and u.offset == grid.cell_faces[i]
      np.testing.assert_allclose(np.asarray(u.data), rec[f'down{f}_{i}'], rtol=0, atol=1e-6)
  # raw arrays with a leading batch axis, host in -> host out
  batch = np.stack([rec['v0_0'], 2 * rec['v0_0']])
  got = cfd.resize.downsample_staggered_velocity_component(batch, 0, 2, ndim=nd)
  assert isinstance(got, np.ndarray) and got.shape == (2,) + tuple(n // 2 for n in rec['shape'])
  np.testing.assert_allclose(got[1], 2 * rec['down2_0'], rtol=0, atol=2e-6)
  with pytest.raises(ValueError):
    cfd.resize.downsample_staggered_velocity_component(rec['v0_0'], 0, 5)  # 48 (or 12, 16) % 5 != 0


def test_trajectory_post_process_runs_on_the_device(cfd):
  """funcutils.trajectory(post_process=...) (funcutils.py:118-121) with the coarse-graining and the
  vorticity evaluated on the device: frames are stacked in device memory and agree with the oracle
  applied to the states of the same trajectory."""
  rec = gu.load('k2d_64x32')
  grid = cfd.grids.Grid(rec['shape'], domain=rec['domain'])
  dst = cfd.grids.Grid((16, 8), domain=rec['domain'])
  step = build_step(cfd, rec, grid)
  v = wrap(cfd, grid, [rec[f'v0_{i}'] for i in range(2)])

  def post(state):
    coarse = cfd.resize.downsample_staggered_velocity(grid, dst, state)
    return coarse + (cfd.resize.vorticity_2d(state),)

  final, frames = cfd.funcutils.trajectory(cfd.funcutils.repeated(step, 2), 3, post_process=post)(v)
  assert isinstance(frames[0].data, cfd.DeviceArray) and frames[0].data.shape == (3, 16, 8)
  assert frames[2].data.shape == (3, 64, 32) and frames[2].offset == (1.0, 1.0)
  # last frame vs the oracle on the final state
  fin = to_np(final)
  for i in range(2):
    want = cfd_oracle.downsample_staggered_velocity_component(fin[i], i, 4)
    np.testing.assert_allclose(np.asarray(frames[i].data[-1]), want, rtol=0, atol=1e-6)
  wz = cfd_oracle.vorticity_2d(fin[0], fin[1], grid.step[0], grid.step[1])
  np.testing.assert_allclose(np.asarray(frames[2].data[-1]), wz, rtol=0, atol=1e-5 * np.abs(wz).max())


def test_readme_example_flow(cfd):
  """The user-facing flow of README.md at a small size: device initial condition, repeated steps in
  one call, a trajectory with a device post-process, diagnostics."""
  grid = cfd.grids.Grid((512, 512), domain=((0, 2 * np.pi), (0, 2 * np.pi)))
  v0 = cfd.initial_conditions.filtered_velocity_field(0, grid, maximum_velocity=7.0, peak_wavenumber=4)
  dt = cfd.equations.stable_time_step(7.0, 0.5, 1e-3, grid)
  forcing = cfd.forcings.sum_forcings(cfd.forcings.kolmogorov_forcing(grid, scale=1.0, k=4),
                                      cfd.forcings.linear_forcing(grid, -0.1))
  step_fn = cfd.equations.semi_implicit_navier_stokes(density=1.0, viscosity=1e-3, dt=dt, grid=grid,
                                                      forcing=forcing)
  v = cfd.funcutils.repeated(step_fn, 20)(v0)
  dst = cfd.grids.Grid((64, 64), domain=grid.domain)
  coarse = lambda s: cfd.resize.downsample_staggered_velocity(grid, dst, s)
  v, frames = cfd.funcutils.trajectory(cfd.funcutils.repeated(step_fn, 5), 4, post_process=coarse)(v)
  assert frames[0].data.shape == (4, 64, 64) and frames[0].grid == dst
  d = cfd.diagnostics(v)
  assert np.isfinite(d['kinetic_energy']) and d['max_abs_div'] < 2e-3
  # the coarse frames of a divergence-free state are divergence free on the coarse grid
  last = [np.asarray(f.data[-1]) for f in frames]
  assert np.abs(cfd_oracle.divergence(last, dst.step)).max() < 2e-3
