"""Parity at the REAL sizes of the BASELINE.json configs (driver-run, `-m gpu`).

  K8192   8192^2 Kolmogorov step            vs oracle/cpu_baseline.CpuStep (golden-pinned C port)
  E1024   1024 x 256^2 ensemble, 10 steps   vs cfd_oracle.step on sampled members + bitwise vs
                                             single-member runs
  TGV     128^3 / 256^3 Taylor-Green + Smagorinsky vs the NumPy oracle; 512^3 through
          size-independent invariants; thin 3-D shapes with 512-point lines on every axis
  slab    world-size-2 (and 4 / 8 when present) bitwise check, spawned under torchrun

Bar: per-step relative L2 <= 1e-5 on every velocity component and on q.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import cfd_oracle
import golden_util as gu

pytestmark = pytest.mark.gpu

TOL = 1e-5
TWO_PI = 2 * np.pi
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def cfd():
  import jax_cfd_b200 as m
  m._lib.require_device()
  return m


def wrap(cfd, grid, arrays):
  bc = cfd.boundaries.periodic_boundary_conditions(grid.ndim)
  return tuple(cfd.grids.GridVariable(
      cfd.grids.GridArray(cfd.DeviceArray.from_numpy(np.ascontiguousarray(a, np.float32)), o, grid), bc)
               for a, o in zip(arrays, grid.cell_faces))


def to_np(v):
  return [np.asarray(u.data) for u in v]


def multiscale_field(shape, seed, vmax):
  """Cheap at any size: a few smooth modes on the staggered grid plus grid-scale noise, so that
  both branches of the upwind selection and the limiter are exercised everywhere."""
  nx, ny = shape
  rs = np.random.RandomState(seed)
  x = (np.arange(nx, dtype=np.float64) + 1.0) * (TWO_PI / nx)
  xc = (np.arange(nx, dtype=np.float64) + 0.5) * (TWO_PI / nx)
  y = (np.arange(ny, dtype=np.float64) + 1.0) * (TWO_PI / ny)
  yc = (np.arange(ny, dtype=np.float64) + 0.5) * (TWO_PI / ny)
  u = np.zeros(shape, np.float32)
  v = np.zeros(shape, np.float32)
  for a, b, amp in ((3, 4, 1.0), (17, 9, 0.5), (130, 77, 0.25), (1021, 640, 0.1)):
    u += (amp * np.sin(a * x)[:, None] * np.cos(b * yc)[None, :]).astype(np.float32)
    v -= (amp * np.cos(a * xc)[:, None] * np.sin(b * y)[None, :]).astype(np.float32)
  u += 0.05 * rs.standard_normal(shape).astype(np.float32)
  v += 0.05 * rs.standard_normal(shape).astype(np.float32)
  s = np.float32(vmax / 2.0)
  return u * s, v * s


def test_k8192_step_matches_cpu_oracle(cfd):
  """BASELINE headline size: 8192^2 steps with the K8192 physics (bench.py) against the OpenMP-C +
  pocketfft port of the oracle (tests/test_oracle_c.py pins it to the golden vectors).

  The initial condition is made divergence free first, like every initial condition of the
  reference (initial_conditions.py:112-121) and like bench.py's.  (On a field with O(1) divergence
  a float32 solve at h = 7.7e-4 leaves eps * |q| / h ~ 1e-5 of rounding noise in the velocity --
  the reference's own float32 run is that far from its float64 run; measured 2.7e-6 at 2048^2.)
  `q` of a divergence-free flow is dt * pressure ~ 1e-4 and is itself dominated by the rounding of
  div(u*): it is held to 3x the oracle's own sensitivity to a half-ulp perturbation of the input."""
  import cpu_baseline
  shape = (8192, 8192)
  dom = ((0.0, TWO_PI), (0.0, TWO_PI))
  grid = cfd.grids.Grid(shape, domain=dom)
  nu, vmax = 1e-4, 7.0
  dt = cfd.equations.stable_time_step(vmax, 0.5, nu, grid)
  u0, v0 = to_np(cfd.pressure.projection(wrap(cfd, grid, multiscale_field(shape, 0, vmax))))
  forcing = cfd.forcings.sum_forcings(cfd.forcings.kolmogorov_forcing(grid, scale=1.0, k=4),
                                      cfd.forcings.linear_forcing(grid, -0.1))
  step = cfd.equations.semi_implicit_navier_stokes(1.0, nu, dt, grid, forcing=forcing)
  got, q = step.advance(wrap(cfd, grid, [u0, v0]), 1, return_q=True)
  got = to_np(got)
  q = np.asarray(q)
  cs = cpu_baseline.CpuStep(shape, grid.step, dt, 1.0, nu,
                            cfd_oracle.kolmogorov_field(shape, dom, 1.0, 4), -0.1)
  wu, wv = (a.copy() for a in cs.step(u0, v0))
  wq = cs.q.copy()
  eu, ev, eq = gu.rel_l2(got[0], wu), gu.rel_l2(got[1], wv), gu.rel_l2(q, wq)
  # conditioning of q: the oracle's q for an input perturbed by half an ulp
  rs = np.random.RandomState(1)
  pert = [(a * (1 + np.float32(6e-8) * rs.choice(np.float32([-1, 1]), size=shape))).astype(np.float32)
          for a in (u0, v0)]
  cs.step(pert[0], pert[1])
  floor_q = gu.rel_l2(cs.q, wq)
  del pert
  print(f'\nK8192 one step: rel-L2 u {eu:.2e} v {ev:.2e} q {eq:.2e} (q floor {floor_q:.2e})')
  assert eu < TOL and ev < TOL
  assert eq < max(TOL, 3 * floor_q)
  # the chained (lazy-projection) path must agree with single steps at this size too
  two = to_np(cfd.funcutils.repeated(step, 2)(wrap(cfd, grid, [u0, v0])))
  wu2, wv2 = cs.step(wu, wv)
  assert gu.rel_l2(two[0], wu2) < TOL
  assert gu.rel_l2(two[1], wv2) < TOL
  # divergence-free residual: rounding noise / h (h = 7.7e-4), held to the oracle's own residual
  ref_div = np.abs(cfd_oracle.divergence([wu2, wv2], grid.step)).max()
  assert np.abs(cfd_oracle.divergence(two, grid.step)).max() < max(2e-3, 3 * ref_div)


def batch_fields(n, shape, seed, vmax, kpeak=4.0):
  """n distinct filtered-noise members (log-normal spectrum), one batched real FFT."""
  import scipy.fft
  nx, ny = shape
  rs = np.random.RandomState(seed)
  kx = TWO_PI * np.fft.fftfreq(nx, TWO_PI / nx)
  ky = TWO_PI * np.fft.rfftfreq(ny, TWO_PI / ny)
  k = np.sqrt(kx[:, None] ** 2 + ky[None, :] ** 2)
  with np.errstate(divide='ignore', invalid='ignore'):
    logk = np.log(k)
    filt = np.exp(-(np.log(kpeak) + 0.25 - logk) ** 2 / 0.5 - logk) / k
  filt[0, 0] = 0.0
  out = []
  for _ in range(2):
    noise = rs.standard_normal((n,) + tuple(shape)).astype(np.float32)
    spec = scipy.fft.rfft2(noise, workers=-1)
    spec *= filt.astype(np.float32)
    f = scipy.fft.irfft2(spec, s=shape, workers=-1).astype(np.float32)
    f *= (vmax / np.abs(f).reshape(n, -1).max(axis=1)).astype(np.float32)[:, None, None]
    out.append(f)
  return out


def test_e1024_ensemble_members_match_oracle_and_single_runs(cfd):
  """BASELINE config #3: 1024 Kolmogorov 256^2 members, 10 steps in one batched call."""
  shape, nb, nsteps = (256, 256), 1024, 10
  dom = ((0.0, TWO_PI), (0.0, TWO_PI))
  grid = cfd.grids.Grid(shape, domain=dom)
  nu, vmax = 1e-3, 7.0
  dt = cfd.equations.stable_time_step(vmax, 0.5, nu, grid)
  u0, v0 = batch_fields(nb, shape, 5, 3.0)
  forcing = cfd.forcings.sum_forcings(cfd.forcings.kolmogorov_forcing(grid, scale=1.0, k=4),
                                      cfd.forcings.linear_forcing(grid, -0.1))
  step = cfd.equations.semi_implicit_navier_stokes(1.0, nu, dt, grid, forcing=forcing)
  out = to_np(cfd.funcutils.repeated(step, nsteps)(wrap(cfd, grid, [u0, v0])))
  assert out[0].shape == (nb,) + shape
  of = cfd_oracle.Forcing((('const', cfd_oracle.kolmogorov_field(shape, dom, 1.0, 4)), ('linear', -0.1)))
  diag = cfd_oracle.pinv_diagonals(shape, grid.step, np.float32)
  members = [0, 1, 511, 512, 777, 1023] + list(np.random.RandomState(1).randint(0, nb, 2))
  for m in members:
    w = (u0[m], v0[m])
    for _ in range(nsteps):
      w = cfd_oracle.step(w, dt, grid.step, 1.0, nu, of, diag=diag)
    for c in range(2):
      assert gu.rel_l2(out[c][m], w[c]) < TOL, (m, c)
    single = to_np(cfd.funcutils.repeated(step, nsteps)(wrap(cfd, grid, [u0[m], v0[m]])))
    for c in range(2):
      np.testing.assert_array_equal(out[c][m], single[c])


def taylor_green_3d(cfd, grid):
  """u = sin x cos y cos z, v = -cos x sin y cos z, w = 0 at grid.cell_faces."""
  ax = [grid.axes(o) for o in grid.cell_faces]
  f32 = np.float32
  u = (np.sin(ax[0][0])[:, None, None] * np.cos(ax[0][1])[None, :, None] * np.cos(ax[0][2])[None, None, :]).astype(f32)
  v = (-np.cos(ax[1][0])[:, None, None] * np.sin(ax[1][1])[None, :, None] * np.cos(ax[1][2])[None, None, :]).astype(f32)
  return [u, v, np.zeros(grid.shape, f32)]


@pytest.mark.parametrize('n,nsteps', [(128, 2), (256, 1)])
def test_tgv_smagorinsky_matches_oracle(cfd, n, nsteps):
  """BASELINE config #5 physics (Taylor-Green vortex, nu = 1/1600, Smagorinsky cs = 0.2) at 128^3
  (two steps) and 256^3 (one step: the NumPy oracle needs ~30 s per 256^3 step)."""
  shape = (n, n, n)
  grid = cfd.grids.Grid(shape, domain=((0.0, TWO_PI),) * 3)
  v0 = taylor_green_3d(cfd, grid)
  rs = np.random.RandomState(2)
  v0 = [a + 0.02 * rs.standard_normal(shape).astype(np.float32) for a in v0]
  nu = 1.0 / 1600
  dt = cfd.equations.stable_time_step(1.0, 0.5, nu, grid)
  step = cfd.subgrid_models.explicit_smagorinsky_navier_stokes(
      dt=dt, cs=0.2, forcing=None, density=1.0, viscosity=nu, grid=grid)
  got, q = step.advance(wrap(cfd, grid, v0), nsteps, return_q=True)
  of = cfd_oracle.Forcing((('smagorinsky', 0.2),))
  want = tuple(v0)
  for _ in range(nsteps):
    want, wq = cfd_oracle.step(want, dt, grid.step, 1.0, nu, of, return_q=True)
  for a, b in zip(to_np(got), want):
    assert gu.rel_l2(a, b) < TOL
  assert gu.rel_l2(np.asarray(q), wq) < TOL


@pytest.mark.parametrize('shape', [(512, 16, 32), (16, 512, 32), (16, 16, 512)])
def test_3d_512_point_lines_match_oracle(cfd, shape):
  """Every 3-D FFT sweep at the 512-point line length of TGV512, on a thin grid the oracle
  finishes in a second."""
  dom = ((0.0, TWO_PI),) * 3
  grid = cfd.grids.Grid(shape, domain=dom)
  v0 = cfd_oracle.filtered_velocity_field(9, shape, dom, 1.0, 2)
  dt = 0.25 * min(grid.step)
  nu = 1.0 / 1600
  step = cfd.subgrid_models.explicit_smagorinsky_navier_stokes(
      dt=dt, cs=0.2, forcing=None, density=1.0, viscosity=nu, grid=grid)
  got, q = step.advance(wrap(cfd, grid, v0), 2, return_q=True)
  of = cfd_oracle.Forcing((('smagorinsky', 0.2),))
  want = v0
  for _ in range(2):
    want, wq = cfd_oracle.step(want, dt, grid.step, 1.0, nu, of, return_q=True)
  for a, b in zip(to_np(got), want):
    assert gu.rel_l2(a, b) < TOL
  w64 = tuple(a.astype(np.float64) for a in v0)
  for _ in range(2):
    w64, q64 = cfd_oracle.step(w64, dt, grid.step, 1.0, nu, of, return_q=True)
  floor = gu.rel_l2(wq, q64)  # 32:1 cells: the reference's own f32-vs-f64 distance on q
  assert gu.rel_l2(np.asarray(q), q64) < max(TOL, 3 * floor)


def test_tgv512_invariants(cfd):
  """BASELINE config #5 at full size (512^3, Smagorinsky): divergence-free residual, momentum
  conservation and monotone kinetic-energy decay over 20 steps (equations_test.py:84-163 checks
  the same invariants; the oracle cannot run this size in test time)."""
  shape = (512, 512, 512)
  grid = cfd.grids.Grid(shape, domain=((0.0, TWO_PI),) * 3)
  v0 = taylor_green_3d(cfd, grid)
  nu = 1.0 / 1600
  dt = cfd.equations.stable_time_step(1.0, 0.5, nu, grid)
  step = cfd.subgrid_models.explicit_smagorinsky_navier_stokes(
      dt=dt, cs=0.2, forcing=None, density=1.0, viscosity=nu, grid=grid)
  v = wrap(cfd, grid, v0)
  mean0 = [float(a.mean(dtype=np.float64)) for a in v0]
  del v0
  ke = [cfd.diagnostics(v)['kinetic_energy']]
  assert abs(ke[0] - 0.125) < 1e-4  # <0.5 (u^2 + v^2)> of the Taylor-Green vortex
  for _ in range(4):
    v = cfd.funcutils.repeated(step, 5)(v)
    d = cfd.diagnostics(v)
    ke.append(d['kinetic_energy'])
    assert d['max_abs_div'] < 2e-3  # equations_test.py:99,127
  assert all(b < a for a, b in zip(ke, ke[1:])), ke
  assert ke[-1] > 0.99 * ke[0]  # 20 steps of dt = 6.1e-3: about 0.03 % viscous decay
  for u, m0 in zip(v, mean0):
    assert abs(float(np.asarray(u.data).mean(dtype=np.float64)) - m0) < 1e-5


def _device_count():
  try:
    import jax_cfd_b200 as m
    return m._lib.lib().cfd_device_count()
  except Exception:
    return 0


@pytest.mark.parametrize('world,shape,nsteps,env', [
    (2, (256, 512), 6, {}), (2, (2048, 1024), 4, {}), (4, (1024, 512), 4, {}), (8, (2048, 1024), 4, {}),
    # 32768-point x lines: the cluster / DSMEM kernel writing to the peers, plain and pair-interleaved layout
    (2, (32768, 64), 3, {}), (2, (32768, 64), 3, {'CFD_T_PAIRED': '1'}), (8, (32768, 256), 2, {'CFD_T_PAIRED': '1'}),
    (2, (2048, 1024), 4, {'CFD_DIST_MODE': 'pull'}),
    # 3-D with the Smagorinsky closure (config #5)
    (2, (64, 32, 64), 3, {}), (8, (128, 128, 64), 2, {})])
def test_slab_decomposition_is_bitwise_equal_to_one_gpu(world, shape, nsteps, env):
  """Row (e): the slab-decomposed step on `world` GPUs reproduces the single-GPU result bit for
  bit (tests/mgpu_worker.py under torchrun, one rank per GPU).  Skipped on boxes with fewer GPUs."""
  if _device_count() < world:
    pytest.skip(f'needs {world} GPUs')
  port = 29600 + world
  cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}',
         '--master-addr', '127.0.0.1', '--master-port', str(port),
         os.path.join(ROOT, 'tests', 'mgpu_worker.py'), *[str(n) for n in shape], str(nsteps)]
  out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env={**os.environ, **env})
  assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
  assert 'MGPU PASS' in out.stdout
  assert out.stdout.count('bitwise=True') == len(shape) + 1, out.stdout
