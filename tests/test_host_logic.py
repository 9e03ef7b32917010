"""Host-side mirror of the reference interface: grid / offset / boundary / forcing / stepper logic.
CPU only; mirrors the reference's own KATs."""
import os
import sys

import numpy as np
import pytest

import jax_cfd_b200 as cfd
import cfd_oracle
import golden_util as gu


def test_grid_matches_reference_conventions():
  """grids.py:499-675."""
  g = cfd.grids.Grid((4, 8), domain=((0, 1), (0, 4)))
  assert g.step == (0.25, 0.5) and g.ndim == 2
  assert g.cell_center == (0.5, 0.5)
  assert g.cell_faces == ((1.0, 0.5), (0.5, 1.0))
  assert cfd.grids.Grid((4, 4, 4)).cell_faces == ((1.0, .5, .5), (.5, 1.0, .5), (.5, .5, 1.0))
  ax = g.axes((1.0, 0.5))
  assert ax[0].dtype == np.float32
  np.testing.assert_array_equal(ax[0], np.float32([0.25, 0.5, 0.75, 1.0]))
  with pytest.raises(TypeError):
    cfd.grids.Grid((4, 4), step=1.0, domain=((0, 1), (0, 1)))
  with pytest.raises(ValueError):
    cfd.grids.Grid((4, 4), domain=((0, 1),))
  assert cfd.grids.Grid((4, 4), step=0.5).domain == ((0.0, 2.0), (0.0, 2.0))


def test_shift_periodic_kat():
  """boundaries_test.py:150-209: out[i] = in[(i+k) mod N], offset += k."""
  g = cfd.grids.Grid((4,), step=1.0)
  bc = cfd.boundaries.periodic_boundary_conditions(1)
  u = cfd.grids.GridVariable(cfd.grids.GridArray(np.array([11, 12, 13, 14]), (0.5,), g), bc)
  s = u.shift(+1, 0)
  np.testing.assert_array_equal(s.data, [12, 13, 14, 11])
  assert s.offset == (1.5,)
  s = u.shift(-1, 0)
  np.testing.assert_array_equal(s.data, [14, 11, 12, 13])
  assert s.offset == (-0.5,)
  assert u.impose_bc().bc == bc
  assert cfd.boundaries.has_all_periodic_boundary_conditions(u)


def test_control_volume_offsets_and_consistency():
  """grids.py:448-482."""
  g = cfd.grids.Grid((4, 4))
  c = cfd.grids.GridArray(np.zeros((4, 4)), (1.0, 0.5), g)
  assert cfd.grids.control_volume_offsets(c) == ((1.5, 0.5), (1.0, 1.0))
  d = cfd.grids.GridArray(np.zeros((4, 4)), (0.5, 1.0), g)
  with pytest.raises(cfd.grids.InconsistentOffsetError):
    _ = c + d
  e = c + c
  assert e.offset == (1.0, 0.5) and e.grid == g
  with pytest.raises(cfd.grids.InconsistentGridError):
    cfd.grids.consistent_grid(c, cfd.grids.GridArray(np.zeros((4, 4)), (1.0, 0.5), cfd.grids.Grid((4, 4), step=2)))


@pytest.mark.parametrize('name', ['k2d_64x32', 'k2d_32x64_rho', 'tg2d_32', 's3d_16x8x32_kolm'])
def test_forcing_descriptors_bit_exact_vs_reference(name):
  """forcings_test.py:41-90 + bit-exact constant fields (reference arrays in the golden file)."""
  rec = gu.load(name)
  grid = cfd.grids.Grid(rec['shape'], domain=rec['domain'])
  fs = []
  for kind, arg in rec['forcing_spec']:
    if kind == 'kolmogorov':
      fs.append(cfd.forcings.kolmogorov_forcing(grid, **arg))
    elif kind == 'taylor_green':
      fs.append(cfd.forcings.taylor_green_forcing(grid, **arg))
  assert len(fs) == 1
  bc = cfd.boundaries.periodic_boundary_conditions(grid.ndim)
  v = tuple(cfd.grids.GridVariable(cfd.grids.GridArray(rec[f'v0_{i}'], o, grid), bc)
            for i, o in enumerate(grid.cell_faces))
  out = fs[0](v)
  for i, a in enumerate(out):
    assert a.offset == grid.cell_faces[i]          # forcing lives at the velocity offsets
    np.testing.assert_array_equal(a.data, rec[f'f32_constforce_{i}'])


def test_sum_forcings_order_and_linear():
  g = cfd.grids.Grid((8, 8), domain=((0, 2 * np.pi), (0, 2 * np.pi)))
  f = cfd.forcings.simple_turbulence_forcing(g, constant_magnitude=1.0, constant_wavenumber=2,
                                             linear_coefficient=-0.1)
  kinds = [t.kind for t in f.terms]
  assert kinds == [cfd._lib.FORCE_LINEAR, cfd._lib.FORCE_SEPARABLE]   # forcings.py:178: linear first
  with pytest.raises(ValueError):
    cfd.forcings.simple_turbulence_forcing(g, forcing_type='nope')
  bc = cfd.boundaries.periodic_boundary_conditions(2)
  rs = np.random.RandomState(0)
  v = tuple(cfd.grids.GridVariable(cfd.grids.GridArray(rs.standard_normal((8, 8)).astype(np.float32), o, g), bc)
            for o in g.cell_faces)
  out = cfd.forcings.linear_forcing(g, 0.5)(v)
  np.testing.assert_array_equal(out[0].data, np.float32(0.5) * v[0].data)


def test_stable_time_step():
  """equations.py:45-59; advection.py:398-416; diffusion.py:40-57."""
  g = cfd.grids.Grid((256, 256), domain=((0, 2 * np.pi), (0, 2 * np.pi)))
  dt = cfd.equations.stable_time_step(7.0, 0.5, 1e-3, g)
  assert abs(dt - 0.5 * (2 * np.pi / 256) / 7.0) < 1e-15
  with pytest.raises(ValueError):
    cfd.equations.stable_time_step(0.01, 0.5, 1.0, g)
  assert cfd.diffusion.stable_time_step(0.0, g) == float('inf')


def test_rk_tableau_logic_on_host_ode():
  """time_stepping_test.py:27-82 (harmonic oscillator, no projection)."""
  ts = cfd.time_stepping

  class Osc(ts.ExplicitNavierStokesODE):
    def __init__(self):
      super().__init__(lambda s: (s[1], -s[0]), lambda s: s)

  for method, tol in ((ts.forward_euler, 2e-2), (ts.midpoint_rk2, 2e-4), (ts.heun_rk2, 2e-4),
                      (ts.classic_rk4, 1e-8)):
    dt = 1e-2
    step = method(Osc(), dt)
    s = (np.float64(1.0), np.float64(0.0))
    for _ in range(100):
      s = step(s)
    assert abs(s[0] - np.cos(1.0)) < tol and abs(s[1] + np.sin(1.0)) < tol
  with pytest.raises(ValueError):
    ts.ButcherTableau(a=[[1]], b=[1])


def test_unsupported_builder_options_raise_at_build_time():
  g = cfd.grids.Grid((64, 64), domain=((0, 1), (0, 1)))
  with pytest.raises(NotImplementedError):
    cfd.equations.semi_implicit_navier_stokes(1.0, 1e-3, 0.01, g, convect=lambda v: v)
  with pytest.raises(NotImplementedError):
    cfd.equations.semi_implicit_navier_stokes(1.0, 1e-3, 0.01, g, pressure_solve=lambda v: v)
  with pytest.raises(NotImplementedError):
    cfd.equations.semi_implicit_navier_stokes(1.0, 1e-3, 0.01, g, forcing=lambda v: v)


def test_repeated_and_trajectory_semantics():
  """funcutils_test.py:26-50."""
  f = cfd.funcutils.repeated(lambda x: x + 1, 5)
  assert f(0) == 5
  final, traj = cfd.funcutils.trajectory(lambda x: x + 1, 4)(np.float32(0))
  assert final == 4
  np.testing.assert_array_equal(traj, [1, 2, 3, 4])
  final, traj = cfd.funcutils.trajectory(lambda x: x + 1, 4, start_with_input=True)(np.float32(0))
  np.testing.assert_array_equal(traj, [0, 1, 2, 3])


def test_reference_arm_prints_contract_line_under_a_torchrun_like_environment():
  """bench.py --impl reference: rank 0 prints ONE JSON line with the contract keys even when the
  launcher capped OMP_NUM_THREADS (it re-runs itself with a clean environment); other ranks
  print nothing and exit 0."""
  import json
  import subprocess
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  cmd = [sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--gpus', '2',
         '--workload', 'K256', '--steps', '1', '--warmup', '1']
  env = dict(os.environ, OMP_NUM_THREADS='1', RANK='0', WORLD_SIZE='2', LOCAL_RANK='0')
  out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
  assert out.returncode == 0, out.stderr
  lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
  assert len(lines) == 1
  d = json.loads(lines[0])
  assert d['impl'] == 'reference' and d['n_gpus'] == 2 and d['unit'] == 'Gcell*step/s'
  assert d['value'] > 0 and d['cpu_baseline']['cores'] == os.cpu_count()
  assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['value'] == d['value']
  env['RANK'] = '1'
  out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
  assert out.returncode == 0 and not [l for l in out.stdout.splitlines() if l.startswith('{')]


def test_strided_device_views_are_rejected():
  """the kernels read raw base pointers, so a device array that reports non-dense
  strides (a slice / transpose) must be refused instead of being misread."""
  import jax_cfd_b200 as cfd
  from jax_cfd_b200 import _engine, _lib

  class FakeCai:
    def __init__(self, shape, strides):
      self.shape, self.dtype = shape, np.dtype(np.float32)
      self.__cuda_array_interface__ = {'shape': shape, 'typestr': '<f4', 'data': (4096, False),
                                       'version': 3, 'strides': strides}
  assert _lib.is_c_contiguous(FakeCai((8, 16), None))
  assert _lib.is_c_contiguous(FakeCai((8, 16), (64, 4)))
  assert _lib.is_c_contiguous(FakeCai((1, 8, 16), (999, 64, 4)))
  assert not _lib.is_c_contiguous(FakeCai((8, 16), (128, 4)))   # every other row
  assert not _lib.is_c_contiguous(FakeCai((8, 16), (4, 32)))    # transposed
  grid = cfd.grids.Grid((8, 16), domain=((0, 1), (0, 1)))
  bc = cfd.boundaries.periodic_boundary_conditions(2)
  v = tuple(cfd.grids.GridVariable(cfd.grids.GridArray(FakeCai((8, 16), (128, 4)), o, grid), bc)
            for o in grid.cell_faces)
  with pytest.raises(ValueError, match='C-contiguous'):
    _engine.validate_velocity(v, grid)


def test_forward_euler_with_a_different_time_step_keeps_the_convection_dt():
  """navier_stokes_rk's `time_step` and the dt bound into convect by the equation
  builder are separate in the reference (time_stepping.py:59-106 vs equations.py:127-128)."""
  import jax_cfd_b200 as cfd
  grid = cfd.grids.Grid((32, 32), domain=((0, 1), (0, 1)))
  step = cfd.equations.semi_implicit_navier_stokes(
      1.0, 1e-3, 0.01, grid, time_stepper=lambda ode, dt: cfd.time_stepping.forward_euler(ode, dt / 2))
  assert step.dt == 0.005 and step.convect_dt == 0.01
  same = cfd.equations.semi_implicit_navier_stokes(1.0, 1e-3, 0.01, grid)
  assert same.dt == 0.01 and same.convect_dt is None


def test_jax_ffi_attributes_describe_the_equation_without_pointers():
  """The attribute dictionary of the XLA custom call (jax_ffi.step_attrs) carries exactly what struct
  cfd_params carries, as typed scalars / small arrays (csrc/xla_ffi_shim.cc, CFD_STEP_ATTRS) --
  serialisable, no host pointers -- and the forcing tables travel as operands."""
  import jax_cfd_b200 as cfd
  from jax_cfd_b200 import jax_ffi, _lib
  grid = cfd.grids.Grid((64, 32), domain=((0.0, 2 * np.pi), (0.0, 2 * np.pi)))
  forcing = cfd.forcings.sum_forcings(cfd.forcings.kolmogorov_forcing(grid, scale=2.0, k=4),
                                      cfd.forcings.linear_forcing(grid, -0.1))
  a = jax_ffi.step_attrs(grid, 0.01, 1.5, None, forcing, nsteps=7)
  expected = {'step', 'nsteps', 'implementation', 'dt', 'convect_dt', 'density', 'viscosity', 'linear_coef',
              'smagorinsky_cs', 'terms', 'sep_scale', 'sep_mask', 'sep_has', 'field_mask'}
  assert set(a) == expected
  assert all(isinstance(v, (np.generic, np.ndarray)) for v in a.values())
  assert a['viscosity'] < 0 and a['nsteps'] == 7 and a['density'] == 1.5
  assert list(a['terms']) == [_lib.FORCE_SEPARABLE, _lib.FORCE_LINEAR]
  assert a['linear_coef'] == -0.1 and a['sep_scale'][0] == 2.0
  assert a['sep_has'] == 1          # only u is forced (forcings.py:88-99)
  sep, fld = jax_ffi.forcing_operands(grid, forcing)
  assert sep.shape == (bin(int(a['sep_mask'])).count('1'), 64) and sep.dtype == np.float32
  assert fld.shape == (0, 64, 32)
  # the profile rows are the term's own tables, in (component, axis) order of the set bits
  term = forcing.terms[0]
  rows = [np.asarray(term.profiles[c][j], np.float32) for c in range(2) for j in range(2)
          if term.profiles[c][j] is not None]
  for r, want in zip(sep, rows):
    np.testing.assert_array_equal(r[:want.size], want)
  # 3-D with the Smagorinsky closure: one term, one scalar
  g3 = cfd.grids.Grid((32, 32, 64), domain=((0.0, 2 * np.pi),) * 3)
  f3 = cfd._engine.ForcingFn([cfd._engine.SmagorinskyTerm(0.2)])
  a3 = jax_ffi.step_attrs(g3, 0.01, 1.0, 1e-3, f3)
  assert list(a3['terms']) == [_lib.FORCE_SMAGORINSKY] and a3['smagorinsky_cs'] == 0.2
  assert a3['step'].shape == (3,) and a3['viscosity'] == 1e-3
  assert set(jax_ffi.TARGETS.values()) == {'B200CfdStep2D', 'B200CfdStep3D', 'B200CfdProject2D',
                                           'B200CfdProject3D'}
