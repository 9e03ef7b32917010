"""The oracle restatement (oracle/cfd_oracle.py) against golden vectors produced by the
reference's own code (oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest

import cfd_oracle
import golden_util as gu


@pytest.mark.parametrize('name', gu.step_cases())
@pytest.mark.parametrize('prec', ['f32', 'f64'])
def test_step_matches_reference(name, prec):
  rec = gu.load(name)
  dtype = np.float32 if prec == 'f32' else np.float64
  # f32: same arithmetic as the reference up to reassociation -> a few ulp; f64: 1e-12 class,
  # except that `diagonals` are complex64 in both (fast_diagonalization.py:214) -> ~1e-7.
  tol = 2e-6 if prec == 'f32' else 1e-6
  d = rec['ndim']
  v = tuple(rec[f'v0_{i}'].astype(dtype) for i in range(d))
  forcing = gu.oracle_forcing(rec, dtype)
  a, b = gu.TABLEAUS[rec['stepper']]
  diag = cfd_oracle.pinv_diagonals(rec['shape'], rec['h'], np.float32)
  done = 0
  for n in rec['nsteps']:
    for _ in range(n - done):
      if rec['stepper'] == 'forward_euler':
        out = cfd_oracle.step(v, rec['dt'], rec['h'], rec['density'], rec['viscosity'], forcing,
                              diag=diag, return_q=True, return_ustar=True)
        if done == 0:
          vnew, q, ustar = out
          for i in range(d):
            assert gu.rel_l2(ustar[i], rec[f'{prec}_ustar_{i}']) < tol
          want_q = rec[f'{prec}_q']
          if rec['shape'][-1] % 2:
            # odd last axis: the reference falls back to the matmul transform
            # (fast_diagonalization.py:101-108), whose eigh-computed zero eigenvalue is ~1e-13, not 0,
            # so in float64 the mean mode of q is amplified rounding noise: compare q up to a constant
            q, want_q = q - q.mean(), want_q - want_q.mean()
          assert gu.rel_l2(q, want_q) < 10 * tol
        v = out[0]
      else:
        v = cfd_oracle.rk_step(v, rec['dt'], rec['h'], a, b, rec['density'], rec['viscosity'],
                               forcing, diag=diag)
      done += 1
    for i in range(d):
      assert v[i].dtype == dtype
      err = gu.rel_l2(v[i], rec[f'{prec}_v{n}_{i}'])
      assert err < tol * max(1, n), (name, n, i, err)


@pytest.mark.parametrize('name', ['k2d_64x32', 'k2d_32x64_rho', 'tg2d_32', 's3d_16x8x32_kolm'])
def test_constant_forcing_field_bit_exact(name):
  """Forcing and grid-indexing logic reproduced bit-exactly (forcings.py:63-104, 35-60)."""
  rec = gu.load(name)
  f = [t for t in gu.oracle_forcing(rec).terms if t[0] == 'const']
  assert len(f) == 1
  for i, a in enumerate(f[0][1]):
    ref = rec[f'f32_constforce_{i}']
    assert a.dtype == np.float32 and ref.dtype == np.float32
    np.testing.assert_array_equal(a, ref)


@pytest.mark.parametrize('name', ['proj2d_64x32', 'proj3d_16x8x32', 'proj2d_step1_30x20'])
def test_projection_matches_reference(name):
  rec = gu.load(name)
  d = rec['ndim']
  for prec, dtype in (('f32', np.float32), ('f64', np.float64)):
    v = tuple(rec[f'v0_{i}'].astype(dtype) for i in range(d))
    vp, q = cfd_oracle.projection(v, rec['h'])
    assert gu.rel_l2(q, rec[f'{prec}_q']) < 2e-6
    for i in range(d):
      assert gu.rel_l2(vp[i], rec[f'{prec}_proj_{i}']) < 2e-6
    div = cfd_oracle.divergence(vp, rec['h'])
    assert np.abs(div).max() < 1e-4 * max(1.0, np.abs(np.asarray(v)).max() / min(rec['h']))


def test_shift_kat():
  """boundaries_test.py:170-209."""
  a = np.array([11, 12, 13, 14])
  np.testing.assert_array_equal(cfd_oracle.shift(a, +1, 0), [12, 13, 14, 11])
  np.testing.assert_array_equal(cfd_oracle.shift(a, -1, 0), [14, 11, 12, 13])
  np.testing.assert_array_equal(cfd_oracle.shift(a, +2, 0), [13, 14, 11, 12])


def test_cell_faces_kat():
  """grids.py:567-572."""
  assert cfd_oracle.cell_faces(2) == ((1.0, 0.5), (0.5, 1.0))
  assert cfd_oracle.cell_faces(3) == ((1.0, 0.5, 0.5), (0.5, 1.0, 0.5), (0.5, 0.5, 1.0))


def test_pinv_poisson_kat():
  """fast_diagonalization_test.py:58-67: A x == b - mean(b)."""
  rs = np.random.RandomState(0)
  shape, h = (16, 12), (0.5, 0.25)
  b = rs.standard_normal(shape)
  diag = cfd_oracle.pinv_diagonals(shape, h, np.float64)
  x = np.fft.irfftn(diag * np.fft.rfftn(b), s=shape, axes=(0, 1))
  ax = cfd_oracle.laplacian(x, h)
  np.testing.assert_allclose(ax, b - b.mean(), atol=1e-5)


@pytest.mark.parametrize('name', gu.implicit_cases())
@pytest.mark.parametrize('prec', ['f32', 'f64'])
def test_implicit_diffusion_step_matches_reference(name, prec):
  """equations.implicit_diffusion_navier_stokes (equations.py:154-195) run by the reference itself
  vs cfd_oracle.implicit_diffusion_step."""
  rec = gu.load(name)
  dtype = np.float32 if prec == 'f32' else np.float64
  tol = 2e-6 if prec == 'f32' else 1e-6
  v = tuple(rec[f'v0_{i}'].astype(dtype) for i in range(rec['ndim']))
  forcing = gu.oracle_forcing(rec, dtype)
  diag = cfd_oracle.pinv_diagonals(rec['shape'], rec['h'], np.float32)
  ddiag = cfd_oracle.diffusion_diagonals(rec['shape'], rec['h'], rec['viscosity'], rec['dt'], np.float32)
  done = 0
  for n in rec['nsteps']:
    for _ in range(n - done):
      v = cfd_oracle.implicit_diffusion_step(v, rec['dt'], rec['h'], rec['density'], rec['viscosity'], forcing,
                                             diag=diag, ddiag=ddiag)
    done = n
    for i in range(rec['ndim']):
      assert v[i].dtype == dtype
      assert gu.rel_l2(v[i], rec[f'{prec}_v{n}_{i}']) < tol * max(1, n), (name, n, i)


@pytest.mark.parametrize('name', ['post2d_32x48', 'post3d_8x12x16'])
def test_downsample_matches_reference(name):
  """resize.downsample_staggered_velocity (resize.py:187-222) run by the reference itself vs the
  oracle's restatement (mean over the block in a possibly different order: 1e-6)."""
  rec = gu.load(name)
  for f in rec['factors']:
    for i in range(rec['ndim']):
      got = cfd_oracle.downsample_staggered_velocity_component(rec[f'v0_{i}'], i, int(f))
      want = rec[f'down{int(f)}_{i}']
      assert got.shape == want.shape and got.dtype == np.float32
      np.testing.assert_allclose(got, want, rtol=0, atol=1e-6)


def test_downsample_kat_of_the_reference_test():
  """resize_test.py:41-56: the reference's own known-answer vectors."""
  u = np.arange(16.0).reshape(4, 4)
  np.testing.assert_array_equal(cfd_oracle.downsample_staggered_velocity_component(u, 0, 2),
                                [[4.5, 6.5], [12.5, 14.5]])
  np.testing.assert_array_equal(cfd_oracle.downsample_staggered_velocity_component(u, 1, 2),
                                [[3.0, 5.0], [11.0, 13.0]])
  with pytest.raises(ValueError):  # array_utils.py:157-160
    cfd_oracle.downsample_staggered_velocity_component(np.zeros((4, 6)), 0, 4)
  # along `direction` itself the strided slice simply drops the remainder (resize.py:72)
  assert cfd_oracle.downsample_staggered_velocity_component(np.zeros((6, 4)), 0, 4).shape == (1, 1)


def test_downsampling_keeps_a_divergence_free_field_divergence_free():
  """The property the procedure exists for (resize.py:49-51)."""
  rec = gu.load('proj2d_64x32')
  v = [rec['f32_proj_0'].astype(np.float64), rec['f32_proj_1'].astype(np.float64)]
  h = rec['h']
  assert np.abs(cfd_oracle.divergence(v, h)).max() < 1e-4
  f = 4
  vc = [cfd_oracle.downsample_staggered_velocity_component(a, i, f) for i, a in enumerate(v)]
  assert np.abs(cfd_oracle.divergence(vc, tuple(f * x for x in h))).max() < 1e-4
