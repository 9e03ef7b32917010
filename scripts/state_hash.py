"""Prints a SHA-256 of the state after a few chained steps of a seeded 2-D case: run it with two
builds of the library (CFD_B200_LIB=...) to check that a kernel change is bit-neutral."""
import hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import jax_cfd_b200 as cfd

shape = (512, 1024)
grid = cfd.grids.Grid(shape, domain=((0, 2 * np.pi), (0, 4 * np.pi)))
rs = np.random.RandomState(0)
v0 = [rs.standard_normal(shape).astype(np.float32) for _ in range(2)]
bc = cfd.boundaries.periodic_boundary_conditions(2)
forcing = cfd.forcings.sum_forcings(cfd.forcings.kolmogorov_forcing(grid, k=4), cfd.forcings.linear_forcing(grid, -0.1))
step = cfd.equations.semi_implicit_navier_stokes(1.0, 1e-3, 2e-3, grid, forcing=forcing)
v = tuple(cfd.grids.GridVariable(cfd.grids.GridArray(cfd.DeviceArray.from_numpy(a), o, grid), bc)
          for a, o in zip(v0, grid.cell_faces))
v = cfd.pressure.projection(v)
out = step.advance(v, 7)
h = hashlib.sha256()
for a in out:
  h.update(np.ascontiguousarray(np.asarray(a.data)).tobytes())
print(os.environ.get('CFD_B200_LIB', 'default').split('/')[-1], h.hexdigest()[:16], float(np.abs(np.asarray(out[0].data)).max()))

# 3-D case with the Smagorinsky closure (marching kernels)
shape3 = (64, 32, 64)
g3 = cfd.grids.Grid(shape3, domain=((0, 2 * np.pi),) * 3)
v3 = [rs.standard_normal(shape3).astype(np.float32) for _ in range(3)]
bc3 = cfd.boundaries.periodic_boundary_conditions(3)
step3 = cfd.subgrid_models.explicit_smagorinsky_navier_stokes(dt=2e-3, cs=0.2, forcing=None, density=1.0, viscosity=1e-3, grid=g3)
w = tuple(cfd.grids.GridVariable(cfd.grids.GridArray(cfd.DeviceArray.from_numpy(a), o, g3), bc3)
          for a, o in zip(v3, g3.cell_faces))
w = cfd.pressure.projection(w)
for _ in range(3):
  w = step3(w)
h = hashlib.sha256()
for a in w:
  h.update(np.ascontiguousarray(np.asarray(a.data)).tobytes())
print('3d', h.hexdigest()[:16], float(np.abs(np.asarray(w[0].data)).max()))
