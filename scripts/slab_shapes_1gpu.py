"""Times the per-GPU kernels of the weak-scaling shapes on ONE GPU (world=1 slab path)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import jax_cfd_b200 as cfd
from jax_cfd_b200 import _lib
for shape in ((8192, 8192), (4096, 16384), (2048, 32768)):
  grid = cfd.grids.Grid(shape, domain=((0, 2 * np.pi * shape[0] / 8192), (0, 2 * np.pi * shape[1] / 8192)))
  dt = cfd.equations.stable_time_step(7.0, 0.5, 1e-4, grid)
  f = cfd.forcings.sum_forcings(cfd.forcings.kolmogorov_forcing(grid, k=4), cfd.forcings.linear_forcing(grid, -0.1))
  st = cfd.distributed.SlabStepper(grid, dt, 1.0, 1e-4, f, rank=0, world=1, device=0, exchange=lambda b: [b])
  u, v = bench.analytic_ic(shape, (0, shape[0]), 7.0)
  st.load([u, v]); st.advance(3); st.sync()
  e0, e1 = _lib.Event(), _lib.Event()
  e0.record(st.stream.handle); st.advance(10); e1.record(st.stream.handle); st.sync()
  ms = e0.elapsed_ms(e1) / 10
  o = st.store()
  print(shape, 'ms/step', round(ms, 3), 'Gcell/s', round(shape[0] * shape[1] / ms / 1e6, 1), 'finite',
        bool(np.isfinite(o[0].numpy()).all()), flush=True)
  st.close()
