"""Multi-GPU worker (run under torchrun, one rank per GPU): the slab-decomposed path must give the
same result as the single-GPU path on the same global field.  Prints PASS/FAIL on rank 0.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 tests/mgpu_worker.py 256 512 6

Four numbers (N0 N1 N2 steps) run the 3-D case: Taylor-Green-like field with the Smagorinsky closure
(BASELINE config #5 in miniature).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
  sys.path.insert(0, p)


def main():
  import torch
  import torch.distributed as dist
  import jax_cfd_b200 as cfd
  import cfd_oracle
  rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
  local_rank = int(os.environ.get('LOCAL_RANK', rank))
  torch.cuda.set_device(local_rank)
  dist.init_process_group('gloo')
  nums = [int(a) for a in sys.argv[1:]]
  shape, nsteps = tuple(nums[:-1]), nums[-1]
  nd = len(shape)
  dom = ((0.0, 2 * np.pi),) * nd
  grid = cfd.grids.Grid(shape, domain=dom)
  if nd == 2:
    v0 = cfd_oracle.filtered_velocity_field(3, shape, dom, 3.0, 3)   # same on every rank
    dt = 0.5 * min(grid.step) / 3.0
    nu = 1e-3
    forcing = cfd.forcings.sum_forcings(cfd.forcings.kolmogorov_forcing(grid, k=4),
                                        cfd.forcings.linear_forcing(grid, -0.1))
    make_step = lambda: cfd.equations.semi_implicit_navier_stokes(1.0, nu, dt, grid, forcing=forcing)
  else:
    v0 = cfd_oracle.filtered_velocity_field(7, shape, dom, 1.0, 2)
    dt = 0.5 * min(grid.step) / 1.0
    nu = 1.0 / 1600
    lin = cfd.forcings.linear_forcing(grid, 0.05)
    forcing = cfd.forcings.sum_forcings(lin, cfd._engine.ForcingFn([cfd._engine.SmagorinskyTerm(0.2)]))
    make_step = lambda: cfd.subgrid_models.explicit_smagorinsky_navier_stokes(
        dt=dt, cs=0.2, forcing=lin, density=1.0, viscosity=nu, grid=grid)
  st = cfd.distributed.SlabStepper(grid, dt, 1.0, nu, forcing, rank=rank, world=world, device=local_rank)
  r0, r1 = st.rows
  st.load([a[r0:r1] for a in v0])
  st.advance(nsteps // 2)
  st.advance(nsteps - nsteps // 2)          # a second call continues the lazy chain
  outs, q = st.store(want_q=True)
  loc = [o.numpy() for o in outs] + [q.numpy()]
  gathered = [None] * world
  dist.all_gather_object(gathered, loc)
  ok = True
  if rank == 0:
    full = [np.concatenate([g[i] for g in gathered], axis=0) for i in range(nd + 1)]
    # single-GPU run of the same global problem on this rank's GPU
    bc = cfd.boundaries.periodic_boundary_conditions(nd)
    step = make_step()
    v = tuple(cfd.grids.GridVariable(cfd.grids.GridArray(cfd.DeviceArray.from_numpy(a), o, grid), bc)
              for a, o in zip(v0, grid.cell_faces))
    ref, rq = step.advance(v, nsteps, return_q=True)
    ref = [np.asarray(u.data) for u in ref] + [np.asarray(rq)]
    for name, a, b in zip(('u', 'v', 'q') if nd == 2 else ('u', 'v', 'w', 'q'), full, ref):
      same = np.array_equal(a, b)
      err = float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b.astype(np.float64)))
      print(f'{name}: bitwise={same} rel_l2={err:.2e}', flush=True)
      ok = ok and err < 1e-6
    print('MGPU PASS' if ok else 'MGPU FAIL', f'world={world} shape={shape} steps={nsteps}', flush=True)
  dist.barrier()
  st.close()
  dist.destroy_process_group()
  sys.exit(0 if ok else 1)


if __name__ == '__main__':
  main()
