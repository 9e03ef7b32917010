#!/usr/bin/env python
"""bench.py -- cell-updates/s of the staggered-grid FVM time step on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload K8192|D2048|E1024|K256]
  python bench.py --impl reference ...     # the CPU implementation (oracle port) timed beside it

A "step" is one forward-Euler projection step (explicit terms + pressure projection) of the whole
grid.  `value` has the state resident in HBM; `e2e` goes through the public host-array API with
H2D/D2H copies of the full state inside every timed step.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# The CPU legs mix an OpenMP runtime (stencils) with scipy's pocketfft thread pool: idle OpenMP
# workers that spin (libgomp's default) starve the FFT threads (+10 % at 8192^2, 75x at 256^2).
# Must be in the environment before libgomp initialises.
os.environ.setdefault('OMP_WAIT_POLICY', 'passive')

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TWO_PI = 2 * np.pi
WORKLOADS = {
    # name: shape, batch, viscosity, vmax, forcing, description
    'K8192': dict(shape=(8192, 8192), batch=1, nu=1e-4, vmax=7.0, kolmogorov=True, kpeak=4,
                  desc='2D Kolmogorov flow 8192x8192 (scale 1, k 4, linear -0.1, nu 1e-4), float32'),
    'D2048': dict(shape=(2048, 2048), batch=1, nu=1e-3, vmax=2.0, kolmogorov=False, kpeak=3,
                  desc='2D decaying turbulence 2048x2048 periodic (nu 1e-3, vmax 2), float32'),
    'K32768': dict(shape=(32768, 32768), batch=1, nu=1e-4, vmax=7.0, kolmogorov=True, kpeak=4,
                   desc='2D Kolmogorov flow 32768x32768 slab-sharded (nu 1e-4, vmax 7), float32'),
    'K256': dict(shape=(256, 256), batch=1, nu=1e-3, vmax=7.0, kolmogorov=True, kpeak=4,
                 desc='demo 2D Kolmogorov flow 256x256, float32'),
    'E1024': dict(shape=(256, 256), batch=1024, nu=1e-3, vmax=7.0, kolmogorov=True, kpeak=4,
                  desc='ensemble of 1024 Kolmogorov 256x256 trajectories, float32'),
    'TGV512': dict(shape=(512, 512, 512), batch=1, nu=1.0 / 1600, vmax=1.0, kolmogorov=False, kpeak=2,
                   smagorinsky=0.2,
                   desc='3D periodic Taylor-Green vortex 512^3 with Smagorinsky closure (cs 0.2), float32'),
    'D3D256': dict(shape=(256, 256, 256), batch=1, nu=1.0 / 1600, vmax=1.0, kolmogorov=False, kpeak=2,
                   desc='3D Taylor-Green vortex 256^3, no closure, float32'),
    'TGV256': dict(shape=(256, 256, 256), batch=1, nu=1.0 / 1600, vmax=1.0, kolmogorov=False, kpeak=2,
                   smagorinsky=0.2, desc='3D Taylor-Green vortex 256^3 with Smagorinsky closure, float32'),
}
# algorithmic HBM bytes per cell of each kernel (its inputs read once + outputs written once)
KERNEL_BYTES = {'explicit_2d': 20.0, 'explicit_2d_lazy': 24.0, 'rfft_rows': 8.0, 'xlines': 8.0,
                'irfft_rows': 8.0, 'correct': 20.0,
                # 3-D (first implementation: unfused divergence / correction, five FFT sweeps)
                'smag_nut': 16.0, 'explicit_3d': 28.0, 'divergence_3d': 16.0, 'rfft_z': 8.0, 'fft_y': 8.0,
                'xlines3': 8.0, 'ifft_y': 8.0, 'irfft_z': 8.0, 'correct_3d': 28.0}
CHAIN_KERNELS = ('explicit_2d_lazy', 'rfft_rows', 'xlines', 'irfft_rows')  # one chained step
STEP_BYTES_PER_CELL = 40.0  # 2-D, working set > L2
STEP_BYTES_PER_CELL_3D = 52.0  # ... 3-D


def synth_ic(shape, batch, seed, vmax, kpeak):
  """Filtered-noise velocity field in the spirit of initial_conditions.filtered_velocity_field
  (log-normal spectrum peaked at kpeak), built with float32 real FFTs on all cores; it is made
  divergence free by the device projection afterwards (see run_gpu)."""
  import scipy.fft
  nx, ny = shape
  rs = np.random.RandomState(seed)
  kx = TWO_PI * np.fft.fftfreq(nx, TWO_PI / nx)
  ky = TWO_PI * np.fft.rfftfreq(ny, TWO_PI / ny)
  k = np.sqrt(kx[:, None] ** 2 + ky[None, :] ** 2).astype(np.float32)
  with np.errstate(divide='ignore', invalid='ignore'):
    logk = np.log(k)
    filt = np.exp(-(np.log(kpeak) + 0.25 - logk) ** 2 / 0.5 - logk) / k
  filt[0, 0] = 0.0
  filt = filt.astype(np.float32)
  out = []
  for _ in range(2):
    comp = np.empty((batch,) + tuple(shape), np.float32)
    for b in range(batch):
      noise = rs.standard_normal(shape).astype(np.float32)
      spec = scipy.fft.rfft2(noise, workers=-1)
      spec *= filt
      comp[b] = scipy.fft.irfft2(spec, s=shape, workers=-1)
    comp *= np.float32(vmax / max(np.abs(comp).max(), 1e-30))
    out.append(comp if batch > 1 else comp[0])
  return out


class ClockSampler(threading.Thread):
  """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
  Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
       'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
       'clocks_event_reasons.sw_power_cap')

  def __init__(self, index=0):
    super().__init__(daemon=True)
    self.index, self.rows, self.proc = index, [], None

  def run(self):
    try:
      self.proc = subprocess.Popen(
          ['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
           '-lms', '200'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      for line in self.proc.stdout:
        self.rows.append([x.strip() for x in line.split(',')])
    except Exception:
      pass

  def stop(self):
    if self.proc is not None:
      self.proc.terminate()
    sm = [float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit()]
    mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    reasons = sorted({n for r in self.rows if len(r) >= 7 for n, f in zip(names, r[3:7]) if f == 'Active'})
    return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
            'reasons': reasons, 'samples': len(sm)}


def peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    with open(path) as f:
      return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
  return 6650.0, 'fallback (B200_PROFILING.md)'


CPU_SAMPLE_CELLS = 1 << 26  # the CPU legs run the whole grid up to 8192^2 cells, a slab of rows beyond


# weak scaling of the slab-decomposed path: 8192^2 cells per GPU
# (measured: 4 GPUs 32768x8192 1.82 ms/step vs 16384x16384 1.90 ms/step)
SLAB_SHAPES = {1: (8192, 8192), 2: (16384, 8192), 4: (32768, 8192), 8: (32768, 16384)}
# the north-star multi-GPU configuration (BASELINE config #4): 32768^2 over the GPUs of the box
SLAB_SHAPES_32K = {1: (32768, 32768), 2: (32768, 32768), 4: (32768, 32768), 8: (32768, 32768)}


def is_slab(name, world):
  return (world > 1 and name in ('K8192', 'TGV512', 'TGV256')) or name == 'K32768'


def workload_grid(wl, name, world):
  """(shape, domain) of the grid the arm runs at `world` GPUs -- shared by the GPU arm and the
  reference arm so that both describe the same configuration."""
  if is_slab(name, world) and len(wl['shape']) == 2:
    shape = (SLAB_SHAPES_32K if name == 'K32768' else SLAB_SHAPES)[world]
    if os.environ.get('CFD_SLAB_SHAPE'):  # tuning aid, e.g. CFD_SLAB_SHAPE=32768x8192
      shape = tuple(int(x) for x in os.environ['CFD_SLAB_SHAPE'].split('x'))
    return shape, tuple((0.0, TWO_PI * n / 8192.0) for n in shape)  # cell size of the 8192^2 case
  return tuple(wl['shape']), ((0.0, TWO_PI),) * len(wl['shape'])


def gpu_config(wl, name, world):
  """The `config` object of the JSON line (identical keys in both arms)."""
  shape, _ = workload_grid(wl, name, world)
  batch = wl['batch'] // world if (wl['batch'] > 1 and world > 1) else wl['batch']
  slab = is_slab(name, world)
  cells_per_gpu = int(np.prod(shape)) // world if slab else int(np.prod(shape)) * batch
  desc = wl['desc'] + (f' -- weak-scaled to {shape[0]}x{shape[1]}' if slab and name == 'K8192' else '')
  if slab and len(shape) == 3:
    desc += f' -- slab-decomposed along axis 0 over {world} GPUs (strong scaling)'
  return {'workload': f'{name}-slab' if slab and name == 'K8192' else name, 'description': desc,
          'grid': list(shape), 'batch': batch, 'cells_per_gpu': cells_per_gpu}


def cpu_reference(wl, steps, warmup, grid=None):
  """The oracle's OpenMP/pocketfft implementation of the same step on the host cores: the FULL
  grid of the workload when it has at most CPU_SAMPLE_CELLS cells (K8192: all of 8192^2), else a
  slab of its rows with the same row length and cell size."""
  sys.path.insert(0, os.path.join(ROOT, 'oracle'))
  import cfd_oracle
  import cpu_baseline
  full, fdom = grid if grid is not None else (tuple(wl['shape']), ((0.0, TWO_PI),) * 2)
  nx, ny = full
  while nx * ny > CPU_SAMPLE_CELLS and nx > 16:
    nx //= 2
  shape = (nx, ny)
  dom = ((0.0, fdom[0][1] * nx / full[0]), fdom[1])
  h = cfd_oracle.grid_step(shape, dom)
  dt = 0.5 * min(h) / wl['vmax']
  u, v = synth_ic(shape, 1, 0, wl['vmax'], wl['kpeak'])
  const = lin = None
  if wl['kolmogorov']:
    const = cfd_oracle.kolmogorov_field(shape, dom, 1.0, 4)
    lin = -0.1
  cores = os.cpu_count()
  cs = cpu_baseline.CpuStep(shape, h, dt, 1.0, wl['nu'], const, lin, workers=cores)
  u2, v2 = np.empty_like(u), np.empty_like(v)
  for _ in range(warmup):
    cs.step(u, v, u2, v2)
    u, u2, v, v2 = u2, u, v2, v
  t0 = time.perf_counter()
  for _ in range(steps):
    cs.step(u, v, u2, v2)
    u, u2, v, v2 = u2, u, v2, v
  dt_s = time.perf_counter() - t0
  assert np.isfinite(u).all()
  value = nx * ny * steps / dt_s / 1e9
  what = (f'the full {nx}x{ny} grid' if shape == full else
          f'a {nx}x{ny} slab of rows of the {full[0]}x{full[1]} grid (same row length and cell size)')
  return dict(value=value, unit='Gcell*step/s', cores=cores, kind='port', full_grid=(shape == full),
              sample=f'{steps} steps (+{warmup} warm-up) of {what}, same physics; OpenMP C stencils + '
                     f'scipy.fft pocketfft on {cores} threads (JAX is not installed, so the '
                     'reference jitted CPU path cannot run)'), dt_s / steps * 1e3


def run_reference(args, wl, name):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  # torchrun exports OMP_NUM_THREADS=1 to its workers, and both the OpenMP stencils and scipy's
  # pocketfft pool honour it from process start: re-run this arm in a child with a clean
  # environment so that the CPU reference really uses every host core.
  omp = os.environ.get('OMP_NUM_THREADS')
  if omp and omp.isdigit() and int(omp) < (os.cpu_count() or 1) and not os.environ.get('CFD_REF_CHILD'):
    env = {k: v for k, v in os.environ.items()
           if k not in ('OMP_NUM_THREADS', 'RANK', 'LOCAL_RANK', 'WORLD_SIZE', 'MASTER_ADDR', 'MASTER_PORT',
                        'GROUP_RANK', 'ROLE_RANK', 'LOCAL_WORLD_SIZE', 'ROLE_WORLD_SIZE')
           and not k.startswith('TORCHELASTIC')}
    env['CFD_REF_CHILD'] = '1'
    cmd = [sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--gpus', str(args.gpus),
           '--steps', str(args.steps), '--warmup', str(args.warmup), '--workload', name]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True)
    sys.stdout.write(out.stdout)
    sys.stdout.flush()
    if out.returncode != 0:
      sys.stderr.write(out.stderr)
      sys.exit(out.returncode)
    return
  if len(wl['shape']) != 2:
    print(json.dumps({'impl': 'reference', 'unavailable': 'the CPU port of the oracle is 2-D only'}), flush=True)
    return
  config = gpu_config(wl, name, args.gpus)
  cb, ms = cpu_reference(wl, args.steps, args.warmup, grid=workload_grid(wl, name, args.gpus))
  line = {
      'impl': 'reference', 'metric': 'cell-updates/sec', 'value': cb['value'], 'unit': 'Gcell*step/s',
      'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
      'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
      'data': 'synthetic', 'config': config,
      'cpu_baseline': cb,
      'e2e': {'value': cb['value'], 'unit': 'Gcell*step/s', 'h2d_bytes_per_step': 0,
              'd2h_bytes_per_step': 0},
  }
  print(json.dumps(line), flush=True)


def analytic_ic(shape, rows, vmax):
  """Smooth, globally consistent, nearly divergence-free field evaluated on a slab of rows
  (sum of Taylor-Green-like modes on the staggered grid); cheap at any size."""
  nx, ny = shape
  r0, r1 = rows
  x = (np.arange(r0, r1, dtype=np.float64) + 1.0) * (TWO_PI / nx)      # u: offset (1, .5)
  xc = (np.arange(r0, r1, dtype=np.float64) + 0.5) * (TWO_PI / nx)
  y = (np.arange(ny, dtype=np.float64) + 1.0) * (TWO_PI / ny)          # v: offset (.5, 1)
  yc = (np.arange(ny, dtype=np.float64) + 0.5) * (TWO_PI / ny)
  u = np.zeros((r1 - r0, ny))
  v = np.zeros((r1 - r0, ny))
  for a, b, amp in ((3, 4, 1.0), (5, 2, 0.6), (7, 9, 0.4), (11, 6, 0.25)):
    u += amp * b * np.sin(a * x)[:, None] * np.cos(b * yc)[None, :] / max(a, b)
    v -= amp * a * np.cos(a * xc)[:, None] * np.sin(b * y)[None, :] / max(a, b)
  scale = vmax / 2.3
  return (u * scale).astype(np.float32), (v * scale).astype(np.float32)


def slab_parity_check(cfd, _lib, dist, rank, world, local_rank):
  """Before timing: 4 steps of a 2048x1024 Kolmogorov problem through SlabStepper on all ranks and
  through the single-GPU path on rank 0 must agree BIT FOR BIT (same kernels, same arithmetic; the
  single-GPU path itself is checked against the oracle by tests/).  Returns the `parity` object."""
  shape, nsteps = (2048, 1024), 4
  dom = ((0.0, TWO_PI), (0.0, TWO_PI))
  grid = cfd.grids.Grid(shape, domain=dom)
  dt = cfd.equations.stable_time_step(7.0, 0.5, 1e-3, grid)
  forcing = cfd.forcings.sum_forcings(cfd.forcings.kolmogorov_forcing(grid, scale=1.0, k=4),
                                      cfd.forcings.linear_forcing(grid, -0.1))
  st = cfd.distributed.SlabStepper(grid, dt, 1.0, 1e-3, forcing, rank=rank, world=world, device=local_rank)
  u0, v0 = analytic_ic(shape, (0, shape[0]), 7.0)
  rs = np.random.RandomState(0)
  u0 = u0 + 0.1 * rs.standard_normal(shape).astype(np.float32)
  v0 = v0 + 0.1 * rs.standard_normal(shape).astype(np.float32)
  r0, r1 = st.rows
  st.load([u0[r0:r1], v0[r0:r1]])
  st.advance(nsteps // 2)
  st.advance(nsteps - nsteps // 2)
  outs, q = st.store(want_q=True)
  loc = [o.numpy() for o in outs] + [q.numpy()]
  gathered = [None] * world
  dist.all_gather_object(gathered, loc)
  st.close()
  res = None
  if rank == 0:
    full = [np.concatenate([g[i] for g in gathered], axis=0) for i in range(3)]
    bc = cfd.boundaries.periodic_boundary_conditions(2)
    step = cfd.equations.semi_implicit_navier_stokes(1.0, 1e-3, dt, grid, forcing=forcing)
    v = tuple(cfd.grids.GridVariable(cfd.grids.GridArray(_lib.DeviceArray.from_numpy(a), o, grid), bc)
              for a, o in zip((u0, v0), grid.cell_faces))
    ref, rq = step.advance(v, nsteps, return_q=True)
    ref = [np.asarray(x.data) for x in ref] + [np.asarray(rq)]
    bitwise = all(np.array_equal(a, b) for a, b in zip(full, ref))
    err = max(float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b.astype(np.float64)))
              for a, b in zip(full, ref))
    res = {'bitwise': bool(bitwise), 'max_rel_l2': err, 'grid': list(shape), 'steps': nsteps,
           'against': 'single-GPU path on rank 0 (itself checked against the oracle in tests/)'}
  flag = [res]
  dist.broadcast_object_list(flag, src=0)
  if not flag[0]['bitwise']:
    raise SystemExit(f'slab-decomposed step differs from the single-GPU step: {flag[0]}')
  return flag[0]


def taylor_green_slab(shape, rows):
  """u = sin x cos y cos z, v = -cos x sin y cos z, w = 0 sampled at grid.cell_faces (SURVEY.md
  section 8(d), TGV512) on the planes [rows) of axis 0."""
  n0, n1, n2 = shape
  r0, r1 = rows
  ax = lambda n, off, lo=0, hi=None: ((np.arange(lo, n if hi is None else hi, dtype=np.float64) + off) * (TWO_PI / n))
  xf, xc = ax(n0, 1.0, r0, r1), ax(n0, 0.5, r0, r1)
  yf, yc = ax(n1, 1.0), ax(n1, 0.5)
  zc = ax(n2, 0.5)
  u = (np.sin(xf)[:, None, None] * np.cos(yc)[None, :, None] * np.cos(zc)[None, None, :]).astype(np.float32)
  v = (-np.cos(xc)[:, None, None] * np.sin(yf)[None, :, None] * np.cos(zc)[None, None, :]).astype(np.float32)
  return u, v, np.zeros_like(u)


def slab_parity_check_3d(cfd, _lib, dist, rank, world, local_rank):
  """3-D counterpart: 2 steps of a 128x128x64 Taylor-Green problem with the Smagorinsky closure
  through SlabStepper on all ranks and through the single-GPU path on rank 0, bit for bit."""
  shape, nsteps = (128, 128, 64), 2
  dom = ((0.0, TWO_PI),) * 3
  grid = cfd.grids.Grid(shape, domain=dom)
  nu, cs = 1.0 / 1600, 0.2
  dt = cfd.equations.stable_time_step(1.0, 0.5, nu, grid)
  forcing = cfd._engine.ForcingFn([cfd._engine.SmagorinskyTerm(cs)])
  st = cfd.distributed.SlabStepper(grid, dt, 1.0, nu, forcing, rank=rank, world=world, device=local_rank)
  v0 = list(taylor_green_slab(shape, (0, shape[0])))
  rs = np.random.RandomState(0)
  v0 = [a + 0.05 * rs.standard_normal(shape).astype(np.float32) for a in v0]
  r0, r1 = st.rows
  st.load([a[r0:r1] for a in v0])
  st.advance(nsteps)
  outs, q = st.store(want_q=True)
  loc = [o.numpy() for o in outs] + [q.numpy()]
  gathered = [None] * world
  dist.all_gather_object(gathered, loc)
  st.close()
  res = None
  if rank == 0:
    full = [np.concatenate([g[i] for g in gathered], axis=0) for i in range(4)]
    bc = cfd.boundaries.periodic_boundary_conditions(3)
    step = cfd.subgrid_models.explicit_smagorinsky_navier_stokes(
        dt=dt, cs=cs, forcing=None, density=1.0, viscosity=nu, grid=grid)
    v = tuple(cfd.grids.GridVariable(cfd.grids.GridArray(_lib.DeviceArray.from_numpy(a), o, grid), bc)
              for a, o in zip(v0, grid.cell_faces))
    ref, rq = step.advance(v, nsteps, return_q=True)
    ref = [np.asarray(x.data) for x in ref] + [np.asarray(rq)]
    bitwise = all(np.array_equal(a, b) for a, b in zip(full, ref))
    err = max(float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b.astype(np.float64)))
              for a, b in zip(full, ref))
    res = {'bitwise': bool(bitwise), 'max_rel_l2': err, 'grid': list(shape), 'steps': nsteps,
           'against': 'single-GPU path on rank 0 (itself checked against the oracle in tests/)'}
  flag = [res]
  dist.broadcast_object_list(flag, src=0)
  if not flag[0]['bitwise']:
    raise SystemExit(f'slab-decomposed 3-D step differs from the single-GPU step: {flag[0]}')
  return flag[0]


def time_slab(cfd, _lib, dist, torch, wl, shape, dom, steps, warmup, rank, world, local_rank, full=True):
  """Times `steps` chained steps of the slab-decomposed grid (CUDA events on the launching stream,
  barrier + device sync both sides, max over ranks)."""
  lib = _lib.lib()
  grid = cfd.grids.Grid(shape, domain=dom)
  dt = cfd.equations.stable_time_step(wl['vmax'], 0.5, wl['nu'], grid)
  nd = len(shape)
  if nd == 2:
    forcing = cfd.forcings.sum_forcings(cfd.forcings.kolmogorov_forcing(grid, scale=1.0, k=4),
                                        cfd.forcings.linear_forcing(grid, -0.1))
  else:
    cs = wl.get('smagorinsky')
    forcing = cfd._engine.ForcingFn([cfd._engine.SmagorinskyTerm(cs)]) if cs else None
  st = cfd.distributed.SlabStepper(grid, dt, 1.0, wl['nu'], forcing, rank=rank, world=world,
                                   device=local_rank)
  ic = analytic_ic(shape, st.rows, wl['vmax']) if nd == 2 else taylor_green_slab(shape, st.rows)
  st.load(list(ic))
  del ic
  out = {'cells_local': int(np.prod(st.local_shape)), 'local_shape': st.local_shape}

  def barrier():
    st.sync()
    _lib.check(lib.cfd_device_sync())
    dist.barrier()

  st.advance(warmup)
  barrier()
  sampler = ClockSampler(local_rank)
  if rank == 0 and full:
    sampler.start()
    time.sleep(0.25)
  e0, e1 = _lib.Event(), _lib.Event()
  launches0 = lib.cfd_launch_count()
  barrier()
  e0.record(st.stream.handle)
  st.advance(steps)
  e1.record(st.stream.handle)
  barrier()
  ms_total = e0.elapsed_ms(e1)
  out['launches'] = lib.cfd_launch_count() - launches0
  out['clocks'] = sampler.stop() if rank == 0 and full else None
  t = torch.tensor([ms_total], device='cuda')
  dist.all_reduce(t, op=dist.ReduceOp.MAX)
  out['ms_total'] = float(t.item())
  out['kern'] = st.profile(2)
  outs = st.store()
  loc = outs[0].numpy()
  if full:
    # e2e at N GPUs: every step copies the rank's slab host->device, steps once, copies it back
    pin_in = [_lib.PinnedArray(st.local_shape) for _ in range(nd)]
    pin_out = [_lib.PinnedArray(st.local_shape) for _ in range(nd)]
    for pa, o in zip(pin_in, outs):
      pa.array[...] = o.numpy()
    e2e_steps = 3
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
      st.load([p.array for p in pin_in])
      st.advance(1)
      st.store(host_out=[p.array for p in pin_out])
      pin_in, pin_out = pin_out, pin_in
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device='cuda')
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    out['e2e_s'], out['e2e_steps'] = float(te.item()), e2e_steps
  finite = bool(np.isfinite(loc).all())
  out['finite'], out['umax'] = finite, float(np.abs(loc).max())
  t2 = torch.tensor([0.0 if finite else 1.0], device='cuda')
  dist.all_reduce(t2, op=dist.ReduceOp.MAX)
  assert float(t2.item()) == 0.0, 'non-finite state after the timed steps'
  dist.barrier()
  st.close()
  return out


def run_gpu_slab(args, wl, name):
  """N > 1: one rank per GPU, slab decomposition along axis 0 (jax_cfd_b200.distributed)."""
  import torch
  import torch.distributed as dist
  import jax_cfd_b200 as cfd
  from jax_cfd_b200 import _lib
  rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
  local_rank = int(os.environ.get('LOCAL_RANK', rank))
  torch.cuda.set_device(local_rank)
  if world == 1:
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29533')
    dist.init_process_group('nccl', rank=0, world_size=1, device_id=torch.device('cuda', local_rank))
  else:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
  lib = _lib.lib()
  _lib.check(lib.cfd_set_device(local_rank))
  shape, dom = workload_grid(wl, name, world)
  nd = len(shape)
  if nd == 3:
    parity = slab_parity_check_3d(cfd, _lib, dist, rank, world, local_rank) if world > 1 else None
  else:
    parity = slab_parity_check(cfd, _lib, dist, rank, world, local_rank) if world > 1 else None
  r = time_slab(cfd, _lib, dist, torch, wl, shape, dom, args.steps, args.warmup, rank, world, local_rank)
  # the north-star multi-GPU configuration itself (32768^2, BASELINE config #4) beside the
  # weak-scaling family member, on the same box in the same run
  k32 = None
  if name == 'K8192' and world == 8 and not os.environ.get('CFD_SKIP_K32768'):
    s32, d32 = workload_grid(WORKLOADS['K32768'], 'K32768', world)
    n32 = max(3, min(args.steps, 10))
    r32 = time_slab(cfd, _lib, dist, torch, WORKLOADS['K32768'], s32, d32, n32, 3, rank, world, local_rank,
                    full=False)
    ms32 = r32['ms_total'] / n32
    k32 = {'grid': list(s32), 'steps': n32, 'warmup': 3, 'ms_per_step': ms32,
           'value': float(np.prod(s32)) / (ms32 * 1e-3) / 1e9, 'unit': 'Gcell*step/s',
           'cells_per_gpu': r32['cells_local'], 'kernel_ms_rank0': r32['kern'],
           'roofline_step_frac': STEP_BYTES_PER_CELL * r32['cells_local'] / (ms32 * 1e-3) / 1e9 / peaks()[0]}
  if rank == 0:
    peak, peak_src = peaks()
    cells_local = r['cells_local']
    ms_step = r['ms_total'] / args.steps
    cells = cells_local * world
    value = cells * args.steps / (r['ms_total'] * 1e-3) / 1e9
    bpc = STEP_BYTES_PER_CELL if nd == 2 else STEP_BYTES_PER_CELL_3D
    step_gbs = bpc * cells_local / (ms_step * 1e-3) / 1e9
    # NVLink bytes per rank per step and direction: (world-1)/world of the packed spectrum
    # (4 B/cell), once to the line owners and once back
    nvl = 2 * cells_local * 4 * (world - 1) / world
    config = gpu_config(wl, name, world)
    config.update({
        'l2_policy': 'per-GPU working set (%d fields x %.0f MB) larger than L2 (126 MB)'
                     % (9 if nd == 2 else 14, cells_local * 4 / 1e6),
        'parallelism': f'slab decomposition along axis 0 over {world} GPUs; halo rows and the FFT '
                       'all-to-all move over NVLink through CUDA-IPC peer mappings driven by this '
                       'library\'s own kernels; device-side flags, no NCCL on the data path'})
    line = {
        'metric': 'cell-updates/sec', 'value': value, 'unit': 'Gcell*step/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step,
        'higher_is_better': True, 'scaling': 'weak' if nd == 2 else 'strong', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'config': config,
        'value_kind': 'chained steps (SlabStepper.advance), state resident in HBM',
        'roofline': {'bound': 'hbm', 'kernel': 'whole step (per GPU)', 'achieved': step_gbs, 'peak': peak,
                     'unit': 'GB/s', 'frac': step_gbs / peak, 'traffic': None, 'peak_source': peak_src,
                     'bytes_per_cell_model': bpc},
        'kernel_ms_rank0': r['kern'],
        'nvlink': {'bytes_per_gpu_per_step_each_direction': nvl,
                   'lower_bound_ms_at_770GBs': nvl / 770e9 * 1e3},
        'parity': parity, 'k32768': k32,
        'cpu_baseline': None,
        'e2e': {'value': cells * r['e2e_steps'] / r['e2e_s'] / 1e9, 'unit': 'Gcell*step/s',
                'h2d_bytes_per_step': nd * cells * 4, 'd2h_bytes_per_step': nd * cells * 4,
                'steps': r['e2e_steps'], 'ms_per_step': r['e2e_s'] / r['e2e_steps'] * 1e3,
                'api': 'SlabStepper.load(pinned numpy) / advance(1) / store(host_out=pinned numpy) on every rank'},
        'gpu_launches': int(r['launches']), 'clocks': r['clocks'],
        'diagnostics_after': {'finite': r['finite'], 'max_abs_u_rank0': r['umax']},
    }
    print(json.dumps(line), flush=True)
  dist.barrier()
  dist.destroy_process_group()


def run_gpu(args, wl, name):
  import jax_cfd_b200 as cfd
  from jax_cfd_b200 import _lib
  rank = int(os.environ.get('RANK', '0'))
  world = int(os.environ.get('WORLD_SIZE', '1'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  if is_slab(name, world):
    return run_gpu_slab(args, wl, name)
  dist = None
  if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
  lib = _lib.lib()
  _lib.check(lib.cfd_set_device(local_rank))
  _lib.require_device()

  shape, batch = wl['shape'], wl['batch']
  ndim = len(shape)
  if batch > 1 and world > 1:
    batch = batch // world  # ensemble members are independent: shard the batch, no collective
  grid = cfd.grids.Grid(shape, domain=((0.0, TWO_PI),) * ndim)
  dt = cfd.equations.stable_time_step(wl['vmax'], 0.5, wl['nu'], grid)
  forcing = None
  if wl['kolmogorov']:
    forcing = cfd.forcings.sum_forcings(cfd.forcings.kolmogorov_forcing(grid, scale=1.0, k=4),
                                        cfd.forcings.linear_forcing(grid, -0.1))
  if wl.get('smagorinsky'):
    step = cfd.subgrid_models.explicit_smagorinsky_navier_stokes(
        dt=dt, cs=wl['smagorinsky'], forcing=forcing, density=1.0, viscosity=wl['nu'], grid=grid)
  else:
    step = cfd.equations.semi_implicit_navier_stokes(1.0, wl['nu'], dt, grid, forcing=forcing)
  bc = cfd.boundaries.periodic_boundary_conditions(ndim)
  if ndim == 3:
    # Taylor-Green vortex sampled at the staggered offsets
    ax = [grid.axes(o) for o in grid.cell_faces]
    host = [(np.sin(ax[0][0])[:, None, None] * np.cos(ax[0][1])[None, :, None] * np.cos(ax[0][2])[None, None, :]).astype(np.float32),
            (-np.cos(ax[1][0])[:, None, None] * np.sin(ax[1][1])[None, :, None] * np.cos(ax[1][2])[None, None, :]).astype(np.float32),
            np.zeros(shape, np.float32)]
  else:
    host = synth_ic(shape, batch, 1000 + rank, wl['vmax'], wl['kpeak'])
  full = ((batch,) if batch > 1 else ()) + tuple(shape)
  cells = int(np.prod(full))

  def wrap(datas):
    return tuple(cfd.grids.GridVariable(cfd.grids.GridArray(d, o, grid), bc)
                 for d, o in zip(datas, grid.cell_faces))

  # device-resident state: project the synthetic field once so it is divergence free
  v = cfd.pressure.projection(wrap([_lib.DeviceArray.from_numpy(a) for a in host]))
  plan = cfd.get_plan(grid, batch, local_rank)
  params = step.params()
  import ctypes
  stream = _lib.Stream()
  a = [u.data for u in v]
  b = [_lib.DeviceArray(full) for _ in a]
  pa, pb = _lib.ptr_array(a), _lib.ptr_array(b)
  in_b = ctypes.c_int(0)

  def advance(n, src_a=True):
    x, y = (pa, pb) if src_a else (pb, pa)
    _lib.check(lib.cfd_repeated(plan.handle, stream.handle, x, y, n, ctypes.byref(params),
                                ctypes.byref(in_b)))
    return src_a != bool(in_b.value)  # True if the result is in a

  def barrier():
    stream.sync()
    _lib.check(lib.cfd_device_sync())
    if dist is not None:
      dist.barrier()

  in_a = advance(args.warmup, True)
  barrier()
  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()
    time.sleep(0.25)
  e0, e1 = _lib.Event(), _lib.Event()
  launches0 = lib.cfd_launch_count()
  barrier()
  e0.record(stream.handle)
  in_a = advance(args.steps, in_a)
  e1.record(stream.handle)
  barrier()
  ms_total = e0.elapsed_ms(e1)
  launches = lib.cfd_launch_count() - launches0
  clocks = sampler.stop() if rank == 0 else None
  if dist is not None:
    import torch
    t = torch.tensor([ms_total], device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
  ms_step = ms_total / args.steps
  value = cells * world * args.steps / (ms_total * 1e-3) / 1e9

  # sanity: the state is finite and divergence free after the timed steps
  diag = cfd.diagnostics(wrap(a if in_a else b))
  assert np.isfinite(diag['kinetic_energy']) and diag['max_abs_div'] < 1.0, diag

  line = None
  if rank == 0:
    peak, peak_src = peaks()
    # per-kernel CUDA-event times of one step (live, same stream)
    names = (ctypes.c_char_p * 8)()
    ms = (ctypes.c_float * 8)()
    nk = ctypes.c_int(0)
    src, dst = (pa, pb) if in_a else (pb, pa)
    _lib.check(lib.cfd_step_profile(plan.handle, stream.handle, src, dst, ctypes.byref(params), 5, 8,
                                    ms, names, ctypes.byref(nk)))
    kern = {names[i].decode(): float(ms[i]) for i in range(nk.value)}
    chain = {k: kern[k] for k in CHAIN_KERNELS if k in kern} or dict(kern)
    kgbs = {k: KERNEL_BYTES.get(k, 0.0) * cells / (t * 1e-3) / 1e9 for k, t in kern.items()}
    dom = max(chain, key=chain.get)       # the kernel with the largest share of the step
    worst = min(chain, key=lambda k: kgbs[k])  # ... and the one furthest from the roofline
    dom_bytes = KERNEL_BYTES.get(dom, 0.0) * cells
    achieved = dom_bytes / (kern[dom] * 1e-3) / 1e9
    # measured DRAM traffic per launch (ncu --set full, dram__bytes_read + write) of the commit
    # named in the file; dropped when the file is for another workload
    traffic = traffic_commit = None
    tfiles = sorted(f for f in os.listdir(os.path.join(ROOT, 'profiles')) if f.endswith('_traffic.json'))
    if tfiles:
      tj = json.load(open(os.path.join(ROOT, 'profiles', tfiles[-1])))
      if tj.get('workload') == name:
        traffic = tj['bytes_per_launch'].get(dom)
        traffic_commit = tj.get('commit', 'round-1 end state (9cb5c0e)')
    roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak, 'traffic': traffic, 'traffic_measured_at': traffic_commit,
                'peak_source': peak_src,
                'algorithmic_bytes_per_launch': dom_bytes, 'kernel_ms': kern[dom],
                'share_of_step': kern[dom] / sum(chain.values())}
    roofline_worst = {'kernel': worst, 'achieved': kgbs[worst], 'frac': kgbs[worst] / peak,
                      'kernel_ms': kern[worst], 'share_of_step': kern[worst] / sum(chain.values())}
    # bytes/cell model of the whole step: 52 in 3-D, 40 in 2-D when the
    # working set exceeds L2, 16 for batched members small enough to stay on chip (256^2)
    member_on_chip = ndim == 2 and batch > 1 and int(np.prod(shape)) * 4 * 3 <= 1 << 20
    bpc = 52.0 if ndim == 3 else (16.0 if member_on_chip else STEP_BYTES_PER_CELL)
    step_gbs = bpc * cells / (ms_step * 1e-3) / 1e9
    roofline_step = {'bound': 'hbm', 'bytes_per_cell_model': bpc,
                     'achieved': step_gbs, 'peak': peak, 'unit': 'GB/s', 'frac': step_gbs / peak,
                     'kernel_ms': kern, 'kernel_gbs': kgbs}
    # the same steps issued one call at a time (cfd_step: every call also materialises the
    # projected state, +20 B/cell) -- what a lone step_fn(v) costs
    sc_steps = max(3, min(args.steps, 10))
    x, y = (pa, pb) if in_a else (pb, pa)
    _lib.check(lib.cfd_step(plan.handle, stream.handle, x, y, None, ctypes.byref(params)))
    stream.sync()
    e0.record(stream.handle)
    for _ in range(sc_steps):
      x, y = y, x
      _lib.check(lib.cfd_step(plan.handle, stream.handle, x, y, None, ctypes.byref(params)))
    e1.record(stream.handle)
    sc_ms = e0.elapsed_ms(e1) / sc_steps
    single_call = {'ms_per_step': sc_ms, 'value': cells / (sc_ms * 1e-3) / 1e9, 'unit': 'Gcell*step/s',
                   'steps': sc_steps, 'api': 'cfd_step per step (step_fn(v) on device arrays)'}

    # e2e: public host-array API, pinned host buffers, H2D + D2H of the whole state every step
    hin = [_lib.PinnedArray(full) for _ in range(ndim)]
    hout = [_lib.PinnedArray(full) for _ in range(ndim)]
    for p, src_arr in zip(hin, host):
      p.array[...] = src_arr
    nb = cells * 4
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_step():
      _lib.check(lib.cfd_step_host(plan.handle, _lib.ptr_array([p.ptr for p in hin]),
                                   _lib.ptr_array([p.ptr for p in hout]), None, 1,
                                   ctypes.byref(params)))
    e2e_step()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
      e2e_step()
      hin, hout = hout, hin
    e2e_s = time.perf_counter() - t0
    e2e = {'value': cells * e2e_steps / e2e_s / 1e9, 'unit': 'Gcell*step/s',
           'h2d_bytes_per_step': ndim * nb, 'd2h_bytes_per_step': ndim * nb, 'steps': e2e_steps,
           'ms_per_step': e2e_s / e2e_steps * 1e3,
           'api': 'cfd_step_host (C ABI, pinned numpy in/out) == step_fn on host arrays'}
    cb = None
    if not args.no_cpu_baseline and ndim == 2:
      # free the pinned staging first: the CPU leg needs the host memory bandwidth to itself
      del hin, hout
      cb, _ = cpu_reference(wl, 5, 2, grid=workload_grid(wl, name, world))
    line = {
        'metric': 'cell-updates/sec', 'value': value, 'unit': 'Gcell*step/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': dict(gpu_config(wl, name, world), **{
            'l2_policy': 'working set (7 fields x %.0f MB) %s L2 (126 MB)' % (
                cells * 4 / 1e6, 'larger than' if cells * 4 * 7 > 126e6 else 'fits in'),
            'parallelism': 'single GPU' if world == 1 else (
                f'ensemble sharded over {world} GPUs ({batch} members each), no collective' if batch > 1
                else f'{world} independent domain replicas')}),
        'value_kind': 'chained steps (cfd_repeated == funcutils.repeated(step_fn, n)), state resident in HBM',
        'single_call': single_call,
        'roofline': roofline, 'roofline_worst': roofline_worst, 'roofline_step': roofline_step,
        'cpu_baseline': cb, 'e2e': e2e,
        'gpu_launches': int(launches), 'clocks': clocks,
        'diagnostics_after': diag,
    }
    print(json.dumps(line), flush=True)
  if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
  return line


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=50)
  ap.add_argument('--warmup', type=int, default=5)
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  ap.add_argument('--workload', default='K8192', choices=sorted(WORKLOADS))
  ap.add_argument('--no-cpu-baseline', action='store_true')
  args = ap.parse_args()
  args.warmup = max(args.warmup, 3)
  wl = WORKLOADS[args.workload]
  if args.impl == 'reference':
    run_reference(args, wl, args.workload)
  else:
    run_gpu(args, wl, args.workload)


if __name__ == '__main__':
  main()
