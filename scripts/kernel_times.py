"""Per-kernel CUDA-event times of one step for a workload (perf iteration helper)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import jax_cfd_b200 as cfd
from jax_cfd_b200 import _lib
name = sys.argv[1] if len(sys.argv) > 1 else 'K8192'
wl = bench.WORKLOADS[name]
lib = _lib.lib()
shape, batch = wl['shape'], wl['batch']
if os.environ.get('CFD_KT_SHAPE'):
  shape = tuple(int(x) for x in os.environ['CFD_KT_SHAPE'].split(','))
grid = cfd.grids.Grid(shape, domain=((0.0, bench.TWO_PI),) * 2)
dt = cfd.equations.stable_time_step(wl['vmax'], 0.5, wl['nu'], grid)
forcing = cfd.forcings.sum_forcings(cfd.forcings.kolmogorov_forcing(grid, k=4), cfd.forcings.linear_forcing(grid, -0.1)) if wl['kolmogorov'] else None
step = cfd.equations.semi_implicit_navier_stokes(1.0, wl['nu'], dt, grid, forcing=forcing)
full = ((batch,) if batch > 1 else ()) + tuple(shape)
rs = np.random.RandomState(0)
k = np.arange(shape[1]) * (2 * np.pi / shape[1])
base = (np.sin(3 * k)[None, :] * np.ones((shape[0], 1))).astype(np.float32)
a = [_lib.DeviceArray.from_numpy(np.broadcast_to(base, full).copy()), _lib.DeviceArray.from_numpy(np.broadcast_to(base, full).copy())]
b = [_lib.DeviceArray(full) for _ in a]
plan = cfd.get_plan(grid, batch)
params = step.params()
st = _lib.Stream()
names = (ctypes.c_char_p * 8)(); ms = (ctypes.c_float * 8)(); nk = ctypes.c_int(0)
in_b = ctypes.c_int(0)
_lib.check(lib.cfd_repeated(plan.handle, st.handle, _lib.ptr_array(a), _lib.ptr_array(b), 4, ctypes.byref(params), ctypes.byref(in_b)))
_lib.check(lib.cfd_step_profile(plan.handle, st.handle, _lib.ptr_array(a), _lib.ptr_array(b), ctypes.byref(params), 10, 8, ms, names, ctypes.byref(nk)))
tot = sum(ms[i] for i in range(nk.value) if names[i].decode() in ('explicit_2d_lazy','rfft_rows','xlines','irfft_rows'))
cells = int(np.prod(full))
print(os.environ.get('CFD_XLINES_LE', '-'), os.environ.get('CFD_ROWS_LE', '-'), name, shape, ' '.join(f'{names[i].decode()}={ms[i]*1e3:.0f}us' for i in range(nk.value)), f'total={tot*1e3:.0f}us', f'{cells/tot/1e6:.1f} Gcell/s')
