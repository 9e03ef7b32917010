"""ctypes binding of libcfd_b200.so (include/cfd_b200.h) + a minimal device-array type.

The product path has NO CPU fallback: if the shared library is missing, or no CUDA device is
usable, every compute entry point raises.
"""
from __future__ import annotations

import ctypes
import os
import threading
from typing import Optional, Sequence

import numpy as np

MAX_DIM = 3
MAX_TERMS = 4
FORCE_SEPARABLE, FORCE_FIELD, FORCE_LINEAR, FORCE_SMAGORINSKY = 1, 2, 3, 4
IMPL_AUTO, IMPL_RFFT, IMPL_MATMUL = 0, 1, 2
# fast_diagonalization.transform's `implementation` strings (fast_diagonalization.py:90-125); 'fft'
# and 'rfft' compute the same transform of real data and share the line-FFT kernels
IMPLEMENTATIONS = {None: IMPL_AUTO, 'rfft': IMPL_RFFT, 'fft': IMPL_RFFT, 'matmul': IMPL_MATMUL}


def implementation_code(implementation) -> int:
  if implementation not in IMPLEMENTATIONS:
    raise ValueError(f'invalid implementation: {implementation}')
  return IMPLEMENTATIONS[implementation]

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('CFD_B200_LIB', os.path.join(_HERE, 'lib', 'libcfd_b200.so'))

c_float_p = ctypes.POINTER(ctypes.c_float)
c_float_pp = ctypes.POINTER(c_float_p)


class CfdError(RuntimeError):
  pass


class Params(ctypes.Structure):
  """struct cfd_params (include/cfd_b200.h)."""
  _fields_ = [
      ('dt', ctypes.c_double),
      ('density', ctypes.c_double),
      ('viscosity', ctypes.c_double),
      ('has_viscosity', ctypes.c_int32),
      ('n_terms', ctypes.c_int32),
      ('term_kind', ctypes.c_int32 * MAX_TERMS),
      ('linear_coef', ctypes.c_double),
      ('smagorinsky_cs', ctypes.c_double),
      ('sep_prof', (ctypes.c_void_p * MAX_DIM) * MAX_DIM),
      ('sep_scale', ctypes.c_float * MAX_DIM),
      ('has_sep', ctypes.c_int32 * MAX_DIM),
      ('field', ctypes.c_void_p * MAX_DIM),
      ('convect_dt', ctypes.c_double),
  ]


class Diag(ctypes.Structure):
  _fields_ = [('kinetic_energy', ctypes.c_double), ('enstrophy', ctypes.c_double),
              ('max_abs_div', ctypes.c_double), ('max_speed_sq', ctypes.c_double)]


_lib = None
_lock = threading.Lock()

_SIGS = {
    'cfd_last_error': (ctypes.c_char_p, []),
    'cfd_version': (ctypes.c_char_p, []),
    'cfd_device_count': (ctypes.c_int, []),
    'cfd_launch_count': (ctypes.c_uint64, []),
    'cfd_plan_create': (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int,
                                       ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_double),
                                       ctypes.c_int, ctypes.c_int]),
    'cfd_plan_create_impl': (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int,
                                            ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_double),
                                            ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    'cfd_plan_implementation': (ctypes.c_int, [ctypes.c_void_p]),
    'cfd_plan_destroy': (None, [ctypes.c_void_p]),
    'cfd_plan_workspace_bytes': (ctypes.c_size_t, [ctypes.c_void_p]),
    'cfd_step': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p),
                                ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p,
                                ctypes.POINTER(Params)]),
    'cfd_repeated': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p),
                                    ctypes.POINTER(ctypes.c_void_p), ctypes.c_int,
                                    ctypes.POINTER(Params), ctypes.POINTER(ctypes.c_int)]),
    'cfd_advance': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p),
                                   ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.POINTER(Params)]),
    'cfd_explicit_terms': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.POINTER(ctypes.c_void_p),
                                          ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(Params)]),
    'cfd_project': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p),
                                   ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p]),
    'cfd_transform_rfft': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p]),
    'cfd_transform_matmul': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p),
                                            ctypes.c_void_p]),
    'cfd_scale': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p),
                                 ctypes.c_double, ctypes.c_double, ctypes.POINTER(ctypes.c_void_p)]),
    'cfd_downsample_component': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                                ctypes.POINTER(ctypes.c_int64), ctypes.c_int, ctypes.c_int,
                                                ctypes.c_int]),
    'cfd_vorticity_2d': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.POINTER(ctypes.c_int64), ctypes.c_int, ctypes.c_double,
                                        ctypes.c_double]),
    'cfd_axpy': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p),
                                ctypes.c_int, ctypes.POINTER(ctypes.POINTER(ctypes.c_void_p)),
                                ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_void_p)]),
    'cfd_diagnostics': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(Diag)]),
    'cfd_step_host': (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p),
                                     ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, ctypes.c_int,
                                     ctypes.POINTER(Params)]),
    'cfd_step_profile': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.POINTER(ctypes.c_void_p),
                                        ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(Params),
                                        ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float),
                                        ctypes.POINTER(ctypes.c_char_p),
                                        ctypes.POINTER(ctypes.c_int)]),
    'cfd_dist_plan_create': (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int64),
                                            ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.c_int,
                                            ctypes.c_int]),
    'cfd_dist_plan_create_nd': (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int,
                                               ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_double),
                                               ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    'cfd_dist_handle_bytes': (ctypes.c_size_t, []),
    'cfd_dist_export': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    'cfd_dist_connect': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    'cfd_dist_load': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
    'cfd_dist_advance': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                        ctypes.POINTER(Params)]),
    'cfd_dist_store': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p),
                                      ctypes.c_void_p]),
    'cfd_dist_check': (ctypes.c_int, [ctypes.c_void_p]),
    'cfd_dist_profile': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(Params),
                                        ctypes.c_int, ctypes.POINTER(ctypes.c_float),
                                        ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_int)]),
    'cfd_malloc': (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]),
    'cfd_free': (ctypes.c_int, [ctypes.c_void_p]),
    'cfd_malloc_host': (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]),
    'cfd_free_host': (ctypes.c_int, [ctypes.c_void_p]),
    'cfd_memcpy_h2d': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    'cfd_memcpy_d2h': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    'cfd_memcpy_d2d': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    'cfd_memset': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p]),
    'cfd_stream_create': (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p)]),
    'cfd_stream_destroy': (ctypes.c_int, [ctypes.c_void_p]),
    'cfd_stream_sync': (ctypes.c_int, [ctypes.c_void_p]),
    'cfd_device_sync': (ctypes.c_int, []),
    'cfd_set_device': (ctypes.c_int, [ctypes.c_int]),
    'cfd_get_device': (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    'cfd_pointer_device': (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]),
    'cfd_event_create': (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p)]),
    'cfd_event_destroy': (ctypes.c_int, [ctypes.c_void_p]),
    'cfd_event_record': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    'cfd_event_elapsed_ms': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.POINTER(ctypes.c_float)]),
}

EXPORTED_SYMBOLS = tuple(sorted(_SIGS))


def lib():
  """Loads libcfd_b200.so (once).  Raises CfdError when it has not been built."""
  global _lib
  if _lib is None:
    with _lock:
      if _lib is None:
        if not os.path.exists(LIB_PATH):
          raise CfdError(f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; '
                         'g.build()"` (there is no CPU fallback)')
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
          fn = getattr(handle, name)
          fn.restype = res
          fn.argtypes = args
        _lib = handle
  return _lib


def check(code: int):
  if code != 0:
    raise CfdError(lib().cfd_last_error().decode())


def require_device():
  if lib().cfd_device_count() < 1:
    raise CfdError('no CUDA device is usable: the B200 path has no CPU fallback')


# ----------------------------------------------------------------------------- device arrays
class DeviceArray:
  """float32 array in device memory obtained through the C ABI (cfd_malloc)."""

  def __init__(self, shape: Sequence[int], dtype=np.float32, ptr: Optional[int] = None,
               owner=None, device: Optional[int] = None):
    self.shape = tuple(int(s) for s in shape)
    self.dtype = np.dtype(dtype)
    self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
    self._owner = owner
    if ptr is None:
      require_device()
      prev = current_device()
      self.device = prev if device is None else int(device)
      if self.device != prev:
        check(lib().cfd_set_device(self.device))
      try:
        p = ctypes.c_void_p()
        check(lib().cfd_malloc(ctypes.byref(p), max(self.nbytes, 16)))
      finally:
        if self.device != prev:
          check(lib().cfd_set_device(prev))
      self.ptr = p.value
      self._owned = True
    else:
      self.ptr = int(ptr)
      self._owned = False
      self.device = device

  @property
  def ndim(self):
    return len(self.shape)

  @property
  def size(self):
    return int(np.prod(self.shape, dtype=np.int64))

  @classmethod
  def from_numpy(cls, a: np.ndarray, stream=None) -> 'DeviceArray':
    a = np.ascontiguousarray(a)
    out = cls(a.shape, a.dtype)
    check(lib().cfd_memcpy_h2d(out.ptr, a.ctypes.data, a.nbytes, stream))
    check(lib().cfd_stream_sync(stream))
    return out

  def numpy(self, stream=None) -> np.ndarray:
    out = np.empty(self.shape, self.dtype)
    check(lib().cfd_memcpy_d2h(out.ctypes.data, self.ptr, self.nbytes, stream))
    check(lib().cfd_stream_sync(stream))
    return out

  def copy_(self, other, stream=None):
    if isinstance(other, np.ndarray):
      other = np.ascontiguousarray(other, dtype=self.dtype)
      assert other.nbytes == self.nbytes
      check(lib().cfd_memcpy_h2d(self.ptr, other.ctypes.data, self.nbytes, stream))
      check(lib().cfd_stream_sync(stream))
    else:
      assert device_nbytes(other) == self.nbytes
      check(lib().cfd_memcpy_d2d(self.ptr, device_ptr(other), self.nbytes, stream))
    return self

  def __array__(self, dtype=None, copy=None):
    a = self.numpy()
    return a if dtype is None else a.astype(dtype)

  def __len__(self):
    return self.shape[0]

  def __getitem__(self, k):
    """Leading-axis integer index -> a view (no copy) that keeps this array alive, e.g. one frame
    of a stacked trajectory."""
    if not isinstance(k, (int, np.integer)) or not self.shape:
      raise TypeError('DeviceArray supports integer indexing of the leading axis only')
    n = self.shape[0]
    k = int(k) + n if k < 0 else int(k)
    if not 0 <= k < n:
      raise IndexError('index out of range')
    sub = self.shape[1:]
    step = int(np.prod(sub, dtype=np.int64)) * self.dtype.itemsize
    return DeviceArray(sub, self.dtype, ptr=self.ptr + k * step, owner=self, device=self.device)

  @property
  def __cuda_array_interface__(self):
    return {'shape': self.shape, 'typestr': self.dtype.str, 'data': (self.ptr, False), 'version': 3}

  def __del__(self):
    if getattr(self, '_owned', False) and self.ptr and _lib is not None:
      try:
        _lib.cfd_free(self.ptr)
      except Exception:  # interpreter shutdown
        pass
      self.ptr = 0

  def __repr__(self):
    return f'DeviceArray(shape={self.shape}, dtype={self.dtype}, ptr=0x{self.ptr:x})'


def is_device_array(x) -> bool:
  return isinstance(x, DeviceArray) or (hasattr(x, '__cuda_array_interface__') and
                                        not isinstance(x, np.ndarray))


def is_c_contiguous(x) -> bool:
  """True when a device array is dense and row-major (the kernels read raw base pointers)."""
  if isinstance(x, DeviceArray):
    return True
  if hasattr(x, 'is_contiguous'):  # torch
    return bool(x.is_contiguous())
  cai = x.__cuda_array_interface__
  strides = cai.get('strides')
  if strides is None:
    return True
  itemsize = np.dtype(cai['typestr']).itemsize
  expect = itemsize
  for n, st in zip(reversed(cai['shape']), reversed(strides)):
    if n != 1 and st != expect:
      return False
    expect *= n
  return True


def device_of(x) -> int:
  """Index of the CUDA device that owns a device array's memory."""
  dev = getattr(x, 'device', None)
  if isinstance(dev, int):  # DeviceArray
    return dev
  if dev is not None and getattr(dev, 'type', None) == 'cuda' and dev.index is not None:  # torch
    return int(dev.index)
  d = ctypes.c_int(0)
  check(lib().cfd_pointer_device(device_ptr(x), ctypes.byref(d)))
  return int(d.value)


def current_device() -> int:
  d = ctypes.c_int(0)
  check(lib().cfd_get_device(ctypes.byref(d)))
  return int(d.value)


def device_ptr(x) -> int:
  if isinstance(x, DeviceArray):
    return x.ptr
  if hasattr(x, 'data_ptr'):  # torch tensor
    return int(x.data_ptr())
  return int(x.__cuda_array_interface__['data'][0])


def device_nbytes(x) -> int:
  if isinstance(x, DeviceArray):
    return x.nbytes
  if hasattr(x, 'element_size'):
    return int(x.numel() * x.element_size())
  cai = x.__cuda_array_interface__
  return int(np.prod(cai['shape'], dtype=np.int64)) * np.dtype(cai['typestr']).itemsize


def empty_like(x):
  """Device output buffer of the same kind as `x` (torch tensor -> torch tensor)."""
  if hasattr(x, 'data_ptr') and hasattr(x, 'new_empty'):
    return x.new_empty(tuple(x.shape))
  return DeviceArray(tuple(x.shape), np.float32, device=device_of(x))


def stream_of(x) -> Optional[int]:
  """The CUDA stream work on `x` should be enqueued on (torch: the current stream)."""
  if hasattr(x, 'data_ptr') and hasattr(x, 'new_empty'):
    import torch  # plumbing only: stream handle of the caller's tensors
    return int(torch.cuda.current_stream(x.device).cuda_stream) or None
  return None


def ptr_array(items) -> ctypes.Array:
  arr = (ctypes.c_void_p * len(items))()
  for i, it in enumerate(items):
    arr[i] = it if isinstance(it, int) or it is None else device_ptr(it)
  return arr


class Stream:
  def __init__(self):
    p = ctypes.c_void_p()
    check(lib().cfd_stream_create(ctypes.byref(p)))
    self.handle = p.value

  def sync(self):
    check(lib().cfd_stream_sync(self.handle))


class Event:
  def __init__(self):
    p = ctypes.c_void_p()
    check(lib().cfd_event_create(ctypes.byref(p)))
    self.handle = p.value

  def record(self, stream=None):
    check(lib().cfd_event_record(self.handle, stream))

  def elapsed_ms(self, stop: 'Event') -> float:
    ms = ctypes.c_float()
    check(lib().cfd_event_elapsed_ms(self.handle, stop.handle, ctypes.byref(ms)))
    return float(ms.value)


class PinnedArray:
  """Pinned host float32 array (cfd_malloc_host) exposed as numpy."""

  def __init__(self, shape, dtype=np.float32):
    self.shape = tuple(shape)
    nbytes = int(np.prod(self.shape, dtype=np.int64)) * np.dtype(dtype).itemsize
    p = ctypes.c_void_p()
    check(lib().cfd_malloc_host(ctypes.byref(p), max(nbytes, 16)))
    self.ptr = p.value
    buf = (ctypes.c_char * nbytes).from_address(self.ptr)
    self.array = np.frombuffer(buf, dtype=dtype).reshape(self.shape)

  def __del__(self):
    if getattr(self, 'ptr', 0) and _lib is not None:
      try:
        self.array = None
        _lib.cfd_free_host(self.ptr)
      except Exception:
        pass
      self.ptr = 0
