"""`jax.numpy` stand-in: NumPy with x64-disabled defaults."""
import numpy as _np
from numpy import *  # noqa
from numpy import fft, linalg  # noqa

ndarray = _np.ndarray
_DOWN = {_np.dtype('float64'): _np.float32, _np.dtype('complex128'): _np.complex64,
         _np.dtype('int64'): _np.int32}


X64 = False  # toggled by jax.config.update('jax_enable_x64', ...)


def _narrow(a):
  a = _np.asarray(a)
  if X64:
    return a
  t = _DOWN.get(a.dtype)
  return a.astype(t) if t is not None else a


class _X32(_np.ndarray):
  """ndarray whose ufunc results are narrowed to 32 bit (mimics jax weak typing of
  python scalars meeting int32 arrays, e.g. `jnp.arange(n) + 0.5` -> float32)."""
  __array_priority__ = 100

  def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
    base = tuple(_np.asarray(x) if isinstance(x, _X32) else x for x in inputs)
    res = getattr(ufunc, method)(*base, **kwargs)
    if isinstance(res, tuple):
      return tuple(_narrow(r).view(_X32) for r in res)
    return _narrow(res).view(_X32) if isinstance(res, _np.ndarray) else res


def asarray(a, dtype=None, **k):
  if dtype is not None:
    return _np.asarray(a, dtype=dtype)
  return _narrow(a)


def array(a, dtype=None, **k):
  if dtype is not None:
    return _np.array(a, dtype=dtype)
  return _narrow(_np.array(a))


def zeros(shape, dtype=None):
  return _np.zeros(shape, dtype=_np.float32 if dtype is None else dtype)


def ones(shape, dtype=None):
  return _np.ones(shape, dtype=_np.float32 if dtype is None else dtype)


def full(shape, fill_value, dtype=None):
  out = _np.full(shape, fill_value, dtype=dtype)
  return out if dtype is not None else _narrow(out)


def arange(*a, dtype=None, **k):
  out = _np.arange(*a, dtype=dtype, **k)
  return out if dtype is not None else _narrow(out).view(_X32)


def linspace(*a, dtype=None, **k):
  out = _np.linspace(*a, dtype=dtype, **k)
  return out if dtype is not None else _narrow(out)


def meshgrid(*a, **k):
  return list(_np.meshgrid(*a, **k))


def tensordot(a, b, axes=2, precision=None):
  return _np.tensordot(a, b, axes=axes)


def matmul(a, b, precision=None):
  return _np.matmul(a, b)


def dot(a, b, precision=None):
  return _np.dot(a, b)


def where(c, x=None, y=None):
  if x is None:
    return _np.where(c)
  out = _np.where(c, x, y)
  # python scalars are weakly typed in jax
  if not isinstance(x, _np.ndarray) and not isinstance(y, _np.ndarray):
    return _narrow(out)
  if isinstance(x, _np.ndarray) and not isinstance(y, (_np.ndarray, _np.generic)):
    return out.astype(x.dtype)
  if isinstance(y, _np.ndarray) and not isinstance(x, (_np.ndarray, _np.generic)):
    return out.astype(y.dtype)
  return out


def _wrap_ufunc(uf):
  def f(*a, **k):
    res = uf(*a, **k)
    if isinstance(res, tuple):
      return tuple(_narrow(r) if isinstance(r, (_np.ndarray, _np.generic)) else r for r in res)
    return _narrow(res) if isinstance(res, (_np.ndarray, _np.generic)) else res
  f.__name__ = uf.__name__
  return f


# GridArray.__array_ufunc__ dispatches through getattr(jnp, ufunc.__name__): narrow there so
# strongly typed np.float64 scalars meeting f32 arrays give f32, as in x64-disabled jax.
for _n in dir(_np):
  _o = getattr(_np, _n)
  if isinstance(_o, _np.ufunc):
    globals()[_n] = _wrap_ufunc(_o)
del _n, _o
