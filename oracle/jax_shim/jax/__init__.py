"""NumPy-backed stand-in for the parts of `jax` used by jax_cfd.base (test infrastructure)."""
import numpy as _np
from . import numpy, lax, tree_util, core, random, scipy, interpreters  # noqa
from . import tree  # noqa
Array = _np.ndarray


def named_call(fun, name=None):
  return fun


def jit(fun, *a, **k):
  return fun


def remat(fun, *a, **k):
  return fun


checkpoint = remat


class _Dev:
  platform = 'cpu'


def local_devices():
  return [_Dev()]


devices = local_devices


def vmap(fun, in_axes=0, out_axes=0):
  def mapped(*args):
    if isinstance(in_axes, int):
      axes = [in_axes] * len(args)
    else:
      axes = list(in_axes)
    n = None
    for a, ax in zip(args, axes):
      if ax is not None:
        leaves = tree_util.tree_leaves(a)
        n = leaves[0].shape[ax]
        break
    outs = []
    for i in range(n):
      sl = [tree_util.tree_map(lambda x: _np.take(x, i, axis=ax), a) if ax is not None else a
            for a, ax in zip(args, axes)]
      outs.append(fun(*sl))
    return tree_util.tree_map(lambda *xs: _np.stack(xs, axis=out_axes), *outs)
  return mapped


class _Config:
  def update(self, name, value):
    if name == 'jax_enable_x64':
      numpy.X64 = bool(value)
      self.jax_enable_x64 = bool(value)

  def parse_flags_with_absl(self):
    pass

  jax_enable_x64 = False


config = _Config()


def device_get(x):
  return x


def device_put(x, *a, **k):
  return x
