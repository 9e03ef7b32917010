import numpy as _np
from . import tree_util as _tu


class Precision:
  HIGHEST = 'highest'
  HIGH = 'high'
  DEFAULT = 'default'


def slice_in_dim(x, start_index, limit_index, stride=1, axis=0):
  sl = [slice(None)] * x.ndim
  sl[axis] = slice(start_index, limit_index, stride)
  return x[tuple(sl)]


def scan(f, init, xs=None, length=None):
  if xs is not None:
    n = len(_tu.tree_leaves(xs)[0])
  else:
    n = length
  carry, ys = init, []
  for i in range(n):
    x = None if xs is None else _tu.tree_map(lambda a: a[i], xs)
    carry, y = f(carry, x)
    ys.append(y)
  if ys and ys[0] is not None:
    ys = _tu.tree_map(lambda *a: _np.stack(a), *ys)
  else:
    ys = None
  return carry, ys
