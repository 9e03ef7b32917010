"""Stand-in for tree_math.{Vector,wrap,unwrap} (test infrastructure)."""
import operator as _op
from jax import tree_util as _tu


class Vector:
  def __init__(self, tree):
    self.tree = tree

  def _bin(self, other, f):
    if isinstance(other, Vector):
      return Vector(_tu.tree_map(f, self.tree, other.tree))
    return Vector(_tu.tree_map(lambda a: f(a, other), self.tree))

  def _rbin(self, other, f):
    return Vector(_tu.tree_map(lambda a: f(other, a), self.tree))

  def __add__(self, o): return self._bin(o, _op.add)
  def __radd__(self, o): return self._rbin(o, _op.add)
  def __sub__(self, o): return self._bin(o, _op.sub)
  def __rsub__(self, o): return self._rbin(o, _op.sub)
  def __mul__(self, o): return self._bin(o, _op.mul)
  def __rmul__(self, o): return self._rbin(o, _op.mul)
  def __truediv__(self, o): return self._bin(o, _op.truediv)
  def __rtruediv__(self, o): return self._rbin(o, _op.truediv)
  def __neg__(self): return Vector(_tu.tree_map(_op.neg, self.tree))


def _vec(x):
  return x if isinstance(x, Vector) else Vector(x)


def _tree(x):
  return x.tree if isinstance(x, Vector) else x


def wrap(fun, vector_argnums=0):
  """Pytree-callable view of a function written on Vectors."""
  def wrapped(v, *a, **k):
    return _tree(fun(_vec(v), *a, **k))
  return wrapped


def unwrap(fun, vector_argnums=0, out_vectors=True):
  """Vector-callable view of a function written on pytrees."""
  def unwrapped(v, *a, **k):
    return _vec(fun(_tree(v), *a, **k))
  return unwrapped
