"""Deterministic stand-in for jax.random (NOT threefry: distribution matches, bits do not)."""
import numpy as _np


def PRNGKey(seed):
  return _np.array([0, seed], dtype=_np.uint32)


key = PRNGKey


def split(key, num=2):
  base = int(key[0]) * 1000003 + int(key[1])
  return [_np.array([(base * 7919 + i + 1) % (2**32), (base * 104729 + 31 * i + 7) % (2**32)],
                    dtype=_np.uint32) for i in range(num)]


def normal(key, shape=(), dtype=_np.float32):
  rs = _np.random.RandomState((int(key[0]) * 2654435761 + int(key[1])) % (2**32))
  return rs.standard_normal(shape).astype(dtype)


def uniform(key, shape=(), dtype=_np.float32, minval=0., maxval=1.):
  rs = _np.random.RandomState((int(key[0]) * 2654435761 + int(key[1])) % (2**32))
  return rs.uniform(minval, maxval, shape).astype(dtype)
